"""Environment sharding over ranks, and the optional end-of-run statistics gather.

Environments never interact, so the multi-GPU story is a partition: rank r of R owns the contiguous block of
global environment indices ``[start, start+count)``; initial states come from the index-keyed generator in
``synth`` (no scatter) and NO collective runs on the step path.  The only communication is an optional
end-of-run reduction of O(100 B) of counters / energy sums, over whatever backend the process group uses
(NCCL on the GPU box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_total, rank, world_size):
    """Contiguous block of rank ``rank``: sizes differ by at most one, earlier ranks get the remainder."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, rem = divmod(int(n_total), int(world_size))
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def local_stats(model, data):
    """[env-substeps placeholder, contacts, impulses, kinetic + potential energy sum, max height] for this rank's
    environments, computed by one pass of the CUDA statistics kernel (rbs_stats) over the SoA state."""
    import ctypes

    from . import _lib
    from .stepper import current_stream, rbs_dtype
    if data.device.type != "cuda":
        raise _lib.RbsError("run statistics are computed on a CUDA device only (no CPU fallback)")
    out = torch.tensor([0.0, 0.0, -1.0e300, 0.0, 0.0], dtype=torch.float64, device=data.device)
    n = data.nenv * data.nfree
    pe = model.per_env
    first = model.free_ids[0]
    per_body = data.layout == "body" or data.nfree == 1          # per-env arrays line up with the flattened columns
    mass_t = pe.get("mass") if per_body else None
    inertia_t = pe.get("inertia") if per_body else None
    if data.layout == "env" and data.nfree > 1 and any(model.body_mass[i] != model.body_mass[first] for i in model.free_ids):
        raise ValueError("statistics of env-layout scenes need identical bodies")
    I3 = (ctypes.c_double * 3)(*[float(v) for v in model.body_inertia[first]])
    G3 = (ctypes.c_double * 3)(*[float(v) for v in model.opt.gravity])
    P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    _lib.check(_lib.load().rbs_stats(rbs_dtype(model.dtype), n, P(data.state), n if data.layout == "env" else data.stride,
                                     P(mass_t), float(model.body_mass[first]), P(inertia_t), n, I3, G3,
                                     P(data.n_contacts), P(data.n_impulses), P(out), current_stream(model.device)))
    return torch.stack([torch.zeros((), dtype=torch.float64, device=data.device), out[3], out[4], out[0] + out[1], out[2]])


def gather_stats(stats, env_substeps):
    """End-of-run reduction: sums (substeps, contacts, impulses, energy) and max (height) over ranks.
    ``stats`` is the tensor from ``local_stats``; works with any initialised process group or none."""
    stats = stats.clone()
    stats[0] = float(env_substeps)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        sums, mx = stats[:4].clone(), stats[4:].clone()
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        stats = torch.cat([sums, mx])
    keys = ("env_substeps", "contacts", "impulses", "energy_sum", "max_height")
    return dict(zip(keys, stats.tolist()))


def max_over_ranks(value, device):
    """max of a python float over the process group (timing is always max-over-ranks)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
