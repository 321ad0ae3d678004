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


def local_stats(data, mass, gravity_z):
    """[n_env_substeps_placeholder, contacts, impulses, kinetic+potential energy sum, max height] for one rank."""
    rows = data.rows(0, 13)                                           # [13, nfree, E]
    v2 = (rows[7:10] ** 2).sum(dim=0)
    # rotational energy needs the inertia; spheres/cubes of the reference are isotropic: callers pass mass only
    ke = 0.5 * mass * v2.sum()
    pe = -mass * gravity_z * rows[2].sum()
    return torch.stack([torch.zeros((), dtype=torch.float64, device=rows.device),
                        data.n_contacts.sum().to(torch.float64), data.n_impulses.sum().to(torch.float64),
                        (ke + pe).to(torch.float64), rows[2].max().to(torch.float64)])


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
