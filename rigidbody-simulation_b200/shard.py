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


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def _nvml_cpu_affinity(bdf):
    """CPUs the NVIDIA driver reports as local to the GPU at PCI address ``bdf`` (empty set when NVML cannot say)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByPciBusId(("0000" + bdf).encode() if len(bdf.split(":")[0]) == 4 else bdf.encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, 16)
        cpus = set()
        for i, w in enumerate(words):
            for bit in range(64):
                if (int(w) >> bit) & 1:
                    cpus.add(64 * i + bit)
        return cpus
    except Exception:
        return set()


def bind_to_gpu_numa(device_index):
    """Pin the calling thread (and every thread it starts afterwards) to the CPUs of the NUMA node that GPU
    ``device_index`` hangs off, so that pinned host buffers allocated from now on are first-touched in the memory next
    to that GPU's PCIe root.  With eight ranks streaming state in and out of one host this keeps each rank's H2D / D2H
    traffic off the inter-socket link.  Best effort: returns what was done (for the bench line) and never raises."""
    import os
    info = {"bound": False}
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        info["pci"] = bdf
        with open(base + "/numa_node") as f:
            info["numa_node"] = int(f.read().strip())
        with open(base + "/local_cpulist") as f:
            local = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus = local & allowed
        info["allowed_cpus"] = len(allowed)
        if info["numa_node"] < 0 or not cpus or cpus == allowed:
            # sysfs has no locality (typical inside a VM): ask the driver for the GPU's ideal CPU set instead
            nv = _nvml_cpu_affinity(bdf)
            if nv:
                info["source"] = "nvml"
                cpus = nv & allowed
        if not cpus or cpus == allowed:
            info["note"] = "single NUMA domain (or no locality information): nothing to bind"
            return info
        os.sched_setaffinity(0, cpus)
        info.update(bound=True, cpus=len(cpus))
    except Exception as exc:                       # sysfs not mounted, odd container: report and carry on unbound
        info["note"] = "not bound: %r" % (exc,)
    return info


def host_link_bandwidth(device, pinned, world_size=1, rank=0, reps=4):
    """Measured pinned-host <-> device copy bandwidth of every rank (GB/s, CUDA events): each rank ALONE (the others
    idle at a barrier) and all ranks TOGETHER -- the figure that says whether the host-buffer path of an N-GPU job is
    limited by this GPU's link or by what the host can feed to all of them at once."""
    dev_buf = torch.empty(pinned.shape, dtype=pinned.dtype, device=device)
    nbytes = pinned.numel() * pinned.element_size()

    def measure():
        out = []
        for src, dst in ((pinned, dev_buf), (dev_buf, pinned)):
            dst.copy_(src, non_blocking=True)                       # warm
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                dst.copy_(src, non_blocking=True)
            b.record()
            b.synchronize()
            out.append(reps * nbytes / (a.elapsed_time(b) * 1e-3) / 1e9)
        return out

    def measure_duplex():
        """both directions at once on two streams (what the chunk pipeline of the host-buffer drivers does): GB/s per direction"""
        other_host = torch.empty_like(pinned).pin_memory()
        other_dev = torch.empty_like(dev_buf)
        s_in, s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        torch.cuda.synchronize(device)
        with torch.cuda.stream(s_in):
            dev_buf.copy_(pinned, non_blocking=True)
            ev[0].record()
            for _ in range(reps):
                dev_buf.copy_(pinned, non_blocking=True)
            ev[1].record()
        with torch.cuda.stream(s_out):
            other_host.copy_(other_dev, non_blocking=True)
            ev[2].record()
            for _ in range(reps):
                other_host.copy_(other_dev, non_blocking=True)
            ev[3].record()
        torch.cuda.synchronize(device)
        return [reps * nbytes / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9, reps * nbytes / (ev[2].elapsed_time(ev[3]) * 1e-3) / 1e9]

    multi = dist.is_available() and dist.is_initialized() and world_size > 1
    alone = [0.0, 0.0]
    for r in range(world_size):
        if multi:
            torch.cuda.synchronize(device)
            dist.barrier()
        if r == rank:
            alone = measure()
    together = alone
    duplex = [0.0, 0.0]
    for r in range(world_size):
        if multi:
            torch.cuda.synchronize(device)
            dist.barrier()
        if r == rank:
            duplex = measure_duplex()
    if multi:
        torch.cuda.synchronize(device)
        dist.barrier()
        together = measure()
        t = torch.tensor(alone + together + duplex, dtype=torch.float64, device=device)
        rows = [torch.empty_like(t) for _ in range(world_size)]
        dist.all_gather(rows, t)
        table = [r.tolist() for r in rows]
    else:
        table = [alone + together + duplex]
    return {"bytes_per_copy": nbytes, "unit": "GB/s",
            "h2d_while_d2h_alone": [round(r[4], 2) for r in table], "d2h_while_h2d_alone": [round(r[5], 2) for r in table],
            "h2d_alone": [round(r[0], 2) for r in table], "d2h_alone": [round(r[1], 2) for r in table],
            "h2d_all_ranks_together": [round(r[2], 2) for r in table], "d2h_all_ranks_together": [round(r[3], 2) for r in table],
            "aggregate_together": round(sum(r[2] + r[3] for r in table), 1)}
