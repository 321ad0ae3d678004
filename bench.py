#!/usr/bin/env python
"""Benchmark of the hot path on B200: batched sphere-on-incline stepping (BASELINE config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over one batch: advance all E environments of this rank by
``--substeps`` integration steps (default 2048, the throughput horizon of SURVEY.md section 8(d)), issued as
launches of ``--fuse`` fused substeps.  Metric: env-substeps/s, whole job (all ranks), device-timed with CUDA
events, max over ranks.  Prints ONE JSON line on rank 0 ("Measurement").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env-substeps/s"
ENVS_PER_GPU = 1 << 20


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="environments per GPU")
    p.add_argument("--substeps", type=int, default=2048, help="integration steps per bench step")
    p.add_argument("--fuse", type=int, default=256, help="substeps fused per kernel launch")
    p.add_argument("--dtype", default="fp64", choices=["fp64", "fp32"])
    p.add_argument("--arith", default="fast", choices=["strict", "fast"],
                   help="strict = the reference's rounding sequence; fast = FMA/reciprocal re-association (<=1e-12/step)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-envs-per-core", type=int, default=64,
                   help="CPU baseline sample: environments per host core (64 x 2048 steps is ~12 s per core)")
    p.add_argument("--cpu-steps", type=int, default=2048, help="CPU baseline sample: steps per environment (= one bench step)")
    p.add_argument("--k1-launches", type=int, default=64, help="launches per step of the 1-substep-per-launch regime")
    return p.parse_args()


def workload_config(args, n_gpus):
    return {"workload": "sphere_incline_1M: models/sphere.xml body (r=0.2, density 50) over a plane tilted 0.7 rad, "
                        "randomised pose/velocity, per-env restitution U(0.5,1) and friction U(0,1), dt=0.009 "
                        "(BASELINE configs[1])",
            "envs_per_gpu": args.envs, "envs_total": args.envs * n_gpus, "substeps_per_step": args.substeps,
            "substeps_fused_per_launch": args.fuse, "arith": args.arith,
            "launch_schedule": "2 independent half-batch chains on 2 streams per GPU", "sharding": f"env-sharded x{n_gpus}, no collective on the step path",
            "l2": "L2 flushed (512 MiB write) between timed steps"}


# ------------------------------------------------------------------------------------------ reference arm
def reference_arm(args):
    """The reference's own CPU implementation of the path on the box's host cores.  The reference is a Python
    package that cannot travel to this box, so its step function is timed through the NumPy port in oracle/
    (same NumPy/SciPy calls, fake MuJoCo for contacts), one process per core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_baseline
    from rigidbody_simulation_b200 import synth
    cores = os.cpu_count() or 1
    sample = synth.sphere_incline(cores * args.cpu_envs_per_core)
    values, walls = [], []
    for i in range(args.warmup + args.steps):
        r = cpu_baseline.python_port_sphere_incline(sample, cores, args.cpu_envs_per_core, args.cpu_steps)
        if i >= args.warmup:
            values.append(r["value"])
            walls.append(r["wall_s"])
    v = sum(values) / len(values)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": METRIC, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(walls) / len(walls), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": v, "unit": METRIC, "cores": cores, "kind": "port", "sample": r["sample"]},
            "e2e": {"value": v, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock, power and throttle reasons every ~5 ms through NVML while the timed region runs."""

    def __init__(self, gpu_index):
        self.samples, self.gpu, self._stop, self._thread, self.err = [], gpu_index, False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as exc:                      # NVML missing: report that, never fake numbers
            self.err = repr(exc)
            return
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self):
        nv = self.nv
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.samples.append((time.time(), sm, pw, rs))
            except Exception as exc:
                self.err = repr(exc)
                return
            time.sleep(0.005)

    def stop(self, t0, t1):
        self._stop = True
        if self._thread is not None:
            self._thread.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable: %s" % self.err]}
        nv = self.nv
        rows = [x for x in self.samples if t0 <= x[0] <= t1] or self.samples[-3:]
        sm = sorted(r[1] for r in rows)
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(k for k, bit in names.items() if any(r[3] & bit for r in rows))
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.max_sm, "power_w_max": max(r[2] for r in rows),
                "samples": len(rows), "reasons": reasons}


# ------------------------------------------------------------------------------------------ B200 arm
def b200_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import rigidbody_simulation_b200 as rb
    from rigidbody_simulation_b200 import scenes, shard, stepper, synth

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator comes up; stdout must carry exactly
        # one JSON line, so fd 1 points at stderr until the communicator exists.
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    rb._lib.load()
    tdtype = torch.float64 if args.dtype == "fp64" else torch.float32
    esize = 8 if args.dtype == "fp64" else 4
    E, S, F = args.envs, args.substeps, args.fuse
    if S % F:
        raise SystemExit("--substeps must be a multiple of --fuse")

    # CPU baseline first (rank 0, before the timed GPU region so the host is quiet during it)
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import cpu_baseline
        cores = os.cpu_count() or 1
        sample = synth.sphere_incline(cores * args.cpu_envs_per_core)
        cpu = cpu_baseline.python_port_sphere_incline(sample, cores, args.cpu_envs_per_core, args.cpu_steps)
        try:
            cpu_native = cpu_baseline.c_port_sphere_incline(synth.sphere_incline(1 << 16), steps=200, threads=cores)
        except Exception as exc:                              # the native port is informative only
            cpu_native = {"error": str(exc)}

    # this rank's shard of the global environment index space (weak scaling: E per GPU)
    s = synth.sphere_incline(E, start=rank * E)
    model = scenes.sphere_on_incline(E, device=dev, dtype=tdtype)
    model.set_per_env(restitution=s["restitution"], friction=s["friction"])
    data = rb.BatchedData(model)
    qpos_h = torch.from_numpy(s["qpos"]).to(tdtype).pin_memory()
    qvel_h = torch.from_numpy(s["qvel"]).to(tdtype).pin_memory()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def reset_state():
        data.set_state(qpos_h, qvel_h)

    # Two independent chains of environment windows on two streams: the ragged last wave of CTAs of one chain's
    # launch is filled by the other chain (environments never interact, so no ordering is needed between them).
    chains = stepper.SplitChains(model, data, parts=2)

    def one_step(fuse):
        chains.fork()
        for _ in range(S // fuse):
            chains.step(dt=s["dt"], restitution=None, friction_coeff=None, contact_threshold=0.0, substeps=fuse,
                        count=False, arith=args.arith)
        chains.join()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        t0 = time.time()
        for a, b in evs:
            flush.fill_(1)                                   # evict L2 between timed steps
            a.record()
            fn()
            b.record()
        barrier()
        t1 = time.time()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        return shard.max_over_ranks(ms, dev), t0, t1

    # --- device-resident throughput (inputs already in HBM) -------------------------------------------
    reset_state()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.05)
    launches0 = rb.launch_count()
    total_ms, t0, t1 = timed(lambda: one_step(F), args.steps, args.warmup)
    launches = rb.launch_count() - launches0 - args.warmup * (S // F) * len(chains.ranges)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = world * E * S / (ms_per_step * 1e-3)

    # --- contact statistics of the timed regime (untimed, counters on) for the algorithmic flop count ----
    data.n_contacts.zero_()
    data.n_impulses.zero_()
    stepper.step_body_plane(model, data, -1, s["dt"], None, None, 0.0, substeps=F, count=True, arith=args.arith)
    torch.cuda.synchronize(dev)
    c_per = float(data.n_contacts.sum().item()) / (E * F)
    i_per = float(data.n_impulses.sum().item()) / (E * F)

    # --- the strict arithmetic policy (both inertia variants), same job, continuing from the same regime ---------
    others = {}
    for label, kw in (("strict", dict(arith="strict")), ("strict_isotropic_shortcut", dict(arith="strict", strict_inertia=False))):
        if args.arith == "strict" and label == "strict":
            kw = dict(arith="fast")
            label = "fast"

        def other_step(kw=kw):
            for _ in range(S // F):
                stepper.step_body_plane(model, data, -1, s["dt"], None, None, 0.0, substeps=F, count=False, **kw)

        n_other = max(2, args.steps // 3)
        other_ms, _, _ = timed(other_step, n_other, 1)
        others[label] = world * E * S / (other_ms / n_other * 1e-3)

    # --- K=1 streaming regime (HBM-bound): one launch per substep ---------------------------------------
    k1_launches = args.k1_launches

    def k1_step():
        for _ in range(k1_launches):
            stepper.step_body_plane(model, data, -1, s["dt"], None, None, 0.0, substeps=1, count=False, arith=args.arith)

    k1_ms, _, _ = timed(k1_step, max(3, args.steps), 3)
    k1_launch_ms = k1_ms / max(3, args.steps) / k1_launches

    # --- end to end through the host-buffer C-ABI call (H2D + S substeps + D2H every step) --------------
    def e2e_step():
        stepper.run_body_plane_host(model, qpos_h, qvel_h, S, dt=s["dt"], restitution=None, friction_coeff=None,
                                    contact_threshold=0.0, substeps=F, arith=args.arith)

    e2e_ms, _, _ = timed(e2e_step, args.steps, args.warmup)
    e2e_value = world * E * S / (e2e_ms / args.steps * 1e-3)

    # --- roofline ----------------------------------------------------------------------------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    fp_peak = stepper.fma_peak(dev, tdtype) / 1e12           # TFLOP/s, FMA = 2 flops, measured on this device
    # algorithmic work per env-substep: free flight 60 flops (incl. 4 div, 1 sqrt);
    # +72 per contact that produces an impulse, +22 per contact found separating (u_n >= 0); mul/add/div/sqrt = 1
    # flop each, i.e. an FMA-capable pipe could retire two of them per lane-cycle.  State 13+13 scalars and the two
    # per-env parameters (restitution, friction) cross HBM once per launch.
    flops_per_substep = 60.0 + 72.0 * i_per + 22.0 * (c_per - i_per)
    bytes_per_launch = E * (26 + 2) * esize
    launch_ms = ms_per_step / (S // F)           # time per F-substep advance of all E envs (two half-batch launches)
    fused_tflops = E * F * flops_per_substep / (launch_ms * 1e-3) / 1e12
    fused_gbs = bytes_per_launch / (launch_ms * 1e-3) / 1e9
    k1_gbs = bytes_per_launch / (k1_launch_ms * 1e-3) / 1e9

    # --- optional end-of-run gather of logged statistics: the only collective in the whole job -------------------
    stats = shard.gather_stats(shard.local_stats(model, data), env_substeps=E * S * args.steps)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if args.dtype == "fp64" else "f32", "data": "synthetic", "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": METRIC, "h2d_bytes_per_step": E * 13 * esize,
                    "d2h_bytes_per_step": E * 13 * esize, "ms_per_step": e2e_ms / args.steps,
                    "call": "rbs_run_body_plane_host via stepper.run_body_plane_host (pinned host qpos/qvel in the "
                            "reference layout)"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "fp64" if args.dtype == "fp64" else "fp32", "kernel": ("rbs::step_sphere_plane_pf_kernel<double,6,COUNT=false,THR=false,UNROLL=4> (plane-frame fast kernel)" if args.arith == "fast" else "rbs::step_body_plane_kernel<double,sphere,schemeA,iso>") if args.dtype == "fp64" else ("rbs::step_sphere_plane_pf2_kernel<6,COUNT=false,THR=false> (plane-frame fast kernel, two envs per thread on packed fp32x2 FFMA2, branch-free contact path)" if args.arith == "fast" and os.environ.get("RBS_PF_PACKED", "1") != "0" else "float instantiation of the fp64 kernel"),
                         "achieved": fused_tflops, "peak": fp_peak, "unit": "TFLOP/s", "frac": fused_tflops / fp_peak,
                         "peak_source": "FMA microbenchmark rbs_fma_probe run in this process (MEASURED_PEAKS.json has no "
                                        "CUDA-core peak)",
                         "flops_per_env_substep": flops_per_substep, "contacts_per_env_substep": c_per,
                         "impulses_per_env_substep": i_per, "launch_ms": launch_ms, "substeps_per_launch": F,
                         "hbm_GBps_of_same_launch": fused_gbs,
                         # dram__bytes_read.sum + dram__bytes_write.sum from the ncu --set full capture in
                         # profiles/r1_ncu_full_pf_kernel.csv: 62.9 + 7.7 MB per half-batch launch, two launches per
                         # advance of all envs (most of the 54.5 MB write-back is still in the 126 MB L2 when a launch
                         # ends).  Valid for the default 1,048,576-env fp64 shape only.
                         "traffic": 141.3e6 if (E == ENVS_PER_GPU and args.dtype == "fp64") else None},
            "roofline_k1": {"bound": "hbm", "kernel": "same kernel, 1 substep per launch (the reference's per-frame call)",
                            "achieved": k1_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": k1_gbs / hbm_peak,
                            "peak_source": hbm_src, "frac_of_nominal_8TBps": k1_gbs / 8000.0,
                            "bytes_per_env": (26 + 2) * esize, "launch_ms": k1_launch_ms,
                            "note": ("back-to-back launches; the state (%.0f MB) is larger than what the 126 MB L2 can keep "
                                     "between a launch's write and the next launch's read (ncu: reads come from DRAM)"
                                     % (E * 13 * esize / 1e6)) if E * 26 * esize > 126e6 else
                                    ("state %.0f MB read + written per launch fits the 126 MB L2: this figure is L2-assisted, "
                                     "not an HBM measurement" % (E * 13 * esize / 1e6)),
                            "env_steps_per_s": world * E / (k1_launch_ms * 1e-3),
                            # ncu capture of the one-substep launches (profiles/r1_summary.md): 125.8 MB read + ~60 MB written
                            "traffic": 185.9e6 if (E == ENVS_PER_GPU and args.dtype == "fp64") else None},
        }
        line["end_of_run_stats"] = stats
        line["other_policies"] = {
            "values": others, "unit": METRIC,
            "note": "strict = the reference's rounding sequence with the literal inv(R diag(I) R^T): bit-for-bit the C oracle "
                    "(profiles/r1_parity_report.md); strict_isotropic_shortcut = same but inv = (1/I)*Id for I1=I2=I3; fast = "
                    "FMA / reciprocal-multiply re-association, <= 1e-12 relative per step (tests/test_gpu_parity.py)"}
        if cpu is not None:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline_native"] = cpu_native
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: `python bench.py --gpus N` re-launches itself under torchrun, one rank per GPU
        import socket
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        os.execv(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                                  "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:])
    if args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()
