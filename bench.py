#!/usr/bin/env python
"""Benchmark of the hot path on B200: the batched impulse / friction steppers on the BASELINE configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config NAME] [--impl reference]

One "step" = one pass of the hot path over one batch: every environment of this rank is reset to the config's
synthetic initial state (outside the timed region) and advanced by ``--substeps`` integration steps (default 2048,
the throughput horizon of SURVEY.md section 8(d)) in launches of ``--fuse`` fused substeps.  Metric: env-substeps/s,
whole job (all ranks), device-timed with CUDA events, max over ranks.  Rank 0 prints ONE JSON line (DESIGN.md,
"Measurement"): the headline config (``--config``, default BASELINE configs[1] = sphere on incline, 1M envs per GPU)
with ``roofline``, ``cpu_baseline``, ``e2e``; plus a ``configs`` object with value / roofline / cpu_baseline / e2e of
EVERY BASELINE config measured in the same run (two balls, cube bounce, cube incline, 64 spheres -- the last one
65,536 environments in total, sharded over the ranks: north_star configs[4]).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env-substeps/s"
NOMINAL_TFLOPS = {"fp64": 37.2, "fp32": 74.4}      # 148 SMs x 64 (128) FMA lanes x 2 flop x 1.965 GHz

# name -> BASELINE.json configs[] entry it measures, environments (per GPU, or in total when the config is
# strong-scaled), bodies per environment, substeps fused per launch
CONFIGS = {
    "sphere_incline": dict(baseline="configs[1]", envs=1 << 20, bodies=1, fuse=256, scaling="weak",
                           workload="sphere_incline_1M: models/sphere.xml body (r=0.2, density 50) over a plane tilted 0.7 rad, "
                                    "randomised pose/velocity, per-env restitution U(0.5,1) and friction U(0,1), dt=0.009 "
                                    "(BASELINE configs[1])"),
    "two_ball": dict(baseline="configs[2]", envs=1 << 20, bodies=2, fuse=256, scaling="weak",
                     workload="two_ball_1M: models/ball_collision.xml, shipped ICs (ball_collision.py:31-34) + perturbations, "
                              "sphere-sphere impulse + friction, e=1.0, mu=0.3, dt=0.01 (BASELINE configs[2])"),
    "cube_bounce": dict(baseline="configs[3]", envs=1 << 20, bodies=1, fuse=128, scaling="weak",
                        workload="cube_bounce_1M: models/cube.xml body (half 0.4) over a flat plane, random pose/velocity, "
                                 "8 vertex-plane candidates per step, e=0.2, mu=0.6, thr=1e-4, dt=0.009 (BASELINE configs[3])"),
    "cube_incline": dict(baseline="configs[3]", envs=1 << 20, bodies=1, fuse=128, scaling="weak",
                         workload="cube_incline_1M: models/cube.xml, plane and cube tilted 0.7 rad, shipped pose + perturbation, "
                                  "from rest (2-4 contacts per step), e=0.2, mu=0.6, thr=1e-4, dt=0.009 (BASELINE configs[3])"),
    "multi_sphere64": dict(baseline="configs[4]", envs=1 << 16, bodies=64, fuse=256, scaling="strong",
                           workload="multi_sphere64_65k: models/multi_sphere.xml scaled to 64 spheres per env (r=0.1) on a jittered "
                                    "4x4x4 lattice, all-pairs contacts, e=1.0, mu=0.0, dt=0.01; 65,536 envs IN TOTAL sharded over the "
                                    "GPUs (BASELINE configs[4])"),
}


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--config", default="sphere_incline", choices=list(CONFIGS), help="headline config of the JSON line")
    p.add_argument("--envs", type=int, default=None, help="environments per GPU (total for multi_sphere64); default: the BASELINE size")
    p.add_argument("--substeps", type=int, default=2048, help="integration steps per bench step")
    p.add_argument("--fuse", type=int, default=None, help="substeps fused per kernel launch (default per config)")
    p.add_argument("--dtype", default="fp64", choices=["fp64", "fp32"])
    p.add_argument("--arith", default="fast", choices=["strict", "fast"],
                   help="strict = the reference's rounding sequence; fast = FMA/reciprocal re-association (<=1e-12/step)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-other-configs", action="store_true", help="measure the headline config only")
    p.add_argument("--other-steps", type=int, default=3, help="timed steps of each non-headline config")
    p.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU work per core of the headline CPU baseline")
    p.add_argument("--k1-launches", type=int, default=64, help="launches per step of the 1-substep-per-launch regime")
    return p.parse_args()


def workload_config(name, args, n_gpus, envs_per_gpu, fuse):
    c = CONFIGS[name]
    return {"workload": c["workload"], "config": name, "envs_per_gpu": envs_per_gpu,
            "envs_total": envs_per_gpu * n_gpus if c["scaling"] == "weak" else c["envs"] if args.envs is None else args.envs,
            "bodies_per_env": c["bodies"], "substeps_per_step": args.substeps, "substeps_fused_per_launch": fuse, "arith": args.arith,
            "reset": "every step starts from the config's initial state (reset outside the timed region)",
            "sharding": f"env-sharded x{n_gpus} ({c['scaling']} scaling), no collective on the step path",
            "l2": "L2 flushed (512 MiB write) between timed steps"}


# ------------------------------------------------------------------------------------------ CPU side (oracle/)
# kind, environments per core and steps of the FULL sample: ~10-12 s of CPU work per core (measured: the Python step
# costs ~110 us per sphere step, ~200 us per cube step, ~120 us per two-ball step, ~0.25 s per 64-sphere step).
# ``scale`` < 1 takes fewer environments per core (fewer steps once there is only one environment left).
CPU_SAMPLES = {"sphere_incline": ("sphere_incline", 48, 2048), "two_ball": ("two_ball", 48, 2048), "cube_bounce": ("cube", 24, 2048),
               "cube_incline": ("cube", 24, 2048), "multi_sphere64": ("multi_sphere", 1, 48)}


def cpu_baseline_of(name, scale=1.0, cores=None):
    """The reference's CPU implementation of config ``name`` on all host cores, bounded sample (oracle/cpu_baseline.py):
    the installed reference (baseline/_ref, kind "reference") when present, else the NumPy port (kind "port")."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_baseline
    from rigidbody_simulation_b200 import synth
    kind, per_core, steps = CPU_SAMPLES[name]
    if per_core * scale < 1:
        steps = max(2, int(round(steps * per_core * scale)))
    per_core = max(1, int(round(per_core * scale)))
    cores = cores or os.cpu_count() or 1
    n = cores * per_core
    if name == "sphere_incline":
        sample, extra = synth.sphere_incline(n), {}
    elif name == "two_ball":
        sample, extra = synth.two_ball(n), {}
    elif name.startswith("cube"):
        sample = synth.cube(n, kind=name.split("_")[1])
        extra = {"theta": sample["theta"]}
    else:
        sample, extra = synth.multi_sphere(n, n_body=64, friction=0.0), {"n_body": 64, "friction": 0.0}
    return cpu_baseline.run_config(kind, sample, cores=cores, envs_per_core=per_core, steps=steps, extra=extra)


def reference_arm(args):
    """``--impl reference``: the reference's own CPU implementation of the path on the box's host cores, on the headline
    config.  Each step is a bounded sample of the workload; rank 0 alone runs and prints."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    name = args.config
    c = CONFIGS[name]
    envs = args.envs if args.envs is not None else c["envs"]
    values, walls, r = [], [], None
    scale = 0.25 * args.cpu_seconds / 12.0              # ~3 s of CPU work per core and step: K steps stay within minutes
    for i in range(args.warmup + args.steps):
        r = cpu_baseline_of(name, scale=scale * (0.25 if i < args.warmup else 1.0))
        if i >= args.warmup:
            values.append(r["value"])
            walls.append(r["wall_s"])
    v = sum(values) / len(values)
    per_gpu = envs if c["scaling"] == "weak" else -(-envs // args.gpus)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": METRIC, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(walls) / len(walls), "higher_is_better": True,
            "scaling": c["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(name, args, args.gpus, per_gpu, args.fuse or c["fuse"]),
            "cpu_baseline": {"value": v, "unit": METRIC, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
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


# ------------------------------------------------------------------------------------------ workloads
class Workload:
    """One BASELINE config on this rank: scene, device state, pinned host copies of the initial state, the launch
    schedule of one bench step, the host-buffer (e2e) call, and the algorithmic work model of DESIGN.md section 3."""

    def __init__(self, name, args, rank, world, dev, tdtype):
        import torch
        import rigidbody_simulation_b200 as rb
        from rigidbody_simulation_b200 import scenes, shard, stepper, synth
        from rigidbody_simulation_b200.src.simulation import ball_collision, multi_sphere_bounce
        self.name, self.args, self.dev, self.stepper, self.torch = name, args, dev, stepper, torch
        c = CONFIGS[name]
        self.bodies, self.S = c["bodies"], args.substeps
        self.F = args.fuse if (args.fuse and name == args.config) else c["fuse"]
        envs = args.envs if (args.envs is not None and name == args.config) else c["envs"]
        if c["scaling"] == "weak":
            start, self.E = rank * envs, envs
        else:
            start, self.E = shard.shard_range(envs, rank, world)
        self.E_job = envs * world if c["scaling"] == "weak" else envs
        if self.S % self.F:
            raise SystemExit("--substeps must be a multiple of the fuse factor")
        E = self.E
        if name == "sphere_incline":
            s = synth.sphere_incline(E, start=start)
            self.model = scenes.sphere_on_incline(E, device=dev, dtype=tdtype)
            self.model.set_per_env(restitution=s["restitution"], friction=s["friction"])
            self.data = rb.BatchedData(self.model)
            self.step_kw = dict(dt=s["dt"], restitution=None, friction_coeff=None, contact_threshold=0.0)
        elif name.startswith("cube"):
            s = synth.cube(E, start=start, kind=name.split("_")[1])
            self.model = scenes.cube_on_plane(E, theta=s["theta"], device=dev, dtype=tdtype)
            self.data = rb.BatchedData(self.model)
            self.step_kw = dict(dt=s["dt"], restitution=0.2, friction_coeff=0.6, contact_threshold=1e-4)
        elif name == "two_ball":
            s = synth.two_ball(E, start=start)
            self.model, self.data = ball_collision.build(E, device=dev, dtype=tdtype)
        else:
            s = synth.multi_sphere(E, n_body=64, start=start, friction=0.0)
            self.model, self.data = multi_sphere_bounce.build(E, device=dev, dtype=tdtype, n_body=64)
        self.qpos0 = torch.from_numpy(s["qpos"]).to(tdtype).pin_memory()
        self.qvel0 = torch.from_numpy(s["qvel"]).to(tdtype).pin_memory()
        self.qpos_h, self.qvel_h = self.qpos0.clone().pin_memory(), self.qvel0.clone().pin_memory()
        self.chains = stepper.SplitChains(self.model, self.data, parts=2) if self.bodies == 1 else None
        self.esize = 8 if tdtype == torch.float64 else 4

    # -- state ---------------------------------------------------------------------------------------
    def reset(self):
        self.data.set_state(self.qpos0, self.qvel0)
        self.qpos_h.copy_(self.qpos0)
        self.qvel_h.copy_(self.qvel0)

    def zero_counters(self):
        self.data.n_contacts.zero_()
        self.data.n_impulses.zero_()

    # -- one bench step ------------------------------------------------------------------------------
    def launches_per_step(self):
        return (self.S // self.F) * (len(self.chains.ranges) if self.chains else 1)

    def launch(self, k, count=False, arith=None, **kw):
        """advance every environment by k substeps with ONE launch per chain"""
        st, arith = self.stepper, arith or self.args.arith
        if self.chains is not None and not kw:
            self.chains.step(substeps=k, count=count, arith=arith, **self.step_kw)
        elif self.bodies == 1:
            st.step_body_plane(self.model, self.data, -1, self.step_kw["dt"], self.step_kw["restitution"], self.step_kw["friction_coeff"],
                               self.step_kw["contact_threshold"], substeps=k, count=count, arith=arith, **kw)
        elif self.name == "two_ball":
            st.step_two_ball(self.model, self.data, 0.01, 1.0, 0.3, radius=0.1, substeps=k, count=count, arith=arith)
        else:
            st.step_multi_sphere(self.model, self.data, 0.01, 1.0, 0.0, substeps=k, count=count, arith=arith, **kw)

    def step(self, count=False, arith=None, **kw):
        if self.chains is not None and not kw:
            self.chains.fork()
        for _ in range(self.S // self.F):
            self.launch(self.F, count, arith, **kw)
        if self.chains is not None and not kw:
            self.chains.join()

    def e2e_step(self):
        """The reference-facing call with HOST buffers: H2D + S substeps + D2H inside (include/rbsim_b200.h rbs_run_*_host)."""
        st, a = self.stepper, self.args
        if self.bodies == 1:
            st.run_body_plane_host(self.model, self.qpos_h, self.qvel_h, self.S, substeps=self.F, arith=a.arith, **self.step_kw)
        elif self.name == "two_ball":
            st.run_two_ball_host(self.model, self.qpos_h, self.qvel_h, self.S, dt=0.01, restitution=1.0, friction=0.3, radius=0.1,
                                 substeps=self.F, arith=a.arith)
        else:
            st.run_multi_sphere_host(self.model, self.qpos_h, self.qvel_h, self.S, dt=0.01, restitution=1.0, friction=0.0,
                                     substeps=self.F, arith=a.arith)

    def e2e_reset(self):
        self.qpos_h.copy_(self.qpos0)
        self.qvel_h.copy_(self.qvel0)

    # -- work model (mul / add / div / sqrt = 1 flop each, FMA = 2) --------------------
    def rates(self):
        """(contacts, impulses) per env-substep from the counters of a counted run (two balls: ground hits, pair hits)"""
        n = float(self.E * self.S)
        return float(self.data.n_contacts.sum().item()) / n, float(self.data.n_impulses.sum().item()) / n

    def flops_per_env_substep(self, c, i):
        if self.name == "sphere_incline":     # free flight 60; +72 per impulse, +22 per contact found separating
            return 60.0 + 72.0 * i + 22.0 * (c - i)
        if self.name.startswith("cube"):      # + rotation 42 and 8 x 23 for the vertex scan; per contact arm 9 + impulse 72 / separating 22
            return 60.0 + 226.0 + 81.0 * i + 31.0 * (c - i)
        if self.name == "two_ball":           # gravity 6 + ground tests 2 + pair test 12 + integrate 12; ~120 per ground impulse, ~190 per pair hit
            return 32.0 + 120.0 * c + 190.0 * i
        # 64 bodies: free flight 60 + 63 pair rejects x 9 per body (SURVEY 8(d) all-pairs accounting) + 97 per impulse, 47 per separating contact
        return 64.0 * (60.0 + 567.0) + 97.0 * i + 47.0 * (c - i)

    def bytes_per_launch(self):
        per_env = {"sphere_incline": 26 + 2, "two_ball": 36, "cube_bounce": 26, "cube_incline": 26, "multi_sphere64": 64 * 26}[self.name]
        return self.E * per_env * self.esize

    def kernel(self):
        a, f64 = self.args, self.esize == 8
        t = "double" if f64 else "float"
        if a.arith == "strict":
            return {"sphere_incline": f"rbs::step_body_plane_kernel<{t},sphere,schemeA,literal inertia>",
                    "two_ball": f"rbs::step_two_ball_kernel<{t}>", "multi_sphere64": f"rbs::step_multi_sphere_kernel<{t},literal inertia,256>"}.get(
                        self.name, f"rbs::step_body_plane_kernel<{t},box,schemeA,literal inertia>")
        if self.name == "sphere_incline":
            return ("rbs::step_sphere_plane_pf_kernel<double,6,COUNT=false,THR=false,UNROLL=4> (plane-frame fast kernel)" if f64 else
                    "rbs::step_sphere_plane_pf2_kernel<6,COUNT=false,THR=false> (plane-frame fast kernel, two envs per thread on packed "
                    "fp32x2 FFMA2, branch-free contact path)")
        return {"two_ball": f"rbs::step_two_ball_fast_kernel<{t},GZ=true>", "multi_sphere64": f"rbs::step_multi_sphere_fast_kernel<{t},256>"}.get(
            self.name, f"rbs::step_box_plane_pf_kernel<{t},6>")


def ncu_figures():
    """per-kernel ncu figures (FP64 pipe utilisation, DRAM traffic per launch) of the committed captures, keyed by
    config and dtype: profiles/ncu_figures.json, written by hand from the ncu CSVs it cites"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_figures.json")) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


# ------------------------------------------------------------------------------------------ B200 arm
def b200_arm(args):
    import torch
    import torch.distributed as dist
    import rigidbody_simulation_b200 as rb
    from rigidbody_simulation_b200 import shard, stepper

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
    names = [args.config] + ([] if args.no_other_configs else [n for n in CONFIGS if n != args.config])

    # CPU baselines first (rank 0, all host cores, before this process pins itself to the GPU's NUMA node and before
    # the timed GPU regions, so the host is quiet during them)
    cpu = {}
    if rank == 0 and not args.no_cpu_baseline:
        for n in names:
            try:
                cpu[n] = cpu_baseline_of(n, scale=(args.cpu_seconds / 12.0) if n == args.config else 0.1)
            except Exception as exc:
                cpu[n] = {"error": repr(exc)}
        try:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import cpu_baseline
            from rigidbody_simulation_b200 import synth
            cpu_native = cpu_baseline.c_port_sphere_incline(synth.sphere_incline(1 << 16), steps=200, threads=os.cpu_count() or 1)
        except Exception as exc:                              # the native port is informative only
            cpu_native = {"error": str(exc)}

    # pinned host buffers are allocated after binding to the cores (and thereby the memory) next to this rank's GPU
    numa = shard.bind_to_gpu_numa(local)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup, before=None):
        """CUDA-event time of ``steps`` calls of fn (ms, max over ranks); ``before`` runs ahead of every call, untimed"""
        for _ in range(warmup):
            if before:
                before()
            fn()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        t0 = time.time()
        for a, b in evs:
            if before:
                before()
            flush.fill_(1)                                   # evict L2 between timed steps
            a.record()
            fn()
            b.record()
        barrier()
        t1 = time.time()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        return shard.max_over_ranks(ms, dev), t0, t1

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    fp_peak = stepper.fma_peak(dev, tdtype) / 1e12           # TFLOP/s, FMA = 2 flops, measured on this device
    ncu = ncu_figures()

    def measure(name, headline):
        """value, roofline, e2e (and for the headline: clocks, launches, other policies, K=1 regime) of one config"""
        w = Workload(name, args, rank, world, dev, tdtype)
        steps, warm = (args.steps, args.warmup) if headline else (args.other_steps, max(1, min(args.warmup, 2)))
        out = {}
        sampler = ClockSampler(local)
        if headline and rank == 0:
            sampler.start()
            time.sleep(0.05)
        launches0 = rb.launch_count()
        total_ms, t0, t1 = timed(w.step, steps, warm, before=w.reset)
        launches = rb.launch_count() - launches0 - warm * w.launches_per_step()
        if headline:
            out["clocks"] = sampler.stop(t0, t1) if rank == 0 else None
            out["gpu_launches"] = launches
        ms = total_ms / steps
        out["ms_per_step"] = ms
        out["value"] = w.E_job * w.S / (ms * 1e-3)
        # contact statistics of the measured horizon (first S substeps from the reset state) and of the next one
        w.reset(); w.zero_counters(); w.step(count=True); torch.cuda.synchronize(dev)
        c0, i0 = w.rates()
        w.zero_counters(); w.step(count=True); torch.cuda.synchronize(dev)
        c1, i1 = w.rates()
        flops = w.flops_per_env_substep(c0, i0)
        launch_ms = ms / (w.S // w.F)                          # one F-substep advance of all of this rank's environments
        tflops = w.E * w.F * flops / (launch_ms * 1e-3) / 1e12
        fig = ncu.get(name, {}).get(args.dtype, {}) if args.arith == "fast" else {}
        out["roofline"] = {
            "bound": args.dtype, "kernel": w.kernel(), "achieved": tflops, "peak": fp_peak, "unit": "TFLOP/s",
            "frac": tflops / fp_peak, "frac_of_nominal": tflops / NOMINAL_TFLOPS[args.dtype],
            "peak_source": "FMA microbenchmark rbs_fma_probe run in this process (MEASURED_PEAKS.json has no CUDA-core peak); "
                           "frac_of_nominal is against %.1f TFLOP/s" % NOMINAL_TFLOPS[args.dtype],
            "flops_per_env_substep": flops, "contacts_per_env_substep": c0, "impulses_per_env_substep": i0,
            "contacts_per_env_substep_next_horizon": c1, "impulses_per_env_substep_next_horizon": i1,
            "launch_ms": launch_ms, "substeps_per_launch": w.F, "hbm_GBps_of_same_launch": w.bytes_per_launch() / (launch_ms * 1e-3) / 1e9,
            "pipe_fp64_active_pct": fig.get("pipe_fp64_active_pct"), "ncu_source": fig.get("source"),
            "traffic": fig.get("dram_bytes_per_advance") if w.E == CONFIGS[name]["envs"] else None}
        if name == "multi_sphere64":
            out["body_substeps_per_s"] = out["value"] * 64
            if args.arith == "fast" and fig.get("issue_active_pct") is not None:
                # The plane-frame multi-sphere kernel is bound by INSTRUCTION ISSUE (integer / FP32 partner-list work; FP64
                # pipe ~13 % busy), so the fraction that means something is the issue-slot utilisation of the committed ncu
                # capture.  The algorithmic-flop view (SURVEY 8(d)'s all-pairs accounting: 63 pair rejects per body-substep,
                # most of which the lists never execute) is kept alongside and can exceed 1 for that reason.
                r = out["roofline"]
                out["roofline"] = {"bound": "issue", "kernel": r["kernel"], "achieved": fig["issue_active_pct"], "peak": 100.0,
                                   "unit": "% of issue slots (ncu, profiles/ncu_figures.json)", "frac": fig["issue_active_pct"] / 100.0,
                                   "pipe_fp64_active_pct": r["pipe_fp64_active_pct"], "ncu_source": r["ncu_source"], "traffic": r["traffic"],
                                   "launch_ms": r["launch_ms"], "substeps_per_launch": r["substeps_per_launch"],
                                   "contacts_per_env_substep": r["contacts_per_env_substep"], "impulses_per_env_substep": r["impulses_per_env_substep"],
                                   "contacts_per_env_substep_next_horizon": r["contacts_per_env_substep_next_horizon"],
                                   "impulses_per_env_substep_next_horizon": r["impulses_per_env_substep_next_horizon"],
                                   "algorithmic_flops_view": {"achieved_TFLOPs": r["achieved"], "peak_TFLOPs": r["peak"], "frac": r["frac"],
                                                              "frac_of_nominal": r["frac_of_nominal"], "flops_per_env_substep": r["flops_per_env_substep"],
                                                              "note": "SURVEY 8(d) all-pairs accounting; the partner lists skip most of those tests"}}
        # the bit-faithful policy on the same job (exact contact-event counts by construction, profiles/r2_parity_report.md)
        other = "strict" if args.arith == "fast" else "fast"
        o_ms, _, _ = timed(lambda: w.step(arith=other), 1 if not headline else max(2, steps // 3), 1, before=w.reset)
        out["other_policies"] = {"unit": METRIC, "values": {other: w.E_job * w.S / (o_ms / (1 if not headline else max(2, steps // 3)) * 1e-3)}}
        if name in ("cube_bounce", "cube_incline") and other == "strict":
            # the thread-per-environment strict kernel beside the default (rolled resident kernel with a per-CTA density
            # vote, DESIGN.md section 3, profiles/r2_ab_strict_hybrid.jsonl); same bits either way
            old_opt = rb._lib.set_option("strict_compact", -1)
            try:
                o_ms, _, _ = timed(lambda: w.step(arith=other), 1, 1, before=w.reset)
            finally:
                rb._lib.set_option("strict_compact", old_opt)
            out["other_policies"]["values"]["strict_thread_per_env (option strict_compact=-1)"] = w.E_job * w.S / (o_ms * 1e-3)
        if headline and name == "sphere_incline" and args.arith == "fast":
            def iso_step():
                for _ in range(w.S // w.F):
                    w.launch(w.F, False, "strict", strict_inertia=False)
            n_o = max(2, steps // 3)
            o_ms, _, _ = timed(iso_step, n_o, 1, before=w.reset)
            out["other_policies"]["values"]["strict_isotropic_shortcut"] = w.E_job * w.S / (o_ms / n_o * 1e-3)
        out["other_policies"]["note"] = ("strict = the reference's rounding sequence with the literal inv(R diag(I) R^T): bit-for-bit the C "
                                         "oracle, exact contact-event counts; fast = FMA / reciprocal-multiply re-association, <= 1e-12 "
                                         "relative per step (tests/test_gpu_parity.py, tests/test_bench_parity.py)")
        if headline and w.bodies == 1:
            # K=1 streaming regime (HBM-bound): one launch per substep -- the reference's per-frame call
            k1 = args.k1_launches

            def k1_step():
                for _ in range(k1):
                    w.launch(1, False, None, env_range=None)
            k1_ms, _, _ = timed(k1_step, max(3, steps), 3)
            k1_launch_ms = k1_ms / max(3, steps) / k1
            k1_gbs = w.bytes_per_launch() / (k1_launch_ms * 1e-3) / 1e9
            state_mb = w.E * 13 * w.esize / 1e6
            out["roofline_k1"] = {
                "bound": "hbm", "kernel": "same stepper, 1 substep per launch (the reference's per-frame call)", "achieved": k1_gbs,
                "peak": hbm_peak, "unit": "GB/s", "frac": k1_gbs / hbm_peak, "peak_source": hbm_src, "frac_of_nominal_8TBps": k1_gbs / 8000.0,
                "bytes_per_env": w.bytes_per_launch() // w.E, "launch_ms": k1_launch_ms,
                "note": ("back-to-back launches; the state (%.0f MB) is larger than what the 126 MB L2 can keep between a launch's write "
                         "and the next launch's read (ncu: reads come from DRAM)" % state_mb) if w.E * 26 * w.esize > 126e6 else
                        ("state %.0f MB read + written per launch fits the 126 MB L2: this figure is L2-assisted, not an HBM measurement" % state_mb),
                "env_steps_per_s": w.E_job / (k1_launch_ms * 1e-3), "traffic": fig.get("k1_dram_bytes_per_launch") if w.E == CONFIGS[name]["envs"] else None}
        # end to end through the host-buffer C-ABI call (H2D + S substeps + D2H every step)
        e_steps, e_warm = (steps, warm) if headline else (2, 1)
        e2e_ms, _, _ = timed(w.e2e_step, e_steps, e_warm, before=w.e2e_reset)
        nbytes = w.E * w.bodies * 13 * w.esize
        out["e2e"] = {"value": w.E_job * w.S / (e2e_ms / e_steps * 1e-3), "unit": METRIC, "h2d_bytes_per_step": nbytes,
                      "d2h_bytes_per_step": nbytes, "ms_per_step": e2e_ms / e_steps,
                      "call": {1: "rbs_run_body_plane_host via stepper.run_body_plane_host", 2: "rbs_run_two_ball_host via stepper.run_two_ball_host",
                               64: "rbs_run_multi_sphere_host via stepper.run_multi_sphere_host"}[w.bodies] +
                              " (pinned host qpos/qvel in the reference layout)"}
        if headline:
            out["end_of_run_stats"] = shard.gather_stats(shard.local_stats(w.model, w.data), env_substeps=w.E * w.S * steps)
            out["host_link"] = shard.host_link_bandwidth(dev, w.qpos_h, world, rank)
        out["config"] = workload_config(name, args, world, w.E, w.F)
        out["scaling"] = CONFIGS[name]["scaling"]
        if name in cpu:
            out["cpu_baseline"] = {k: cpu[name][k] for k in ("value", "unit", "cores", "kind", "sample") if k in cpu[name]} or cpu[name]
        del w
        torch.cuda.empty_cache()
        return out

    results = {n: measure(n, n == args.config) for n in names}

    def measure_mixed_pile():
        """SURVEY section 8(f) row N4 (not a BASELINE config): eight spheres and boxes per environment, 131,072 environments
        per GPU, 512 substeps from the scenario's initial pile in launches of 64; strict policy (the only one this stepper
        has), counters on; the C oracle's restatement of the same loop on all host cores beside it (rank 0, bounded sample)."""
        from rigidbody_simulation_b200.src.simulation import mixed_pile
        E, B, S, F = 131072, 8, 512, 64
        model, data = mixed_pile.build(E, device=dev, dtype=tdtype, n_body=B, start=rank * E)
        s0 = data.state.clone()

        def reset():
            data.state.copy_(s0)

        def step():
            for _ in range(S // F):
                stepper.step_multi_body(model, data, mixed_pile.timestep, mixed_pile.restitution_coefficient,
                                        mixed_pile.friction_coefficient, substeps=F, count=True)
        ms, _, _ = timed(step, 2, 1, before=reset)
        calls, imps = data.counters()
        out = {"workload": "mixed_pile: 4 boxes (one cube of models/cube.xml, three anisotropic) + 4 spheres per environment, random pile "
                           "over a flat plane, plane-sphere / plane-box / sphere-sphere / sphere-box / box-box contacts, e=0.2, mu=0.6, dt=0.005",
               "envs_per_gpu": E, "bodies_per_env": B, "substeps_per_step": S, "substeps_fused_per_launch": F, "arith": "strict",
               "value": world * E * S / (ms / 2 * 1e-3), "unit": METRIC, "body_substeps_per_s": world * E * B * S / (ms / 2 * 1e-3),
               "contacts_per_body_substep": float(calls.sum()) / (3.0 * E * B * S), "impulses_per_body_substep": float(imps.sum()) / (3.0 * E * B * S),
               "kernel": "step_multi_body_kernel"}
        # end to end through the host-buffer C-ABI call (pinned reference-layout arrays, H2D + S substeps + D2H inside)
        data.state.copy_(s0)
        qp0, qv0 = data.qpos.torch().cpu().pin_memory(), data.qvel.torch().cpu().pin_memory()
        qp_h, qv_h = qp0.clone().pin_memory(), qv0.clone().pin_memory()

        def e2e_reset():
            qp_h.copy_(qp0)
            qv_h.copy_(qv0)

        def e2e_step():
            stepper.run_multi_body_host(model, qp_h, qv_h, S, dt=mixed_pile.timestep, restitution=mixed_pile.restitution_coefficient,
                                        friction=mixed_pile.friction_coefficient, substeps=F)
        e_ms, _, _ = timed(e2e_step, 2, 1, before=e2e_reset)
        nbytes = E * B * 13 * (8 if tdtype == torch.float64 else 4)
        out["e2e"] = {"value": world * E * S / (e_ms / 2 * 1e-3), "unit": METRIC, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
                      "ms_per_step": e_ms / 2, "call": "rbs_run_multi_body_host via stepper.run_multi_body_host (pinned host qpos/qvel)"}
        if rank == 0 and not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import c_oracle as co
            import numpy as np
            n = 32 * (os.cpu_count() or 1)
            tab = stepper.body_table(model)
            data.state.copy_(s0)
            q = data.qpos.torch()[:n].cpu().numpy().astype(np.float64).reshape(n, B, 7).copy()
            v = data.qvel.torch()[:n].cpu().numpy().astype(np.float64).reshape(n, B, 6).copy()
            t0 = time.time()
            co.step_multi_body(q, v, 256, gtype=tab[:, 0].astype(np.int32), mass=tab[:, 4], inertia=tab[:, 5:8], size=tab[:, 1:4],
                               plane_pos=[0, 0, 0], plane_normal=[0, 0, 1], gravity=[0, 0, -9.8], dt=mixed_pile.timestep,
                               restitution=mixed_pile.restitution_coefficient, friction=mixed_pile.friction_coefficient)
            out["cpu_baseline"] = {"value": n * 256 / (time.time() - t0), "unit": METRIC, "cores": co.max_threads(), "kind": "port",
                                   "sample": f"{n} environments x 256 steps of the C restatement (OpenMP); the reference has no such scene"}
        return out

    extra = {}
    if not args.no_other_configs:
        extra["mixed_pile"] = measure_mixed_pile()

    if rank == 0:
        h = results[args.config]
        line = {"metric": METRIC, "value": h["value"], "unit": METRIC, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": h["ms_per_step"], "higher_is_better": True, "scaling": h["scaling"], "vs_baseline": None,
                "dtype": "f64" if args.dtype == "fp64" else "f32", "data": "synthetic", "config": h["config"], "e2e": h["e2e"],
                "gpu_launches": h["gpu_launches"], "clocks": h["clocks"], "roofline": h["roofline"]}
        for k in ("roofline_k1", "end_of_run_stats", "other_policies", "host_link", "cpu_baseline", "body_substeps_per_s"):
            if k in h:
                line[k] = h[k]
        line["numa"] = numa
        if cpu:
            line["cpu_baseline_native"] = cpu_native
        line["configs"] = {n: {k: v for k, v in r.items() if k not in ("clocks", "gpu_launches", "end_of_run_stats", "host_link")}
                           for n, r in results.items()}
        if extra:
            line["extra"] = extra                            # beyond BASELINE.json: SURVEY section 8(f) rows
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
