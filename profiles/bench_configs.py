"""Throughput of every BASELINE config at its stated size (device-timed, CUDA events, state resident in HBM).
Not the bench line (that is config 2, bench.py) -- a record of where the other steppers stand.
    python profiles/bench_configs.py [--quick]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import rigidbody_simulation_b200 as rb
from rigidbody_simulation_b200 import scenes, stepper, synth
from rigidbody_simulation_b200.src.simulation import ball_collision, multi_sphere_bounce

quick = "--quick" in sys.argv
dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ms = []
    for _ in range(reps):
        flush.fill_(0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


PEAK = {}


def report(name, E, B, K, ms, extra="", flops=None):
    """flops = approximate algorithmic flops per body-substep (mul/add/div/sqrt = 1 each) from the measured contact
    rates; the fraction is against the FMA-probe peak measured in this process."""
    out = {"config": name, "envs": E, "bodies_per_env": B, "substeps_per_launch": K, "launch_ms": round(ms, 4),
           "env_substeps_per_s": E * K / (ms * 1e-3), "body_substeps_per_s": E * B * K / (ms * 1e-3), "note": extra}
    if flops is not None:
        tf = out["body_substeps_per_s"] * flops / 1e12
        dt = torch.float32 if "fp32" in name else torch.float64
        if dt not in PEAK:
            PEAK[dt] = stepper.fma_peak(dev, dt) / 1e12
        out.update(flops_per_body_substep=round(flops, 1), tflops=round(tf, 2), fp_roofline_frac=round(tf / PEAK[dt], 3))
    print(json.dumps(out), flush=True)


def rates(data, n_units, K, fn):
    """contacts / impulses per unit-substep over one counted launch of K substeps"""
    data.n_contacts.zero_(); data.n_impulses.zero_()
    fn()
    torch.cuda.synchronize()
    return float(data.n_contacts.sum()) / (n_units * K), float(data.n_impulses.sum()) / (n_units * K)


for dtype, tag in ((torch.float64, "fp64"), (torch.float32, "fp32")):
    E = (1 << 20) if not quick else (1 << 16)
    # config 2 (both policies) --------------------------------------------------------------------------------
    s = synth.sphere_incline(E)
    model = scenes.sphere_on_incline(E, device=dev, dtype=dtype)
    model.set_per_env(restitution=s["restitution"], friction=s["friction"])
    data = rb.BatchedData(model)
    data.set_state(s["qpos"], s["qvel"])
    for arith in ("strict", "fast"):
        for K in (1, 128):
            f = lambda: stepper.step_body_plane(model, data, -1, s["dt"], None, None, 0.0, substeps=K, count=False, arith=arith)
            for _ in range(8):
                stepper.step_body_plane(model, data, -1, s["dt"], None, None, 0.0, substeps=128, count=False, arith=arith)
            c, i = rates(data, E, K, lambda: stepper.step_body_plane(model, data, -1, s["dt"], None, None, 0.0, substeps=K, count=True, arith=arith))
            report(f"cfg2 sphere_incline {tag} {arith}", E, 1, K, timed(f), flops=60 + 72 * i + 22 * (c - i))
    # config 4 cube ------------------------------------------------------------------------------------------
    for kind in ("bounce", "incline"):
        s = synth.cube(E, kind=kind)
        model = scenes.cube_on_plane(E, theta=s["theta"], device=dev, dtype=dtype)
        data = rb.BatchedData(model)
        data.set_state(s["qpos"], s["qvel"])
        for arith in ("strict", "fast"):
            for K in (1, 128):
                f = lambda: stepper.step_body_plane(model, data, -1, s["dt"], 0.2, 0.6, 1e-4, substeps=K, count=False, arith=arith)
                c, i = rates(data, E, K, lambda: stepper.step_body_plane(model, data, -1, s["dt"], 0.2, 0.6, 1e-4, substeps=K, count=True, arith=arith))
                # free flight 60 + vertex scan (rotation 42 + 8 x 23) + per contact: arm 9 + impulse 72 / separating 22
                report(f"cfg4 cube_{kind} {tag} {arith}", E, 1, K, timed(f), flops=60 + 226 + 81 * i + 31 * (c - i))
    # config 3 two balls ---------------------------------------------------------------------------------------
    s = synth.two_ball(E)
    model, data = ball_collision.build(E, device=dev, dtype=dtype)
    data.set_state(s["qpos"], s["qvel"])
    for arith in ("strict", "fast"):
        for K in (1, 256):
            f = lambda: stepper.step_two_ball(model, data, 0.01, 1.0, 0.3, radius=0.1, substeps=K, count=False, arith=arith)
            c, i = rates(data, E, K, lambda: stepper.step_two_ball(model, data, 0.01, 1.0, 0.3, radius=0.1, substeps=K, count=True, arith=arith))
            # per env: gravity 6 + ground tests 2 + pair test 12 + integrate 12; ~120 per ground impulse, ~190 per pair hit
            report(f"cfg3 two_ball {tag} {arith}", E, 2, K, timed(f), flops=(32 + 120 * c + 190 * i) / 2)
    # config 5 multi sphere (8192 envs = the per-GPU shard of 65536 over 8 GPUs; and the whole 65536) ------------
    # Two regimes: "early" = the first launches from the lattice (dense, falling), "steady" = after 512 substeps (what a
    # 2048-substep horizon mostly sees).  The flop count is SURVEY 8(d)'s all-pairs accounting (63 pair rejects per
    # body-substep); the partner lists skip most of those tests, so the fraction is throughput in the reference
    # algorithm's units, not FP-pipe utilisation.
    for E5 in ((8192, 65536) if not quick else (1024,)):
        s = synth.multi_sphere(E5, n_body=64, friction=0.0)
        for arith in ("strict", "fast"):
            for K in (1, 128):
                model, data = multi_sphere_bounce.build(E5, device=dev, dtype=dtype, n_body=64)
                data.set_state(s["qpos"], s["qvel"])
                f = lambda: stepper.step_multi_sphere(model, data, 0.01, 1.0, 0.0, substeps=K, count=False, arith=arith)
                fc = lambda: stepper.step_multi_sphere(model, data, 0.01, 1.0, 0.0, substeps=K, count=True, arith=arith)
                if K > 1:
                    c, i = rates(data, E5 * 64, K, fc)
                    report(f"cfg5 multi_sphere64 {tag} {arith} early", E5, 64, K, timed(f, reps=1, warm=0), flops=60 + 567 + 97 * i + 47 * (c - i))
                    for _ in range(2):
                        f()
                c, i = rates(data, E5 * 64, K, fc)
                report(f"cfg5 multi_sphere64 {tag} {arith}" + (" steady" if K > 1 else ""), E5, 64, K, timed(f, reps=3, warm=1),
                       flops=60 + 567 + 97 * i + 47 * (c - i))
