"""A/B of the strict (bit-faithful, literal-inertia) single-body stepper's occupancy variants (option strict_minb = 2 /
4 / 5 / 6 resident CTAs per SM) on configs 2 and 4 (1,048,576 envs, fp64): 512 substeps from the initial state in 4
launches of 128; one JSON line per run, with a bitwise comparison of the final state against the first variant.
    python profiles/ab_strict.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import rigidbody_simulation_b200 as rb
from rigidbody_simulation_b200 import scenes, stepper, synth

dev = torch.device("cuda:0")
E, K, L = 1 << 20, 128, 4
for name in ("sphere_incline", "cube_bounce", "cube_incline"):
    if name == "sphere_incline":
        s = synth.sphere_incline(E)
        model = scenes.sphere_on_incline(E, device=dev)
        model.set_per_env(restitution=s["restitution"], friction=s["friction"])
        kw = dict(restitution=None, friction_coeff=None, contact_threshold=0.0)
    else:
        s = synth.cube(E, kind=name.split("_")[1])
        model = scenes.cube_on_plane(E, theta=s["theta"], device=dev)
        kw = dict(restitution=0.2, friction_coeff=0.6, contact_threshold=1e-4)
    data = rb.BatchedData(model)
    ref = None
    for compact, minb in ((-1, 2), (-1, 4), (-1, 5), (4, 0), (5, 0), (6, 0), (8, 0), (25, 0), (44, 0), (54, 0), (64, 0), (65, 0), (74, 0), (124, 0), (134, 0), (135, 0), (264, 0), (274, 0), (364, 0), (374, 0), (0, 0)):   # the instantiations kept in the library
        rb._lib.set_option("strict_compact", compact)
        rb._lib.set_option("strict_minb", minb)
        best = None
        for rep in range(2):
            data.set_state(s["qpos"], s["qvel"])
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(L + 1)]
            for i in range(L):
                ev[i].record()
                stepper.step_body_plane(model, data, -1, s["dt"], kw["restitution"], kw["friction_coeff"], kw["contact_threshold"],
                                        substeps=K, count=False, arith="strict")
            ev[L].record()
            torch.cuda.synchronize()
            ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(L)]
            if best is None or sum(ms) < sum(best):
                best = ms
        same = None
        if ref is None:
            ref = data.state.clone()
        else:
            same = bool(torch.equal(ref, data.state))
        print(json.dumps({"config": name, "strict_compact": compact, "strict_minb": minb, "launch_ms": [round(m, 3) for m in best],
                          "env_substeps_per_s": E * K * L / (sum(best) * 1e-3), "state_bitwise_equal_to_first_variant": same}), flush=True)
rb._lib.set_option("strict_compact", 0)
rb._lib.set_option("strict_minb", 0)

# strict two-ball (config 3) and strict multi-sphere (config 5, 8,192 envs) ----------------------------------------------
from rigidbody_simulation_b200.src.simulation import ball_collision, multi_sphere_bounce

s = synth.two_ball(E)
model, data = ball_collision.build(E, device=dev)
ref = None
for minb in (3, 4, 5):
    rb._lib.set_option("strict_tb_minb", minb)
    best = None
    for rep in range(2):
        data.set_state(s["qpos"], s["qvel"])
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(L + 1)]
        for i in range(L):
            ev[i].record()
            stepper.step_two_ball(model, data, 0.01, 1.0, 0.3, radius=0.1, substeps=K, count=False, arith="strict")
        ev[L].record()
        torch.cuda.synchronize()
        ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(L)]
        if best is None or sum(ms) < sum(best):
            best = ms
    same = None if ref is None else bool(torch.equal(ref, data.state))
    ref = data.state.clone() if ref is None else ref
    print(json.dumps({"config": "two_ball", "strict_tb_minb": minb, "launch_ms": [round(m, 3) for m in best],
                      "env_substeps_per_s": E * K * L / (sum(best) * 1e-3), "state_bitwise_equal_to_first_variant": same}), flush=True)
rb._lib.set_option("strict_tb_minb", 0)
E5 = 8192
s = synth.multi_sphere(E5, n_body=64, friction=0.0)
model, data = multi_sphere_bounce.build(E5, device=dev, n_body=64)
ref = None
for regs in (168, 128, 96):
    rb._lib.set_option("strict_ms_regs", regs)
    best = None
    for rep in range(2):
        data.set_state(s["qpos"], s["qvel"])
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(L + 1)]
        for i in range(L):
            ev[i].record()
            stepper.step_multi_sphere(model, data, 0.01, 1.0, 0.0, substeps=K, count=False, arith="strict")
        ev[L].record()
        torch.cuda.synchronize()
        ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(L)]
        if best is None or sum(ms) < sum(best):
            best = ms
    same = None if ref is None else bool(torch.equal(ref, data.state))
    ref = data.state.clone() if ref is None else ref
    print(json.dumps({"config": "multi_sphere64_8192", "strict_ms_regs": regs, "launch_ms": [round(m, 3) for m in best],
                      "body_substeps_per_s": E5 * 64 * K * L / (sum(best) * 1e-3), "state_bitwise_equal_to_first_variant": same}), flush=True)
rb._lib.set_option("strict_ms_regs", 0)
