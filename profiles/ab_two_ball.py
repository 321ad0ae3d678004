"""A/B of the two-ball fast kernel's occupancy variants (option tb_minb = 5 / 6 / 8 resident CTAs per SM; tb_packed = the
fp32x2 kernel with two envs per thread, float only) on config 3
(1,048,576 envs), fp64 and fp32: 2048 substeps from the initial state in 8 launches of 256; one JSON line per run.
    python profiles/ab_two_ball.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import rigidbody_simulation_b200 as rb
from rigidbody_simulation_b200 import stepper, synth
from rigidbody_simulation_b200.src.simulation import ball_collision

dev = torch.device("cuda:0")
E, K, L = 1 << 20, 256, 8
s = synth.two_ball(E)
for dtype, tag in ((torch.float64, "fp64"), (torch.float32, "fp32")):
    model, data = ball_collision.build(E, device=dev, dtype=dtype)
    ref = None
    for minb, packed, uniform in ((5, 0, 0), (6, 0, 0), (8, 0, 0), (5, 0, 1), (6, 0, 1), (8, 0, 1)) + (((6, 1, 0), (8, 1, 0)) if tag == "fp32" else ()):
        rb._lib.set_option("tb_minb", minb)
        rb._lib.set_option("tb_packed", packed)
        rb._lib.set_option("tb_uniform", uniform)
        best = None
        for rep in range(3):
            data.set_state(s["qpos"], s["qvel"])
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(L + 1)]
            for i in range(L):
                ev[i].record()
                stepper.step_two_ball(model, data, 0.01, 1.0, 0.3, radius=0.1, substeps=K, count=False, arith="fast")
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
        print(json.dumps({"dtype": tag, "tb_minb": minb, "tb_packed": packed, "tb_uniform": uniform, "launch_ms": [round(m, 3) for m in best],
                          "env_substeps_per_s_2048": E * K * L / (sum(best) * 1e-3), "state_bitwise_equal_to_first_variant": same}), flush=True)
rb._lib.set_option("tb_minb", 0)
rb._lib.set_option("tb_packed", 0)
rb._lib.set_option("tb_uniform", 1)
