"""A/B of the plane-frame multi-sphere kernel's TIGHT span (option ms_tight_span) and register cap (ms_regs) on config 5
(65,536 x 64 spheres, fp64, mu = 0): 2048 substeps from the lattice in 8 launches of 256; one JSON line per run.
    python profiles/ab_multi_sphere_tight.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import rigidbody_simulation_b200 as rb
from rigidbody_simulation_b200 import stepper, synth
from rigidbody_simulation_b200.src.simulation import multi_sphere_bounce

dev = torch.device("cuda:0")
E, K, L = 65536, 256, 8
s = synth.multi_sphere(E, n_body=64, friction=0.0)
model, data = multi_sphere_bounce.build(E, device=dev, dtype=torch.float64, n_body=64)
ref = None
for regs, span in ((96, 32), (96, 0), (96, 8), (96, 64), (96, 128), (96, 256), (128, 32), (128, 128)):
    rb._lib.set_option("ms_regs", regs)
    rb._lib.set_option("ms_tight_span", span)
    best = None
    for rep in range(2):
        data.set_state(s["qpos"], s["qvel"])
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(L + 1)]
        for i in range(L):
            ev[i].record()
            stepper.step_multi_sphere(model, data, 0.01, 1.0, 0.0, substeps=K, count=False, arith="fast")
        ev[L].record()
        torch.cuda.synchronize()
        ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(L)]
        if best is None or sum(ms) < sum(best):
            best = ms
    same = None if ref is None else bool(torch.equal(ref, data.state))
    ref = data.state.clone() if ref is None else ref
    n = E * 64 * K
    print(json.dumps({"ms_regs": regs, "ms_tight_span": span, "launch_ms": [round(m, 2) for m in best],
                      "body_substeps_per_s_first_launch": n / (best[0] * 1e-3), "body_substeps_per_s_2048": n * L / (sum(best) * 1e-3),
                      "state_bitwise_equal_to_first_variant": same}), flush=True)
rb._lib.set_option("ms_regs", 96)
rb._lib.set_option("ms_tight_span", 32)
