"""A/B of the strict multi-sphere stepper with queued contacts (option strict_ms_queue) on config 5: 65,536 and 8,192
environments x 64 spheres, fp64, mu = 0 and mu = 0.3, 512 substeps from the lattice in 4 launches of 128.
    python profiles/ab_strict_ms_queue.py
RECORD ONLY: the queued variant (contacts of a body queued in local memory, one impulse loop, as in step_multi_body_kernel)
measured SLOWER here (profiles/r2_ab_strict_ms_queue.jsonl: 4.06 vs 4.59e9 body-substeps/s at mu = 0, 1.57 vs 2.74e9 at
mu = 0.3; bit-identical) and was taken out of the library again; the option this script sets no longer exists.
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
K, L = 128, 4
for E5 in (8192, 65536):
    for mu in (0.0, 0.3):
        s = synth.multi_sphere(E5, n_body=64, friction=mu)
        model, data = multi_sphere_bounce.build(E5, device=dev, n_body=64)
        ref = None
        for queue, regs in ((0, 128), (1, 128), (0, 168), (1, 168)):
            rb._lib.set_option("strict_ms_queue", queue)
            rb._lib.set_option("strict_ms_regs", regs)
            best = None
            for rep in range(2):
                data.set_state(s["qpos"], s["qvel"])
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(L + 1)]
                for i in range(L):
                    ev[i].record()
                    stepper.step_multi_sphere(model, data, 0.01, 1.0, mu, substeps=K, count=False, arith="strict")
                ev[L].record()
                torch.cuda.synchronize()
                ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(L)]
                if best is None or sum(ms) < sum(best):
                    best = ms
            same = None if ref is None else bool(torch.equal(ref, data.state))
            ref = data.state.clone() if ref is None else ref
            print(json.dumps({"envs": E5, "mu": mu, "strict_ms_queue": queue, "strict_ms_regs": regs, "launch_ms": [round(m, 2) for m in best],
                              "body_substeps_per_s": E5 * 64 * K * L / (sum(best) * 1e-3), "state_bitwise_equal_to_first_variant": same}), flush=True)
rb._lib.set_option("strict_ms_queue", 0)
rb._lib.set_option("strict_ms_regs", 0)
