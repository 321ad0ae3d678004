"""A/B of the multi-sphere fast kernels on config 5 (64 spheres per env): first-generation kernel (option ms_kernel=1)
against the plane-frame kernel (ms_kernel=2), mu = 0 (shipped) and mu = 0.3, 8,192 and 65,536 envs.  Every run starts
from the lattice and advances 2048 substeps in 16 launches of 128; prints one JSON line per run with the per-launch
times (early dense phase first) and the whole-horizon throughput.  Device-timed, state resident in HBM.
    python profiles/ab_multi_sphere.py [--quick]
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

quick = "--quick" in sys.argv
dev = torch.device("cuda:0")
K, L = 128, 16
for dtype, tag in ((torch.float64, "fp64"), (torch.float32, "fp32")):
    for E in ((8192,) if quick else (8192, 65536)):
        for mu in (0.0, 0.3):
            s = synth.multi_sphere(E, n_body=64, friction=mu)
            for kernel in (1, 2):
                for walk in ((20,) if kernel == 1 else (20, 40, 80)):
                    rb._lib.set_option("ms_kernel", kernel)
                    rb._lib.set_option("ms_walk_cost", walk)
                    model, data = multi_sphere_bounce.build(E, device=dev, dtype=dtype, n_body=64)
                    best = None
                    for rep in range(2):
                        data.set_state(s["qpos"], s["qvel"])
                        ev = [torch.cuda.Event(enable_timing=True) for _ in range(L + 1)]
                        for i in range(L):
                            ev[i].record()
                            stepper.step_multi_sphere(model, data, 0.01, 1.0, mu, substeps=K, count=False, arith="fast")
                        ev[L].record()
                        torch.cuda.synchronize()
                        ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(L)]
                        if best is None or sum(ms) < sum(best):
                            best = ms
                    data.set_state(s["qpos"], s["qvel"])
                    data.n_contacts.zero_(); data.n_impulses.zero_()
                    for i in range(L):
                        stepper.step_multi_sphere(model, data, 0.01, 1.0, mu, substeps=K, count=True, arith="fast")
                    torch.cuda.synchronize()
                    n = E * 64 * K
                    print(json.dumps({"dtype": tag, "envs": E, "mu": mu, "ms_kernel": kernel, "walk_cost": walk,
                                      "launch_ms": [round(m, 3) for m in best],
                                      "body_substeps_per_s_first_launch": n / (best[0] * 1e-3),
                                      "body_substeps_per_s_last_launch": n / (best[-1] * 1e-3),
                                      "body_substeps_per_s_2048": n * L / (sum(best) * 1e-3),
                                      "contacts_per_body_substep": float(data.n_contacts.sum()) / (n * L),
                                      "impulses_per_body_substep": float(data.n_impulses.sum()) / (n * L)}), flush=True)
rb._lib.set_option("ms_kernel", 2)
