"""Partner-list skin sweep for the multi-sphere steppers (config 5: 64 spheres per env, all-pairs contacts).
    python profiles/tune_skin.py [--envs 65536] [--quick]
Prints one JSON line per (policy, dtype, skin, substeps per launch): launch time and body-substeps/s, early in the run
(lattice still falling) and late (after ~512 substeps: a dilute gas of bouncing spheres, e = 1).  skin -1 = no lists
(every partner scanned every substep), 0 = adaptive per CTA (the default).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from rigidbody_simulation_b200 import stepper, synth
from rigidbody_simulation_b200.src.simulation import multi_sphere_bounce

quick = "--quick" in sys.argv
E = int(sys.argv[sys.argv.index("--envs") + 1]) if "--envs" in sys.argv else (4096 if quick else 65536)
B = 64
dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=3):
    ms = []
    for _ in range(reps):
        flush.fill_(0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


s = synth.multi_sphere(E, n_body=B, friction=0.0)
for dtype, tag in ((torch.float64, "fp64"), (torch.float32, "fp32")):
    for arith in ("fast", "strict"):
        for K in ((16, 128) if not quick else (16,)):
            for skin in (-1, 0, 50, 100, 250, 500, 1000):
                model, data = multi_sphere_bounce.build(E, device=dev, dtype=dtype, n_body=B)
                data.set_state(s["qpos"], s["qvel"])
                f = lambda: stepper.step_multi_sphere(model, data, 0.01, 1.0, 0.0, substeps=K, count=False, arith=arith,
                                                      list_skin_percent=skin)
                early = timed(f)                                   # substeps 0 .. 3K: the lattice is falling
                for _ in range(max(0, 512 // K - 3)):
                    f()
                late = timed(f)                                    # after ~512 substeps: the pile
                print(json.dumps({"dtype": tag, "arith": arith, "substeps_per_launch": K, "skin_percent": skin,
                                  "early_ms": round(early, 4), "late_ms": round(late, 4),
                                  "early_body_substeps_per_s": E * B * K / (early * 1e-3),
                                  "late_body_substeps_per_s": E * B * K / (late * 1e-3)}), flush=True)
