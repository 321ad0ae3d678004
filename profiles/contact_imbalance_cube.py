"""CPU analysis (C oracle, no GPU): config 4, contacts handed to the impulse routine per env-substep against the maximum
over the 32 environments of a warp (the cube kernels run their per-candidate loop max-over-lanes times).
    python profiles/contact_imbalance_cube.py [envs] [steps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np

import c_oracle as co
from rigidbody_simulation_b200 import synth

E = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 768
for kind in ("bounce", "incline"):
    s = synth.cube(E, kind=kind)
    qp, qv = s["qpos"].copy(), s["qvel"].copy()
    m = 50 * 0.8 ** 3
    inertia = [m / 3 * 2 * 0.16] * 3
    calls, imps = np.zeros(E, np.uint32), np.zeros(E, np.uint32)
    prev = calls.copy()
    rows = []
    for k in range(STEPS):
        co.step_body_plane(qp, qv, 1, geom="box", mass=m, inertia=inertia, size=s["half"], plane_pos=[0, 0, 0],
                           plane_normal=s["plane_normal"], gravity=[0, 0, -9.8], dt=s["dt"], restitution=s["restitution"],
                           friction=s["friction"], threshold=s["threshold"], counters=(calls, imps))
        c = (calls - prev).astype(np.int64)
        prev = calls.copy()
        rows.append((c.mean(), c.reshape(-1, 32).max(axis=1).mean(), np.sort(c).reshape(-1, 32).max(axis=1).mean(), (c > 0).mean()))
    rows = np.array(rows)
    print(f"config 4 cube {kind}: {E} envs, contacts per env-substep / per warp-substep (max over its 32 lanes)")
    for lo, hi in ((0, 128), (128, 256), (256, 512), (512, STEPS)):
        if lo >= STEPS:
            break
        a = rows[lo:min(hi, STEPS)].mean(axis=0)
        print(f"  substeps {lo:4d}-{min(hi, STEPS):4d}: mean {a[0]:.3f}  warp max {a[1]:.3f}  (x{a[1] / max(a[0], 1e-9):.1f} the mean)  "
              f"if environments were ordered by contact count {a[2]:.3f} (x{a[2] / max(a[0], 1e-9):.1f})  envs in contact {100 * a[3]:.0f} %")
