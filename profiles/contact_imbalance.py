"""CPU analysis (C oracle, no GPU): how unevenly do contacts fall on the 32 bodies of a warp in config 5?

The multi-sphere kernels give one thread to a body and resolve its contacts one after the other (each impulse reads
the velocities the previous one wrote), so a warp-substep costs max-over-lanes(contacts) passes of the impulse code,
while the useful work is the mean.  This script steps the config-5 scene with the oracle one step at a time, takes the
per-body contact counts of each step, groups bodies as the kernel does (thread = env*B + body, warps of 32) and prints
mean, warp-max and the ratio per phase, plus what sorting the bodies of a CTA (2 envs = 128 bodies) by contact count
before resolving would give (heavy bodies share warps).
    python profiles/contact_imbalance.py [envs] [steps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np

import c_oracle as co
from rigidbody_simulation_b200 import synth

E = int(sys.argv[1]) if len(sys.argv) > 1 else 256
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 768
B = 64
for mu in (0.0, 0.3):
    s = synth.multi_sphere(E, n_body=B, friction=mu)
    qp, qv = s["qpos"].reshape(E, B, 7).copy(), s["qvel"].reshape(E, B, 6).copy()
    r = 0.1
    m = 50 * 4 / 3 * np.pi * r ** 3
    inertia = np.full((E, B, 3), 0.4 * m * r * r)
    calls = np.zeros(E * B, np.uint32)
    imps = np.zeros(E * B, np.uint32)
    prev = calls.copy()
    rows = []
    for k in range(STEPS):
        co.step_multi_sphere(qp, qv, 1, mass=m, inertia=inertia, radius=r, plane_pos=[0, 0, 0], plane_normal=[0, 0, 1],
                             gravity=[0, 0, -9.8], dt=s["dt"], restitution=s["restitution"], friction=mu, counters=(calls, imps))
        c = (calls - prev).astype(np.int64)
        prev = calls.copy()
        warp_max = c.reshape(-1, 32).max(axis=1)                       # kernel's mapping: 2 warps per env
        cta = np.sort(c.reshape(-1, 128), axis=1)                      # sorted within a CTA of 2 envs
        sorted_max = cta.reshape(-1, 4, 32).max(axis=2)
        rows.append((c.mean(), warp_max.mean(), sorted_max.mean(), (c > 0).mean()))
    rows = np.array(rows)
    print(f"config 5, mu = {mu}: {E} envs x {B} spheres, contacts per body-substep / per warp-substep (max over its 32 lanes)")
    for lo, hi in ((0, 128), (128, 256), (256, 512), (512, STEPS)):
        if lo >= STEPS:
            break
        a = rows[lo:min(hi, STEPS)].mean(axis=0)
        print(f"  substeps {lo:4d}-{min(hi, STEPS):4d}: mean {a[0]:.3f}  warp max {a[1]:.3f}  (x{a[1] / max(a[0], 1e-9):.1f} the mean)  "
              f"sorted within the CTA {a[2]:.3f} (x{a[2] / max(a[0], 1e-9):.1f})  bodies in contact {100 * a[3]:.0f} %")
