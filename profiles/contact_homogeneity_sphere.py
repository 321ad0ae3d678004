"""CPU analysis (C oracle, no GPU): config 2, how often a warp of the headline kernel has to run the contact path, as the
environments are laid out today (32 consecutive envs per warp) and if every CTA ordered its environments by how often they
touched the plane in the PREVIOUS launch (a predictor the kernel could have: one counter per env).
    python profiles/contact_homogeneity_sphere.py [envs] [substeps] [fuse]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np

import c_oracle as co
from rigidbody_simulation_b200 import synth

E = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
FUSE = int(sys.argv[3]) if len(sys.argv) > 3 else 256
s = synth.sphere_incline(E)
qp, qv = s["qpos"].copy(), s["qvel"].copy()
m = 50 * 4 / 3 * np.pi * 0.2 ** 3
inertia = [0.4 * m * 0.04] * 3
calls, imps = np.zeros(E, np.uint32), np.zeros(E, np.uint32)
prev = calls.copy()
hit = np.zeros((STEPS, E), bool)
for k in range(STEPS):
    co.step_body_plane(qp, qv, 1, geom="sphere", mass=m, inertia=inertia, size=s["radius"], plane_pos=[0, 0, 0],
                       plane_normal=s["plane_normal"], gravity=[0, 0, -9.8], dt=s["dt"], restitution=s["restitution"],
                       friction=s["friction"], threshold=s["threshold"], counters=(calls, imps))
    hit[k] = calls != prev
    prev = calls.copy()
print(f"config 2: {E} envs x {STEPS} substeps, launches of {FUSE}; contacts per env-substep {hit.mean():.3f}")
print("fraction of warp-substeps in which some lane takes the contact path:")
for cta in (32, 128, 256, 512, 1024):
    out = []
    for l in range(STEPS // FUSE):
        h = hit[l * FUSE:(l + 1) * FUSE]
        if cta == 32:
            order = np.arange(E)
        else:
            key = hit[(l - 1) * FUSE:l * FUSE].sum(0) if l > 0 else np.zeros(E, int)   # previous launch's count
            order = np.concatenate([c0 + np.argsort(key[c0:c0 + cta], kind="stable") for c0 in range(0, E, cta)])
            oracle_key = h.sum(0)
            best = np.concatenate([c0 + np.argsort(oracle_key[c0:c0 + cta], kind="stable") for c0 in range(0, E, cta)])
        w = h[:, order].reshape(FUSE, -1, 32).any(axis=2).mean()
        wb = h[:, best].reshape(FUSE, -1, 32).any(axis=2).mean() if cta != 32 else w
        out.append((w, wb))
    label = "as laid out" if cta == 32 else f"CTA of {cta} ordered by previous launch's count (by this launch's own count)"
    print(f"  {label}: " + "  ".join(f"{a:.2f}({b:.2f})" for a, b in out))
# distribution of per-launch counts
for l in range(STEPS // FUSE):
    c = hit[l * FUSE:(l + 1) * FUSE].sum(0)
    print(f"  launch {l}: envs with 0 contacts {np.mean(c == 0):.2f}, 1-8 {np.mean((c > 0) & (c <= 8)):.2f}, 9-64 {np.mean((c > 8) & (c <= 64)):.2f}, "
          f"65-200 {np.mean((c > 64) & (c <= 200)):.2f}, >200 {np.mean(c > 200):.2f}")
