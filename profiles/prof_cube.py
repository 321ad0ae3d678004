"""Timing / ncu target: config 4 (cube on a plane, 1,048,576 envs), fast policy, fp64.
    python profiles/prof_cube.py [bounce|incline] [substeps per launch] [fast|strict]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import rigidbody_simulation_b200 as rb
from rigidbody_simulation_b200 import scenes, stepper, synth

kind = sys.argv[1] if len(sys.argv) > 1 else "bounce"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 128
ARITH = sys.argv[3] if len(sys.argv) > 3 else "fast"
E = 1 << 20
s = synth.cube(E, kind=kind)
model = scenes.cube_on_plane(E, theta=s["theta"], device=torch.device("cuda:0"), dtype=torch.float64)
data = rb.BatchedData(model)
data.set_state(s["qpos"], s["qvel"])
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
for i in range(5):
    ev[i].record()
    stepper.step_body_plane(model, data, -1, s["dt"], 0.2, 0.6, 1e-4, substeps=K, count=(i == 4), arith=ARITH)
ev[5].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(5)]
print(kind, "launch ms:", [round(m, 3) for m in ms], "env-substeps/s (last):", E * K / (ms[-1] * 1e-3))
c, i = data.counters()
print("contacts / impulses per env-substep in the last launch:", c.sum() / (E * K), i.sum() / (E * K))
