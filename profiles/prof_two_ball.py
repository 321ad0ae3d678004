"""ncu target: config 3 (two balls per env, 1,048,576 envs), fast policy, fp64.  Two warm launches of 256 substeps carry
the scene past the ball-ball collision (step ~100); the third launch is the one to capture:
    ncu --set full --clock-control none --import-source on -k regex:step_two_ball_fast -s 2 -c 1 \
        -o gpurun_out/prof_tb python profiles/prof_two_ball.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from rigidbody_simulation_b200 import stepper, synth
from rigidbody_simulation_b200.src.simulation import ball_collision

E = 1 << 20
K = int(sys.argv[1]) if len(sys.argv) > 1 else 256
s = synth.two_ball(E)
model, data = ball_collision.build(E, device=torch.device("cuda:0"), dtype=torch.float64)
data.set_state(s["qpos"], s["qvel"])
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
for i in range(4):
    ev[i].record()
    stepper.step_two_ball(model, data, 0.01, 1.0, 0.3, radius=0.1, substeps=K, count=(i == 3), arith="fast")
ev[4].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
print("launch ms:", [round(m, 3) for m in ms], "env-substeps/s (last):", E * K / (ms[-1] * 1e-3))
g, p = data.counters()
print("ground / pair events per env-substep in the last launch:", g.sum() / (E * K), p.sum() / (E * K))
