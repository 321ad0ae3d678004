"""ncu target: config 5 (64 spheres per env), fast policy, fp64.  The first launch of 128 substeps is the dense early
phase (the lattice collapsing), the fifth the steady regime:
    ncu --set full --clock-control none --import-source on -k regex:step_multi_sphere -s 0 -c 1 \
        -o gpurun_out/prof_ms python profiles/prof_multi_sphere.py [envs] [mu] [fast|strict] [substeps per launch]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from rigidbody_simulation_b200 import stepper, synth
from rigidbody_simulation_b200.src.simulation import multi_sphere_bounce

E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
MU = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
ARITH = sys.argv[3] if len(sys.argv) > 3 else "fast"
K = int(sys.argv[4]) if len(sys.argv) > 4 else 128
s = synth.multi_sphere(E, n_body=64, friction=MU)
model, data = multi_sphere_bounce.build(E, device=torch.device("cuda:0"), dtype=torch.float64, n_body=64)
data.set_state(s["qpos"], s["qvel"])
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
for i in range(5):
    ev[i].record()
    stepper.step_multi_sphere(model, data, 0.01, 1.0, MU, substeps=K, count=(i == 4), arith=ARITH)
ev[5].record()
torch.cuda.synchronize()
print("launch ms:", [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(5)])
calls, imps = data.counters()
print("contacts / impulses per body-substep in the last launch:", calls.sum() / (E * 64 * K), imps.sum() / (E * 64 * K))
