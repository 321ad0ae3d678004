"""A few launches of the multi-body stepper (N4) on the bench pile, for ncu: python profiles/prof_multi_body.py [envs] [bodies]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import rigidbody_simulation_b200 as rb  # noqa: F401
from rigidbody_simulation_b200 import stepper
from rigidbody_simulation_b200.src.simulation import mixed_pile

E = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
model, data = mixed_pile.build(E, device="cuda:0", dtype=torch.float64, n_body=B)
for _ in range(4):
    stepper.step_multi_body(model, data, mixed_pile.timestep, 0.2, 0.6, substeps=64, count=True)
torch.cuda.synchronize()
