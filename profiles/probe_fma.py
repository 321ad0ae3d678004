"""Scratch probe (kept for the record): FMA-peak microbenchmark sensitivity to iterations / occupancy."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rigidbody_simulation_b200 import stepper
for dt in (torch.float64, torch.float32):
    for iters in (4096, 32768, 131072):
        for bps in (4, 8):
            print(dt, iters, bps, "%.2f TFLOP/s" % (stepper.fma_peak("cuda:0", dt, iters, bps) / 1e12), flush=True)
