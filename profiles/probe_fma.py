"""Scratch probe (kept for the record): FMA-peak microbenchmark, operand forms and chain counts.
Run each mode in its own process: RBS_PROBE_MODE=0|1|2 python profiles/probe_fma.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rigidbody_simulation_b200 import stepper
for dt in (torch.float64, torch.float32):
    for bps in (2, 4, 8):
        print("mode", os.environ.get("RBS_PROBE_MODE", "1"), dt, "blocks/SM", bps, "%.2f TFLOP/s" % (stepper.fma_peak("cuda:0", dt, 32768, bps) / 1e12), flush=True)
