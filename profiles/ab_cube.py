"""A/B of the plane-frame box kernels on config 4 (cube on a plane, 1,048,576 envs): thread-per-environment kernel
(option box_compact=0) against the CTA-compacting one (box_compact=1), bounce and incline, fp64 and fp32, 5 and 6
resident CTAs per SM.  Every run starts from the config's initial state and advances 2048 substeps in 16 launches of
128; one JSON line per run.  Device-timed, state resident in HBM.
    python profiles/ab_cube.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import rigidbody_simulation_b200 as rb
from rigidbody_simulation_b200 import scenes, stepper, synth

dev = torch.device("cuda:0")
E, K, L = 1 << 20, 128, 16
for dtype, tag in ((torch.float64, "fp64"), (torch.float32, "fp32")):
    for kind in ("bounce", "incline"):
        s = synth.cube(E, kind=kind)
        model = scenes.cube_on_plane(E, theta=s["theta"], device=dev, dtype=dtype)
        data = rb.BatchedData(model)
        ref = None
        for compact in (0, 1):
            for minb in (5, 6):
                rb._lib.set_option("box_compact", compact)
                rb._lib.set_option("box_minb", minb)
                best = None
                for rep in range(2):
                    data.set_state(s["qpos"], s["qvel"])
                    ev = [torch.cuda.Event(enable_timing=True) for _ in range(L + 1)]
                    for i in range(L):
                        ev[i].record()
                        stepper.step_body_plane(model, data, -1, s["dt"], 0.2, 0.6, 1e-4, substeps=K, count=False, arith="fast")
                    ev[L].record()
                    torch.cuda.synchronize()
                    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(L)]
                    if best is None or sum(ms) < sum(best):
                        best = ms
                same = None
                if ref is None:
                    ref = data.state.clone()
                else:
                    same = bool(torch.equal(ref, data.state))
                print(json.dumps({"dtype": tag, "kind": kind, "box_compact": compact, "box_minb": minb,
                                  "launch_ms": [round(m, 3) for m in best], "env_substeps_per_s_2048": E * K * L / (sum(best) * 1e-3),
                                  "env_substeps_per_s_last_launch": E * K / (best[-1] * 1e-3),
                                  "state_bitwise_equal_to_first_variant": same}), flush=True)
rb._lib.set_option("box_compact", 0)
rb._lib.set_option("box_minb", 6)
