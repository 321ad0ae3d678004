"""A/B of the host-buffer pipeline (rbs_run_body_plane_host / rbs_run_two_ball_host): compute streams (option
host_streams), chunk count (host_chunks) and chunk quantum (host_wave_ctas) on config 2 and config 3 at 1,048,576 envs,
2048 substeps per call, pinned host buffers; best of 3 calls, device-resident time of the same work beside it.
    python profiles/ab_host_pipeline.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import rigidbody_simulation_b200 as rb
from rigidbody_simulation_b200 import scenes, stepper, synth
from rigidbody_simulation_b200.src.simulation import ball_collision

dev = torch.device("cuda:0")
E, S, F = 1 << 20, 2048, 256
combos = [(2, 16, 0), (3, 16, 0), (4, 16, 0), (2, 24, 0), (3, 24, 0), (4, 32, 0), (2, 16, 3), (3, 16, 3), (3, 24, 3), (4, 32, 3), (3, 16, 2), (4, 32, 2), (2, 16, 6)]
for name in ("sphere_incline", "two_ball"):
    if name == "sphere_incline":
        s = synth.sphere_incline(E)
        model = scenes.sphere_on_incline(E, device=dev, dtype=torch.float64)
        model.set_per_env(restitution=s["restitution"], friction=s["friction"])
        call = lambda qp, qv: stepper.run_body_plane_host(model, qp, qv, S, substeps=F, arith="fast", dt=s["dt"], restitution=None,
                                                          friction_coeff=None, contact_threshold=0.0)
    else:
        s = synth.two_ball(E)
        model, _ = ball_collision.build(E, device=dev, dtype=torch.float64)
        call = lambda qp, qv: stepper.run_two_ball_host(model, qp, qv, S, dt=0.01, restitution=1.0, friction=0.3, radius=0.1,
                                                        substeps=F, arith="fast")
    qp0 = torch.from_numpy(s["qpos"]).pin_memory()
    qv0 = torch.from_numpy(s["qvel"]).pin_memory()
    qp, qv = qp0.clone().pin_memory(), qv0.clone().pin_memory()
    ref = None
    for streams, chunks, wave in combos:
        if name == "two_ball" and wave:
            continue
        rb._lib.set_option("host_streams", streams)
        rb._lib.set_option("host_chunks", chunks)
        rb._lib.set_option("host_wave_ctas", wave)
        best = None
        for rep in range(4):
            qp.copy_(qp0); qv.copy_(qv0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            call(qp, qv)
            dt = time.perf_counter() - t0
            if rep and (best is None or dt < best):
                best = dt
        same = None
        if ref is None:
            ref = (qp.clone(), qv.clone())
        else:
            same = bool(torch.equal(ref[0], qp) and torch.equal(ref[1], qv))
        print(json.dumps({"config": name, "host_streams": streams, "host_chunks": chunks, "host_wave_ctas": wave, "ms_per_call": round(best * 1e3, 3),
                          "env_substeps_per_s_e2e": E * S / best, "same_result_as_first": same}), flush=True)
rb._lib.set_option("host_streams", 3)
rb._lib.set_option("host_chunks", 16)
rb._lib.set_option("host_wave_ctas", 0)
