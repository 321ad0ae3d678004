"""N4 (spheres and boxes in one scene): throughput of rbs_step_multi_body on one B200 with the C oracle (OpenMP, all
host cores) timed beside it on a sample; register-cap variants (option mb_minb); one JSON line per run.
    python profiles/bench_multi_body.py [envs] [substeps_per_launch] [launches]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import rigidbody_simulation_b200 as rb
import rigidbody_simulation_b200.mj as mj
from rigidbody_simulation_b200 import scenes, stepper, synth

E = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
K = int(sys.argv[2]) if len(sys.argv) > 2 else 64
L = int(sys.argv[3]) if len(sys.argv) > 3 else 8
BODIES = [{"type": "box", "size": [0.4, 0.4, 0.4]}, {"type": "sphere", "size": [0.2]}, {"type": "box", "size": [0.3, 0.2, 0.25]},
          {"type": "sphere", "size": [0.25]}, {"type": "box", "size": [0.35, 0.35, 0.15]}, {"type": "sphere", "size": [0.15]},
          {"type": "box", "size": [0.2, 0.45, 0.3]}, {"type": "sphere", "size": [0.3]}]
dev = torch.device("cuda:0")
for B in (8, 32):
    bodies = (BODIES * (B // 8))[:B]
    s = synth.multi_body(E, bodies, pitch=0.8)
    for dtype, tag in ((torch.float64, "fp64"), (torch.float32, "fp32")):
        model = mj.MjModel.from_xml_string(scenes.multi_body_xml(bodies), nenv=E, dtype=dtype, device=dev)
        data = mj.MjData(model, layout="body")
        ref = None
        for minb in (1, 2, 3):
            rb._lib.set_option("mb_minb", minb)
            best = None
            for rep in range(2):
                data.set_state(s["qpos"], s["qvel"])
                data.n_contacts.zero_()
                data.n_impulses.zero_()
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(L + 1)]
                for i in range(L):
                    ev[i].record()
                    stepper.step_multi_body(model, data, 0.005, 0.2, 0.6, substeps=K, count=True)
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
            calls, imps = data.counters()
            print(json.dumps({"bodies_per_env": B, "envs": E, "dtype": tag, "mb_minb": minb, "launch_ms": [round(m, 3) for m in best],
                              "body_substeps_per_s": E * B * K * L / (sum(best) * 1e-3), "env_substeps_per_s": E * K * L / (sum(best) * 1e-3),
                              "contacts_per_body_substep": float(calls.sum()) / (E * B * K * L),
                              "impulses_per_body_substep": float(imps.sum()) / (E * B * K * L),
                              "state_bitwise_equal_to_first_variant": same}), flush=True)
        rb._lib.set_option("mb_minb", 0)
    # the C oracle on all host cores, a sample of the same scene
    import c_oracle as co
    n = 64 * (os.cpu_count() or 1)
    tab = stepper.body_table(model)
    qp = s["qpos"][:n].reshape(n, B, 7).copy()
    qv = s["qvel"][:n].reshape(n, B, 6).copy()
    t0 = time.perf_counter()
    co.step_multi_body(qp, qv, 256, gtype=tab[:, 0].astype(np.int32), mass=tab[:, 4], inertia=tab[:, 5:8], size=tab[:, 1:4],
                       plane_pos=[0, 0, 0], plane_normal=[0, 0, 1], gravity=[0, 0, -9.8], dt=0.005, restitution=0.2, friction=0.6)
    dt = time.perf_counter() - t0
    print(json.dumps({"bodies_per_env": B, "cpu_oracle_openmp": {"cores": co.max_threads(), "envs": n, "steps": 256,
                                                                  "body_substeps_per_s": n * B * 256 / dt}}), flush=True)
