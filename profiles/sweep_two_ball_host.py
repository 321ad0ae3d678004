import json, os, sys, time
sys.path.insert(0, '.')
import torch
import rigidbody_simulation_b200 as rb
from rigidbody_simulation_b200 import stepper, synth
from rigidbody_simulation_b200.src.simulation import ball_collision
dev = torch.device("cuda:0")
E = 1 << 20
s = synth.two_ball(E)
model, _ = ball_collision.build(E, device=dev, dtype=torch.float64)
qp0 = torch.from_numpy(s["qpos"]).pin_memory(); qv0 = torch.from_numpy(s["qvel"]).pin_memory()
qp, qv = qp0.clone().pin_memory(), qv0.clone().pin_memory()
for chunks, wave in ((16, 0),):
    rb._lib.set_option("host_chunks", chunks)
    for S, F in ((1, 1), (256, 256), (1024, 256), (2048, 256), (4096, 256)):
        best = None
        for rep in range(4):
            qp.copy_(qp0); qv.copy_(qv0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            stepper.run_two_ball_host(model, qp, qv, S, dt=0.01, restitution=1.0, friction=0.3, radius=0.1, substeps=F, arith="fast")
            dt = time.perf_counter() - t0
            if rep and (best is None or dt < best): best = dt
        print(json.dumps({"host_chunks": chunks, "total_steps": S, "ms_per_call": round(best * 1e3, 3)}), flush=True)
