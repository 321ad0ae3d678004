"""Scratch experiment (kept for the record): how much would contact-rate-aware binning of environments into warps
buy?  Same 1,048,576-env workload, (a) in generator order, (b) sorted by restitution, (c) sorted by the per-env
contact count observed over the previous 2048 substeps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rigidbody_simulation_b200 as rb
from rigidbody_simulation_b200 import scenes, stepper, synth
E = 1 << 20
s = synth.sphere_incline(E)
dev = torch.device("cuda:0")
def run(order, label):
    model = scenes.sphere_on_incline(E, device=dev)
    model.set_per_env(restitution=s["restitution"][order], friction=s["friction"][order])
    data = rb.BatchedData(model)
    data.set_state(s["qpos"][order], s["qvel"][order])
    kw = dict(dt=s["dt"], restitution=None, friction_coeff=None, contact_threshold=0.0, substeps=128, arith="fast")
    for _ in range(64):                                   # reach the steady regime of the bench (8192 substeps)
        stepper.step_body_plane(model, data, -1, count=False, **kw)
    ms = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(16):
            stepper.step_body_plane(model, data, -1, count=False, **kw)
        b.record(); b.synchronize(); ms.append(a.elapsed_time(b))
    data.n_contacts.zero_()
    for _ in range(16):
        stepper.step_body_plane(model, data, -1, count=True, **kw)
    torch.cuda.synchronize()
    c = data.n_contacts.cpu().numpy().astype(np.int64)
    print("%-28s %.3f ms per 2048 substeps  (%.3e env-substeps/s)  mean contacts/env-substep %.3f" %
          (label, np.median(ms), E * 2048 / (np.median(ms) * 1e-3), c.mean() / 2048), flush=True)
    return c
ident = np.arange(E)
c0 = run(ident, "generator order")
run(np.argsort(s["restitution"], kind="stable"), "sorted by restitution")
run(np.argsort(c0, kind="stable"), "sorted by contact count")
hist = np.histogram(c0 / 2048.0, bins=[0, 0.001, 0.01, 0.05, 0.2, 0.5, 0.9, 1.01])[0] / E
print("share of envs by contacts per substep [0,.001,.01,.05,.2,.5,.9,1]:", np.round(hist, 3))
