"""Scratch experiment (kept for the record): does splitting the batch into independent chains on separate streams
recover the ragged last wave of each fused launch?  1,048,576 envs, fp64 fast, 16 x 128 substeps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rigidbody_simulation_b200 as rb
from rigidbody_simulation_b200 import scenes, stepper, synth
E = 1 << 20
s = synth.sphere_incline(E)
dev = torch.device("cuda:0")
model = scenes.sphere_on_incline(E, device=dev)
model.set_per_env(restitution=s["restitution"], friction=s["friction"])
data = rb.BatchedData(model)
data.set_state(s["qpos"], s["qvel"])
kw = dict(dt=s["dt"], restitution=None, friction_coeff=None, contact_threshold=0.0, substeps=128, count=False, arith="fast")
def whole():
    for _ in range(16):
        stepper.step_body_plane(model, data, -1, **kw)
def split(parts):
    ch = stepper.SplitChains(model, data, parts)
    def run():
        ch.fork()
        for _ in range(16):
            ch.step(**kw)
        ch.join()
    return run
def timed(fn, reps=6):
    for _ in range(3): fn()
    ms = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ms.append(a.elapsed_time(b))
    return np.median(ms)
for _ in range(4): whole()
print("whole   %.3f ms" % timed(whole))
for parts in (2, 3, 4, 8):
    print("split %d %.3f ms" % (parts, timed(split(parts))))
print("whole   %.3f ms" % timed(whole))
