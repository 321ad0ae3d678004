import os, sys
sys.path.insert(0, '.')
import torch
import rigidbody_simulation_b200 as rb
from rigidbody_simulation_b200 import scenes, stepper, synth
dev = torch.device("cuda:0")
E = 1 << 19
compact = int(sys.argv[1])
s = synth.sphere_incline(E)
model = scenes.sphere_on_incline(E, device=dev)
model.set_per_env(restitution=s["restitution"], friction=s["friction"])
data = rb.BatchedData(model)
data.set_state(s["qpos"], s["qvel"])
rb._lib.set_option("strict_compact", compact)
for _ in range(3):
    stepper.step_body_plane(model, data, -1, s["dt"], None, None, 0.0, substeps=128, count=False, arith="strict")
torch.cuda.synchronize()
