"""Host-side cost of one per-frame call (K = 1) at a size where the kernel is shorter than the call: 4096 envs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rigidbody_simulation_b200 as rb
from rigidbody_simulation_b200 import scenes, stepper, synth
E = 4096
s = synth.sphere_incline(E)
model = scenes.sphere_on_incline(E, device="cuda:0")
model.set_per_env(restitution=s["restitution"], friction=s["friction"])
data = rb.BatchedData(model)
data.set_state(s["qpos"], s["qvel"])
def loop(n):
    for _ in range(n):
        stepper.step_body_plane(model, data, -1, s["dt"], None, None, 0.0, substeps=1, count=False, arith="fast")
loop(200); torch.cuda.synchronize()
t0 = time.perf_counter(); loop(5000); torch.cuda.synchronize(); t = time.perf_counter() - t0
print("per-frame call, cached args: %.1f us" % (t / 5000 * 1e6))
from rigidbody_simulation_b200.src.physics.collision import custom_step_with_impulse_collision_friction as step
t0 = time.perf_counter()
for _ in range(2000):
    step(model, "ball", data, dt=s["dt"], restitution=None, friction_coeff=None, arith="fast")
torch.cuda.synchronize(); t = time.perf_counter() - t0
print("reference-named facade call (returns the [E,3] position view): %.1f us" % (t / 2000 * 1e6))
g = torch.cuda.CUDAGraph()
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    loop(3)
    with torch.cuda.graph(g, stream=st):
        loop(100)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50): g.replay()
    torch.cuda.synchronize(); t = time.perf_counter() - t0
print("CUDA graph of 100 per-frame launches: %.2f us per step" % (t / 5000 * 1e6))
