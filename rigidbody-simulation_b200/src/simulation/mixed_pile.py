"""Headless scenario of SURVEY.md section 8(f) row N4: spheres AND boxes in one scene (new; the reference has no such
script -- its ``multi_sphere_bounce.py`` is the closest, and the step below is that script's loop, body by body
(src/simulation/multi_sphere_bounce.py:42-92, repaired as in ``multi_sphere_bounce`` here) with the reference's impulse
(src/physics/collision.py:7-48), its application (src/physics/physics_utils.py:25-49) and world inertia
(collision.py:51-53) per contact, over a contact set widened by sphere-box and box-box pairs (DESIGN.md "N4")."""
import numpy as np
import torch

import rigidbody_simulation_b200.mj as mj
from rigidbody_simulation_b200 import scenes, stepper, synth

restitution_coefficient = 0.2          # the cube scenario's values (src/config/sim_overrides.py:8-15)
friction_coefficient = 0.6
timestep = 0.005
# Eight bodies: the cube of models/cube.xml, the sphere of models/sphere.xml, and anisotropic / other-size relatives.  The
# boxes are at least as heavy as the shipped cube on purpose: the reference's impulse uses the constant effective mass
# k = 1/m + 1/18 (collision.py:36) instead of 1/m + |r x n|^2 / I, which over-answers corner hits of a box by
# (1/m + 3/m) / k -- below m ~ 25 (the shipped cube has 25.6) that factor times (1 + e) exceeds 2 and bounces gain energy.
BODIES = [{"type": "box", "size": [0.4, 0.4, 0.4]}, {"type": "sphere", "size": [0.2]}, {"type": "box", "size": [0.5, 0.4, 0.35]},
          {"type": "sphere", "size": [0.25]}, {"type": "box", "size": [0.5, 0.5, 0.3]}, {"type": "sphere", "size": [0.15]},
          {"type": "box", "size": [0.35, 0.55, 0.4]}, {"type": "sphere", "size": [0.3]}]
PITCH = 1.2                            # lattice pitch of the initial pile


def bodies_for(n_body):
    return [dict(BODIES[i % len(BODIES)]) for i in range(n_body)]


def build(nenv=1, device=None, dtype=torch.float64, n_body=8, start=0, seed=synth.SEED):
    bodies = bodies_for(n_body)
    model = mj.MjModel.from_xml_string(scenes.multi_body_xml(bodies, timestep=timestep), nenv=nenv, device=device, dtype=dtype)
    data = mj.MjData(model, layout="body")
    s = synth.multi_body(nenv, bodies, start=start, seed=seed, pitch=PITCH)
    data.set_state(s["qpos"], s["qvel"])
    # per-body mass / inertia columns for the statistics kernel (the stepper itself reads the body table)
    tab = stepper.body_table(model)
    model.set_per_env(mass=np.tile(tab[:, 4], nenv), inertia=np.tile(tab[:, 5:8].T, (1, nenv)))
    return model, data


def custom_step_multi_body(model, data, dt=timestep, restitution=restitution_coefficient, substeps=1, friction=None):
    """Per body: gravity, impulses for every start-of-step contact touching the body, pose integration; None like
    ``custom_step_multi_sphere`` (multi_sphere_bounce.py:92)."""
    mj.mj_forward(model, data)
    stepper.step_multi_body(model, data, dt, restitution, friction_coefficient if friction is None else friction, substeps=substeps)
    return None


def run_headless(steps=400, nenv=1, device=None, dtype=torch.float64, substeps=1, n_body=8, start=0, seed=synth.SEED):
    model, data = build(nenv, device, dtype, n_body, start, seed)
    done = 0
    while done < steps:
        k = min(substeps, steps - done)
        custom_step_multi_body(model, data, model.opt.timestep, substeps=k)
        done += k
    return model, data, None


if __name__ == "__main__":
    _, d, _ = run_headless()
    print("final qpos", np.asarray(d.qpos))
