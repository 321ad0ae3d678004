"""Headless counterpart of the reference's ``src/simulation/ball_collision.py`` (config 3).

``compute_inverse_inertia`` (:39-41), ``compute_collision_impulse`` (:53-68) and
``step_with_custom_collisions`` (:73-125) keep their names and signatures.  The reference reads masses, inverse
inertias, restitution, friction and the hard-coded ``ball_radius = 0.1`` (:23) from module globals; here the
same values come from the model / config, with keyword overrides.  Initial conditions as shipped (:31-34).
The reference only steps while ``running`` is True (toggled by SPACE, :131-141); headless runs are un-paused."""
import numpy as np
import torch

import rigidbody_simulation_b200.mj as mj
from rigidbody_simulation_b200 import scenes, stepper
from rigidbody_simulation_b200.free_functions import compute_collision_impulse, compute_inverse_inertia  # noqa: F401

from ..config import load_sim_config

config = load_sim_config("ball_collision")
friction_coefficient = config["FRICTION_COEFFICIENT"]
restitution = config["RESTITUTION"]
timestep = config["TIMESTEP"]
ball_radius = 0.1                                           # :23 hard-coded geometry parameter


def build(nenv=1, device=None, dtype=torch.float64):
    model = mj.MjModel.from_xml_path(scenes.model_path("ball_collision"), nenv=nenv, device=device, dtype=dtype)
    data = mj.MjData(model)
    qpos = np.tile(np.array([-1.0, 0.0, 1.0, 1, 0, 0, 0, 1.0, 0.0, 1.0, 1, 0, 0, 0]), (nenv, 1))    # :31-32
    qvel = np.tile(np.array([1.0, 0.0, 0.5, 0, 0, 0, -1.0, 0.0, 0.5, 0, 0, 0]), (nenv, 1))          # :33-34
    data.set_state(qpos, qvel)
    return model, data


def step_with_custom_collisions(model, data, dt=0.01, substeps=1, restitution=restitution,
                                friction_coefficient=friction_coefficient, ball_radius=ball_radius, arith="strict"):
    """Gravity on both balls, ball-ground impulses with the z clamp, the one-sided ball-ball impulse with the
    symmetric positional correction, explicit position integration; quaternions untouched (:73-125).
    Returns both ball positions like the reference (:125)."""
    mj.mj_forward(model, data)                               # :74 (no effect on the results there either)
    stepper.step_two_ball(model, data, dt, restitution, friction_coefficient, radius=ball_radius, substeps=substeps,
                          arith=arith)
    rows = data.rows(0, 3)                                   # [3, 2, E]
    if data.squeeze:
        host = rows[:, :, 0].cpu().numpy()
        return host[:, 0].copy(), host[:, 1].copy()
    return rows[:, 0, :].t(), rows[:, 1, :].t()


def ball_collision_step(model, data, dt):
    step_with_custom_collisions(model, data, dt)
    return None                                              # two separate logs in the reference (:146-154)


def run_headless(steps=500, nenv=1, device=None, dtype=torch.float64, substeps=1, arith="strict"):
    model, data = build(nenv, device, dtype)
    done = 0
    while done < steps:                                     # launches of `substeps`, ragged last one
        k = min(substeps, steps - done)
        step_with_custom_collisions(model, data, model.opt.timestep, substeps=k, arith=arith)
        done += k
    return model, data, None


if __name__ == "__main__":
    _, d, _ = run_headless()
    print("final qpos", np.asarray(d.qpos))
