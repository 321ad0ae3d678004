"""Headless counterpart of the reference's ``src/simulation/multi_sphere_bounce.py`` (config 5).

The shipped file cannot run: (1) it uses the MuJoCo *body id* (world = 0, balls 1..4) as a 0-based joint index,
so ball4 reads the empty slice ``qpos[28:35]`` and ``compute_inertia_tensor_world`` raises IndexError on the first
step (:47-55); (2) ``model.id2name`` does not exist on ``mujoco.MjModel`` and the geoms are named
``ball_geomN`` / ``ground``, so the ownership filter could never match (:66); (3) ``glfw`` is used without being
imported (:98).  This module implements the evidently intended semantics:
ball b <-> slices 7b / 6b; a contact belongs to ball b iff one of its geoms is ball b's; contacts are visited in
MuJoCo order (ground first, then partners by ascending index); the normal is used as generated (lower index ->
higher index, never flipped) and the partner is treated as static, exactly as ``compute_collision_impulse_friction``
does (src/physics/collision.py:27)."""
import numpy as np
import torch

import rigidbody_simulation_b200.mj as mj
from rigidbody_simulation_b200 import scenes, stepper

from ..config import load_sim_config

config = load_sim_config("multi_sphere_bounce")
friction_coefficient = config["FRICTION_COEFFICIENT"]
restitution_coefficient = config["RESTITUTION"]
timestep = config["TIMESTEP"]
ball_names = ["ball1", "ball2", "ball3", "ball4"]            # :29


def build(nenv=1, device=None, dtype=torch.float64, n_body=None):
    if n_body is None:
        model = mj.MjModel.from_xml_path(scenes.model_path("multi_sphere"), nenv=nenv, device=device, dtype=dtype)
    else:
        model = mj.MjModel.from_xml_string(scenes.multi_sphere_xml(n_body), nenv=nenv, device=device, dtype=dtype)
    return model, mj.MjData(model, layout="body")


def custom_step_multi_sphere(model, data, dt=timestep, restitution=restitution_coefficient, substeps=1,
                             friction=None, arith="strict"):
    """Per ball: gravity, impulses for every start-of-step contact touching the ball, pose integration (:42-92).
    Returns None like the reference (:92)."""
    mj.mj_forward(model, data)                               # :43
    stepper.step_multi_sphere(model, data, dt, restitution, friction_coefficient if friction is None else friction,
                              substeps=substeps, arith=arith)
    return None


def run_headless(steps=300, nenv=1, device=None, dtype=torch.float64, substeps=1, n_body=None, arith="strict"):
    model, data = build(nenv, device, dtype, n_body)
    done = 0
    while done < steps:                                     # launches of `substeps`, ragged last one
        k = min(substeps, steps - done)
        custom_step_multi_sphere(model, data, model.opt.timestep, substeps=k, arith=arith)
        done += k
    return model, data, None


if __name__ == "__main__":
    _, d, _ = run_headless()
    print("final qpos", np.asarray(d.qpos))
