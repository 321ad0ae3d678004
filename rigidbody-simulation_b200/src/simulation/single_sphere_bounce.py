"""Headless counterpart of the reference's ``src/simulation/single_sphere_bounce.py`` (config 1).

Same configuration sources (``load_sim_config("single_sphere_bounce")`` :14, ``models/sphere.xml`` with the
``{INCLINE_ANGLE}`` / ``{TIMESTEP}`` templating :26-36), same initial conditions (zero linear velocity, spin
(2, 2, 0), :40-41), same step wrapper (:65-70, including obj="sphere", which is not a body of the scene: the
lookup returns -1 and selects the last body, as shipped).  Window, video and plots are out of scope."""
import functools

import numpy as np
import torch

import rigidbody_simulation_b200.mj as mj
from rigidbody_simulation_b200 import scenes
from rigidbody_simulation_b200.headless import TrajectoryLog, start_main_loop

from ..config import load_sim_config
from ..physics.collision import custom_step_with_impulse_collision_friction

config = load_sim_config("single_sphere_bounce")
friction_coefficient = config["FRICTION_COEFFICIENT"]
restitution = config["RESTITUTION"]
timestep = config["TIMESTEP"]
incline_angle_rad = config["INCLINE_ANGLE_RAD"]


def build(nenv=1, device=None, dtype=torch.float64):
    model = mj.MjModel.from_xml_path(scenes.model_path("sphere"), nenv=nenv, device=device, dtype=dtype,
                                     incline_angle=incline_angle_rad, timestep=timestep)
    data = mj.MjData(model)
    if nenv == 1:
        data.qvel[:3] = np.array([0.0, 0.0, 0.0])        # :40
        data.qvel[3:6] = np.array([2.0, 2.0, 0.0])       # :41
    else:
        data.qvel[:, 0:3] = 0.0
        data.qvel[:, 3:6] = torch.tensor([2.0, 2.0, 0.0], dtype=dtype)
    return model, data


def sphere_simulation_step(model, data, dt, substeps=1, trajectory=None, arith="strict"):
    return custom_step_with_impulse_collision_friction(model, "sphere", data, dt=dt, restitution=restitution,
                                                       friction_coeff=friction_coefficient, substeps=substeps,
                                                       trajectory=trajectory, arith=arith)


def run_headless(steps=2000, nenv=1, device=None, dtype=torch.float64, log=True, substeps_per_launch=1, arith="strict"):
    model, data = build(nenv, device, dtype)
    logger = TrajectoryLog(steps, min(nenv, 4), model.device, dtype) if log else None
    step = sphere_simulation_step if arith == "strict" else functools.partial(sphere_simulation_step, arith=arith)
    start_main_loop(model, data, step, steps, logger, substeps_per_launch)
    if logger is not None:
        logger.finish()
    return model, data, logger


if __name__ == "__main__":
    _, d, lg = run_headless()
    print("final qpos", np.asarray(d.qpos), "contacts/impulses", d.counters())
