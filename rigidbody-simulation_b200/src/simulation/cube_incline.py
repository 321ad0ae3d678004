"""Headless counterpart of the reference's ``src/simulation/cube_incline.py`` (config 4, shipped pose).

``load_sim_config("cube_incline")`` (:14), ``models/cube.xml`` (:33-42), cube from rest (:46), step wrapper with
``timestep_integration`` passing dt, restitution and friction but NOT the threshold, so its default 1e-4 applies
(:70-78)."""
import functools

import numpy as np
import torch

import rigidbody_simulation_b200.mj as mj
from rigidbody_simulation_b200 import scenes
from rigidbody_simulation_b200.headless import TrajectoryLog, start_main_loop

from ..config import load_sim_config
from ..physics.time_integeration import timestep_integration

config = load_sim_config("cube_incline")
friction_coefficient = config["FRICTION_COEFFICIENT"]
restitution = config["RESTITUTION"]
timestep = config["TIMESTEP"]
incline_angle_rad = config["INCLINE_ANGLE_RAD"]
obj = "cube"


def build(nenv=1, device=None, dtype=torch.float64):
    model = mj.MjModel.from_xml_path(scenes.model_path(obj), nenv=nenv, device=device, dtype=dtype,
                                     incline_angle=incline_angle_rad, timestep=timestep)
    data = mj.MjData(model)
    if nenv == 1:
        data.qvel[:6] = 0.0                                # :46
    return model, data


def cube_incline_step(model, data, dt, substeps=1, trajectory=None, arith="strict"):
    return timestep_integration(model, obj, data, dt=dt, restitution=restitution, friction_coeff=friction_coefficient,
                                substeps=substeps, trajectory=trajectory, arith=arith)


def run_headless(steps=240, nenv=1, device=None, dtype=torch.float64, log=True, substeps_per_launch=1, arith="strict"):
    model, data = build(nenv, device, dtype)
    logger = TrajectoryLog(steps, min(nenv, 4), model.device, dtype) if log else None
    step = cube_incline_step if arith == "strict" else functools.partial(cube_incline_step, arith=arith)
    start_main_loop(model, data, step, steps, logger, substeps_per_launch)
    if logger is not None:
        logger.finish()
    return model, data, logger


if __name__ == "__main__":
    _, d, lg = run_headless()
    print("final qpos", np.asarray(d.qpos), "contacts/impulses", d.counters())
