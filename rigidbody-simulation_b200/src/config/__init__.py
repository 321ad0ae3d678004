"""Same keys and values as the reference's ``src/config`` (``load_sim_config``, src/config/__init__.py:7-19):
global defaults (global_sim_params.py:1-5) overlaid by per-simulation overrides (sim_overrides.py:1-28).
Camera and recording entries belong to the viewer, which is out of scope; the keys exist and are empty.
Note that the physics never reads TIMESTEP: the loop passes ``model.opt.timestep`` from the XML
(src/viewer/mujoco_viewer.py:113)."""

GLOBAL_DEFAULTS = {"FRICTION_COEFFICIENT": 0.5, "RESTITUTION": 0.9, "TIMESTEP": 0.01, "INCLINE_ANGLE_RAD": 0.0,
                   "RECORD_VIDEO": True}

SIMULATION_OVERRIDES = {
    "single_sphere_bounce": {"FRICTION_COEFFICIENT": 0.5, "RESTITUTION": 1.0, "TIMESTEP": 0.01,
                             "INCLINE_ANGLE_RAD": 0.0, "RECORD_VIDEO": True},
    "cube_incline": {"FRICTION_COEFFICIENT": 0.6, "RESTITUTION": 0.2, "TIMESTEP": 0.009, "INCLINE_ANGLE_RAD": 0.7,
                     "RECORD_VIDEO": True},
    "ball_collision": {"FRICTION_COEFFICIENT": 0.3, "RESTITUTION": 1.0, "TIMESTEP": 0.01, "RECORD_VIDEO": True},
    "multi_sphere_bounce": {"FRICTION_COEFFICIENT": 0.0, "RESTITUTION": 1.0, "TIMESTEP": 0.01, "RECORD_VIDEO": True},
}


def load_sim_config(simulation_name):
    config = dict(GLOBAL_DEFAULTS)
    config["CAMERA"] = {}
    config["RECORDING_PATH"] = None
    config.update(SIMULATION_OVERRIDES.get(simulation_name, {}))
    return config
