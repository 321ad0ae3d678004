"""Drop-in for the reference's ``src/physics/collision.py`` (rows A1, A4, A5 of SURVEY.md section 8).

``compute_collision_impulse_friction`` (:7-48), ``compute_inertia_tensor_world`` (:51-53) and
``custom_step_with_impulse_collision_friction`` (:56-102) keep their names, argument order and defaults;
``model`` / ``data`` are ``rigidbody_simulation_b200.mj.MjModel`` / ``MjData`` objects holding any number of
environments on the GPU.  Extra keywords: ``substeps`` fuses that many integration steps into the one launch;
``trajectory`` (device tensor ``[substeps, n, 3]``) receives the position of the first n environments after every one
of them -- what the reference's per-frame ``logger.record`` would have seen.
"""
import numpy as np

import rigidbody_simulation_b200.mj as mj
from rigidbody_simulation_b200 import stepper
from rigidbody_simulation_b200._lib import RBS_SCHEME_A
from rigidbody_simulation_b200.free_functions import compute_collision_impulse_friction, compute_inertia_tensor_world

from .physics_utils import apply_impulse, apply_impulse_friction  # noqa: F401  (re-exported like the reference, :2)

__all__ = ["compute_collision_impulse_friction", "compute_inertia_tensor_world",
           "custom_step_with_impulse_collision_friction", "apply_impulse", "apply_impulse_friction"]


def _position(data):
    """pos_new as the reference returns it: a fresh (3,) array for one environment; for a batch the [E,3]
    view of the device state (no copy, no sync)."""
    rows = data.rows(0, 3)[:, 0, :]                    # [3, E]
    if data.squeeze:
        return rows[:, 0].cpu().numpy().astype(np.float64)
    return rows.t()


def custom_step_with_impulse_collision_friction(model, obj, data, dt=0.01, restitution=1.0, friction_coeff=1.0,
                                                contact_threshold=0, substeps=1, arith="strict", trajectory=None):
    """One step of scheme A for every environment in ``data`` (in place), reference lines :56-102:
    contacts of the start-of-step pose, gravity / applied wrench, sequential per-contact impulses
    (normal + Coulomb-clamped tangential), then position and first-order quaternion integration."""
    mj.mj_forward(model, data)                                              # :57 (contacts are generated in-kernel)
    body_id = mj.mj_name2id(model, mj.mjtObj.mjOBJ_BODY, f"{obj}")          # :58 (-1 -> last body, as shipped)
    stepper.step_body_plane(model, data, body_id, dt, restitution, friction_coeff, contact_threshold,
                            scheme=RBS_SCHEME_A, substeps=substeps, arith=arith, trajectory=trajectory)
    return _position(data)
