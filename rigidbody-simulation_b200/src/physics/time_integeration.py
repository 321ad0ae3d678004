"""Drop-in for the reference's ``src/physics/time_integeration.py`` (rows A4, A6, A7 of SURVEY.md section 8;
the file name keeps the reference's spelling).

``timestep_integration`` (:13-72) is scheme A with the defaults friction_coeff=0.5, contact_threshold=1e-4;
``general`` (:75-141) moves the position with the OLD velocity, then updates velocities and resolves contacts,
and never integrates the quaternion."""
import rigidbody_simulation_b200.mj as mj
from rigidbody_simulation_b200 import stepper
from rigidbody_simulation_b200._lib import RBS_SCHEME_A, RBS_SCHEME_GENERAL
from rigidbody_simulation_b200.free_functions import compute_inertia_tensor_world  # noqa: F401  (:8-10)

from .collision import _position, compute_collision_impulse_friction  # noqa: F401  (:4)
from .physics_utils import apply_impulse_friction  # noqa: F401  (:5)


def timestep_integration(model, obj, data, dt=0.01, restitution=1.0, friction_coeff=0.5, contact_threshold=1e-4,
                         substeps=1, arith="strict", trajectory=None):
    mj.mj_forward(model, data)                                              # :29
    body_id = mj.mj_name2id(model, mj.mjtObj.mjOBJ_BODY, f"{obj}")          # :30
    stepper.step_body_plane(model, data, body_id, dt, restitution, friction_coeff, contact_threshold,
                            scheme=RBS_SCHEME_A, substeps=substeps, arith=arith, trajectory=trajectory)
    return _position(data)


def general(model, obj, data, dt=0.01, restitution=1.0, friction_coeff=0.5, contact_threshold=1e-4, substeps=1,
            trajectory=None):
    mj.mj_forward(model, data)                                              # :95
    body_id = mj.mj_name2id(model, mj.mjtObj.mjOBJ_BODY, f"{obj}")          # :96
    stepper.step_body_plane(model, data, body_id, dt, restitution, friction_coeff, contact_threshold,
                            scheme=RBS_SCHEME_GENERAL, substeps=substeps, trajectory=trajectory)
    return _position(data)
