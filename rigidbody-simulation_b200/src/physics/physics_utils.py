"""Drop-in for the reference's ``src/physics/physics_utils.py`` (rows A2, A3 of SURVEY.md section 8):
``apply_impulse`` (:4-22) and ``apply_impulse_friction`` (:25-49), computed by CUDA kernels
(rbs_apply_impulse / rbs_apply_impulse_friction).  Same positional signatures, same return arity; a leading
batch dimension / CUDA tensors are accepted on every array argument."""
from rigidbody_simulation_b200.free_functions import apply_impulse, apply_impulse_friction

__all__ = ["apply_impulse", "apply_impulse_friction"]
