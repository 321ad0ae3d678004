"""rigidbody-simulation_b200: a B200-native (sm_100a) batched drop-in for the impulse / friction hot path of
pratyay2510/RigidBody-Simulation.

Import name: ``rigidbody_simulation_b200`` (the directory name carries a hyphen, so the repo root has a
small shim package of that name that points here).  Layout:

  csrc/             hand-written CUDA kernels + the C ABI (include/rbsim_b200.h) -> lib/librbsim_b200.so
  _lib.py           ctypes binding (fails loudly if the library is missing; no CPU fallback)
  mjcf.py           MuJoCo-XML subset parser (host, once per scene)
  batched.py        BatchedModel / BatchedData: device-resident SoA state for E environments
  stepper.py        argument marshalling + launches for the fused steppers and host-buffer drivers
  free_functions.py batched versions of the reference's free functions
  mj.py             the handful of ``mujoco`` symbols the reference's physics path uses
  synth.py          counter-based synthetic initial states for the BASELINE configs
  shard.py          environment sharding over ranks + optional end-of-run NCCL statistics gather
  headless.py       fixed-step headless driver (the start_main_loop contract without a window)
  src/              mirror of the reference's module paths (src.physics.collision, ...)
"""
from . import _lib, mjcf
from ._lib import LIB_PATH, RbsError, launch_count
from .batched import BatchedData, BatchedModel
from .free_functions import (apply_impulse, apply_impulse_friction, compute_collision_impulse,
                             compute_collision_impulse_friction, compute_inertia_tensor_world, compute_inverse_inertia)

__all__ = ["BatchedData", "BatchedModel", "LIB_PATH", "RbsError", "launch_count", "mjcf", "apply_impulse",
           "apply_impulse_friction", "compute_collision_impulse", "compute_collision_impulse_friction",
           "compute_inertia_tensor_world", "compute_inverse_inertia"]
