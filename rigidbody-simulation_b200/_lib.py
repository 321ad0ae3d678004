"""ctypes binding of librbsim_b200.so (the C ABI declared in include/rbsim_b200.h).

The library is the product: if it is missing this module raises -- there is no CPU or PyTorch
fallback anywhere in the package.  Build it with ``python rigidbody-simulation_b200/csrc/build.py``
(or ``__graft_entry__.build()``).
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_int, c_long, c_ubyte, c_uint, c_ulonglong, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "librbsim_b200.so")

RBS_F32, RBS_F64 = 0, 1
RBS_GEOM_SPHERE, RBS_GEOM_BOX = 0, 1
RBS_SCHEME_A, RBS_SCHEME_GENERAL = 0, 1
RBS_INERTIA_GENERAL, RBS_INERTIA_ISOTROPIC = 0, 1
RBS_ARITH_STRICT, RBS_ARITH_FAST = 0, 1
RBS_OK, RBS_EINVAL, RBS_ECUDA, RBS_ENOMEM = 0, -1, -2, -3

D3 = c_double * 3
D2 = c_double * 2


class BodyPlaneArgs(Structure):
    """struct rbs_body_plane_args"""
    _fields_ = [
        ("dtype", c_int), ("geom", c_int), ("scheme", c_int), ("inertia_mode", c_int),
        ("n_env", c_long), ("stride", c_long), ("param_stride", c_long), ("substeps", c_int), ("arith", c_int),
        ("state", c_void_p),
        ("mass", c_void_p), ("mass_u", c_double),
        ("inertia", c_void_p), ("inertia_u", D3),
        ("size", c_void_p), ("size_u", D3),
        ("restitution", c_void_p), ("restitution_u", c_double),
        ("friction", c_void_p), ("friction_u", c_double),
        ("xfrc", c_void_p),
        ("plane_point", D3), ("plane_normal", D3), ("gravity", D3),
        ("dt", c_double), ("contact_threshold", c_double),
        ("n_contacts", c_void_p), ("n_impulses", c_void_p),
        ("trajectory", c_void_p), ("trajectory_envs", c_long),
        ("stream", c_void_p),
    ]


class TwoBallArgs(Structure):
    """struct rbs_two_ball_args"""
    _fields_ = [
        ("dtype", c_int), ("substeps", c_int), ("arith", c_int), ("reserved", c_int),
        ("n_env", c_long), ("stride", c_long),
        ("state", c_void_p),
        ("mass", c_void_p), ("mass_u", D2),
        ("radius", c_void_p), ("radius_u", c_double),
        ("gravity", D3), ("dt", c_double), ("restitution", c_double), ("friction", c_double),
        ("n_ground_hits", c_void_p), ("n_pair_hits", c_void_p),
        ("stream", c_void_p),
    ]


class MultiSphereArgs(Structure):
    """struct rbs_multi_sphere_args"""
    _fields_ = [
        ("dtype", c_int), ("substeps", c_int), ("n_body", c_int), ("inertia_mode", c_int),
        ("arith", c_int), ("list_skin_percent", c_int),
        ("n_env", c_long), ("stride", c_long),
        ("state", c_void_p),
        ("mass", c_void_p), ("mass_u", c_double),
        ("inertia", c_void_p), ("inertia_u", D3),
        ("radius", c_void_p), ("radius_u", c_double),
        ("plane_point", D3), ("plane_normal", D3), ("gravity", D3),
        ("dt", c_double), ("restitution", c_double), ("friction", c_double),
        ("n_contacts", c_void_p), ("n_impulses", c_void_p),
        ("stream", c_void_p),
    ]


class MultiBodyArgs(Structure):
    """struct rbs_multi_body_args"""
    _fields_ = [
        ("dtype", c_int), ("substeps", c_int), ("n_body", c_int), ("has_offset", c_int),
        ("n_env", c_long), ("stride", c_long),
        ("state", c_void_p), ("body_table", c_void_p),
        ("plane_point", D3), ("plane_normal", D3), ("gravity", D3),
        ("dt", c_double), ("restitution", c_double), ("friction", c_double),
        ("n_contacts", c_void_p), ("n_impulses", c_void_p),
        ("stream", c_void_p),
    ]


RBS_BODY_TABLE_WIDTH = 16

# name -> (restype, argtypes); this table is also what tests/test_cabi.py checks against the header
PROTOTYPES = {
    "rbs_version": (c_int, []),
    "rbs_last_error": (c_char_p, []),
    "rbs_launch_count": (c_ulonglong, []),
    "rbs_device_count": (c_int, []),
    "rbs_set_option": (c_int, [c_char_p, c_long]),
    "rbs_get_option": (c_long, [c_char_p]),
    "rbs_impulse_friction": (c_int, [c_int, c_long, c_void_p, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_double, c_void_p, c_double, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rbs_apply_impulse_friction": (c_int, [c_int, c_long, c_void_p, c_void_p, c_void_p, c_double, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rbs_apply_impulse": (c_int, [c_int, c_long, c_void_p, c_void_p, c_void_p, c_double, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_double, c_void_p, c_void_p, c_void_p]),
    "rbs_inertia_world": (c_int, [c_int, c_long, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rbs_two_ball_impulse": (c_int, [c_int, c_long, c_void_p, c_double, c_void_p, c_double, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_double, c_void_p, c_double, c_void_p, c_void_p]),
    "rbs_step_body_plane": (c_int, [POINTER(BodyPlaneArgs)]),
    "rbs_step_two_ball": (c_int, [POINTER(TwoBallArgs)]),
    "rbs_step_multi_sphere": (c_int, [POINTER(MultiSphereArgs)]),
    "rbs_step_multi_body": (c_int, [POINTER(MultiBodyArgs)]),
    "rbs_pack_state": (c_int, [c_int, c_long, c_int, c_int, c_void_p, c_void_p, c_void_p, c_long, c_void_p]),
    "rbs_unpack_state": (c_int, [c_int, c_long, c_int, c_int, c_void_p, c_long, c_void_p, c_void_p, c_void_p]),
    "rbs_reset_envs": (c_int, [c_int, c_long, c_int, c_int, c_void_p, c_long, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rbs_run_body_plane_host": (c_int, [POINTER(BodyPlaneArgs), c_void_p, c_void_p, c_long]),
    "rbs_run_two_ball_host": (c_int, [POINTER(TwoBallArgs), c_void_p, c_void_p, c_long]),
    "rbs_run_multi_sphere_host": (c_int, [POINTER(MultiSphereArgs), c_void_p, c_void_p, c_long]),
    "rbs_run_multi_body_host": (c_int, [POINTER(MultiBodyArgs), c_void_p, c_void_p, c_long]),
    "rbs_release_workspace": (c_int, []),
    "rbs_stats": (c_int, [c_int, c_long, c_void_p, c_long, c_void_p, c_double, c_void_p, c_long, POINTER(c_double),
                          POINTER(c_double), c_void_p, c_void_p, c_void_p, c_void_p]),
    "rbs_fma_probe": (c_int, [c_int, c_long, c_int, c_void_p, c_void_p]),
}

_lib = None


class RbsError(RuntimeError):
    pass


def load():
    """Load the CUDA library (once).  Raises if it has not been built: no fallback exists."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RbsError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(run `python rigidbody-simulation_b200/csrc/build.py`). There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in PROTOTYPES.items():
            fn = getattr(lib, name)          # AttributeError here = header/library mismatch
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.rbs_version() != 1:
            raise RbsError(f"ABI version mismatch: library reports {lib.rbs_version()}, binding expects 1")
        _lib = lib
    return _lib


def check(rc):
    """Translate a C status into the exceptions the reference-side callers expect."""
    if rc == RBS_OK:
        return
    msg = load().rbs_last_error().decode("utf-8", "replace")
    if rc == RBS_EINVAL:
        raise ValueError(msg)
    if rc == RBS_ENOMEM:
        raise MemoryError(msg)
    raise RbsError(msg)


def launch_count():
    return int(load().rbs_launch_count())


def set_option(name, value):
    """Tuning knob of the launch dispatch (include/rbsim_b200.h: rbs_set_option); returns the previous value."""
    lib = load()
    old = int(lib.rbs_get_option(name.encode()))
    check(lib.rbs_set_option(name.encode(), int(value)))
    return old


def get_option(name):
    v = int(load().rbs_get_option(name.encode()))
    if v == -(1 << 63):
        raise ValueError(f"unknown option {name!r}")
    return v
