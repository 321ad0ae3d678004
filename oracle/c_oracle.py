"""ctypes front-end of the C oracle (oracle/rb_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may import this.
All arrays are NumPy, in the REFERENCE's layout (qpos[E,7] xyz+wxyz, qvel[E,6]); the step
functions update qpos/qvel in place and accumulate the event counters.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "librb_oracle.so")
        if not os.path.isfile(path):
            import importlib.util
            spec = importlib.util.spec_from_file_location("_rbo_build", os.path.join(_HERE, "build_oracle.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build()
        _LIB = ctypes.CDLL(path)
        _LIB.rbo_max_threads.restype = ctypes.c_int
    return _LIB


def max_threads():
    return int(lib().rbo_max_threads())


def set_threads(n):
    lib().rbo_set_threads(ctypes.c_int(int(n)))


def _suffix(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "_f64", ctypes.c_double
    if dtype == np.float32:
        return "_f32", ctypes.c_float
    raise TypeError(dtype)


def _arr(x, dtype, shape=None):
    a = np.ascontiguousarray(x, dtype=dtype)
    if shape is not None:
        a = np.ascontiguousarray(np.broadcast_to(a, shape))
    return a


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def impulse_friction(mass, vel, omega, r, normal, restitution, friction, dtype=np.float64):
    """A1 batched: returns (jn[n], jt[n,3])."""
    suf, _ = _suffix(dtype)
    vel = _arr(vel, dtype).reshape(-1, 3)
    n = vel.shape[0]
    args = [_arr(mass, dtype, (n,)), vel, _arr(omega, dtype, (n, 3)), _arr(r, dtype, (n, 3)),
            _arr(normal, dtype, (n, 3)), _arr(restitution, dtype, (n,)), _arr(friction, dtype, (n,))]
    jn = np.zeros(n, dtype)
    jt = np.zeros((n, 3), dtype)
    getattr(lib(), "rbo_impulse_friction" + suf)(ctypes.c_long(n), *[_p(a) for a in args], _p(jn), _p(jt))
    return jn, jt


def apply_impulse_friction(vel, omega, mass, inertia_world, r, normal, jn, jt, dtype=np.float64):
    """A2 batched: returns (vel'[n,3], omega'[n,3])."""
    suf, _ = _suffix(dtype)
    vel = _arr(vel, dtype).reshape(-1, 3)
    n = vel.shape[0]
    vo, wo = np.zeros((n, 3), dtype), np.zeros((n, 3), dtype)
    getattr(lib(), "rbo_apply_impulse_friction" + suf)(
        ctypes.c_long(n), _p(vel), _p(_arr(omega, dtype, (n, 3))), _p(_arr(mass, dtype, (n,))),
        _p(_arr(inertia_world, dtype, (n, 3, 3))), _p(_arr(r, dtype, (n, 3))), _p(_arr(normal, dtype, (n, 3))),
        _p(_arr(jn, dtype, (n,))), _p(_arr(jt, dtype, (n, 3))), _p(vo), _p(wo))
    return vo, wo


def apply_impulse(vel, omega, mass, inertia_world, r, normal, impulse, dtype=np.float64):
    """A3 batched."""
    suf, _ = _suffix(dtype)
    vel = _arr(vel, dtype).reshape(-1, 3)
    n = vel.shape[0]
    vo, wo = np.zeros((n, 3), dtype), np.zeros((n, 3), dtype)
    getattr(lib(), "rbo_apply_impulse" + suf)(
        ctypes.c_long(n), _p(vel), _p(_arr(omega, dtype, (n, 3))), _p(_arr(mass, dtype, (n,))),
        _p(_arr(inertia_world, dtype, (n, 3, 3))), _p(_arr(r, dtype, (n, 3))), _p(_arr(normal, dtype, (n, 3))),
        _p(_arr(impulse, dtype, (n,))), _p(vo), _p(wo))
    return vo, wo


def inertia_world(inertia_diag, quat, dtype=np.float64):
    """A4 batched: returns [n,3,3]."""
    suf, _ = _suffix(dtype)
    quat = _arr(quat, dtype).reshape(-1, 4)
    n = quat.shape[0]
    out = np.zeros((n, 3, 3), dtype)
    getattr(lib(), "rbo_inertia_world" + suf)(ctypes.c_long(n), _p(_arr(inertia_diag, dtype, (n, 3))), _p(quat), _p(out))
    return out


def two_ball_impulse(mass, inv_inertia, v, w, r, n, restitution, friction, dtype=np.float64):
    """A10 batched: returns J[n,3]."""
    suf, _ = _suffix(dtype)
    v = _arr(v, dtype).reshape(-1, 3)
    cnt = v.shape[0]
    J = np.zeros((cnt, 3), dtype)
    getattr(lib(), "rbo_two_ball_impulse" + suf)(
        ctypes.c_long(cnt), _p(_arr(mass, dtype, (cnt,))), _p(_arr(inv_inertia, dtype, (cnt,))), _p(v),
        _p(_arr(w, dtype, (cnt, 3))), _p(_arr(r, dtype, (cnt, 3))), _p(_arr(n, dtype, (cnt, 3))),
        _p(_arr(restitution, dtype, (cnt,))), _p(_arr(friction, dtype, (cnt,))), _p(J))
    return J


def step_body_plane(qpos, qvel, steps, *, geom, mass, inertia, size, plane_pos, plane_normal, gravity, dt,
                    restitution, friction, threshold, scheme="A", xfrc=None, counters=None):
    """A5/A6 (scheme 'A') or A7 (scheme 'general') for E independent single-body envs.
    qpos[E,7], qvel[E,6] are updated in place; ``counters`` = (calls[E], impulses[E]) uint32 or None."""
    dtype = qpos.dtype
    suf, creal = _suffix(dtype)
    assert qpos.flags.c_contiguous and qvel.flags.c_contiguous and qvel.dtype == dtype
    E = qpos.shape[0]
    size3 = _arr(size, dtype, (E, 3))          # scalar, (3,), (E,1) or (E,3); sphere radius = column 0
    calls, imps = counters if counters is not None else (None, None)
    getattr(lib(), "rbo_step_body_plane" + suf)(
        ctypes.c_long(E), ctypes.c_int(int(steps)), ctypes.c_int(0 if scheme == "A" else 1),
        ctypes.c_int({"sphere": 0, "box": 1}[geom]), _p(qpos), _p(qvel), _p(_arr(mass, dtype, (E,))),
        _p(_arr(inertia, dtype, (E, 3))), _p(size3), _p(_arr(plane_pos, dtype, (3,))),
        _p(_arr(plane_normal, dtype, (3,))), _p(_arr(gravity, dtype, (3,))),
        _p(None if xfrc is None else _arr(xfrc, dtype, (E, 6))), creal(dt), _p(_arr(restitution, dtype, (E,))),
        _p(_arr(friction, dtype, (E,))), creal(threshold), _p(calls), _p(imps))


def step_multi_sphere(qpos, qvel, steps, *, mass, inertia, radius, plane_pos, plane_normal, gravity, dt,
                      restitution, friction, counters=None):
    """A9 (repaired) for E envs of B spheres.  qpos[E,B,7], qvel[E,B,6] in place."""
    dtype = qpos.dtype
    suf, creal = _suffix(dtype)
    assert qpos.flags.c_contiguous and qvel.flags.c_contiguous
    E, B = qpos.shape[0], qpos.shape[1]
    calls, imps = counters if counters is not None else (None, None)
    getattr(lib(), "rbo_step_multi_sphere" + suf)(
        ctypes.c_long(E), ctypes.c_int(B), ctypes.c_int(int(steps)), _p(qpos), _p(qvel),
        _p(_arr(mass, dtype, (E, B))), _p(_arr(inertia, dtype, (E, B, 3))), _p(_arr(radius, dtype, (E, B))),
        _p(_arr(plane_pos, dtype, (3,))), _p(_arr(plane_normal, dtype, (3,))), _p(_arr(gravity, dtype, (3,))),
        creal(dt), creal(restitution), creal(friction), _p(calls), _p(imps))


def step_multi_body(qpos, qvel, steps, *, gtype, mass, inertia, size, plane_pos, plane_normal, gravity, dt, restitution,
                    friction, geom_pos=None, geom_quat=None, counters=None):
    """N4: the repaired A9 loop for E envs of B bodies, each a sphere (gtype 0, size[0] = radius) or a box (gtype 1, half
    extents); gtype[B], mass[B], inertia[B,3], size[B,3], geom_pos[B,3] / geom_quat[B,4] (or None) are shared by all
    environments.  qpos[E,B,7], qvel[E,B,6] in place."""
    dtype = qpos.dtype
    suf, creal = _suffix(dtype)
    assert qpos.flags.c_contiguous and qvel.flags.c_contiguous
    E, B = qpos.shape[0], qpos.shape[1]
    calls, imps = counters if counters is not None else (None, None)
    gt = np.ascontiguousarray(gtype, dtype=np.int32)
    assert gt.shape == (B,)
    gp = None if geom_pos is None else _arr(geom_pos, dtype, (B, 3))
    gq = None if geom_quat is None else _arr(geom_quat, dtype, (B, 4))
    if (gp is None) != (gq is None):
        gp = _arr(np.zeros((B, 3)), dtype) if gp is None else gp
        gq = _arr(np.tile([1.0, 0, 0, 0], (B, 1)), dtype) if gq is None else gq
    getattr(lib(), "rbo_step_multi_body" + suf)(
        ctypes.c_long(E), ctypes.c_int(B), ctypes.c_int(int(steps)), _p(qpos), _p(qvel), _p(gt),
        _p(_arr(mass, dtype, (B,))), _p(_arr(inertia, dtype, (B, 3))), _p(_arr(size, dtype, (B, 3))), _p(gp), _p(gq),
        _p(_arr(plane_pos, dtype, (3,))), _p(_arr(plane_normal, dtype, (3,))), _p(_arr(gravity, dtype, (3,))),
        creal(dt), creal(restitution), creal(friction), _p(calls), _p(imps))


def step_two_ball(qpos, qvel, steps, *, mass, radius, gravity, dt, restitution, friction, counters=None):
    """A11 for E two-ball envs.  qpos[E,14], qvel[E,12] in place; counters = (ground_hits[E], pair_hits[E])."""
    dtype = qpos.dtype
    suf, creal = _suffix(dtype)
    assert qpos.flags.c_contiguous and qvel.flags.c_contiguous
    E = qpos.shape[0]
    gh, ph = counters if counters is not None else (None, None)
    getattr(lib(), "rbo_step_two_ball" + suf)(
        ctypes.c_long(E), ctypes.c_int(int(steps)), _p(qpos), _p(qvel), _p(_arr(mass, dtype, (E, 2))),
        _p(_arr(radius, dtype, (E,))), _p(_arr(gravity, dtype, (3,))), creal(dt), creal(restitution),
        creal(friction), _p(gh), _p(ph))
