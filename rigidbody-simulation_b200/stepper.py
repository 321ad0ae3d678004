"""Python side of the steppers: build the C-ABI argument structs from a BatchedModel / BatchedData
pair and enqueue the CUDA kernels on the current torch stream.  No allocation, no host sync and no
fallback on the step path: if the CUDA library is absent ``_lib.load()`` raises.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._lib import (RBS_F32, RBS_F64, RBS_GEOM_BOX, RBS_GEOM_SPHERE, RBS_INERTIA_GENERAL, RBS_INERTIA_ISOTROPIC,
                   RBS_SCHEME_A, BodyPlaneArgs, MultiSphereArgs, TwoBallArgs)


def rbs_dtype(dtype):
    if dtype == torch.float64:
        return RBS_F64
    if dtype == torch.float32:
        return RBS_F32
    raise TypeError(f"unsupported dtype {dtype}")


def current_stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _per_env_scalar(value, model, n):
    """A python scalar stays uniform; a tensor / array of n values becomes a device array."""
    if torch.is_tensor(value) or (isinstance(value, np.ndarray) and value.ndim > 0):
        t = torch.as_tensor(value, dtype=model.dtype).to(model.device).contiguous()
        if t.numel() != n:
            raise ValueError(f"per-environment parameter has {t.numel()} values, expected {n}")
        return t, 0.0
    return None, float(value)


def _isotropic(inertia):
    return bool(inertia[0] == inertia[1] == inertia[2])


def _require_cuda(model, offsets_ok=False):
    if not offsets_ok and getattr(model, "has_offset_geoms", False):
        raise ValueError("this stepper takes every geom at its body's origin; scenes with offset geoms go through step_multi_body")
    if model.device.type != "cuda":
        raise _lib.RbsError("the steppers run on a CUDA device only (no CPU fallback)")


ARITH = {"strict": _lib.RBS_ARITH_STRICT, "fast": _lib.RBS_ARITH_FAST}


def _shift(ptr, elements, itemsize):
    return None if ptr is None or ptr.value is None else ctypes.c_void_p(ptr.value + elements * itemsize)


def body_plane_args(model, data, body_id, dt, restitution, friction_coeff, contact_threshold, scheme, substeps,
                    count=True, strict_inertia=None, arith="strict", env_range=None):
    """rbs_body_plane_args for the single free body ``body_id`` of ``model`` resting on its plane.
    ``env_range=(start, count)`` restricts the call to a window of the environments."""
    _require_cuda(model)
    if model.nfree != 1 or data.layout != "env":
        raise ValueError("this step function handles scenes with exactly one free body (sphere.xml / cube.xml shape)")
    if model.plane_normal is None:
        raise ValueError("scene has no plane geom")
    bid = model.free_ids[0]
    # the reference indexes body_mass[body_id] with body_id == -1 when the name is absent, i.e. the LAST body
    # (src/physics/collision.py:58-60 with src/simulation/single_sphere_bounce.py:67)
    if body_id not in (bid, bid - model.nbody, -1):
        raise ValueError(f"body id {body_id} is not the free body of this scene")
    geom = model.body_geom[bid]
    if geom.type not in ("sphere", "box"):
        raise ValueError(f"free-body geom type {geom.type!r} has no plane contact routine")
    a = BodyPlaneArgs()
    a.dtype = rbs_dtype(model.dtype)
    a.geom = RBS_GEOM_SPHERE if geom.type == "sphere" else RBS_GEOM_BOX
    a.scheme = scheme
    a.arith = ARITH[arith]
    a.n_env, a.stride, a.substeps = data.nenv, data.stride, int(substeps)
    a.state = _ptr(data.state)
    pe = model.per_env
    keep = []                                   # tensors that must outlive the launch
    mass_t = pe.get("mass")
    a.mass, a.mass_u = _ptr(mass_t), float(model.body_mass[bid])
    inertia_t = pe.get("inertia")
    a.inertia = _ptr(inertia_t)
    a.inertia_u = _lib.D3(*[float(v) for v in model.body_inertia[bid]])
    size_t = pe.get("size")
    a.size = _ptr(size_t)
    a.size_u = _lib.D3(*[float(v) for v in geom.size])
    # strict policy: literal inv(R diag(I) R^T) by default (bit-for-bit the oracle); strict_inertia=False opts into the
    # isotropic shortcut (1/I)*Id, which is a few ulp away.  The fast policy always uses the shortcut.
    if strict_inertia is None:
        strict_inertia = arith == "strict"
    iso = inertia_t is None and _isotropic(model.body_inertia[bid]) and not strict_inertia
    a.inertia_mode = RBS_INERTIA_ISOTROPIC if iso else RBS_INERTIA_GENERAL
    # None = "use the per-environment values attached to the model" (randomised configs)
    rt, ru = _per_env_scalar(pe["restitution"] if restitution is None else restitution, model, data.nenv)
    ft, fu = _per_env_scalar(pe["friction"] if friction_coeff is None else friction_coeff, model, data.nenv)
    keep += [rt, ft]
    a.restitution, a.restitution_u = _ptr(rt), ru
    a.friction, a.friction_u = _ptr(ft), fu
    a.xfrc = _ptr(data.xfrc_applied)
    a.plane_point = _lib.D3(*model.plane_point)
    a.plane_normal = _lib.D3(*model.plane_normal)
    a.gravity = _lib.D3(*[float(g) for g in model.opt.gravity])
    a.dt = float(dt)
    a.contact_threshold = float(contact_threshold)
    a.n_contacts = _ptr(data.n_contacts) if count else None
    a.n_impulses = _ptr(data.n_impulses) if count else None
    if env_range is not None:
        off, cnt = int(env_range[0]), int(env_range[1])
        if off < 0 or cnt < 0 or off + cnt > data.nenv:
            raise ValueError(f"env_range {env_range} outside 0..{data.nenv}")
        es = 8 if model.dtype == torch.float64 else 4
        a.n_env, a.param_stride = cnt, data.nenv
        for name in ("state", "mass", "inertia", "size", "restitution", "friction", "xfrc"):
            setattr(a, name, _shift(ctypes.c_void_p(getattr(a, name)), off, es))
        for name in ("n_contacts", "n_impulses"):
            setattr(a, name, _shift(ctypes.c_void_p(getattr(a, name)), off, 4))
    a._keep = keep
    return a


def _key_of(value):
    """cache key of a scalar-or-per-env argument.  Only DEVICE tensors are cacheable (by storage address; the cached
    struct keeps them alive): a host tensor / array is uploaded on every call, so an in-place edit of it is seen."""
    if torch.is_tensor(value):
        return ("t", value.data_ptr(), value.numel()) if value.device.type == "cuda" else None
    if isinstance(value, np.ndarray) and value.ndim > 0:
        return None
    return value


def _refresh_by_value(a, model, dt, contact_threshold, strict_inertia, arith):
    """The by-value fields of a cached struct are rewritten on every call (they are cheap), so MuJoCo-style edits of
    the model between two steps -- ``model.opt.gravity[:] = ...``, ``model.body_mass[i] = ...`` -- take effect."""
    bid = model.free_ids[0]
    a.mass_u = float(model.body_mass[bid])
    a.inertia_u = _lib.D3(*[float(v) for v in model.body_inertia[bid]])
    a.size_u = _lib.D3(*[float(v) for v in model.body_geom[bid].size])
    a.plane_point = _lib.D3(*model.plane_point)
    a.plane_normal = _lib.D3(*model.plane_normal)
    a.gravity = _lib.D3(*[float(g) for g in model.opt.gravity])
    a.dt = float(dt)
    a.contact_threshold = float(contact_threshold)
    if strict_inertia is None:
        strict_inertia = arith == "strict"
    iso = model.per_env.get("inertia") is None and _isotropic(model.body_inertia[bid]) and not strict_inertia
    a.inertia_mode = RBS_INERTIA_ISOTROPIC if iso else RBS_INERTIA_GENERAL


def step_body_plane(model, data, body_id, dt, restitution, friction_coeff, contact_threshold,
                    scheme=RBS_SCHEME_A, substeps=1, count=True, strict_inertia=None, arith="strict", env_range=None,
                    trajectory=None):
    """arith: "strict" reproduces the reference's rounding sequence; "fast" re-associates for the FP pipe
    (scheme A + isotropic inertia only; <= 1e-12 relative per step in fp64).

    The marshalled argument struct is cached per call signature and per identity of every device buffer it points to,
    so a per-frame loop pays for one ctypes call per step, not for rebuilding 30 fields.  Only pointers are trusted
    from the cache: every by-value field is refreshed from the model on each call (_refresh_by_value)."""
    kr, kf = _key_of(restitution), _key_of(friction_coeff)
    a = None
    if not (kr is None and restitution is not None) and not (kf is None and friction_coeff is not None):
        pe = model.per_env
        key = (body_id, kr, kf, scheme, count, arith, env_range, data.state.data_ptr(),
               None if data.xfrc_applied is None else data.xfrc_applied.data_ptr(),
               tuple((k, t.data_ptr()) for k, t in pe.items()))
        cache = data.__dict__.setdefault("_args_cache", {})
        a = cache.get(key)
        if a is None:
            if len(cache) > 64:
                cache.clear()
            a = cache[key] = body_plane_args(model, data, body_id, dt, restitution, friction_coeff, contact_threshold,
                                             scheme, substeps, count, strict_inertia, arith, env_range)
        else:
            _refresh_by_value(a, model, dt, contact_threshold, strict_inertia, arith)
            # python scalars for restitution / friction are by-value fields too (a tensor key pins the pointer)
            if a.restitution is None and restitution is not None:
                a.restitution_u = float(restitution)
            if a.friction is None and friction_coeff is not None:
                a.friction_u = float(friction_coeff)
        a.substeps = int(substeps)
    if a is None:
        a = body_plane_args(model, data, body_id, dt, restitution, friction_coeff, contact_threshold, scheme, substeps,
                            count, strict_inertia, arith, env_range)
    a.stream = current_stream(model.device)
    # trajectory: device tensor [substeps, n, 3] that receives the position of the first n environments after every
    # substep of this launch (the per-frame logger.record of the reference, without leaving the fused launch)
    if trajectory is not None:
        if (not torch.is_tensor(trajectory) or trajectory.device != data.state.device or trajectory.dtype != model.dtype
                or trajectory.dim() != 3 or trajectory.shape[0] != int(substeps) or trajectory.shape[2] != 3
                or not trajectory.is_contiguous()):
            raise ValueError("trajectory must be a contiguous device tensor [substeps, n_sample, 3] of the model's dtype")
        if trajectory.shape[1] > a.n_env:
            raise ValueError(f"trajectory samples {trajectory.shape[1]} environments, the launch steps {a.n_env}")
        a.trajectory, a.trajectory_envs = _ptr(trajectory), int(trajectory.shape[1])
    else:
        a.trajectory, a.trajectory_envs = None, 0
    _lib.check(_lib.load().rbs_step_body_plane(ctypes.byref(a)))


class SplitChains:
    """Run the fused launches of a long horizon as ``parts`` independent chains of environment windows, each on its
    own stream.  Environments never interact, so the chains need no ordering between them; their launches interleave
    on the device and the ragged last wave of CTAs of one chain's launch is filled by the other chain's work.

        chains = SplitChains(model, data); chains.fork()
        for _ in range(n): chains.step(dt=..., substeps=128, ...)      # same keywords as step_body_plane
        chains.join()
    """

    def __init__(self, model, data, parts=2, align=128):
        self.model, self.data = model, data
        per = -(-data.nenv // parts)
        per = -(-per // align) * align
        self.ranges = [(o, min(per, data.nenv - o)) for o in range(0, data.nenv, per)]
        self.streams = [torch.cuda.Stream(device=model.device) for _ in self.ranges]

    def fork(self):
        cur = torch.cuda.current_stream(self.model.device)
        for st in self.streams:
            st.wait_stream(cur)

    def step(self, body_id=-1, **kw):
        for rng, st in zip(self.ranges, self.streams):
            with torch.cuda.stream(st):
                step_body_plane(self.model, self.data, body_id, env_range=rng, **kw)

    def join(self):
        cur = torch.cuda.current_stream(self.model.device)
        for st in self.streams:
            cur.wait_stream(st)


def two_ball_args(model, data, dt, restitution, friction, radius, substeps, count=True, arith="strict"):
    _require_cuda(model)
    if model.nfree != 2 or data.layout != "env":
        raise ValueError("the two-ball step needs a scene with exactly two free bodies (ball_collision.xml shape)")
    a = TwoBallArgs()
    a.dtype, a.substeps, a.arith = rbs_dtype(model.dtype), int(substeps), ARITH[arith]
    a.n_env, a.stride = data.nenv, data.stride
    a.state = _ptr(data.state)
    pe = model.per_env
    a.mass = _ptr(pe.get("mass"))                       # [2, E] when per-env
    a.mass_u = _lib.D2(*[float(model.body_mass[i]) for i in model.free_ids])
    a.radius, a.radius_u = _ptr(pe.get("radius")), float(radius)
    a.gravity = _lib.D3(*[float(g) for g in model.opt.gravity])
    a.dt, a.restitution, a.friction = float(dt), float(restitution), float(friction)
    a.n_ground_hits = _ptr(data.n_contacts) if count else None
    a.n_pair_hits = _ptr(data.n_impulses) if count else None
    return a


def step_two_ball(model, data, dt, restitution, friction, radius=0.1, substeps=1, count=True, arith="strict"):
    a = two_ball_args(model, data, dt, restitution, friction, radius, substeps, count, arith)
    a.stream = current_stream(model.device)
    _lib.check(_lib.load().rbs_step_two_ball(ctypes.byref(a)))


def multi_sphere_args(model, data, dt, restitution, friction, substeps, count=True, strict_inertia=None, arith="strict",
                      list_skin_percent=0):
    _require_cuda(model)
    if data.layout != "body":
        raise ValueError("the multi-sphere step needs BatchedData(model, layout='body')")
    if model.plane_normal is None:
        raise ValueError("scene has no plane geom")
    geoms = [model.body_geom[i] for i in model.free_ids]
    if any(g.type != "sphere" for g in geoms):
        raise ValueError("the multi-sphere step handles sphere bodies only")
    pe = model.per_env
    first = model.free_ids[0]
    uniform = all(model.body_mass[i] == model.body_mass[first] and geoms[k].size[0] == geoms[0].size[0]
                  for k, i in enumerate(model.free_ids))
    if not uniform and not all(k in pe for k in ("mass", "radius", "inertia")):
        # heterogeneous spheres: expand the per-body values once
        E, B = data.nenv, data.nfree
        m = torch.tensor([model.body_mass[i] for i in model.free_ids], dtype=model.dtype, device=model.device)
        r = torch.tensor([g.size[0] for g in geoms], dtype=model.dtype, device=model.device)
        I = torch.tensor([model.body_inertia[i] for i in model.free_ids], dtype=model.dtype, device=model.device)
        pe.setdefault("mass", m.repeat(E).contiguous())
        pe.setdefault("radius", r.repeat(E).contiguous())
        pe.setdefault("inertia", I.t().repeat(1, E).contiguous())
    a = MultiSphereArgs()
    a.dtype, a.substeps, a.n_body = rbs_dtype(model.dtype), int(substeps), data.nfree
    a.list_skin_percent = int(list_skin_percent)
    a.n_env, a.stride = data.nenv, data.stride
    a.state = _ptr(data.state)
    a.mass, a.mass_u = _ptr(pe.get("mass")), float(model.body_mass[first])
    a.inertia = _ptr(pe.get("inertia"))
    a.inertia_u = _lib.D3(*[float(v) for v in model.body_inertia[first]])
    a.radius, a.radius_u = _ptr(pe.get("radius")), float(geoms[0].size[0])
    if strict_inertia is None:
        strict_inertia = arith == "strict"
    a.inertia_mode = RBS_INERTIA_GENERAL if strict_inertia else RBS_INERTIA_ISOTROPIC   # spheres: I1 = I2 = I3
    a.arith = ARITH[arith]
    a.plane_point = _lib.D3(*model.plane_point)
    a.plane_normal = _lib.D3(*model.plane_normal)
    a.gravity = _lib.D3(*[float(g) for g in model.opt.gravity])
    a.dt, a.restitution, a.friction = float(dt), float(restitution), float(friction)
    a.n_contacts = _ptr(data.n_contacts) if count else None
    a.n_impulses = _ptr(data.n_impulses) if count else None
    return a


def step_multi_sphere(model, data, dt, restitution, friction, substeps=1, count=True, strict_inertia=None, arith="strict",
                      list_skin_percent=0):
    """``list_skin_percent``: partner-list skin in % of the radius (0 = library default, < 0 = scan all partners every
    substep); a tuning knob only, results do not depend on it."""
    a = multi_sphere_args(model, data, dt, restitution, friction, substeps, count, strict_inertia, arith, list_skin_percent)
    a.stream = current_stream(model.device)
    _lib.check(_lib.load().rbs_step_multi_sphere(ctypes.byref(a)))


# --------------------------------------------------------------------------------------------------
# N4: multi-body scenes with spheres and boxes (new behaviour, DESIGN.md "N4")
# --------------------------------------------------------------------------------------------------
def body_table(model):
    """[nfree, RBS_BODY_TABLE_WIDTH] float64 host table of the scene's free bodies (include/rbsim_b200.h)."""
    rows = []
    for i in model.free_ids:
        g = model.body_geom[i]
        if g.type not in ("sphere", "box"):
            raise ValueError(f"geom type {g.type!r} cannot be a free body's geom")
        size = [float(g.size[0]), 0.0, 0.0] if g.type == "sphere" else [float(v) for v in g.size[:3]]
        bound = size[0] if g.type == "sphere" else float(np.sqrt(sum(v * v for v in size)))
        rows.append([0.0 if g.type == "sphere" else 1.0] + size + [float(model.body_mass[i])]
                    + [float(v) for v in model.body_inertia[i]] + [float(v) for v in g.pos] + [float(v) for v in g.quat]
                    + [bound * (1.0 + 1e-6)])
    return np.asarray(rows, dtype=np.float64)


def multi_body_args(model, data, dt, restitution, friction, substeps, count=True):
    _require_cuda(model, offsets_ok=True)
    if data.layout != "body":
        raise ValueError("the multi-body step needs BatchedData(model, layout='body')")
    if model.plane_normal is None:
        raise ValueError("scene has no plane geom")
    if data.nfree > 256:
        raise ValueError("the multi-body step handles at most 256 bodies per environment")
    # The table is rebuilt (and re-uploaded) only when the model's mass / inertia arrays have changed since the last call:
    # model.body_mass[...] = ... between steps takes effect, a per-frame loop does not pay O(bodies) of Python per step.
    cached = getattr(model, "_body_table_cache", None)
    if cached is None or not (np.array_equal(cached[0], model.body_mass) and np.array_equal(cached[1], model.body_inertia)):
        host = body_table(model)
        cached = (model.body_mass.copy(), model.body_inertia.copy(),
                  torch.as_tensor(host, dtype=model.dtype).to(model.device).contiguous())
        model._body_table_cache = cached
    cache = cached[2]
    a = _lib.MultiBodyArgs()
    a.dtype, a.substeps, a.n_body = rbs_dtype(model.dtype), int(substeps), data.nfree
    a.has_offset = int(model.has_offset_geoms)
    a.n_env, a.stride = data.nenv, data.stride
    a.state = _ptr(data.state)
    a.body_table = _ptr(cache)
    a.plane_point = _lib.D3(*model.plane_point)
    a.plane_normal = _lib.D3(*model.plane_normal)
    a.gravity = _lib.D3(*[float(g) for g in model.opt.gravity])
    a.dt, a.restitution, a.friction = float(dt), float(restitution), float(friction)
    a.n_contacts = _ptr(data.n_contacts) if count else None
    a.n_impulses = _ptr(data.n_impulses) if count else None
    return a


def step_multi_body(model, data, dt, restitution, friction, substeps=1, count=True):
    """The repaired ``custom_step_multi_sphere`` loop (multi_sphere_bounce.py:42-92) for scenes whose free bodies are
    spheres AND boxes, geoms optionally offset in their body's frame; strict arithmetic, literal world inertia."""
    a = multi_body_args(model, data, dt, restitution, friction, substeps, count)
    a.stream = current_stream(model.device)
    _lib.check(_lib.load().rbs_step_multi_body(ctypes.byref(a)))


# --------------------------------------------------------------------------------------------------
# host-buffer drivers (reference layout in, reference layout out)
# --------------------------------------------------------------------------------------------------
def _host_ptr(arr, dtype, shape):
    if torch.is_tensor(arr):
        if arr.device.type != "cpu" or not arr.is_contiguous() or arr.dtype != dtype or tuple(arr.shape) != shape:
            raise ValueError(f"host buffer must be a contiguous CPU tensor of {dtype} with shape {shape}")
        return ctypes.c_void_p(arr.data_ptr())
    np_dtype = np.float64 if dtype == torch.float64 else np.float32
    if not (isinstance(arr, np.ndarray) and arr.flags.c_contiguous and arr.dtype == np_dtype and arr.shape == shape):
        raise ValueError(f"host buffer must be a C-contiguous {np_dtype.__name__} array with shape {shape}")
    return ctypes.c_void_p(arr.ctypes.data)


def run_body_plane_host(model, qpos, qvel, total_steps, body_id=-1, dt=None, restitution=1.0, friction_coeff=1.0,
                        contact_threshold=0.0, scheme=RBS_SCHEME_A, substeps=32, strict_inertia=None, arith="strict"):
    """Advance host arrays qpos[E,7], qvel[E,6] (in place) by ``total_steps`` steps of A5/A6/A7.
    H2D, the launches and D2H all happen inside; returns after the stream is synchronised."""
    data = _HostShim(model, 1)
    a = body_plane_args(model, data, body_id, model.opt.timestep if dt is None else dt, restitution, friction_coeff,
                        contact_threshold, scheme, substeps, count=False, strict_inertia=strict_inertia, arith=arith)
    a.stream = current_stream(model.device)
    E = model.nenv
    _lib.check(_lib.load().rbs_run_body_plane_host(ctypes.byref(a), _host_ptr(qpos, model.dtype, (E, 7)),
                                                   _host_ptr(qvel, model.dtype, (E, 6)), int(total_steps)))


def run_two_ball_host(model, qpos, qvel, total_steps, dt=None, restitution=1.0, friction=0.3, radius=0.1, substeps=32,
                      arith="strict"):
    data = _HostShim(model, 2)
    a = two_ball_args(model, data, model.opt.timestep if dt is None else dt, restitution, friction, radius, substeps,
                      count=False, arith=arith)
    a.stream = current_stream(model.device)
    E = model.nenv
    _lib.check(_lib.load().rbs_run_two_ball_host(ctypes.byref(a), _host_ptr(qpos, model.dtype, (E, 14)),
                                                 _host_ptr(qvel, model.dtype, (E, 12)), int(total_steps)))


def run_multi_sphere_host(model, qpos, qvel, total_steps, dt=None, restitution=1.0, friction=0.0, substeps=8,
                          arith="strict"):
    data = _HostShim(model, model.nfree, layout="body")
    a = multi_sphere_args(model, data, model.opt.timestep if dt is None else dt, restitution, friction, substeps,
                          count=False, arith=arith)
    a.stream = current_stream(model.device)
    E, B = model.nenv, model.nfree
    _lib.check(_lib.load().rbs_run_multi_sphere_host(ctypes.byref(a), _host_ptr(qpos, model.dtype, (E, 7 * B)),
                                                     _host_ptr(qvel, model.dtype, (E, 6 * B)), int(total_steps)))


def run_multi_body_host(model, qpos, qvel, total_steps, dt=None, restitution=0.2, friction=0.6, substeps=8):
    """Advance host arrays qpos[E, 7B], qvel[E, 6B] (in place) by ``total_steps`` steps of the multi-body stepper (N4)."""
    data = _HostShim(model, model.nfree, layout="body")
    a = multi_body_args(model, data, model.opt.timestep if dt is None else dt, restitution, friction, substeps, count=False)
    a.stream = current_stream(model.device)
    E, B = model.nenv, model.nfree
    _lib.check(_lib.load().rbs_run_multi_body_host(ctypes.byref(a), _host_ptr(qpos, model.dtype, (E, 7 * B)),
                                                   _host_ptr(qvel, model.dtype, (E, 6 * B)), int(total_steps)))


class _HostShim:
    """Carries the sizes the *_args builders read from a BatchedData; the host drivers use the library's own
    device workspace, so no state tensor exists on the Python side."""

    def __init__(self, model, nfree, layout="env"):
        self.nenv, self.nfree, self.layout = model.nenv, nfree, layout
        self.stride = model.nenv if layout == "env" else model.nenv * nfree
        self.state = None
        self.xfrc_applied = None
        self.n_contacts = self.n_impulses = None


def fma_peak(device, dtype=torch.float64, iters=4096, blocks_per_sm=8):
    """Measured FMA throughput (FLOP/s) of the CUDA cores of ``device`` for ``dtype`` (bench helper)."""
    lib = _lib.load()
    props = torch.cuda.get_device_properties(device)
    n_threads = props.multi_processor_count * blocks_per_sm * 256
    sink = torch.empty(n_threads, dtype=dtype, device=device)
    stream = current_stream(device)
    code = rbs_dtype(dtype)
    for _ in range(2):
        _lib.check(lib.rbs_fma_probe(code, n_threads, iters, _ptr(sink), stream))
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 0.0
    for _ in range(5):
        t0.record()
        _lib.check(lib.rbs_fma_probe(code, n_threads, iters, _ptr(sink), stream))
        t1.record()
        t1.synchronize()
        chains = 16 if _lib.get_option("probe_mode") == 2 else 8
        best = max(best, 2.0 * chains * n_threads * iters / (t0.elapsed_time(t1) * 1e-3))
    return best
