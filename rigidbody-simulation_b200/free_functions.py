"""Batched CUDA versions of the reference's free functions.

Each function accepts what the reference accepts -- Python floats and NumPy ``(3,)`` / ``(3,3)`` arrays --
and returns the same types and shapes for that single-item case (the work still runs as a one-thread CUDA
launch: there is no CPU path).  With a leading batch dimension, or with torch tensors, the inputs are
treated as N work items and device tensors come back.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .stepper import current_stream, rbs_dtype


class _Batch:
    """Normalises one call's arguments to contiguous device arrays of a common length and dtype."""

    def __init__(self, *probe, device=None, dtype=None):
        tensors = [p for p in probe if torch.is_tensor(p)]
        self.as_torch = bool(tensors)
        if dtype is None:
            dtype = torch.float32 if tensors and all(t.dtype == torch.float32 for t in tensors if t.is_floating_point()) else torch.float64
        self.dtype = dtype
        if device is None:
            cuda = [t.device for t in tensors if t.device.type == "cuda"]
            device = cuda[0] if cuda else torch.device("cuda")
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise _lib.RbsError("free functions run on a CUDA device only (no CPU fallback)")
        self.n = None
        self.batched = False
        self.keep = []

    def vec(self, x, width):
        """[n, *width] array argument (or a single item of shape ``width``)."""
        t = torch.as_tensor(np.asarray(x, dtype=np.float64) if not torch.is_tensor(x) else x)
        t = t.to(device=self.device, dtype=self.dtype)
        nd = len(width)
        if t.dim() == nd:
            t = t.unsqueeze(0)
        elif t.dim() == nd + 1:
            self.batched = True
        else:
            raise ValueError(f"expected shape {width} or (N, {', '.join(map(str, width))}), got {tuple(t.shape)}")
        if tuple(t.shape[1:]) != tuple(width):
            raise ValueError(f"expected trailing shape {width}, got {tuple(t.shape[1:])}")
        t = t.contiguous()
        self._size(t.shape[0])
        self.keep.append(t)
        return t

    def scalar(self, x):
        """python scalar or one-element array -> (NULL, value); array of n -> (device pointer, 0.0).  A per-item array
        must have exactly as many entries as the batch (the kernels index it with the work-item index)."""
        if torch.is_tensor(x) and x.dim() > 0 or isinstance(x, np.ndarray) and x.ndim > 0:
            t = torch.as_tensor(x).to(device=self.device, dtype=self.dtype).contiguous().reshape(-1)
            if t.numel() == 1:                       # uniform value in array clothing: never hand out a 1-element pointer
                return None, float(t.item())
            self.batched = True
            self._size(t.shape[0])
            self.keep.append(t)
            return ctypes.c_void_p(t.data_ptr()), 0.0
        return None, float(x)

    def _size(self, n):
        if n == 1 and self.n is not None:
            return
        if self.n is None or self.n == 1:
            self.n = n
        elif n != self.n:
            raise ValueError(f"batch sizes disagree: {self.n} vs {n}")

    def expand(self, t):
        if t.shape[0] != self.n:
            t = t.expand(self.n, *t.shape[1:]).contiguous()
            self.keep.append(t)
        return ctypes.c_void_p(t.data_ptr())

    def empty(self, *width):
        t = torch.empty((self.n, *width), dtype=self.dtype, device=self.device)
        return t, ctypes.c_void_p(t.data_ptr())

    def out(self, t, scalar=False):
        if self.as_torch or self.batched:
            return t
        a = t[0].cpu().numpy()
        return float(a) if scalar else a


def compute_collision_impulse_friction(mass, inertia_world, vel, omega, contact_point, normal, restitution, friction_coeff):
    """src/physics/collision.py:7-48.  Returns ``(jn, jt)``; ``inertia_world`` is unused, as in the reference
    (k = 1/m + 1/18)."""
    b = _Batch(vel, omega, contact_point, normal, mass)
    v, w, r, n = b.vec(vel, (3,)), b.vec(omega, (3,)), b.vec(contact_point, (3,)), b.vec(normal, (3,))
    (mp, mu_), (ep, eu), (fp, fu) = b.scalar(mass), b.scalar(restitution), b.scalar(friction_coeff)
    jn, jn_p = b.empty()
    jt, jt_p = b.empty(3)
    _lib.check(_lib.load().rbs_impulse_friction(rbs_dtype(b.dtype), b.n, mp, mu_, b.expand(v), b.expand(w), b.expand(r),
                                               b.expand(n), ep, eu, fp, fu, jn_p, jt_p, None, current_stream(b.device)))
    return b.out(jn, scalar=True), b.out(jt)


def apply_impulse_friction(vel, omega, mass, inertia_world, contact_point, normal, jn, jt):
    """src/physics/physics_utils.py:25-49.  Returns NEW ``(vel, omega)``."""
    b = _Batch(vel, omega, inertia_world, contact_point, normal, jt)
    v, w, Iw = b.vec(vel, (3,)), b.vec(omega, (3,)), b.vec(inertia_world, (3, 3))
    r, n, jtv = b.vec(contact_point, (3,)), b.vec(normal, (3,)), b.vec(jt, (3,))
    mp, mu_ = b.scalar(mass)
    jnv = torch.as_tensor(np.asarray(jn, dtype=np.float64) if not torch.is_tensor(jn) else jn).to(device=b.device, dtype=b.dtype).reshape(-1)
    if jnv.numel() > 1:
        b.batched = True
    b._size(jnv.shape[0])
    vo, vo_p = b.empty(3)
    wo, wo_p = b.empty(3)
    _lib.check(_lib.load().rbs_apply_impulse_friction(rbs_dtype(b.dtype), b.n, b.expand(v), b.expand(w), mp, mu_, b.expand(Iw),
                                                     b.expand(r), b.expand(n), b.expand(jnv.contiguous()), b.expand(jtv),
                                                     vo_p, wo_p, current_stream(b.device)))
    return b.out(vo), b.out(wo)


def apply_impulse(vel, omega, mass, inertia_world, contact_point, normal, impulse):
    """src/physics/physics_utils.py:4-22 (normal-only impulse)."""
    b = _Batch(vel, omega, inertia_world, contact_point, normal)
    v, w, Iw = b.vec(vel, (3,)), b.vec(omega, (3,)), b.vec(inertia_world, (3, 3))
    r, n = b.vec(contact_point, (3,)), b.vec(normal, (3,))
    (mp, mu_), (ip, iu) = b.scalar(mass), b.scalar(impulse)
    vo, vo_p = b.empty(3)
    wo, wo_p = b.empty(3)
    _lib.check(_lib.load().rbs_apply_impulse(rbs_dtype(b.dtype), b.n, b.expand(v), b.expand(w), mp, mu_, b.expand(Iw),
                                            b.expand(r), b.expand(n), ip, iu, vo_p, wo_p, current_stream(b.device)))
    return b.out(vo), b.out(wo)


def compute_inertia_tensor_world(inertia_diag, q):
    """src/physics/collision.py:51-53: ``R(q) diag(I) R(q)^T`` with q = wxyz (normalised like SciPy does).
    A zero quaternion raises ValueError like ``scipy.spatial.transform.Rotation.from_quat``."""
    b = _Batch(inertia_diag, q)
    d, qq = b.vec(inertia_diag, (3,)), b.vec(q, (4,))
    if not (isinstance(q, torch.Tensor) and q.device.type == "cuda"):      # host inputs: cheap to validate
        if bool((qq.abs().sum(dim=1) == 0).any()):
            raise ValueError("Found zero norm quaternions in `quat`.")
    out, out_p = b.empty(3, 3)
    _lib.check(_lib.load().rbs_inertia_world(rbs_dtype(b.dtype), b.n, b.expand(d), b.expand(qq), out_p, current_stream(b.device)))
    return b.out(out)


def compute_inverse_inertia(mass, radius):
    """src/simulation/ball_collision.py:39-41: ``eye(3) / (2/5 m r^2)`` -- a host-side constant."""
    return np.eye(3) / ((2.0 / 5.0) * mass * radius ** 2)


def compute_collision_impulse(mass, I_inv, v_lin, v_ang, r, n, restitution, mu):
    """src/simulation/ball_collision.py:53-68.  ``I_inv`` must be a multiple of the identity (a sphere), as
    ``compute_inverse_inertia`` produces; a scalar or per-item array of the diagonal value is also accepted."""
    b = _Batch(v_lin, v_ang, r, n)
    v, w, rr, nn = b.vec(v_lin, (3,)), b.vec(v_ang, (3,)), b.vec(r, (3,)), b.vec(n, (3,))
    iinv = I_inv
    if not torch.is_tensor(iinv):
        iinv = np.asarray(iinv, dtype=np.float64)
    if iinv.ndim >= 2 and tuple(iinv.shape[-2:]) == (3, 3):
        diag = iinv[..., 0, 0]
        off = iinv - diag[..., None, None] * (torch.eye(3, dtype=iinv.dtype, device=iinv.device) if torch.is_tensor(iinv) else np.eye(3))
        if float(abs(off).max()) != 0.0:
            raise ValueError("compute_collision_impulse: I_inv must be isotropic (a multiple of the identity)")
        iinv = diag
    iinv = float(iinv) if getattr(iinv, "ndim", 0) == 0 else iinv
    (mp, mu_), (ip, iu), (ep, eu), (fp, fu) = b.scalar(mass), b.scalar(iinv), b.scalar(restitution), b.scalar(mu)
    J, J_p = b.empty(3)
    _lib.check(_lib.load().rbs_two_ball_impulse(rbs_dtype(b.dtype), b.n, mp, mu_, ip, iu, b.expand(v), b.expand(w), b.expand(rr),
                                               b.expand(nn), ep, eu, fp, fu, J_p, current_stream(b.device)))
    return b.out(J)
