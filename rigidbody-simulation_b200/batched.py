"""Batched stand-ins for ``mujoco.MjModel`` / ``mujoco.MjData`` on the reference's custom-physics path.

The reference keeps one scene in an ``MjData`` (``qpos`` = 7 numbers per free body: xyz + wxyz
quaternion, ``qvel`` = 6: linear + angular) and mutates it in place once per step
(src/physics/collision.py:97-100).  Here one object holds ``nenv`` independent copies of the scene as
a device-resident SoA tensor -- 13 component rows per body, consecutive environments at consecutive
addresses -- which is what the CUDA steppers read and write (layout: include/rbsim_b200.h).

``data.qpos`` / ``data.qvel`` are AoS *views materialised on demand* (for initial conditions, logging and
tests); indexing them is a device gather/scatter, not part of the step path.  With ``nenv == 1`` they hand
back NumPy arrays shaped like the reference's, so the reference's scenario scripts read the same.
"""
import numpy as np
import torch

from . import mjcf

ROWS = 13   # px py pz | qw qx qy qz | vx vy vz | wx wy wz


class _Opt:
    def __init__(self, gravity, timestep):
        self.gravity = np.asarray(gravity, dtype=np.float64)
        self.timestep = float(timestep)


class BatchedModel:
    """Scene description shared by all environments (+ optional per-environment overrides).

    MuJoCo-compatible attributes used by the reference: ``body_mass[nbody]``, ``body_inertia[nbody,3]``,
    ``opt.gravity``, ``opt.timestep`` (src/physics/collision.py:60-66).  ``per_env`` may hold device tensors
    ``mass``, ``inertia`` ([3,E]), ``size`` ([3,E]), ``restitution``, ``friction`` ([E]) that override the
    uniform values environment by environment (randomised configs of SURVEY.md section 8(d)).
    """

    def __init__(self, scene, nenv=1, device=None, dtype=torch.float64):
        if dtype not in (torch.float64, torch.float32):
            raise ValueError("dtype must be torch.float64 (reference precision) or torch.float32")
        self.scene = scene
        self.nenv = int(nenv)
        self.dtype = dtype
        self.device = torch.device(device if device is not None else "cuda")
        self.opt = _Opt(scene.gravity, scene.timestep)
        self.body_names = [b.name for b in scene.bodies]
        self.body_mass = np.array([b.mass for b in scene.bodies], dtype=np.float64)
        self.body_inertia = np.array([b.inertia for b in scene.bodies], dtype=np.float64)
        self.free_ids = scene.free_bodies                     # MuJoCo body ids of the free bodies, in qpos order
        self.nfree = len(self.free_ids)
        self.nq, self.nv = 7 * self.nfree, 6 * self.nfree
        qpos0 = []
        for i in self.free_ids:
            qpos0 += list(scene.bodies[i].pos) + list(scene.bodies[i].quat)
        self.qpos0 = np.array(qpos0, dtype=np.float64)
        planes = scene.planes()
        if len(planes) > 1:
            raise ValueError("scenes with more than one plane are outside the supported subset")
        if planes:
            self.plane_point, self.plane_normal = mjcf.plane_frame(scene, planes[0])
        else:
            self.plane_point, self.plane_normal = None, None
        self.body_geom = {i: scene.geoms[scene.bodies[i].geoms[0]] for i in self.free_ids}
        # geoms placed off their body's origin / axes: only stepper.step_multi_body generates contacts for those
        self.has_offset_geoms = any(any(g.pos) or list(g.quat) != [1.0, 0.0, 0.0, 0.0] for g in self.body_geom.values())
        self.per_env = {}

    # -- construction ------------------------------------------------------------------------------
    @classmethod
    def from_xml_string(cls, text, nenv=1, device=None, dtype=torch.float64):
        return cls(mjcf.parse_string(text), nenv, device, dtype)

    @classmethod
    def from_xml_path(cls, path, nenv=1, device=None, dtype=torch.float64, incline_angle=None, timestep=None):
        return cls(mjcf.parse_file(path, incline_angle, timestep), nenv, device, dtype)

    @property
    def nbody(self):
        return len(self.body_names)

    def set_per_env(self, **tensors):
        """Attach per-environment parameter overrides (host arrays are uploaded once)."""
        for key, val in tensors.items():
            if key not in ("mass", "inertia", "size", "restitution", "friction", "radius"):
                raise KeyError(key)
            t = torch.as_tensor(val, dtype=self.dtype).to(self.device).contiguous()
            self.per_env[key] = t
        return self


class StateView:
    """AoS window (``[nenv, width*nfree]``, the reference's qpos / qvel) onto the SoA device state."""

    def __init__(self, data, first_row, width):
        self._d, self._r0, self._w = data, first_row, width

    def _gather(self):
        d = self._d
        rows = d.rows(self._r0, self._w)                  # [w, nfree, E]
        return rows.permute(2, 1, 0).reshape(d.nenv, d.nfree * self._w)

    def __getitem__(self, idx):
        a = self._gather()
        if self._d.squeeze:                      # reference shape: qpos is (7*nfree,)
            return a[0].detach().cpu().numpy()[idx].copy()
        return a[idx]                            # batched: plain tensor indexing on [nenv, width*nfree]

    def __setitem__(self, idx, value):
        d = self._d
        a = self._gather().clone()
        val = value if torch.is_tensor(value) else torch.as_tensor(np.asarray(value, dtype=np.float64))
        val = val.to(device=d.device, dtype=d.dtype)
        if d.squeeze:
            a[0][idx] = val
        else:
            a[idx] = val
        d.rows(self._r0, self._w).copy_(a.reshape(d.nenv, d.nfree, self._w).permute(2, 1, 0))

    def __array__(self, dtype=None, copy=None):
        a = self._gather().detach().cpu().numpy()
        a = a[0] if self._d.squeeze else a
        return a.astype(dtype) if dtype is not None else a

    def __len__(self):
        return self._d.nfree * self._w if self._d.squeeze else self._d.nenv

    @property
    def shape(self):
        n = self._d.nfree * self._w
        return (n,) if self._d.squeeze else (self._d.nenv, n)

    def torch(self):
        """The full AoS tensor [nenv, width*nfree] (a fresh device tensor)."""
        return self._gather().contiguous()

    def tolist(self):
        return np.asarray(self).tolist()

    def copy(self):
        return np.asarray(self).copy()

    def __repr__(self):
        return f"StateView({np.asarray(self)!r})"


class BatchedData:
    """Device state of ``nenv`` environments (the MjData of the reference, batched).

    layout "env"  : state[13, nfree, stride]   element (row, body, env)   -- thread-per-env steppers
    layout "body" : state[13, stride]          element (row, env*nfree+b) -- thread-per-body stepper
    """

    def __init__(self, model, layout="env"):
        if layout not in ("env", "body"):
            raise ValueError(layout)
        self.model = model
        self.layout = layout
        self.nenv, self.nfree = model.nenv, model.nfree
        self.dtype, self.device = model.dtype, model.device
        self.squeeze = model.nenv == 1
        E, B = self.nenv, self.nfree
        if layout == "env":
            self.state = torch.empty((ROWS, B, E), dtype=self.dtype, device=self.device)
            self.stride = E
        else:
            self.state = torch.empty((ROWS, E * B), dtype=self.dtype, device=self.device)
            self.stride = E * B
        self.xfrc_applied = None              # [6, nenv] (env layout, single body) once set_xfrc() is called
        self.time = 0.0                       # mj_forward never advances it; neither does the custom path
        self.ncon = 0
        self.n_contacts = torch.zeros(E * B, dtype=torch.int32, device=self.device)
        self.n_impulses = torch.zeros(E * B, dtype=torch.int32, device=self.device)
        self.qpos = StateView(self, 0, 7)
        self.qvel = StateView(self, 7, 6)
        self._init_state()

    def rows(self, first, count):
        """state rows [first, first+count) as a [count, nfree, nenv] view (both layouts)."""
        if self.layout == "env":
            return self.state[first:first + count]
        return self.state[first:first + count].view(count, self.nenv, self.nfree).permute(0, 2, 1)

    def _init_state(self):
        """Fresh buffers at construction: qpos0, zero velocities (plain fills -- allocation plumbing, also on a CPU device)."""
        q0 = torch.as_tensor(self.model.qpos0, dtype=self.dtype).to(self.device).view(self.nfree, 7).t()  # [7, nfree]
        rows = self.rows(0, ROWS)
        rows[:7] = q0.unsqueeze(-1)
        rows[7:] = 0
        self.time = 0.0

    def reset(self, env_mask=None):
        """mj_resetData (src/viewer/mujoco_viewer.py:62-65): qpos0, zero velocities, zero event counters; optionally
        only the environments selected by a boolean mask (host or device).  One launch of ``rbs_reset_envs`` on the
        current stream, no host sync."""
        import ctypes
        from . import _lib, stepper
        if self.device.type != "cuda":
            raise _lib.RbsError("reset needs the CUDA extension and a CUDA device (there is no CPU fallback)")
        if getattr(self, "_qpos0_dev", None) is None:
            self._qpos0_dev = torch.as_tensor(self.model.qpos0, dtype=self.dtype).to(self.device).contiguous()
        mask_ptr = None
        if env_mask is not None:
            m = torch.as_tensor(env_mask).to(device=self.device, dtype=torch.bool).contiguous()
            if m.numel() != self.nenv:
                raise ValueError(f"env_mask has {m.numel()} entries for {self.nenv} environments")
            mask_u8 = m.view(torch.uint8)
            mask_ptr = ctypes.c_void_p(mask_u8.data_ptr())
        _lib.check(_lib.load().rbs_reset_envs(
            stepper.rbs_dtype(self.dtype), self.nenv, self.nfree, 1 if self.layout == "body" else 0,
            ctypes.c_void_p(self.state.data_ptr()), self.stride, ctypes.c_void_p(self._qpos0_dev.data_ptr()), mask_ptr,
            ctypes.c_void_p(self.n_contacts.data_ptr()), ctypes.c_void_p(self.n_impulses.data_ptr()),
            stepper.current_stream(self.device)))
        if env_mask is None:
            self.time = 0.0

    def set_state(self, qpos, qvel):
        """Load AoS host/device arrays qpos[nenv, 7*nfree], qvel[nenv, 6*nfree] (reference layout)."""
        qp = torch.as_tensor(qpos, dtype=self.dtype).to(self.device).reshape(self.nenv, self.nfree, 7)
        qv = torch.as_tensor(qvel, dtype=self.dtype).to(self.device).reshape(self.nenv, self.nfree, 6)
        self.rows(0, 7).copy_(qp.permute(2, 1, 0))
        self.rows(7, 6).copy_(qv.permute(2, 1, 0))

    def set_xfrc(self, xfrc):
        """data.xfrc_applied for the (single) free body: [nenv, 6] force + torque (collision.py:66-67)."""
        if self.nfree != 1:
            raise ValueError("applied wrenches are supported for single-body scenes")
        x = torch.as_tensor(xfrc, dtype=self.dtype).to(self.device).reshape(self.nenv, 6)
        self.xfrc_applied = x.t().contiguous()

    def state_dict(self):
        """Checkpoint: the SoA state, the event counters, the applied wrench and the (never advanced) time.
        The reference has no checkpointing (only mj_resetData); resuming from this is bit-exact because the
        steppers keep no hidden state between launches."""
        return {"layout": self.layout, "state": self.state.detach().cpu().clone(), "n_contacts": self.n_contacts.cpu().clone(),
                "n_impulses": self.n_impulses.cpu().clone(), "time": self.time,
                "xfrc_applied": None if self.xfrc_applied is None else self.xfrc_applied.cpu().clone()}

    def load_state_dict(self, sd):
        if sd["layout"] != self.layout or tuple(sd["state"].shape) != tuple(self.state.shape):
            raise ValueError("checkpoint does not match this scene (layout / shape)")
        self.state.copy_(sd["state"].to(self.dtype))
        self.n_contacts.copy_(sd["n_contacts"])
        self.n_impulses.copy_(sd["n_impulses"])
        self.time = float(sd["time"])
        x = sd.get("xfrc_applied")
        self.xfrc_applied = None if x is None else x.to(device=self.device, dtype=self.dtype).contiguous()

    def counters(self):
        """Event counters as int64 host arrays.

        Single-body and multi-sphere steppers: (contacts handed to the impulse routine, impulses applied), one entry
        per body, shape ``[nenv, nfree]``.  Two-ball stepper (env layout, two free bodies): the kernel counts per
        ENVIRONMENT -- (ground hits of either ball, ball-ball hits), shape ``[nenv, 1]`` -- because
        ``step_with_custom_collisions`` has no per-ball notion of a pair hit (ball_collision.py:103-118)."""
        nc, ni = (t.cpu().numpy().astype(np.int64) for t in (self.n_contacts, self.n_impulses))
        if self.layout == "env" and self.nfree > 1:
            return nc[:self.nenv].reshape(self.nenv, 1), ni[:self.nenv].reshape(self.nenv, 1)
        return nc.reshape(self.nenv, self.nfree), ni.reshape(self.nenv, self.nfree)
