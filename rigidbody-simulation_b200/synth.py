"""Deterministic synthetic initial states for the BASELINE configs.

Every number is a pure function of ``(seed, global environment index, field id)`` through a
counter-based hash (splitmix64 finaliser), so a shard of environments [start, start+count) gets the
same values no matter how many GPUs the job is split over.  Generated in float64 on the host, in the
reference's layout: qpos[E, 7*B] (xyz + wxyz), qvel[E, 6*B] (linear + angular).
"""
import math

import numpy as np

SEED = 20261018
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(z):
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uniform01(seed, field, index):
    """U[0,1) with 53 random bits, keyed by (seed, field, index[...])."""
    with np.errstate(over="ignore"):
        idx = np.asarray(index, dtype=np.uint64)
        key = _mix(np.uint64(seed) + np.uint64(field) * np.uint64(0x9E3779B97F4A7C15))
        z = _mix(idx * np.uint64(0xD1342543DE82EF95) + key)
        z = _mix(z + np.uint64(0x9E3779B97F4A7C15))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


class _Fields:
    def __init__(self, seed, start, count, base):
        self.seed, self.idx, self.k = seed, np.arange(start, start + count, dtype=np.uint64), base

    def u(self, lo=0.0, hi=1.0, shape=None):
        """next field: [count] (or [count, *shape]) uniform in [lo, hi)"""
        if shape is None:
            self.k += 1
            return lo + (hi - lo) * uniform01(self.seed, self.k, self.idx)
        cols = [self.u(lo, hi) for _ in range(int(np.prod(shape)))]
        return np.stack(cols, axis=-1).reshape((self.idx.size,) + tuple(shape))

    def unit_quat(self):
        """uniform on S^3: normalised 4-Gaussian (Box-Muller on four uniform pairs)"""
        g = []
        for _ in range(2):
            u1, u2 = self.u(), self.u()
            rad = np.sqrt(-2.0 * np.log(1.0 - u1))
            g += [rad * np.cos(2 * math.pi * u2), rad * np.sin(2 * math.pi * u2)]
        q = np.stack(g, axis=-1)
        return q / np.linalg.norm(q, axis=-1, keepdims=True)


def incline_normal(theta):
    return np.array([0.0, -math.sin(theta), math.cos(theta)])


def sphere_incline(count, start=0, seed=SEED, theta=0.7, radius=0.2):
    """config 2: models/sphere.xml body over a plane tilted ``theta`` about x; per-env restitution, friction."""
    f = _Fields(seed, start, count, 200)
    n = incline_normal(theta)
    tx, ty = np.array([1.0, 0.0, 0.0]), np.cross(n, [1.0, 0.0, 0.0])
    h, a, b = f.u(0.25, 2.5), f.u(-1, 1), f.u(-1, 1)
    pos = h[:, None] * n + a[:, None] * tx + b[:, None] * ty
    qpos = np.concatenate([pos, f.unit_quat()], axis=1)
    qvel = np.concatenate([f.u(-2, 2, (3,)), f.u(-5, 5, (3,))], axis=1)
    return dict(qpos=np.ascontiguousarray(qpos), qvel=np.ascontiguousarray(qvel), restitution=f.u(0.5, 1.0),
                friction=f.u(0.0, 1.0), plane_normal=n, radius=radius, dt=0.009, threshold=0.0)


def two_ball(count, start=0, seed=SEED):
    """config 3: models/ball_collision.xml with the shipped ICs (ball_collision.py:31-34) + perturbations."""
    f = _Fields(seed, start, count, 300)
    d = f.u(-0.05, 0.05, (2, 3))
    qpos = np.zeros((count, 14))
    qpos[:, 0:3] = np.array([-1.0, 0.0, 1.0]) + d[:, 0]
    qpos[:, 7:10] = np.array([1.0, 0.0, 1.0]) + d[:, 1]
    qpos[:, 3] = qpos[:, 10] = 1.0
    qvel = np.zeros((count, 12))
    qvel[:, 0:3] = np.array([1.0, 0.0, 0.5]) + f.u(-0.2, 0.2, (3,))
    qvel[:, 6:9] = np.array([-1.0, 0.0, 0.5]) + f.u(-0.2, 0.2, (3,))
    qvel[:, 3:6], qvel[:, 9:12] = f.u(-2, 2, (3,)), f.u(-2, 2, (3,))
    return dict(qpos=qpos, qvel=qvel, restitution=1.0, friction=0.3, radius=0.1, dt=0.01)


def cube(count, start=0, seed=SEED, kind="bounce"):
    """config 4: models/cube.xml body; 'bounce' = flat plane, random pose and velocity; 'incline' = 0.7 rad,
    shipped pose (cube.xml:33) plus a small perturbation, from rest."""
    f = _Fields(seed, start, count, 400 if kind == "bounce" else 450)
    if kind == "bounce":
        pos = np.stack([f.u(-1, 1), f.u(-1, 1), f.u(0.8, 2.0)], axis=1)
        quat = f.unit_quat()
        qvel = np.concatenate([f.u(-1, 1, (3,)), f.u(-3, 3, (3,))], axis=1)
        theta = 0.0
    elif kind == "incline":
        pos = np.array([0.0, 0.0, 0.4]) + f.u(-0.02, 0.02, (3,))
        quat = np.array([math.cos(0.35), math.sin(0.35), 0.0, 0.0]) + f.u(-0.01, 0.01, (4,))
        quat /= np.linalg.norm(quat, axis=1, keepdims=True)
        qvel = np.zeros((count, 6))
        theta = 0.7
    else:
        raise ValueError(kind)
    return dict(qpos=np.ascontiguousarray(np.concatenate([pos, quat], axis=1)), qvel=np.ascontiguousarray(qvel),
                restitution=0.2, friction=0.6, threshold=1e-4, dt=0.009, theta=theta,
                plane_normal=incline_normal(theta), half=[0.4, 0.4, 0.4])


def multi_sphere(count, n_body=64, start=0, seed=SEED, friction=0.0):
    """config 5: ``n_body`` spheres (r = 0.1) on a jittered cubic lattice of pitch 0.3, lowest layer at z ~ 0.3,
    random linear velocities, so that ball-ball contacts happen early."""
    side = int(round(n_body ** (1.0 / 3.0)))
    while side ** 3 < n_body:
        side += 1
    cells = np.stack(np.meshgrid(*[np.arange(side)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n_body]
    f = _Fields(seed, start, count, 500)
    pos = cells[None].astype(np.float64) * 0.3 + f.u(-0.04, 0.04, (n_body, 3))
    pos[:, :, 2] += 0.3
    qpos = np.zeros((count, n_body, 7))
    qpos[:, :, :3] = pos
    qpos[:, :, 3] = 1.0
    qvel = np.zeros((count, n_body, 6))
    qvel[:, :, :3] = f.u(-1, 1, (n_body, 3))
    return dict(qpos=qpos.reshape(count, 7 * n_body), qvel=qvel.reshape(count, 6 * n_body), restitution=1.0,
                friction=friction, radius=0.1, dt=0.01, n_body=n_body)


def multi_body(count, bodies, start=0, seed=SEED, pitch=1.0):
    """N4 test / bench scene: len(bodies) spheres and boxes (the dicts of scenes.multi_body_xml) on a jittered cubic
    lattice of ``pitch`` above the plane, random orientations, linear and angular velocities, so that sphere-box and
    box-box contacts happen within the first few hundred steps."""
    n_body = len(bodies)
    side = 1
    while side ** 3 < n_body:
        side += 1
    cells = np.stack(np.meshgrid(*[np.arange(side)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n_body]
    f = _Fields(seed, start, count, 700)
    pos = cells[None].astype(np.float64) * pitch + f.u(-0.1 * pitch, 0.1 * pitch, (n_body, 3))
    pos[:, :, 2] += 0.9 * pitch
    qpos = np.zeros((count, n_body, 7))
    qpos[:, :, :3] = pos
    quat = np.stack([f.unit_quat() for _ in range(n_body)], axis=1)
    qpos[:, :, 3:] = quat
    qvel = np.zeros((count, n_body, 6))
    qvel[:, :, :3] = f.u(-1, 1, (n_body, 3))
    qvel[:, :, 3:] = f.u(-2, 2, (n_body, 3))
    return dict(qpos=qpos.reshape(count, 7 * n_body), qvel=qvel.reshape(count, 6 * n_body), n_body=n_body)

