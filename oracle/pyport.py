"""NumPy port of the reference's Python hot path -- TEST INFRASTRUCTURE ONLY (oracle side).

Why it exists: the reference is a Python package living at BASELINE.json:reference_path, which
does not exist on the GPU box.  This module restates its step functions with the same NumPy /
SciPy primitives (np.cross, np.dot, np.linalg.norm / inv, scipy Rotation), one env per call, on the
fake MuJoCo objects of oracle/fake_backend -- so it costs what the reference costs (tens of
microseconds of interpreter overhead per step) and is what ``bench.py`` times as the CPU baseline
(kind "port").  In the build container tests/test_oracle_golden.py checks it against the golden
vectors produced by the unmodified reference (bit-for-bit on this machine, same NumPy calls).

Citations: paths relative to the reference root.
"""
import os
import sys

import numpy as np
from scipy.spatial.transform import Rotation

_FAKE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fake_backend")


def fake_mujoco():
    """Import the fake ``mujoco`` module without leaving the fake directory on sys.path."""
    if "mujoco" in sys.modules and getattr(sys.modules["mujoco"], "__version__", "") == "0.0-fake":
        return sys.modules["mujoco"]
    sys.path.insert(0, _FAKE)
    try:
        import mujoco
    finally:
        sys.path.remove(_FAKE)
    return mujoco


class Tally:
    """contacts handed to A1 / impulses actually computed (u_n < 0)."""

    def __init__(self):
        self.calls = 0
        self.impulses = 0


def impulse_friction(mass, inertia_world, vel, omega, arm, normal, restitution, friction_coeff, tally=None):
    """A1: src/physics/collision.py:7-48 (inertia_world is accepted and unused, as there)."""
    u = vel + np.cross(omega, arm)                      # :26-27
    u_n = np.dot(u, normal)                             # :28
    u_t = u - u_n * normal                              # :29
    if tally is not None:
        tally.calls += 1
    if u_n >= 0:                                        # :32-33
        return 0.0, np.zeros(3)
    if tally is not None:
        tally.impulses += 1
    k = (1.0 / mass) + (1.0 / 18)                       # :36
    jn = -(1 + restitution) * u_n / k                   # :39
    jt = np.zeros(3)
    if np.linalg.norm(u_t) > 1e-6:                      # :43
        cap = friction_coeff * abs(jn)                  # :44
        jt = -min(cap, np.linalg.norm(u_t)) * (u_t / np.linalg.norm(u_t))   # :45-46
    return jn, jt


def apply_impulse_friction(vel, omega, mass, inertia_world, arm, normal, jn, jt):
    """A2: src/physics/physics_utils.py:25-49"""
    j_n = jn * normal
    dv = (j_n + jt) / mass
    dw = np.linalg.inv(inertia_world) @ np.cross(arm, (j_n + jt))
    return vel + dv, omega + dw


def apply_impulse(vel, omega, mass, inertia_world, arm, normal, impulse):
    """A3: src/physics/physics_utils.py:4-22"""
    dv = (impulse / mass) * normal
    dw = np.linalg.inv(inertia_world) @ np.cross(arm, impulse * normal)
    return vel + dv, omega + dw


def inertia_world(inertia_diag, q):
    """A4: src/physics/collision.py:51-53"""
    rot = Rotation.from_quat(q[[1, 2, 3, 0]]).as_matrix()
    return rot @ np.diag(inertia_diag) @ rot.T


def _integrate_pose(mj, qpos, vel, omega, dt):
    """collision.py:90-95"""
    pos_new = qpos[:3] + vel * dt
    res = np.zeros(4)
    mj.mju_mulQuat(res, np.concatenate([[0], omega]), qpos[3:7])
    quat_new = qpos[3:7] + 0.5 * res * dt
    quat_new /= np.linalg.norm(quat_new)
    return pos_new, quat_new


def step_scheme_a(model, obj, data, dt=0.01, restitution=1.0, friction_coeff=1.0, contact_threshold=0, tally=None):
    """A5 custom_step_with_impulse_collision_friction (collision.py:56-102); with the defaults
    friction_coeff=0.5, contact_threshold=1e-4 it is A6 timestep_integration
    (time_integeration.py:13-72)."""
    mj = fake_mujoco()
    mj.mj_forward(model, data)                                                   # :57
    bid = mj.mj_name2id(model, mj.mjtObj.mjOBJ_BODY, f"{obj}")                   # :58 (-1 -> last body)
    mass = model.body_mass[bid]
    iw = inertia_world(model.body_inertia[bid], data.qpos[3:7])                  # :60-62
    vel = data.qvel[:3]
    omega = data.qvel[3:6]
    force = data.xfrc_applied[bid, :3] + mass * model.opt.gravity                # :66
    torque = data.xfrc_applied[bid, 3:]
    vel += (force / mass) * dt                                                   # :69 (in place on the view)
    omega += np.linalg.inv(iw) @ (torque * dt)                                   # :70
    for i in range(data.ncon):                                                   # :72
        c = data.contact[i]
        if not np.isnan(c.dist) and c.dist < 0:                                  # :74
            arm = c.pos - data.qpos[:3]
            normal = c.frame[:3]
            if abs(c.dist) < contact_threshold:                                  # :79-80
                continue
            jn, jt = impulse_friction(mass, iw, vel, omega, arm, normal, restitution, friction_coeff, tally)
            vel, omega = apply_impulse_friction(vel, omega, mass, iw, arm, normal, jn, jt)
    pos_new, quat_new = _integrate_pose(mj, data.qpos, vel, omega, dt)           # :90-95
    data.qpos[:3] = pos_new
    data.qpos[3:7] = quat_new
    data.qvel[:3] = vel
    data.qvel[3:6] = omega
    return pos_new


def step_general(model, obj, data, dt=0.01, restitution=1.0, friction_coeff=0.5, contact_threshold=1e-4, tally=None):
    """A7 general (time_integeration.py:75-141): position from the OLD velocity, no quaternion update."""
    mj = fake_mujoco()
    mj.mj_forward(model, data)
    bid = mj.mj_name2id(model, mj.mjtObj.mjOBJ_BODY, f"{obj}")
    mass = model.body_mass[bid]
    iw = inertia_world(model.body_inertia[bid], data.qpos[3:7])
    v_old, w_old = data.qvel[:3], data.qvel[3:6]
    pos_pred = data.qpos[:3] + v_old * dt                                        # :106
    force = data.xfrc_applied[bid, :3] + mass * model.opt.gravity
    torque = data.xfrc_applied[bid, 3:]
    vel = v_old + (force / mass) * dt                                            # :112
    omega = w_old + np.linalg.inv(iw) @ (torque * dt)                            # :113
    for i in range(data.ncon):
        c = data.contact[i]
        if not np.isnan(c.dist) and c.dist < 0:
            arm = c.pos - data.qpos[:3]
            normal = c.frame[:3]
            if abs(c.dist) < contact_threshold:
                continue
            jn, jt = impulse_friction(mass, iw, vel, omega, arm, normal, restitution, friction_coeff, tally)
            vel, omega = apply_impulse_friction(vel, omega, mass, iw, arm, normal, jn, jt)
    data.qpos[:3] = pos_pred                                                     # :137
    data.qvel[:3] = vel
    data.qvel[3:6] = omega
    return pos_pred


def step_multi_sphere(model, data, dt, restitution, friction_coeff, tallies=None):
    """A9 custom_step_multi_sphere (src/simulation/multi_sphere_bounce.py:42-92) with the repairs of
    SURVEY section 8 row A9 (0-based slices, ownership by body id, no name lookup)."""
    mj = fake_mujoco()
    mj.mj_forward(model, data)                                                   # :43
    for b in range(model.nq // 7):                                               # :46
        bid = b + 1
        mass = model.body_mass[bid]
        qpos = data.qpos[7 * b: 7 * b + 7]
        qvel = data.qvel[6 * b: 6 * b + 6]
        vel, omega = qvel[:3], qvel[3:6]
        iw = inertia_world(model.body_inertia[bid], qpos[3:7])                   # :55
        force = data.xfrc_applied[bid, :3] + mass * model.opt.gravity
        torque = data.xfrc_applied[bid, 3:]
        vel += (force / mass) * dt                                               # :60
        omega += np.linalg.inv(iw) @ (torque * dt)                               # :61
        for i in range(data.ncon):                                               # :64
            c = data.contact[i]
            if c.dist < 0 and bid in (model.geoms[c.geom1].body, model.geoms[c.geom2].body):   # :66 repaired
                arm = c.pos - qpos[:3]
                normal = c.frame[:3]
                jn, jt = impulse_friction(mass, iw, vel, omega, arm, normal, restitution, friction_coeff,
                                          None if tallies is None else tallies[b])
                vel, omega = apply_impulse_friction(vel, omega, mass, iw, arm, normal, jn, jt)
        pos_new, quat_new = _integrate_pose(mj, qpos, vel, omega, dt)            # :77-82
        data.qpos[7 * b: 7 * b + 3] = pos_new
        data.qpos[7 * b + 3: 7 * b + 7] = quat_new
        data.qvel[6 * b: 6 * b + 3] = vel
        data.qvel[6 * b + 3: 6 * b + 6] = omega


def two_ball_impulse(mass, i_inv, v_lin, v_ang, arm, n, restitution, mu):
    """A10 compute_collision_impulse (src/simulation/ball_collision.py:53-68)"""
    vc = v_lin + np.cross(v_ang, arm)
    v_n = np.dot(vc, n)
    v_t = vc - v_n * n
    t_norm = np.linalg.norm(v_t)
    den_n = (1.0 / mass) + np.dot(n, np.cross(i_inv @ np.cross(arm, n), arm))
    jn = -(1 + restitution) * v_n / den_n
    t_dir = v_t / t_norm if t_norm > 1e-8 else np.zeros(3)
    den_t = (1.0 / mass) + np.dot(t_dir, np.cross(i_inv @ np.cross(arm, t_dir), arm))
    jt = np.clip(-t_norm / den_t, -mu * abs(jn), mu * abs(jn))
    return jn * n + jt * t_dir


def step_two_ball(data, masses, i_invs, gravity, dt, restitution, mu, radius=0.1):
    """A11 step_with_custom_collisions (src/simulation/ball_collision.py:73-125)"""
    for vi in (0, 6):
        data.qvel[vi:vi + 3] += gravity * dt                                     # :77-78
    for (pi, vi, ai), mass, i_inv in zip(((0, 0, 3), (7, 6, 9)), masses, i_invs):  # :81-97
        pos = data.qpos[pi:pi + 3]
        up = np.array([0.0, 0.0, 1.0])
        if pos[2] < radius:
            arm = (pos - radius * up) - pos
            J = two_ball_impulse(mass, i_inv, data.qvel[vi:vi + 3], data.qvel[ai:ai + 3], arm, up, restitution, mu)
            data.qvel[vi:vi + 3] += J / mass
            data.qvel[ai:ai + 3] += i_inv @ np.cross(arm, J)
            data.qpos[pi + 2] = radius
    diff = data.qpos[7:10] - data.qpos[0:3]                                      # :100
    dist = np.linalg.norm(diff)
    tol = 0.01
    if dist < 2 * radius + tol:                                                  # :103
        n = diff / (dist + 1e-8)
        mid = (data.qpos[0:3] + data.qpos[7:10]) / 2.0
        a1, a2 = mid - data.qpos[0:3], mid - data.qpos[7:10]
        J = two_ball_impulse(masses[0], i_invs[0], data.qvel[0:3], data.qvel[3:6], a1, n, restitution, mu)
        data.qvel[0:3] += J / masses[0]
        data.qvel[3:6] += i_invs[0] @ np.cross(a1, J)
        data.qvel[6:9] -= J / masses[1]
        data.qvel[9:12] -= i_invs[1] @ np.cross(a2, J)
        push = (2 * radius + tol - dist) / 2.0                                   # :116
        data.qpos[0:3] -= push * n
        data.qpos[7:10] += push * n
    for pi, vi in ((0, 0), (7, 6)):
        data.qpos[pi:pi + 3] += data.qvel[vi:vi + 3] * dt                        # :121-122
    return data.qpos[0:3], data.qpos[7:10]


# ------------------------------------------------------------------------------- scene XML (same shapes as models/*.xml)
def single_body_xml(geom, size, plane_euler=(0.0, 0.0, 0.0), timestep=0.009, gravity=(0, 0, -9.8), density=50.0):
    f = lambda v: " ".join(repr(float(x)) for x in np.atleast_1d(v))
    return (f'<mujoco><compiler angle="radian" inertiafromgeom="true"/>'
            f'<option gravity="{f(gravity)}" timestep="{timestep!r}"/><worldbody>'
            f'<body name="inclined_plane" pos="0 0 0"><geom name="ground" type="plane" size="5 5 0.1" '
            f'euler="{f(plane_euler)}"/></body>'
            f'<body name="obj" pos="0 0 1"><joint name="j" type="free"/>'
            f'<geom name="g" type="{geom}" size="{f(size)}" density="{density!r}"/></body></worldbody></mujoco>')


def multi_sphere_xml(nball, radius=0.1, timestep=0.01, density=50.0):
    balls = "".join(f'<body name="ball{i + 1}" pos="0 0 {1 + i}"><joint name="bj{i + 1}" type="free"/>'
                    f'<geom name="bg{i + 1}" type="sphere" size="{radius!r}" density="{density!r}"/></body>'
                    for i in range(nball))
    return (f'<mujoco><compiler angle="radian" inertiafromgeom="true"/>'
            f'<option gravity="0 0 -9.8" timestep="{timestep!r}"/><worldbody>'
            f'<geom name="ground" type="plane" size="5 5 0.1"/>{balls}</worldbody></mujoco>')
