"""Run the UNMODIFIED reference under the fake MuJoCo backend -- TEST INFRASTRUCTURE ONLY.

Only usable where the reference checkout exists (``BASELINE.json: reference_path``, i.e. the
build container).  It never travels to the GPU box; what travels are the golden vectors that
``oracle/make_golden.py`` produces with it (tests/golden/*.json) and the restatements
(oracle/pyport.py, oracle/rb_oracle.c) that are pinned against those vectors.

Nothing in the product package may import this module.
"""
import contextlib
import importlib
import json
import os
import runpy
import shutil
import sys
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
FAKE_BACKEND = os.path.join(_HERE, "fake_backend")


def reference_path():
    with open(os.path.join(os.path.dirname(_HERE), "BASELINE.json")) as f:
        return json.load(f).get("reference_path", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(reference_path(), "src", "physics", "collision.py"))


def _purge(prefixes):
    for name in list(sys.modules):
        if any(name == p or name.startswith(p + ".") for p in prefixes):
            del sys.modules[name]


@contextlib.contextmanager
def reference_imports():
    """sys.path = [fake backend, reference root, ...]; the names ``src``, ``mujoco``, ``glfw``,
    ``imageio``, ``matplotlib`` are purged on entry and exit so nothing leaks either way."""
    names = ["src", "mujoco", "glfw", "imageio", "matplotlib", "mpl_toolkits"]
    saved_path = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items()
                  if any(k == p or k.startswith(p + ".") for p in names)}
    _purge(names)
    sys.path[:0] = [FAKE_BACKEND, reference_path()]
    try:
        yield
    finally:
        _purge(names)
        sys.modules.update(saved_mods)
        sys.path[:] = saved_path


class CallCounter:
    """Counts calls to A1 (contacts processed) and how many returned a non-zero normal impulse
    decision (``u_n < 0``  <=>  the early return at collision.py:32-33 was NOT taken)."""

    def __init__(self, fn):
        self.fn = fn
        self.calls = 0
        self.impulses = 0

    def __call__(self, mass, inertia_world, vel, omega, contact_point, normal, restitution, friction_coeff):
        self.calls += 1
        u_n = np.dot(vel + np.cross(omega, contact_point), normal)
        if not (u_n >= 0):
            self.impulses += 1
        return self.fn(mass, inertia_world, vel, omega, contact_point, normal, restitution, friction_coeff)


def install_counter():
    """Patch the reference's A1 in both modules that bind the name; returns the counter."""
    col = importlib.import_module("src.physics.collision")
    ti = importlib.import_module("src.physics.time_integeration")
    orig = getattr(col.compute_collision_impulse_friction, "fn", col.compute_collision_impulse_friction)
    counter = CallCounter(orig)
    col.compute_collision_impulse_friction = counter
    ti.compute_collision_impulse_friction = counter
    return counter


# --------------------------------------------------------------------------- whole scripts
def run_script(script, steps, press_space=False, record=None):
    """runpy-execute ``src/simulation/<script>.py`` unmodified for ``steps`` loop iterations.

    The scripts write ``models/*_temp.xml`` relative to cwd (single_sphere_bounce.py:32-34), so
    they run inside a scratch directory holding a copy of ``models/``.
    Returns dict(qpos, qvel, calls, impulses, log) -- ``log`` is the DataLogger trace if any."""
    ref = reference_path()
    tmp = tempfile.mkdtemp(prefix="rbs_oracle_")
    cwd = os.getcwd()
    try:
        shutil.copytree(os.path.join(ref, "models"), os.path.join(tmp, "models"))
        os.chdir(tmp)
        with reference_imports():
            import glfw
            glfw.reset(steps, press_space)
            counter = install_counter()
            g = runpy.run_path(os.path.join(ref, "src", "simulation", script + ".py"), run_name="__main__")
            out = {
                "qpos": np.array(g["data"].qpos, dtype=float).tolist(),
                "qvel": np.array(g["data"].qvel, dtype=float).tolist(),
                "calls": counter.calls,
                "impulses": counter.impulses,
            }
            lg = g.get("logger")
            if lg is not None and hasattr(lg, "times"):
                out["log_t"] = [float(v) for v in lg.times]
                out["log_x"] = [float(v) for v in lg.x_positions]
                out["log_y"] = [float(v) for v in lg.y_positions]
                out["log_z"] = [float(v) for v in lg.z_positions]
            for k in ("logger_ball1", "logger_ball2"):
                if k in g:
                    out[k + "_z"] = [float(v) for v in g[k].z_positions]
                    out[k + "_x"] = [float(v) for v in g[k].x_positions]
            return out
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)


# --------------------------------------------------------------------------- XML builders
def single_body_xml(geom, size, plane_euler=(0.0, 0.0, 0.0), body_pos=(0, 0, 1), body_euler=(0, 0, 0),
                    timestep=0.009, gravity=(0, 0, -9.8), density=50.0, name="obj"):
    """A scene of the same shape as models/sphere.xml / models/cube.xml (plane body + one free body)."""
    sz = " ".join(repr(float(s)) for s in np.atleast_1d(size))
    f = lambda v: " ".join(repr(float(x)) for x in v)
    return f"""<mujoco>
  <compiler angle="radian" coordinate="local" inertiafromgeom="true"/>
  <option gravity="{f(gravity)}" timestep="{timestep!r}"/>
  <worldbody>
    <body name="inclined_plane" pos="0 0 0">
      <geom name="ground" pos="0 0 0" size="5 5 0.1" type="plane" euler="{f(plane_euler)}"/>
    </body>
    <body name="{name}" pos="{f(body_pos)}" euler="{f(body_euler)}">
      <joint name="j" type="free"/>
      <geom name="g" size="{sz}" type="{geom}" density="{density!r}"/>
    </body>
  </worldbody>
</mujoco>"""


def multi_sphere_xml(nball, radius=0.1, timestep=0.01, gravity=(0, 0, -9.8), density=50.0):
    """Same shape as models/multi_sphere.xml: world plane + ``nball`` free spheres ball1..ballN."""
    f = lambda v: " ".join(repr(float(x)) for x in v)
    balls = "\n".join(
        f'<body name="ball{i+1}" pos="0 0 {1 + i}"><joint name="ball_joint{i+1}" type="free"/>'
        f'<geom name="ball_geom{i+1}" size="{radius!r}" type="sphere" density="{density!r}"/></body>'
        for i in range(nball))
    return f"""<mujoco>
  <compiler angle="radian" coordinate="local" inertiafromgeom="true"/>
  <option gravity="{f(gravity)}" timestep="{timestep!r}"/>
  <worldbody>
    <geom name="ground" pos="0 0 0" size="5 5 0.1" type="plane"/>
    {balls}
  </worldbody>
</mujoco>"""


# --------------------------------------------------------------------------- step functions
def run_single_body(scheme, xml, qpos0, qvel0, steps, dt, restitution, friction, threshold,
                    body_name="obj", snapshots=(), xfrc=None):
    """Step one env with the reference's A5 / A6 / A7 (``scheme`` in {'custom','timestep','general'}).
    Returns dict with final qpos/qvel, counts and the requested per-step snapshots."""
    with reference_imports():
        import mujoco as mj
        counter = install_counter()
        col = importlib.import_module("src.physics.collision")
        ti = importlib.import_module("src.physics.time_integeration")
        fn = {"custom": col.custom_step_with_impulse_collision_friction,
              "timestep": ti.timestep_integration, "general": ti.general}[scheme]
        model = mj.MjModel.from_xml_string(xml)
        data = mj.MjData(model)
        data.qpos[:] = qpos0
        data.qvel[:] = qvel0
        if xfrc is not None:
            data.xfrc_applied[mj.mj_name2id(model, mj.mjtObj.mjOBJ_BODY, body_name)] = xfrc
        snaps = {}
        for s in range(1, steps + 1):
            fn(model, body_name, data, dt=dt, restitution=restitution, friction_coeff=friction,
               contact_threshold=threshold)
            if s in snapshots:
                snaps[str(s)] = {"qpos": data.qpos.tolist(), "qvel": data.qvel.tolist(),
                                 "calls": counter.calls, "impulses": counter.impulses}
        return {"qpos": data.qpos.tolist(), "qvel": data.qvel.tolist(),
                "calls": counter.calls, "impulses": counter.impulses, "snapshots": snaps,
                "mass": float(model.body_mass[-1]), "inertia": model.body_inertia[-1].tolist(),
                "plane_normal": plane_normal_of(mj, model)}


def plane_normal_of(mj, model):
    """The unit normal the fake narrow phase actually uses for the (single) plane geom."""
    g = [x for x in model.geoms if x.type == "plane"][0]
    rb = mj._quat_to_mat(model.body_quat[g.body])
    return (rb @ mj._quat_to_mat(g.quat))[:, 2].tolist() if not (g.quat == [1.0, 0, 0, 0]).all() else rb[:, 2].tolist()


def run_multi_sphere(xml, qpos0, qvel0, steps, dt, restitution, friction, snapshots=()):
    """The loop of src/simulation/multi_sphere_bounce.py:42-92 with the three repairs of SURVEY
    section 8 row A9 (0-based joint slices, contact ownership by body, no glfw), calling the
    reference's own A1 / A2 / A4.  The shipped file cannot run (IndexError on its first step)."""
    with reference_imports():
        import mujoco as mj
        counter = install_counter()
        col = importlib.import_module("src.physics.collision")
        pu = importlib.import_module("src.physics.physics_utils")
        model = mj.MjModel.from_xml_string(xml)
        data = mj.MjData(model)
        data.qpos[:] = qpos0
        data.qvel[:] = qvel0
        nball = model.nq // 7
        per_ball_calls = np.zeros(nball, dtype=int)
        per_ball_imp = np.zeros(nball, dtype=int)
        snaps = {}
        for s in range(1, steps + 1):
            mj.mj_forward(model, data)                                        # :43
            for b in range(nball):                                            # :46
                body_id = b + 1
                mass = model.body_mass[body_id]                               # :48
                inertia_diag = model.body_inertia[body_id]                    # :49
                qpos = data.qpos[7 * b: 7 * b + 7]                            # :50 (repaired index)
                qvel = data.qvel[6 * b: 6 * b + 6]                            # :51
                vel = qvel[:3]
                omega = qvel[3:6]
                inertia_world = col.compute_inertia_tensor_world(inertia_diag, qpos[3:7])   # :55
                force = data.xfrc_applied[body_id, :3] + mass * model.opt.gravity           # :58
                torque = data.xfrc_applied[body_id, 3:]
                vel += (force / mass) * dt                                    # :60
                omega += np.linalg.inv(inertia_world) @ (torque * dt)         # :61
                for i in range(data.ncon):                                    # :64
                    c = data.contact[i]
                    owners = (model.geoms[c.geom1].body, model.geoms[c.geom2].body)
                    if c.dist < 0 and body_id in owners:                      # :66 (repaired filter)
                        r = c.pos - qpos[:3]                                  # :67
                        n = c.frame[:3]                                       # :68
                        c0, i0 = counter.calls, counter.impulses
                        jn, jt = col.compute_collision_impulse_friction(
                            mass, inertia_world, vel, omega, r, n, restitution, friction)   # :69
                        per_ball_calls[b] += counter.calls - c0
                        per_ball_imp[b] += counter.impulses - i0
                        vel, omega = pu.apply_impulse_friction(vel, omega, mass, inertia_world, r, n, jn, jt)  # :72
                pos_new = qpos[:3] + vel * dt                                 # :77
                res = np.zeros(4)
                mj.mju_mulQuat(res, np.concatenate([[0], omega]), qpos[3:7])  # :78-80
                quat_new = qpos[3:7] + 0.5 * res * dt                         # :81
                quat_new /= np.linalg.norm(quat_new)                          # :82
                data.qpos[7 * b: 7 * b + 3] = pos_new                         # :85-88
                data.qpos[7 * b + 3: 7 * b + 7] = quat_new
                data.qvel[6 * b: 6 * b + 3] = vel
                data.qvel[6 * b + 3: 6 * b + 6] = omega
            if s in snapshots:
                snaps[str(s)] = {"qpos": data.qpos.tolist(), "qvel": data.qvel.tolist()}
        return {"qpos": data.qpos.tolist(), "qvel": data.qvel.tolist(),
                "calls": per_ball_calls.tolist(), "impulses": per_ball_imp.tolist(), "snapshots": snaps,
                "mass": float(model.body_mass[1]), "inertia": model.body_inertia[1].tolist()}


def run_two_ball(qpos0, qvel0, steps, dt=0.01, snapshots=()):
    """Execute src/simulation/ball_collision.py with zero loop iterations to obtain its namespace
    (functions, masses, inverse inertias), then drive ``step_with_custom_collisions`` (A11)."""
    ref = reference_path()
    tmp = tempfile.mkdtemp(prefix="rbs_oracle_")
    cwd = os.getcwd()
    try:
        shutil.copytree(os.path.join(ref, "models"), os.path.join(tmp, "models"))
        os.chdir(tmp)
        with reference_imports():
            import glfw
            glfw.reset(0, False)
            g = runpy.run_path(os.path.join(ref, "src", "simulation", "ball_collision.py"), run_name="__main__")
            model, data = g["model"], g["data"]
            data.qpos[:] = qpos0
            data.qvel[:] = qvel0
            snaps = {}
            for s in range(1, steps + 1):
                g["step_with_custom_collisions"](model, data, dt)
                if s in snapshots:
                    snaps[str(s)] = {"qpos": data.qpos.tolist(), "qvel": data.qvel.tolist()}
            return {"qpos": data.qpos.tolist(), "qvel": data.qvel.tolist(), "snapshots": snaps,
                    "mass": float(g["mass1"]), "inv_inertia": float(g["I_inv_ball1"][0, 0])}
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)


# --------------------------------------------------------------------------- free functions
def free_functions():
    """Returns the reference's free functions A1-A4 and A10 as a dict of callables (valid only
    inside the ``reference_imports`` context that the caller holds)."""
    col = importlib.import_module("src.physics.collision")
    pu = importlib.import_module("src.physics.physics_utils")
    return {"A1": getattr(col.compute_collision_impulse_friction, "fn", col.compute_collision_impulse_friction),
            "A2": pu.apply_impulse_friction, "A3": pu.apply_impulse, "A4": col.compute_inertia_tensor_world}
