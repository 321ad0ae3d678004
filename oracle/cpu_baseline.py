"""CPU baseline runner -- TEST / MEASUREMENT INFRASTRUCTURE (oracle side).

Times the reference's Python step function -- restated in oracle/pyport.py with the same NumPy / SciPy calls,
because the reference checkout does not exist on the GPU box -- on a bounded sample of the bench workload,
one process per host core.  Kind "port".  Collision detection inside the step is the pure-Python fake MuJoCo
(roughly a fifth of the free-flight step time; real MuJoCo would spend a few microseconds in C instead).

Also offers the C oracle (OpenMP) as a second, much stronger CPU baseline ("native port").
"""
import os
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)


# ------------------------------------------------------------------------------------------------
# The reference itself, when a copy was installed with
#   pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <reference checkout>
# (git-ignored, travels with the gpurun snapshot).  Its setup.py maps package_dir {"": "src"}, so the packages land
# as baseline/_ref/{physics,simulation,config,...} while the sources import each other as ``src.physics...``; a
# namespace alias ``src`` -> baseline/_ref makes the UNMODIFIED files importable.  ``mujoco`` is the fake backend.
# ------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")
FAKE_DIR = os.path.join(_HERE, "fake_backend")


def reference_installed():
    return os.path.isfile(os.path.join(REF_DIR, "physics", "collision.py"))


def _reference_modules():
    """(fake mujoco, src.physics.collision, src.physics.time_integeration, src.physics.physics_utils) of the installed copy"""
    import importlib
    import types
    if "src" not in sys.modules or getattr(sys.modules["src"], "__path__", None) != [REF_DIR]:
        for name in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[name]
        alias = types.ModuleType("src")
        alias.__path__ = [REF_DIR]
        sys.modules["src"] = alias
    import pyport
    mj = pyport.fake_mujoco()                   # registers the fake as sys.modules["mujoco"]: collision.py does ``import mujoco as mj``
    if FAKE_DIR not in sys.path:
        sys.path.append(FAKE_DIR)               # glfw / imageio / matplotlib stubs for the viewer imports of the scripts
    col = importlib.import_module("src.physics.collision")
    ti = importlib.import_module("src.physics.time_integeration")
    pu = importlib.import_module("src.physics.physics_utils")
    return mj, col, ti, pu


def _worker_reference(job):
    """The installed reference's own step functions on a slice of the workload; returns (env_steps, seconds).
    sphere_incline -> custom_step_with_impulse_collision_friction (collision.py:56-102); cube -> timestep_integration
    (time_integeration.py:13-72) called as cube_incline.py:75-77 does; multi_sphere -> the loop of
    multi_sphere_bounce.py:42-92 (repaired indices) around the reference's own A1 / A2 / A4;
    two_ball -> step_with_custom_collisions (ball_collision.py:73-125) taken from the script's namespace."""
    import pyport
    kind, qpos, qvel, extra, steps = job
    mj, col, ti, pu = _reference_modules()
    if kind == "sphere_incline":
        model = mj.MjModel.from_xml_string(pyport.single_body_xml("sphere", [0.2], plane_euler=(extra["theta"], 0, 0)))
        rest, fric = extra["restitution"], extra["friction"]
        t0 = time.perf_counter()
        for i in range(qpos.shape[0]):
            data = mj.MjData(model)
            data.qpos[:], data.qvel[:] = qpos[i], qvel[i]
            for _ in range(steps):
                col.custom_step_with_impulse_collision_friction(model, "obj", data, dt=extra["dt"], restitution=rest[i],
                                                                friction_coeff=fric[i], contact_threshold=0.0)
    elif kind == "cube":
        model = mj.MjModel.from_xml_string(pyport.single_body_xml("box", [0.4, 0.4, 0.4], plane_euler=(extra["theta"], 0, 0)))
        t0 = time.perf_counter()
        for i in range(qpos.shape[0]):
            data = mj.MjData(model)
            data.qpos[:], data.qvel[:] = qpos[i], qvel[i]
            for _ in range(steps):
                ti.timestep_integration(model, "obj", data, dt=0.009, restitution=0.2, friction_coeff=0.6)
    elif kind == "multi_sphere":
        B = extra["n_body"]
        model = mj.MjModel.from_xml_string(pyport.multi_sphere_xml(B))
        t0 = time.perf_counter()
        for i in range(qpos.shape[0]):
            data = mj.MjData(model)
            data.qpos[:], data.qvel[:] = qpos[i], qvel[i]
            for _ in range(steps):
                _reference_multi_sphere_step(mj, col, pu, model, data, 0.01, 1.0, extra["friction"])
    else:
        ns = _reference_two_ball_namespace()
        model, step = ns["model"], ns["step_with_custom_collisions"]
        t0 = time.perf_counter()
        for i in range(qpos.shape[0]):
            data = mj.MjData(model)
            data.qpos[:], data.qvel[:] = qpos[i], qvel[i]
            for _ in range(steps):
                step(model, data, 0.01)
    return qpos.shape[0] * steps, time.perf_counter() - t0


def _reference_multi_sphere_step(mj, col, pu, model, data, dt, restitution, friction):
    """multi_sphere_bounce.py:42-92 with the index / ownership repairs, calling the installed reference's A1, A2, A4"""
    mj.mj_forward(model, data)                                                   # :43
    for b in range(model.nq // 7):                                               # :46
        bid = b + 1
        mass, idiag = model.body_mass[bid], model.body_inertia[bid]
        qpos, qvel = data.qpos[7 * b: 7 * b + 7], data.qvel[6 * b: 6 * b + 6]
        vel, omega = qvel[:3], qvel[3:6]
        iw = col.compute_inertia_tensor_world(idiag, qpos[3:7])                  # :55
        force = data.xfrc_applied[bid, :3] + mass * model.opt.gravity
        torque = data.xfrc_applied[bid, 3:]
        vel += (force / mass) * dt                                               # :60
        omega += np.linalg.inv(iw) @ (torque * dt)                               # :61
        for i in range(data.ncon):                                               # :64
            c = data.contact[i]
            if c.dist < 0 and bid in (model.geoms[c.geom1].body, model.geoms[c.geom2].body):   # :66 repaired
                arm, normal = c.pos - qpos[:3], c.frame[:3]
                jn, jt = col.compute_collision_impulse_friction(mass, iw, vel, omega, arm, normal, restitution, friction)
                vel, omega = pu.apply_impulse_friction(vel, omega, mass, iw, arm, normal, jn, jt)
        pos_new = qpos[:3] + vel * dt                                            # :77
        res = np.zeros(4)
        mj.mju_mulQuat(res, np.concatenate([[0], omega]), qpos[3:7])
        quat_new = qpos[3:7] + 0.5 * res * dt
        quat_new /= np.linalg.norm(quat_new)
        data.qpos[7 * b: 7 * b + 3], data.qpos[7 * b + 3: 7 * b + 7] = pos_new, quat_new
        data.qvel[6 * b: 6 * b + 3], data.qvel[6 * b + 3: 6 * b + 6] = vel, omega


_TWO_BALL_NS = None


def _reference_two_ball_namespace():
    """Run the installed simulation/ball_collision.py with a zero-iteration viewer loop (fake glfw) in a scratch cwd
    holding models/ball_collision.xml, and keep its namespace: model, masses, step_with_custom_collisions."""
    global _TWO_BALL_NS
    if _TWO_BALL_NS is None:
        import runpy
        import shutil
        import tempfile
        import pyport
        tmp = tempfile.mkdtemp(prefix="rbs_ref_")
        cwd = os.getcwd()
        try:
            os.makedirs(os.path.join(tmp, "models"))
            with open(os.path.join(tmp, "models", "ball_collision.xml"), "w") as f:
                f.write(pyport.multi_sphere_xml(2))          # same physics-relevant content as models/ball_collision.xml
            os.chdir(tmp)
            import contextlib
            import glfw
            glfw.reset(0, False)
            with open(os.devnull, "w") as sink, contextlib.redirect_stdout(sink):   # the script reports its (stubbed) plots and video on stdout
                _TWO_BALL_NS = runpy.run_path(os.path.join(REF_DIR, "simulation", "ball_collision.py"), run_name="__main__")
        finally:
            os.chdir(cwd)
            shutil.rmtree(tmp, ignore_errors=True)
    return _TWO_BALL_NS


def _warm_reference(_):
    _reference_modules()
    return 0


def run_config(kind, sample, cores=None, envs_per_core=4, steps=100, extra=None, prefer_reference=True):
    """Aggregate env-steps/s on ``cores`` worker processes for one BASELINE config: the installed reference
    (kind "reference") when baseline/_ref exists, else the NumPy port (kind "port").
    ``kind`` in {'sphere_incline', 'cube', 'two_ball', 'multi_sphere'}."""
    import multiprocessing as mp
    _require_importable_main()
    cores = cores or os.cpu_count() or 1
    n = cores * envs_per_core
    if sample["qpos"].shape[0] < n:
        raise ValueError("sample too small")
    use_ref = prefer_reference and reference_installed()
    extra = dict(extra or {})
    jobs = []
    for c in range(cores):
        sl = slice(c * envs_per_core, (c + 1) * envs_per_core)
        ex = dict(extra)
        if kind == "sphere_incline":
            ex.update(restitution=np.asarray(sample["restitution"])[sl].copy(), friction=np.asarray(sample["friction"])[sl].copy(),
                      dt=sample["dt"], theta=0.7)
        jobs.append((kind, sample["qpos"][sl].copy(), sample["qvel"][sl].copy(), ex, steps))
    for var in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.setdefault(var, "1")
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_warm_reference if use_ref else _warm, range(cores), chunksize=1)
        t0 = time.perf_counter()
        if use_ref:
            res = pool.map(_worker_reference, jobs, chunksize=1)
        elif kind == "sphere_incline":
            res = pool.map(_worker, [(j[1], j[2], j[3]["restitution"], j[3]["friction"], 0.7, j[3]["dt"], steps) for j in jobs], chunksize=1)
        else:
            res = pool.map(_worker_config, jobs, chunksize=1)
        wall = time.perf_counter() - t0
    env_steps = sum(r[0] for r in res)
    what = ("the installed reference's own step function (baseline/_ref, unmodified) under the fake MuJoCo" if use_ref else
            "Python/NumPy port of the reference step under the fake MuJoCo")
    return {"value": env_steps / wall, "unit": "env-substeps/s", "cores": cores, "kind": "reference" if use_ref else "port",
            "sample": f"{n} envs x {steps} steps of the {kind} workload, {what}, {cores} processes, wall {wall:.2f} s",
            "per_core": env_steps / wall / cores, "wall_s": wall}


def _worker(job):
    """Steps ``n_env`` sphere-on-incline envs for ``steps`` steps with the Python port; returns (env_steps, seconds)."""
    import pyport
    qpos, qvel, rest, fric, theta, dt, steps = job
    mj = pyport.fake_mujoco()
    model = mj.MjModel.from_xml_string(pyport.single_body_xml("sphere", [0.2], plane_euler=(theta, 0, 0)))
    t0 = time.perf_counter()
    for i in range(qpos.shape[0]):
        data = mj.MjData(model)
        data.qpos[:] = qpos[i]
        data.qvel[:] = qvel[i]
        for _ in range(steps):
            pyport.step_scheme_a(model, "obj", data, dt=dt, restitution=rest[i], friction_coeff=fric[i], contact_threshold=0.0)
    return qpos.shape[0] * steps, time.perf_counter() - t0


def _worker_config(job):
    """Generic worker: kind in {'cube', 'two_ball', 'multi_sphere'}; returns (env_steps, seconds)."""
    import pyport
    kind, qpos, qvel, extra, steps = job
    mj = pyport.fake_mujoco()
    t0 = time.perf_counter()
    if kind == "cube":
        model = mj.MjModel.from_xml_string(pyport.single_body_xml("box", [0.4, 0.4, 0.4], plane_euler=(extra["theta"], 0, 0)))
        for i in range(qpos.shape[0]):
            data = mj.MjData(model)
            data.qpos[:], data.qvel[:] = qpos[i], qvel[i]
            for _ in range(steps):      # cube_incline.py:75-77: dt, restitution, friction passed; threshold left at 1e-4
                pyport.step_scheme_a(model, "obj", data, dt=0.009, restitution=0.2, friction_coeff=0.6, contact_threshold=1e-4)
    elif kind == "two_ball":
        model = mj.MjModel.from_xml_string(pyport.multi_sphere_xml(2))
        m = float(model.body_mass[1])
        iinv = np.eye(3) / (0.4 * m * 0.01)
        g = np.array([0.0, 0.0, -9.8])
        for i in range(qpos.shape[0]):
            data = mj.MjData(model)
            data.qpos[:], data.qvel[:] = qpos[i], qvel[i]
            for _ in range(steps):
                mj.mj_forward(model, data)                       # ball_collision.py:74 (the reference calls it too)
                pyport.step_two_ball(data, (m, m), (iinv, iinv), g, 0.01, 1.0, 0.3, 0.1)
    else:
        B = extra["n_body"]
        model = mj.MjModel.from_xml_string(pyport.multi_sphere_xml(B))
        for i in range(qpos.shape[0]):
            data = mj.MjData(model)
            data.qpos[:], data.qvel[:] = qpos[i], qvel[i]
            for _ in range(steps):
                pyport.step_multi_sphere(model, data, 0.01, 1.0, extra["friction"])
    return qpos.shape[0] * steps, time.perf_counter() - t0


def python_port_config(kind, sample, cores=None, envs_per_core=4, steps=100, extra=None):
    """Aggregate env-steps/s of the Python port for another BASELINE config (cube / two_ball / multi_sphere)."""
    import multiprocessing as mp
    _require_importable_main()
    cores = cores or os.cpu_count() or 1
    n = cores * envs_per_core
    jobs = [(kind, sample["qpos"][c * envs_per_core:(c + 1) * envs_per_core].copy(),
             sample["qvel"][c * envs_per_core:(c + 1) * envs_per_core].copy(), extra or {}, steps) for c in range(cores)]
    for var in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.setdefault(var, "1")
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_warm, range(cores), chunksize=1)
        t0 = time.perf_counter()
        res = pool.map(_worker_config, jobs, chunksize=1)
        wall = time.perf_counter() - t0
    env_steps = sum(r[0] for r in res)
    return {"value": env_steps / wall, "unit": "env-substeps/s", "cores": cores, "kind": "port",
            "sample": f"{n} envs x {steps} steps, Python/NumPy port under the fake MuJoCo, {cores} processes, wall {wall:.2f} s"}


def _require_importable_main():
    """The pools use the spawn start method: the parent's __main__ must be a real file (not stdin / -c), otherwise every
    worker dies on start-up and the pool respawns them forever."""
    main = sys.modules.get("__main__")
    path = getattr(main, "__file__", None)
    if path is None or not os.path.isfile(path):
        raise RuntimeError("cpu_baseline needs to be driven from a script file (multiprocessing 'spawn')")


def _warm(_):
    """Import everything and run a few steps so that interpreter start-up is outside the timed map."""
    import pyport
    mj = pyport.fake_mujoco()
    model = mj.MjModel.from_xml_string(pyport.single_body_xml("sphere", [0.2]))
    data = mj.MjData(model)
    for _ in range(5):
        pyport.step_scheme_a(model, "obj", data, dt=0.009)
    return 0


def python_port_sphere_incline(sample, cores=None, envs_per_core=16, steps=400):
    """``sample``: dict from synth.sphere_incline (at least cores*envs_per_core envs).  Returns a dict with the
    aggregate env-steps/s over ``cores`` worker processes (wall clock around the parallel map)."""
    import multiprocessing as mp
    _require_importable_main()
    cores = cores or os.cpu_count() or 1
    n = cores * envs_per_core
    if sample["qpos"].shape[0] < n:
        raise ValueError("sample too small")
    jobs = []
    for c in range(cores):
        sl = slice(c * envs_per_core, (c + 1) * envs_per_core)
        jobs.append((sample["qpos"][sl].copy(), sample["qvel"][sl].copy(), sample["restitution"][sl].copy(),
                     sample["friction"][sl].copy(), 0.7, sample["dt"], steps))
    # one BLAS thread per worker process: the reference's arrays are 3-vectors, threads only add contention
    for var in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.setdefault(var, "1")
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_warm, range(cores), chunksize=1)       # start and warm every worker before the clock
        t0 = time.perf_counter()
        res = pool.map(_worker, jobs, chunksize=1)
        wall = time.perf_counter() - t0
    env_steps = sum(r[0] for r in res)
    return {"value": env_steps / wall, "unit": "env-substeps/s", "cores": cores, "kind": "port",
            "sample": f"{n} envs x {steps} steps of the sphere-on-incline workload, Python/NumPy port of the reference step "
                      f"under the fake MuJoCo, {cores} processes, wall {wall:.2f} s",
            "per_core": env_steps / wall / cores, "wall_s": wall}


def c_port_sphere_incline(sample, steps=200, threads=None):
    """The C oracle (gcc -O2, OpenMP) on the same workload: what a compiled CPU implementation achieves."""
    import c_oracle as co
    threads = threads or co.max_threads()
    co.set_threads(threads)
    qp, qv = sample["qpos"].copy(), sample["qvel"].copy()
    E = qp.shape[0]
    I = 0.4 * (50 * 4 / 3 * np.pi * 0.2 ** 3) * 0.04
    kw = dict(geom="sphere", mass=50 * 4 / 3 * np.pi * 0.2 ** 3, inertia=[I] * 3, size=0.2, plane_pos=[0, 0, 0],
              plane_normal=sample["plane_normal"], gravity=[0, 0, -9.8], dt=sample["dt"], restitution=sample["restitution"],
              friction=sample["friction"], threshold=0.0)
    co.step_body_plane(qp, qv, 2, **kw)
    t0 = time.perf_counter()
    co.step_body_plane(qp, qv, steps, **kw)
    wall = time.perf_counter() - t0
    return {"value": E * steps / wall, "unit": "env-substeps/s", "cores": threads, "kind": "port-native",
            "sample": f"{E} envs x {steps} steps, C restatement (gcc -O2 -ffp-contract=off, OpenMP {threads} threads), wall {wall:.2f} s"}
