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
