"""GPU parity tests (run on the B200 box with ``-m gpu``): the CUDA path, called through the Python
host layer and the C ABI, against

  * the golden vectors produced by the unmodified reference (tests/golden/),
  * the C oracle on the same seeded inputs at sizes it finishes in seconds,
  * size-independent properties at BASELINE sizes (substep-fusion equivalence, shard invariance,
    unit quaternions, determinism).

Tolerances (north_star): one step from the same state <= 1e-12 relative per component in fp64 and
<= 1e-5 in fp32 (relative to max(|ref|, 1e-3)); contact-event counts over a horizon must match exactly.
"""
import os
import sys

import numpy as np
import pytest

import c_oracle as co
from helpers import comp_rel_err

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = [0.0, 0.0, -9.8]
F64_STEP, F32_STEP = 1e-12, 1e-5


@pytest.fixture(scope="module")
def rb():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import rigidbody_simulation_b200 as rb_
    rb_._lib.load()                      # raises if the extension is missing: never fall back
    return rb_


def col(rows, key):
    return np.array([r[key] for r in rows], dtype=np.float64)


def tdt(dtype):
    return torch.float64 if dtype == np.float64 else torch.float32


def make_single(rb, geom, size, normal_theta, qpos, qvel, dtype=np.float64, e=None, mu=None, mass=None, inertia=None):
    from rigidbody_simulation_b200 import scenes
    import rigidbody_simulation_b200.mj as mj
    E = qpos.shape[0]
    xml = scenes.single_body_xml(geom, size, plane_euler=(normal_theta, 0, 0))
    model = mj.MjModel.from_xml_string(xml, nenv=E, dtype=tdt(dtype))
    data = mj.MjData(model)
    data.set_state(qpos, qvel)
    return model, data


def state_of(data):
    return (data.qpos.torch().cpu().numpy().astype(np.float64), data.qvel.torch().cpu().numpy().astype(np.float64))


# ------------------------------------------------------------------------------------ free functions
def test_free_functions_golden_and_oracle(rb, golden):
    g = golden("free_functions")
    R = g["random"]
    dev = "cuda"
    T = lambda a: torch.tensor(a, dtype=torch.float64, device=dev)
    jn, jt = rb.compute_collision_impulse_friction(T(col(R, "mass")), None, T(col(R, "v")), T(col(R, "w")), T(col(R, "r")),
                                                   T(col(R, "n")), T(col(R, "e")), T(col(R, "mu")))
    assert comp_rel_err(jn.cpu().numpy(), col(R, "jn"), 1e-6) <= F64_STEP
    assert comp_rel_err(jt.cpu().numpy(), col(R, "jt"), 1e-6) <= F64_STEP
    assert ((jn.cpu().numpy() == 0) == (col(R, "jn") == 0)).all()
    ojn, ojt = co.impulse_friction(col(R, "mass"), col(R, "v"), col(R, "w"), col(R, "r"), col(R, "n"), col(R, "e"), col(R, "mu"))
    assert (jn.cpu().numpy() == ojn).all() and (jt.cpu().numpy() == ojt).all()       # same rounding sequence
    Iw = rb.compute_inertia_tensor_world(T(col(R, "inertia_diag")), T(col(R, "q")))
    assert comp_rel_err(Iw.cpu().numpy(), col(R, "Iw"), 1e-6) <= F64_STEP
    assert (Iw.cpu().numpy() == co.inertia_world(col(R, "inertia_diag"), col(R, "q"))).all()
    v2, w2 = rb.apply_impulse_friction(T(col(R, "v")), T(col(R, "w")), T(col(R, "mass")), T(col(R, "Iw")), T(col(R, "r")),
                                       T(col(R, "n")), T(col(R, "jn")), T(col(R, "jt")))
    assert comp_rel_err(v2.cpu().numpy(), col(R, "v_out"), 1e-6) <= F64_STEP
    assert comp_rel_err(w2.cpu().numpy(), col(R, "w_out"), 1e-6) <= F64_STEP
    v3, w3 = rb.apply_impulse(T(col(R, "v")), T(col(R, "w")), T(col(R, "mass")), T(col(R, "Iw")), T(col(R, "r")),
                              T(col(R, "n")), T(col(R, "impulse")))
    assert comp_rel_err(v3.cpu().numpy(), col(R, "v3"), 1e-6) <= F64_STEP
    assert comp_rel_err(w3.cpu().numpy(), col(R, "w3"), 1e-6) <= F64_STEP
    A = g["A10_random"]
    J = rb.compute_collision_impulse(T(col(A, "mass")), T(col(A, "inv_inertia")), T(col(A, "v")), T(col(A, "w")),
                                     T(col(A, "r")), T(col(A, "n")), T(col(A, "e")), T(col(A, "mu")))
    assert comp_rel_err(J.cpu().numpy(), col(A, "J"), 1e-6) <= F64_STEP


def test_free_functions_reference_calling_convention(rb, golden):
    """Python floats / (3,) NumPy arrays in, ``(float, ndarray(3,))`` out -- the reference's own call shape."""
    g = golden("free_functions")
    Iw = g["inertia"] * np.eye(3)
    for k in g["kats"]:
        v, w, r, n = (np.array(k[x], dtype=float) for x in "vwrn")
        jn, jt = rb.compute_collision_impulse_friction(g["mass"], Iw, v, w, r, n, k["e"], k["mu"])
        assert isinstance(jn, float) and isinstance(jt, np.ndarray) and jt.shape == (3,)
        assert jn == pytest.approx(k["jn"], rel=1e-13, abs=1e-300)
        assert jt == pytest.approx(k["jt"], rel=1e-13, abs=1e-18)
        v2, w2 = rb.apply_impulse_friction(v, w, g["mass"], Iw, r, n, jn, jt)
        assert v2.shape == (3,) and v2 == pytest.approx(k["v_out"], rel=1e-13, abs=1e-18)
        assert w2 == pytest.approx(k["w_out"], rel=1e-13, abs=1e-18)
    a3 = g["A3"]
    v3, w3 = rb.apply_impulse(np.array(a3["v"]), np.array(a3["w"]), g["mass"], Iw, np.array(a3["r"]), np.array(a3["n"]), 0.7)
    assert v3 == pytest.approx(a3["v_out"], rel=1e-14)
    a4 = g["A4"]
    out = rb.compute_inertia_tensor_world(np.array(a4["inertia_diag"]), np.array(a4["q"]))
    assert out.shape == (3, 3) and out == pytest.approx(np.array(a4["out"]), rel=1e-13)
    with pytest.raises(ValueError):
        rb.compute_inertia_tensor_world(np.array([1.0, 2.0, 3.0]), np.zeros(4))      # SciPy raises on a zero quaternion
    k = g["A10_kat"]
    J = rb.compute_collision_impulse(k["mass"], rb.compute_inverse_inertia(k["mass"], 0.1), np.array(k["v"]),
                                     np.array(k["w"]), np.array(k["r"]), np.array(k["n"]), k["e"], k["mu"])
    assert J == pytest.approx(k["J"], rel=1e-13)


def test_free_functions_large_vs_oracle(rb):
    rng = np.random.default_rng(5)
    n = 200_000
    mass, e, mu = rng.uniform(0.1, 30, n), rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    v, w, r = rng.uniform(-2, 2, (n, 3)), rng.uniform(-5, 5, (n, 3)), rng.uniform(-0.5, 0.5, (n, 3))
    nn = rng.normal(size=(n, 3))
    nn /= np.linalg.norm(nn, axis=1, keepdims=True)
    for dtype, tol in ((np.float64, 0.0), (np.float32, 0.0)):
        T = lambda a: torch.tensor(a, dtype=tdt(dtype), device="cuda")
        jn, jt = rb.compute_collision_impulse_friction(T(mass), None, T(v), T(w), T(r), T(nn), T(e), T(mu))
        ojn, ojt = co.impulse_friction(mass, v, w, r, nn, e, mu, dtype=dtype)
        assert np.max(np.abs(jn.cpu().numpy() - ojn)) <= tol and np.max(np.abs(jt.cpu().numpy() - ojt)) <= tol


# ------------------------------------------------------------------------------------ single body + plane
def _step_snapshots(model, data, case, envs, fn, **kw):
    devs, done = {}, 0
    for s in sorted(int(k) for k in envs[0]["snapshots"]) + [case["steps"]]:
        fn(model, "obj", data, dt=case["dt"], substeps=s - done, **kw)
        done = s
        if s == case["steps"]:
            rq, rv = col(envs, "qpos"), col(envs, "qvel")
        else:
            rq = np.array([r["snapshots"][str(s)]["qpos"] for r in envs])
            rv = np.array([r["snapshots"][str(s)]["qvel"] for r in envs])
        qp, qv = state_of(data)
        devs[s] = max(comp_rel_err(qp, rq, 1e-3), comp_rel_err(qv, rv, 1e-3))
    return devs


def test_sphere_incline_golden(rb, golden):
    from rigidbody_simulation_b200.src.physics.collision import custom_step_with_impulse_collision_friction as step
    g = golden("sphere_incline_random")
    envs = g["envs"]
    for strict in (False, True):
        model, data = make_single(rb, "sphere", [g["radius"]], g["theta"], col(envs, "qpos0"), col(envs, "qvel0"))
        assert model.body_mass[-1] == g["mass"] and list(model.body_inertia[-1]) == g["inertia"]
        assert list(model.plane_normal) == g["plane_normal"]
        model.set_per_env(restitution=col(envs, "e"), friction=col(envs, "mu"))
        devs = _step_snapshots(model, data, g, envs, step, restitution=None, friction_coeff=None, contact_threshold=g["thr"])
        assert devs[1] <= F64_STEP, devs
        assert devs[g["steps"]] <= 1e-7, devs
        calls, imps = data.counters()
        assert (calls[:, 0] == col(envs, "calls")).all() and (imps[:, 0] == col(envs, "impulses")).all()


def test_cube_golden(rb, golden):
    from rigidbody_simulation_b200.src.physics.time_integeration import timestep_integration as step
    g = golden("cube_random")
    for kind in ("bounce", "incline"):
        c = g[kind]
        envs = c["envs"]
        model, data = make_single(rb, "box", c["half"], c["theta"], col(envs, "qpos0"), col(envs, "qvel0"))
        assert model.body_mass[-1] == c["mass"] and list(model.body_inertia[-1]) == c["inertia"]
        devs = _step_snapshots(model, data, c, envs, step, restitution=c["e"], friction_coeff=c["mu"])   # default thr 1e-4
        assert devs[1] <= F64_STEP, (kind, devs)
        assert devs[c["steps"]] <= 1e-6, (kind, devs)
        calls, imps = data.counters()
        assert (calls[:, 0] == col(envs, "calls")).all() and (imps[:, 0] == col(envs, "impulses")).all()


def test_general_scheme_and_applied_wrench_golden(rb, golden):
    from rigidbody_simulation_b200.src.physics.collision import custom_step_with_impulse_collision_friction as step_a
    from rigidbody_simulation_b200.src.physics.time_integeration import general
    g = golden("general_and_xfrc")
    for key, fn in (("general", general), ("xfrc", step_a)):
        for r in g[key]:
            model, data = make_single(rb, "sphere" if r["geom"] == "sphere" else "box", r["size"], 0.3,
                                      np.array([r["qpos0"]]), np.array([r["qvel0"]]))
            if "xfrc" in r:
                data.set_xfrc(np.array([r["xfrc"]]))
            devs = _step_snapshots(model, data, r, [r], fn, restitution=r["e"], friction_coeff=r["mu"],
                                   contact_threshold=r["thr"])
            assert devs[1] <= F64_STEP, (key, devs)
            assert devs[r["steps"]] <= 1e-8, (key, devs)
            calls, imps = data.counters()
            assert int(calls[0, 0]) == r["calls"] and int(imps[0, 0]) == r["impulses"]


def test_shipped_single_sphere_and_cube_scripts(rb, golden):
    """configs[0]: the reference scripts as shipped, 1 env, through the mirrored scenario modules."""
    from rigidbody_simulation_b200.src.simulation import cube_incline, single_sphere_bounce
    s = golden("script_single_sphere_2000")
    model, data, log = single_sphere_bounce.run_headless(2000)
    qpos, qvel = np.asarray(data.qpos), np.asarray(data.qvel)
    assert qpos.shape == (7,) and qvel.shape == (6,)
    assert np.max(np.abs(qpos - s["qpos"])) < 1e-9 and np.max(np.abs(qvel - s["qvel"])) < 1e-9
    calls, imps = data.counters()
    assert (int(calls.sum()), int(imps.sum())) == (131, 125)
    z = np.array(log.z_positions)
    assert np.max(np.abs(z[49::50] - np.array(s["z_every_50"]))) < 1e-9
    assert log.times[-1] == pytest.approx(2000 * 0.009)
    c = golden("script_cube_incline_240")
    model, data, log = cube_incline.run_headless(240)
    assert np.max(np.abs(np.asarray(data.qpos) - c["qpos"])) < 1e-9
    assert np.max(np.abs(np.array(log.z_positions) - np.array(c["log_z"]))) < 1e-9
    calls, imps = data.counters()
    assert (int(calls.sum()), int(imps.sum())) == (957, 723)


def _oracle_vs_gpu_single(rb, geom, kind_fn, E, steps, dtype, strict, arith="strict"):
    from rigidbody_simulation_b200 import stepper, synth
    s = kind_fn(E)
    size = [s["radius"]] if geom == "sphere" else s["half"]
    theta = s.get("theta", 0.7)
    model, data = make_single(rb, geom, size, theta, s["qpos"], s["qvel"], dtype=dtype)
    e = s["restitution"] if np.ndim(s["restitution"]) else np.full(E, s["restitution"])
    mu = s["friction"] if np.ndim(s["friction"]) else np.full(E, s["friction"])
    model.set_per_env(restitution=e, friction=mu)
    qp, qv = s["qpos"].astype(dtype), s["qvel"].astype(dtype)
    cnt = (np.zeros(E, np.uint32), np.zeros(E, np.uint32))
    okw = dict(geom="sphere" if geom == "sphere" else "box", mass=model.body_mass[-1], inertia=model.body_inertia[-1],
               size=size if geom == "box" else size[0], plane_pos=[0, 0, 0], plane_normal=model.plane_normal, gravity=G,
               dt=s["dt"], restitution=e, friction=mu, threshold=s["threshold"], counters=cnt)
    out = {}
    done = 0
    for upto in (1, steps):
        co.step_body_plane(qp, qv, upto - done, **okw)
        stepper.step_body_plane(model, data, -1, s["dt"], None, None, s["threshold"], substeps=upto - done,
                                strict_inertia=strict, arith=arith)
        done = upto
        gq, gv = state_of(data)
        out[upto] = (max(comp_rel_err(gq, qp, 1e-3), comp_rel_err(gv, qv, 1e-3)),
                     float(np.mean(gq == qp.astype(np.float64))), float(np.mean(gv == qv.astype(np.float64))))
    calls, imps = data.counters()
    return out, (calls[:, 0] == cnt[0]).all() and (imps[:, 0] == cnt[1]).all(), int(cnt[0].sum())


@pytest.mark.parametrize("dtype,tol", [(np.float64, F64_STEP), (np.float32, F32_STEP)])
def test_sphere_incline_100k_vs_oracle(rb, dtype, tol):
    from rigidbody_simulation_b200 import synth
    for strict in (False, True):
        out, counts_ok, ncalls = _oracle_vs_gpu_single(rb, "sphere", lambda E: synth.sphere_incline(E), 100_000, 200, dtype, strict)
        assert out[1][0] <= tol, out
        assert ncalls > 100_000
        if dtype == np.float64:
            assert counts_ok                 # exact event counts over the horizon
            assert out[200][0] <= 1e-9, out
        if strict:                           # literal inertia path: same rounding sequence as the oracle
            assert out[1][1] == 1.0 and out[1][2] == 1.0, out


@pytest.mark.parametrize("dtype,tol", [(np.float64, F64_STEP), (np.float32, F32_STEP)])
def test_sphere_incline_fast_policy_vs_oracle(rb, dtype, tol):
    """The re-associated ("fast") arithmetic policy meets the same bar: one step <= tolerance, and in fp64 the
    contact-event counts over the horizon are exactly the oracle's."""
    from rigidbody_simulation_b200 import synth
    out, counts_ok, ncalls = _oracle_vs_gpu_single(rb, "sphere", lambda E: synth.sphere_incline(E), 200_000, 300, dtype,
                                                   False, arith="fast")
    assert out[1][0] <= tol, out
    assert ncalls > 200_000
    if dtype == np.float64:
        assert counts_ok
        assert out[300][0] <= 1e-6, out        # measured divergence: 6e-15 @1, 2e-12 @10, 4e-11 @100, 1e-8 @300 steps


def test_fast_policy_golden_and_scope(rb, golden):
    from rigidbody_simulation_b200 import stepper
    from rigidbody_simulation_b200.src.physics.collision import custom_step_with_impulse_collision_friction as step
    g = golden("sphere_incline_random")
    envs = g["envs"]
    model, data = make_single(rb, "sphere", [g["radius"]], g["theta"], col(envs, "qpos0"), col(envs, "qvel0"))
    model.set_per_env(restitution=col(envs, "e"), friction=col(envs, "mu"))
    devs = _step_snapshots(model, data, g, envs, step, restitution=None, friction_coeff=None, contact_threshold=g["thr"],
                           arith="fast")
    assert devs[1] <= F64_STEP and devs[g["steps"]] <= 1e-7, devs
    calls, imps = data.counters()
    assert (calls[:, 0] == col(envs, "calls")).all() and (imps[:, 0] == col(envs, "impulses")).all()
    # shipped single-sphere script (config 1) under the fast policy: same counts, same trajectory to 1e-9
    s = golden("script_single_sphere_2000")
    model, data = make_single(rb, "sphere", [0.2], 0.0, np.array([s["qpos0"]]), np.array([s["qvel0"]]))
    step(model, "obj", data, dt=0.009, restitution=1.0, friction_coeff=0.5, substeps=2000, arith="fast")
    calls, imps = data.counters()
    assert (int(calls.sum()), int(imps.sum())) == (131, 125)
    assert np.max(np.abs(np.asarray(data.qpos) - s["qpos"])) < 1e-9
    # the policy exists for scheme A + isotropic inertia; anything else is refused, not silently strict
    c = golden("cube_random")["bounce"]
    model, data = make_single(rb, "box", c["half"], 0.0, col(c["envs"], "qpos0"), col(c["envs"], "qvel0"))
    with pytest.raises(ValueError):
        stepper.step_body_plane(model, data, -1, 0.009, 0.2, 0.6, 1e-4, arith="fast", scheme=rb._lib.RBS_SCHEME_GENERAL)
    with pytest.raises(ValueError):
        stepper.step_body_plane(model, data, -1, 0.009, 0.2, 0.6, 1e-4, arith="fast", strict_inertia=True)
    model, data = make_single(rb, "box", [0.3, 0.2, 0.1], 0.0, col(c["envs"], "qpos0"), col(c["envs"], "qvel0"))
    with pytest.raises(ValueError):                       # anisotropic box
        stepper.step_body_plane(model, data, -1, 0.009, 0.2, 0.6, 1e-4, arith="fast")


@pytest.mark.parametrize("dtype,tol", [(np.float64, F64_STEP), (np.float32, F32_STEP)])
@pytest.mark.parametrize("kind", ["bounce", "incline"])
@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_cube_50k_vs_oracle(rb, dtype, tol, kind, arith):
    from rigidbody_simulation_b200 import synth
    out, counts_ok, ncalls = _oracle_vs_gpu_single(rb, "box", lambda E: synth.cube(E, kind=kind), 50_000, 120, dtype, False,
                                                   arith=arith)
    assert out[1][0] <= tol, out
    assert ncalls > 50_000
    if dtype == np.float64:
        assert counts_ok and out[120][0] <= (1e-8 if arith == "strict" else 1e-6), out


def test_cube_golden_fast_policy(rb, golden):
    from rigidbody_simulation_b200.src.physics.time_integeration import timestep_integration as step
    g = golden("cube_random")
    for kind in ("bounce", "incline"):
        c = g[kind]
        envs = c["envs"]
        model, data = make_single(rb, "box", c["half"], c["theta"], col(envs, "qpos0"), col(envs, "qvel0"))
        devs = _step_snapshots(model, data, c, envs, step, restitution=c["e"], friction_coeff=c["mu"], arith="fast")
        assert devs[1] <= F64_STEP, (kind, devs)
        assert devs[c["steps"]] <= 1e-5, (kind, devs)
        calls, imps = data.counters()
        assert (calls[:, 0] == col(envs, "calls")).all() and (imps[:, 0] == col(envs, "impulses")).all()


# ------------------------------------------------------------------------------------ two balls
def test_box_compaction_is_bit_identical(rb):
    """The compacting plane-frame box kernel (step_box_plane_pfc_kernel: hit environments queue their contact work in
    shared memory and the first `count` threads of the CTA resolve it) runs, per environment, exactly the statements of
    the thread-per-environment kernel: states and event counters must be the same bits -- bounce and incline (where most
    of a CTA is hit and the in-place path is taken), fp64 and fp32, ragged sizes, with per-environment parameters."""
    from rigidbody_simulation_b200 import scenes, stepper, synth
    dev = torch.device("cuda:0")
    for kind, E, dtype, per_env in (("bounce", 100_003, np.float64, False), ("incline", 65_536 + 77, np.float64, False),
                                    ("bounce", 50_001, np.float32, False), ("bounce", 20_000, np.float64, True), ("incline", 90, np.float64, False)):
        s = synth.cube(E, kind=kind)
        res = {}
        for compact in (0, 1):
            old = rb._lib.set_option("box_compact", compact)
            try:
                model = scenes.cube_on_plane(E, theta=s["theta"], device=dev, dtype=tdt(dtype))
                e, mu = 0.2, 0.6
                if per_env:
                    rng = np.random.default_rng(3)
                    model.set_per_env(restitution=rng.uniform(0.0, 0.6, E), friction=rng.uniform(0.1, 0.9, E))
                    e = mu = None
                data = rb.BatchedData(model)
                data.set_state(s["qpos"], s["qvel"])
                for K in (4, 130, 66):
                    stepper.step_body_plane(model, data, -1, s["dt"], e, mu, 1e-4, substeps=K, count=True, arith="fast")
                res[compact] = state_of(data) + tuple(c.copy() for c in data.counters())
            finally:
                rb._lib.set_option("box_compact", old)
        for a, b in zip(res[0], res[1]):
            assert np.array_equal(a, b), (kind, E, dtype, per_env)
        assert res[1][2].sum() > E // 4                        # the horizon does exercise contacts


def test_strict_compaction_is_bit_identical(rb):
    """The strict single-body stepper with its contact path compacted across the CTA (step_body_plane_compact_kernel:
    every thread parks its state in shared memory, the first `count` threads resolve the queued environments) runs, per
    environment, exactly the statements of the thread-per-environment kernel: states and event counters must be the same
    bits -- sphere and cube, fp64 and fp32, ragged sizes, per-environment parameters -- and, like it, bit for bit the oracle."""
    from rigidbody_simulation_b200 import scenes, stepper, synth
    dev = torch.device("cuda:0")
    cases = (("sphere", 100_003, np.float64), ("bounce", 50_001, np.float64), ("incline", 33_000, np.float64),
             ("sphere", 20_001, np.float32), ("bounce", 20_000, np.float32), ("sphere", 77, np.float64))
    for kind, E, dtype in cases:
        if kind == "sphere":
            s = synth.sphere_incline(E)
            build = lambda: scenes.sphere_on_incline(E, device=dev, dtype=tdt(dtype))
            e, mu, thr = s["restitution"], s["friction"], 0.0
        else:
            s = synth.cube(E, kind=kind)
            build = lambda: scenes.cube_on_plane(E, theta=s["theta"], device=dev, dtype=tdt(dtype))
            rng = np.random.default_rng(4)
            e, mu, thr = rng.uniform(0.0, 0.5, E), rng.uniform(0.2, 0.9, E), 1e-4
        res = {}
        for compact in (-1, 4, 5, 25, 44, 54, 64, 74, 124, 134, 264, 374):   # 2x / 4x: K environments per thread in registers; 5x-7x: resident in shared memory; 1KM: one queue per warp
            old = rb._lib.set_option("strict_compact", compact)
            try:
                model = build()
                model.set_per_env(restitution=e, friction=mu)
                data = rb.BatchedData(model)
                data.set_state(s["qpos"], s["qvel"])
                for K in (1, 90, 37):
                    stepper.step_body_plane(model, data, -1, s["dt"], None, None, thr, substeps=K, count=True, arith="strict")
                res[compact] = state_of(data) + tuple(c.copy() for c in data.counters())
            finally:
                rb._lib.set_option("strict_compact", old)
        for c in (4, 5, 25, 44, 54, 64, 74, 124, 134, 264, 374):
            for a, b in zip(res[-1], res[c]):
                assert np.array_equal(a, b), (kind, E, dtype, c)
        assert res[4][2].sum() > E // 4                        # the horizon does exercise contacts
        if dtype == np.float64 and kind == "sphere" and E > 1000:
            qp, qv = s["qpos"].copy(), s["qvel"].copy()
            cnt = (np.zeros(E, np.uint32), np.zeros(E, np.uint32))
            model = build()
            co.step_body_plane(qp, qv, 128, geom="sphere", mass=model.body_mass[-1], inertia=model.body_inertia[-1], size=0.2,
                               plane_pos=[0, 0, 0], plane_normal=model.plane_normal, gravity=G, dt=s["dt"], restitution=e, friction=mu,
                               threshold=0.0, counters=cnt)
            assert np.array_equal(res[4][0], qp) and np.array_equal(res[4][1], qv)
            assert np.array_equal(res[4][2][:, 0], cnt[0]) and np.array_equal(res[4][3][:, 0], cnt[1])


def test_two_ball_golden_and_script(rb, golden):
    from rigidbody_simulation_b200.src.simulation import ball_collision
    s = golden("script_ball_collision_500")
    model, data, _ = ball_collision.run_headless(500)
    assert np.max(np.abs(np.asarray(data.qpos) - s["qpos"])) < 1e-10
    assert np.max(np.abs(np.asarray(data.qvel) - s["qvel"])) < 1e-10
    g = golden("two_ball_random")
    envs = g["envs"]
    model, data = ball_collision.build(len(envs))
    assert model.body_mass[1] == g["mass"]
    data.set_state(col(envs, "qpos0"), col(envs, "qvel0"))
    done = 0
    for s_ in (1, 10, 100, g["steps"]):
        p1, p2 = ball_collision.step_with_custom_collisions(model, data, g["dt"], substeps=s_ - done)
        done = s_
        rq = col(envs, "qpos") if s_ == g["steps"] else np.array([r["snapshots"][str(s_)]["qpos"] for r in envs])
        rv = col(envs, "qvel") if s_ == g["steps"] else np.array([r["snapshots"][str(s_)]["qvel"] for r in envs])
        qp, qv = state_of(data)
        tol = F64_STEP if s_ == 1 else 1e-8
        assert comp_rel_err(qp, rq, 1e-3) <= tol and comp_rel_err(qv, rv, 1e-3) <= tol, s_
    assert np.allclose(p1.cpu().numpy(), col(envs, "qpos")[:, 0:3], atol=1e-8)


@pytest.mark.parametrize("dtype,tol", [(np.float64, F64_STEP), (np.float32, F32_STEP)])
def test_two_ball_100k_vs_oracle(rb, dtype, tol):
    from rigidbody_simulation_b200 import stepper, synth
    from rigidbody_simulation_b200.src.simulation import ball_collision
    E = 100_000
    s = synth.two_ball(E)
    model, data = ball_collision.build(E, dtype=tdt(dtype))
    data.set_state(s["qpos"], s["qvel"])
    qp, qv = s["qpos"].astype(dtype), s["qvel"].astype(dtype)
    hits = (np.zeros(E, np.uint32), np.zeros(E, np.uint32))
    m = float(model.body_mass[1])
    done = 0
    for upto in (1, 150):
        co.step_two_ball(qp, qv, upto - done, mass=[m, m], radius=0.1, gravity=G, dt=0.01, restitution=1.0, friction=0.3, counters=hits)
        stepper.step_two_ball(model, data, 0.01, 1.0, 0.3, radius=0.1, substeps=upto - done)
        done = upto
        gq, gv = state_of(data)
        err = max(comp_rel_err(gq, qp, 1e-3), comp_rel_err(gv, qv, 1e-3))
        if upto == 1:
            assert err <= tol
        elif dtype == np.float64:
            assert err == 0.0            # identical rounding sequence: bit-for-bit over the horizon
            assert (data.n_contacts[:E].cpu().numpy() == hits[0]).all()
            assert (data.n_impulses[:E].cpu().numpy() == hits[1]).all()
    assert hits[1].sum() > E // 2        # most envs did collide


@pytest.mark.parametrize("dtype,tol", [(np.float64, F64_STEP), (np.float32, F32_STEP)])
def test_two_ball_fast_policy_vs_oracle(rb, dtype, tol, golden):
    from rigidbody_simulation_b200 import stepper, synth
    from rigidbody_simulation_b200.src.simulation import ball_collision
    E = 100_000
    s = synth.two_ball(E)
    model, data = ball_collision.build(E, dtype=tdt(dtype))
    data.set_state(s["qpos"], s["qvel"])
    qp, qv = s["qpos"].astype(dtype), s["qvel"].astype(dtype)
    hits = (np.zeros(E, np.uint32), np.zeros(E, np.uint32))
    m = float(model.body_mass[1])
    done = 0
    floor = 1e-3 if dtype == np.float64 else 1e-2
    for upto in (1, 150):
        co.step_two_ball(qp, qv, upto - done, mass=[m, m], radius=0.1, gravity=G, dt=0.01, restitution=1.0, friction=0.3, counters=hits)
        stepper.step_two_ball(model, data, 0.01, 1.0, 0.3, radius=0.1, substeps=upto - done, arith="fast")
        done = upto
        gq, gv = state_of(data)
        err = max(comp_rel_err(gq, qp, floor), comp_rel_err(gv, qv, floor))
        if upto == 1:
            assert err <= tol, (upto, err)
        elif dtype == np.float64:          # fp32 trajectories part ways at the ball-ball impact (7 significant digits)
            assert err <= 1e-7, (upto, err)
        else:
            assert np.isfinite(gq).all() and np.isfinite(gv).all()
    if dtype == np.float64:
        same = (data.n_contacts[:E].cpu().numpy() == hits[0]) & (data.n_impulses[:E].cpu().numpy() == hits[1])
        assert same.mean() >= 0.9999, same.mean()
        # shipped script under the fast policy: same final state to 1e-9
        g = golden("script_ball_collision_500")
        model, data = ball_collision.build(1)
        ball_collision.step_with_custom_collisions(model, data, 0.01, substeps=500, arith="fast")
        assert np.max(np.abs(np.asarray(data.qpos) - g["qpos"])) < 1e-9


# ------------------------------------------------------------------------------------ multi sphere
def test_multi_sphere_golden(rb, golden):
    from rigidbody_simulation_b200.src.simulation import multi_sphere_bounce as ms
    g = golden("multi_sphere")
    for case in [g["shipped4"]] + g["dense"]:
        B = case["B"]
        model, data = ms.build(1, n_body=B)
        assert model.body_mass[1] == case["mass"]
        data.set_state(np.array([case["qpos0"]]), np.array([case["qvel0"]]))
        done = 0
        for s_ in (1, 10, 100, case["steps"]):
            ms.custom_step_multi_sphere(model, data, case["dt"], case["e"], substeps=s_ - done, friction=case["mu"])
            done = s_
            ref = case if s_ == case["steps"] else case["snapshots"][str(s_)]
            tol = {1: F64_STEP, 10: F64_STEP, 100: 1e-8}.get(s_, 1e-2)       # chaotic: see test_oracle_golden.py
            qp, qv = state_of(data)
            assert comp_rel_err(qp.ravel(), ref["qpos"], 1e-3) <= tol, (B, s_)
            assert comp_rel_err(qv.ravel(), ref["qvel"], 1e-3) <= tol, (B, s_)
        calls, imps = data.counters()
        assert calls[0].tolist() == case["calls"] and imps[0].tolist() == case["impulses"]


@pytest.mark.parametrize("arith,mu", [("strict", 0.3), ("fast", 0.3), ("fast", 0.0), ("strict", 0.0)])
@pytest.mark.parametrize("dtype,tol", [(np.float64, F64_STEP), (np.float32, F32_STEP)])
@pytest.mark.parametrize("B", [64, 27, 5])
def test_multi_sphere_vs_oracle(rb, dtype, tol, B, arith, mu):
    """mu = 0.3 exercises the friction path; mu = 0.0 is the shipped multi_sphere config (sim_overrides.py:22-27), which
    the fast policy runs through its frictionless instantiation (normal impulses only, closed-form orientation)."""
    from rigidbody_simulation_b200 import stepper, synth
    from rigidbody_simulation_b200.src.simulation import multi_sphere_bounce as ms
    E = 1500 if B == 64 else 4000
    s = synth.multi_sphere(E, n_body=B, friction=mu)
    model, data = ms.build(E, n_body=B, dtype=tdt(dtype))
    data.set_state(s["qpos"], s["qvel"])
    qp = s["qpos"].astype(dtype).reshape(E, B, 7).copy()
    qv = s["qvel"].astype(dtype).reshape(E, B, 6).copy()
    cnt = (np.zeros((E, B), np.uint32), np.zeros((E, B), np.uint32))
    okw = dict(mass=model.body_mass[1], inertia=model.body_inertia[1], radius=0.1, plane_pos=[0, 0, 0], plane_normal=[0, 0, 1],
               gravity=G, dt=0.01, restitution=1.0, friction=mu, counters=cnt)
    done = 0
    for upto in (1, 10, 60):
        co.step_multi_sphere(qp, qv, upto - done, **okw)
        stepper.step_multi_sphere(model, data, 0.01, 1.0, mu, substeps=upto - done, arith=arith)
        done = upto
        gq, gv = state_of(data)
        # components that cancel to ~0 carry the absolute rounding error of O(1) intermediates; in fp32 under the
        # re-associated policy the relative measure therefore uses a floor of 1e-2 (positions/velocities are O(1))
        floor = 1e-2 if (dtype == np.float32 and arith == "fast") else 1e-3
        err = max(comp_rel_err(gq.ravel(), qp.ravel(), floor), comp_rel_err(gv.ravel(), qv.ravel(), floor))
        if upto == 1:
            assert err <= tol, (upto, err)
        elif upto == 10:        # collisions amplify rounding differences: 10x per decade for strict, more for fast
            assert err <= tol * (10 if arith == "strict" else 1000), (upto, err)
    if dtype == np.float64:
        calls, imps = data.counters()
        mismatch = (calls != cnt[0]).any(axis=1) | (imps != cnt[1]).any(axis=1)
        if arith == "strict":
            assert int(mismatch.sum()) == 0                 # same rounding sequence as the oracle: exact event counts
            assert np.array_equal(gq, qp.reshape(E, -1)) and np.array_equal(gv, qv.reshape(E, -1))   # bit for bit at step 60
        else:
            assert mismatch.mean() <= 1e-3, mismatch.mean() # re-associated arithmetic on chaotic piles: a vanishing fraction
    assert cnt[0].sum() > E * B                              # contacts are exercised


def test_multi_sphere_tilted_ground_and_mixed_radii_vs_oracle(rb):
    """The plane-frame multi-sphere kernels on a tilted ground (so the frame rotation is not the identity) with
    per-body radii, with and without friction, against the oracle: one step within the bar, a short horizon within
    what collisions amplify, exact event counts after 10 steps; the strict policy stays bit for bit."""
    from rigidbody_simulation_b200 import scenes, stepper, synth
    import rigidbody_simulation_b200.mj as mj
    E, B = 800, 40
    s = synth.multi_sphere(E, n_body=B, friction=0.0)
    theta = 0.35
    normal = [0.0, -np.sin(theta), np.cos(theta)]
    radii = 0.08 + 0.04 * (np.arange(E * B) % B) / B
    for mu in (0.0, 0.4):
        for arith in ("fast", "strict"):
            model = mj.MjModel.from_xml_string(scenes.multi_sphere_xml(B, plane_euler=(theta, 0, 0)), nenv=E)
            data = mj.MjData(model, layout="body")
            m, I = float(model.body_mass[1]), [float(x) for x in model.body_inertia[1]]
            r = torch.tensor(radii, dtype=torch.float64, device="cuda")
            model.set_per_env(radius=r, mass=torch.full_like(r, m), inertia=torch.full((3, E * B), I[0], dtype=torch.float64, device="cuda"))
            data.set_state(s["qpos"], s["qvel"])
            qp, qv = s["qpos"].reshape(E, B, 7).copy(), s["qvel"].reshape(E, B, 6).copy()
            cnt = (np.zeros((E, B), np.uint32), np.zeros((E, B), np.uint32))
            assert np.allclose(model.plane_normal, normal, atol=1e-15)
            okw = dict(mass=m, inertia=I, radius=radii.reshape(E, B), plane_pos=[0, 0, 0], plane_normal=model.plane_normal, gravity=G,
                       dt=0.01, restitution=1.0, friction=mu, counters=cnt)
            done = 0
            for upto in (1, 10, 40):
                co.step_multi_sphere(qp, qv, upto - done, **okw)
                stepper.step_multi_sphere(model, data, 0.01, 1.0, mu, substeps=upto - done, arith=arith)
                done = upto
                gq, gv = state_of(data)
                # Spin components that are analytically zero carry the oracle's own rounding noise: inv(R diag(I) R^T) @ (arm x J)
                # with arm parallel to J is ~1e-17 / I = ~3e-14 (1/I = 1194 for these spheres), where the fast kernels drop the
                # torque-free normal impulse exactly.  The spin is therefore measured against a floor of 0.1 rad/s (|w| is
                # O(1) once there is friction); positions, orientations and linear velocities against 1e-3.
                g6, r6 = gv.reshape(E, B, 6), qv.reshape(E, B, 6)
                err = max(comp_rel_err(gq.ravel(), qp.ravel(), 1e-3), comp_rel_err(g6[:, :, :3].ravel(), r6[:, :, :3].ravel(), 1e-3),
                          comp_rel_err(g6[:, :, 3:].ravel(), r6[:, :, 3:].ravel(), 1e-1))
                ev = np.abs(gv.ravel() - qv.ravel()) / np.maximum(np.abs(qv.ravel()), 1e-3)
                ep = np.abs(gq.ravel() - qp.ravel()) / np.maximum(np.abs(qp.ravel()), 1e-3)
                iv, ip = int(np.argmax(ev)), int(np.argmax(ep))
                where = ("qvel", iv // (6 * B), (iv // 6) % B, iv % 6, gv.ravel()[iv], qv.ravel()[iv], float(ev[iv]),
                         "qpos", ip // (7 * B), (ip // 7) % B, ip % 7, gq.ravel()[ip], qp.ravel()[ip], float(ep[ip]),
                         "counts", int(cnt[0].sum()), int(cnt[1].sum()))
                if arith == "strict":
                    assert err == 0.0, (mu, upto, err, where)
                elif upto == 1:
                    assert err <= F64_STEP, (mu, upto, err, where)
                elif upto == 10:
                    assert err <= 1e-9, (mu, upto, err)
                    calls, imps = data.counters()
                    assert (calls == cnt[0]).all() and (imps == cnt[1]).all(), mu
            assert cnt[0].sum() > E * B // 2 and cnt[1].sum() > 0
            assert np.isfinite(gq).all() and np.abs(np.sqrt((gq.reshape(E, B, 7)[:, :, 3:] ** 2).sum(-1)) - 1).max() < 1e-12


# ------------------------------------------------------------------------------------ properties at full size
def test_substep_fusion_and_shard_invariance_1m(rb):
    """BASELINE size (1,048,576 envs, config 2): K fused substeps == K single-step launches bit-for-bit;
    stepping two half shards == stepping the whole; quaternions stay unit; the run is deterministic."""
    from rigidbody_simulation_b200 import stepper, synth
    E = 1 << 20
    s = synth.sphere_incline(E)

    def run(lo, hi, schedule):
        model, data = make_single(rb, "sphere", [0.2], 0.7, s["qpos"][lo:hi], s["qvel"][lo:hi])
        model.set_per_env(restitution=s["restitution"][lo:hi], friction=s["friction"][lo:hi])
        for k in schedule:
            stepper.step_body_plane(model, data, -1, s["dt"], None, None, 0.0, substeps=k)
        return data

    whole = run(0, E, [64])
    ref = whole.state.clone()
    assert torch.equal(run(0, E, [1] * 64).state, ref)
    assert torch.equal(run(0, E, [64]).state, ref)
    half_a, half_b = run(0, E // 2, [32, 32]), run(E // 2, E, [16, 48])
    assert torch.equal(torch.cat([half_a.state, half_b.state], dim=2), ref)
    qn = (ref[3:7, 0, :] ** 2).sum(dim=0).sqrt()
    assert float((qn - 1).abs().max()) < 1e-14
    assert torch.isfinite(ref).all()
    calls, imps = whole.counters()
    assert calls.sum() > E // 2 and (imps <= calls).all()


def test_shard_generator_is_index_keyed():
    from rigidbody_simulation_b200 import synth
    a = synth.sphere_incline(1000, start=0)
    b = synth.sphere_incline(300, start=500)
    assert (a["qpos"][500:800] == b["qpos"]).all() and (a["restitution"][500:800] == b["restitution"]).all()


# ------------------------------------------------------------------------------------ boundary behaviour
def test_host_buffer_driver_matches_device_path(rb):
    from rigidbody_simulation_b200 import stepper, synth
    E = 20_000
    s = synth.sphere_incline(E)
    model, data = make_single(rb, "sphere", [0.2], 0.7, s["qpos"], s["qvel"])
    model.set_per_env(restitution=s["restitution"], friction=s["friction"])
    stepper.step_body_plane(model, data, -1, s["dt"], None, None, 0.0, substeps=100)
    qp_h = torch.from_numpy(s["qpos"].copy()).pin_memory()
    qv_h = torch.from_numpy(s["qvel"].copy()).pin_memory()
    stepper.run_body_plane_host(model, qp_h, qv_h, 100, dt=s["dt"], restitution=None, friction_coeff=None, substeps=32)
    gq, gv = state_of(data)
    assert (qp_h.numpy() == gq).all() and (qv_h.numpy() == gv).all()


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_host_buffer_driver_pipelined_chunks(rb, arith):
    """Several pipeline chunks (ragged last one), per-env parameters, a step count that is not a multiple of the
    fuse factor: the host-buffer call returns exactly what the device-resident path computes."""
    from rigidbody_simulation_b200 import stepper, synth
    E = 400_003
    s = synth.sphere_incline(E)
    model, data = make_single(rb, "sphere", [0.2], 0.7, s["qpos"], s["qvel"])
    model.set_per_env(restitution=s["restitution"], friction=s["friction"])
    for k in (32, 32, 11):            # the host driver issues the same launches (the fast policy's plane-frame kernel
        stepper.step_body_plane(model, data, -1, s["dt"], None, None, 0.0, substeps=k, arith=arith)   # rounds per launch)
    qp_h, qv_h = s["qpos"].copy(), s["qvel"].copy()            # pageable NumPy buffers work too
    stepper.run_body_plane_host(model, qp_h, qv_h, 75, dt=s["dt"], restitution=None, friction_coeff=None, substeps=32, arith=arith)
    gq, gv = state_of(data)
    assert (qp_h == gq).all() and (qv_h == gv).all()


def test_pack_unpack_roundtrip_and_empty(rb):
    import ctypes
    lib = rb._lib.load()
    for body_fastest, B in ((0, 1), (0, 2), (1, 7)):
        E = 1000
        qp = torch.randn(E, 7 * B, dtype=torch.float64, device="cuda")
        qv = torch.randn(E, 6 * B, dtype=torch.float64, device="cuda")
        st = torch.empty(13 * E * B, dtype=torch.float64, device="cuda")
        stride = E * B if body_fastest else E
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        rb._lib.check(lib.rbs_pack_state(1, E, B, body_fastest, P(qp), P(qv), P(st), stride, None))
        qp2, qv2 = torch.empty_like(qp), torch.empty_like(qv)
        rb._lib.check(lib.rbs_unpack_state(1, E, B, body_fastest, P(st), stride, P(qp2), P(qv2), None))
        torch.cuda.synchronize()
        assert torch.equal(qp, qp2) and torch.equal(qv, qv2)
    # n_env == 0 is a successful no-op
    a = rb._lib.BodyPlaneArgs()
    a.dtype, a.substeps, a.n_env = 1, 1, 0
    assert lib.rbs_step_body_plane(ctypes.byref(a)) == 0


def test_error_behaviour(rb):
    import ctypes
    lib = rb._lib.load()
    a = rb._lib.BodyPlaneArgs()
    a.dtype, a.substeps, a.n_env, a.stride = 7, 1, 4, 4
    assert lib.rbs_step_body_plane(ctypes.byref(a)) == rb._lib.RBS_EINVAL
    assert b"dtype" in lib.rbs_last_error()
    with pytest.raises(ValueError):
        rb._lib.check(rb._lib.RBS_EINVAL)
    a.dtype, a.substeps = 1, 0
    assert lib.rbs_step_body_plane(ctypes.byref(a)) == rb._lib.RBS_EINVAL
    a.substeps, a.stride = 1, 2
    a.state = 1
    assert lib.rbs_step_body_plane(ctypes.byref(a)) == rb._lib.RBS_EINVAL and b"stride" in lib.rbs_last_error()
    with pytest.raises(ValueError):
        rb.compute_collision_impulse_friction(1.0, None, np.zeros(4), np.zeros(3), np.zeros(3), np.zeros(3), 1.0, 0.5)


# ------------------------------------------------------------------------------------ streams, graphs, edge cases
def test_cuda_graph_capture_and_side_stream(rb):
    """Every entry point only enqueues on the caller's stream, so a loop of one-substep launches (the reference's
    per-frame call) can be captured in a CUDA graph and replayed; results equal the eager launches."""
    from rigidbody_simulation_b200 import stepper, synth
    E = 50_000
    s = synth.sphere_incline(E)

    def fresh():
        model, data = make_single(rb, "sphere", [0.2], 0.7, s["qpos"], s["qvel"])
        model.set_per_env(restitution=s["restitution"], friction=s["friction"])
        return model, data

    model, data = fresh()
    for _ in range(48):
        stepper.step_body_plane(model, data, -1, s["dt"], None, None, 0.0, substeps=1)
    eager = data.state.clone()

    model, data = fresh()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        stepper.step_body_plane(model, data, -1, s["dt"], None, None, 0.0, substeps=1)     # warm-up outside capture
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for _ in range(16):
                stepper.step_body_plane(model, data, -1, s["dt"], None, None, 0.0, substeps=1)
        data.set_state(s["qpos"], s["qvel"])
        data.n_contacts.zero_()
        data.n_impulses.zero_()
        for _ in range(3):
            graph.replay()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    assert torch.equal(data.state, eager)


def test_nan_and_degenerate_inputs_propagate_like_the_reference(rb):
    """The reference filters only isnan(dist) (collision.py:74): a NaN env stays NaN and does not disturb its
    neighbours; a sphere exactly touching (dist == 0) is not a contact (dist < 0 is required)."""
    from rigidbody_simulation_b200 import stepper
    E = 256
    qpos = np.tile(np.array([0, 0, 1.0, 1, 0, 0, 0.0]), (E, 1))
    qvel = np.zeros((E, 6))
    qpos[7, 2] = np.nan
    qpos[9, 2] = 0.2                      # exactly touching the flat plane, at rest
    model, data = make_single(rb, "sphere", [0.2], 0.0, qpos, qvel)
    stepper.step_body_plane(model, data, -1, 0.009, 1.0, 0.5, 0.0, substeps=1)
    qp, qv = state_of(data)
    assert np.isnan(qp[7, 2]) and np.isfinite(np.delete(qp, 7, axis=0)).all()
    calls, _ = data.counters()
    assert calls.sum() == 0               # nobody is in contact on the first step, including env 9
    ref_p, ref_v = qpos[:1].copy(), qvel[:1].copy()
    co.step_body_plane(ref_p, ref_v, 1, geom="sphere", mass=model.body_mass[-1], inertia=model.body_inertia[-1], size=0.2,
                       plane_pos=[0, 0, 0], plane_normal=[0, 0, 1], gravity=G, dt=0.009, restitution=1.0, friction=0.5, threshold=0.0)
    assert (qp[0] == ref_p[0]).all() and (qv[0] == ref_v[0]).all()


def test_multi_sphere_ragged_and_maximum_body_counts(rb):
    """B that does not divide the CTA (ragged lanes), B = 1 (no partner) and the ABI maximum B = 1024."""
    from rigidbody_simulation_b200 import stepper, synth
    from rigidbody_simulation_b200.src.simulation import multi_sphere_bounce as ms
    for B, E, steps in ((1, 300, 40), (3, 1001, 40), (100, 37, 25), (1024, 2, 3)):
        s = synth.multi_sphere(E, n_body=B, friction=0.2)
        model, data = ms.build(E, n_body=B)
        data.set_state(s["qpos"], s["qvel"])
        qp = s["qpos"].reshape(E, B, 7).copy()
        qv = s["qvel"].reshape(E, B, 6).copy()
        cnt = (np.zeros((E, B), np.uint32), np.zeros((E, B), np.uint32))
        co.step_multi_sphere(qp, qv, steps, mass=model.body_mass[1], inertia=model.body_inertia[1], radius=0.1,
                             plane_pos=[0, 0, 0], plane_normal=[0, 0, 1], gravity=G, dt=0.01, restitution=1.0, friction=0.2, counters=cnt)
        stepper.step_multi_sphere(model, data, 0.01, 1.0, 0.2, substeps=steps)
        gq, gv = state_of(data)
        assert comp_rel_err(gq.ravel(), qp.ravel(), 1e-3) <= 1e-9, B
        calls, imps = data.counters()
        assert (calls == cnt[0]).all() and (imps == cnt[1]).all(), B


def test_in_kernel_trajectory_equals_per_frame_logging(rb, tmp_path, golden):
    """A fused launch with ``trajectory=`` writes, for the sampled environments, exactly the positions the per-frame
    loop (one launch and one logger.record per step, mujoco_viewer.py:113-119) logs: config 1 (2000 steps, the
    height-vs-time curve of data/plots/single_sphere/height_vs_time.png) and the cube on the incline, strict policy
    bit for bit; fast policy and a batch with a window, within the per-step bar.  Also the CLI's --log."""
    import json
    from rigidbody_simulation_b200 import headless, stepper, synth
    from rigidbody_simulation_b200.src import simulate
    from rigidbody_simulation_b200.src.simulation import cube_incline, single_sphere_bounce
    for mod, steps in ((single_sphere_bounce, 2000), (cube_incline, 240)):
        _, _, per_frame = mod.run_headless(steps, nenv=3)
        _, d_fused, fused = mod.run_headless(steps, nenv=3, substeps_per_launch=173)
        a, b = per_frame.finish(), fused.finish()
        assert a.shape == b.shape == (steps, 3, 3) and np.array_equal(a, b)
        assert np.allclose(per_frame.times, fused.times, rtol=0, atol=1e-12)
        assert fused.z_positions == per_frame.z_positions
    # known answer of SURVEY section 4: first rebound peak of the shipped sphere scene, 1.4776 at t = 1.116 s
    _, _, lg = single_sphere_bounce.run_headless(2000, nenv=1, substeps_per_launch=500)
    lg.finish()
    z, t = np.array(lg.z_positions), np.array(lg.times)
    k = 60 + int(np.argmax(z[60:200]))
    assert z[k] == pytest.approx(1.4776, abs=2e-4) and t[k] == pytest.approx(1.116, abs=0.01)
    # a batch, fast policy, sampled window of 5 envs out of 4096: equals stepping one substep at a time
    E, K = 4096, 37
    s = synth.sphere_incline(E)
    for arith, tol in (("strict", 0.0), ("fast", 1e-12)):
        model, data = make_single(rb, "sphere", [0.2], 0.7, s["qpos"], s["qvel"])
        model.set_per_env(restitution=s["restitution"], friction=s["friction"])
        traj = torch.full((K, 5, 3), float("nan"), dtype=torch.float64, device="cuda")
        stepper.step_body_plane(model, data, -1, s["dt"], None, None, 0.0, substeps=K, arith=arith, trajectory=traj)
        model2, data2 = make_single(rb, "sphere", [0.2], 0.7, s["qpos"], s["qvel"])
        model2.set_per_env(restitution=s["restitution"], friction=s["friction"])
        want = []
        for _ in range(K):
            stepper.step_body_plane(model2, data2, -1, s["dt"], None, None, 0.0, substeps=1, arith=arith)
            want.append(data2.qpos.torch()[:5, :3].clone())
        want = torch.stack(want)
        if tol == 0.0:
            assert torch.equal(traj, want)
        else:
            assert float(((traj - want).abs() / want.abs().clamp_min(1e-3)).max()) <= tol * K
        assert comp_rel_err(data.qpos.torch().cpu().numpy(), data2.qpos.torch().cpu().numpy(), 1e-3) <= tol * K
    with pytest.raises(ValueError):
        stepper.step_body_plane(model, data, -1, s["dt"], None, None, 0.0, substeps=K + 1, trajectory=traj)
    out = tmp_path / "traj.npz"
    simulate.main(["--sim", "single_sphere", "--headless", "--steps", "300", "--substeps-per-launch", "64", "--log", str(out)])
    z = np.load(out)
    assert z["positions"].shape == (300, 1, 3) and z["times"].shape == (300,) and np.isfinite(z["positions"]).all()


def test_reset_envs_kernel(rb):
    """reset(env_mask) = mj_resetData on a subset (mujoco_viewer.py:62-65): masked environments go back to qpos0 with zero
    velocity and zero event counters, the others are untouched bit for bit; both state layouts, fp64 and fp32,
    host and device masks, and the unmasked form."""
    from rigidbody_simulation_b200 import scenes, stepper
    for layout, name, dtype in (("env", "sphere", torch.float64), ("env", "ball_collision", torch.float32),
                                ("body", "multi_sphere", torch.float64)):
        E = 1003
        m = rb.BatchedModel.from_xml_path(scenes.model_path(name), nenv=E, device="cuda", dtype=dtype)
        d = rb.BatchedData(m, layout=layout)
        B = m.nfree
        rng = np.random.default_rng(5)
        qp, qv = rng.normal(size=(E, 7 * B)), rng.normal(size=(E, 6 * B))
        d.set_state(qp, qv)
        d.n_contacts.fill_(7)
        d.n_impulses.fill_(3)
        before_p, before_v = d.qpos.torch().clone(), d.qvel.torch().clone()
        mask = rng.uniform(size=E) < 0.3
        d.reset(mask if layout == "env" else torch.as_tensor(mask, device="cuda"))
        q, v = d.qpos.torch(), d.qvel.torch()
        sel = torch.as_tensor(mask, device="cuda")
        q0 = torch.as_tensor(np.asarray(m.qpos0), dtype=dtype, device="cuda")
        assert torch.equal(q[sel], q0.expand(int(sel.sum()), -1)) and (v[sel] == 0).all()
        assert torch.equal(q[~sel], before_p[~sel]) and torch.equal(v[~sel], before_v[~sel])
        if layout == "body":
            nc = d.n_contacts.view(E, B)
            assert (nc[sel] == 0).all() and (nc[~sel] == 7).all() and (d.n_impulses.view(E, B)[sel] == 0).all()
        elif B == 1:
            assert (d.n_contacts[sel] == 0).all() and (d.n_contacts[~sel] == 7).all()
        d.reset()
        assert torch.equal(d.qpos.torch(), q0.expand(E, -1)) and (d.qvel.torch() == 0).all()
        assert (d.n_contacts == 0).all() and (d.n_impulses == 0).all()
        with pytest.raises(ValueError):
            d.reset(np.ones(E + 1, bool))


def test_plane_frame_box_kernel_arbitrary_plane(rb):
    """The cube's fused fast launches also work in the plane frame: planes tilted about two axes and not through the
    origin, cubes dropped from just above the plane with random orientation and spin, threshold 1e-4.  Four fused
    substeps stay within the per-step bar of the oracle; over 150 steps the event counts are exact."""
    from rigidbody_simulation_b200 import scenes, stepper
    import rigidbody_simulation_b200.mj as mj
    rng = np.random.default_rng(33)
    E = 20_000
    half = [0.3, 0.3, 0.3]
    for euler, ppos in (((0.3, -0.4, 0.0), (0.2, -0.1, 0.05)), ((-0.8, 0.25, 0.0), (0.0, 0.0, -0.3)), ((0.0, 0.0, 0.0), (0.0, 0.0, 0.0))):
        xml = scenes.single_body_xml("box", half, plane_euler=euler)
        xml = xml.replace('<geom name="ground" type="plane"', f'<geom name="ground" pos="{ppos[0]} {ppos[1]} {ppos[2]}" type="plane"')
        model = mj.MjModel.from_xml_string(xml, nenv=E)
        n, pp = np.array(model.plane_normal), np.array(model.plane_point)
        h = rng.uniform(0.25, 0.9, E)                                   # some start penetrating, most just above
        tang = rng.normal(size=(E, 3))
        pos = pp + h[:, None] * n + (tang - (tang @ n)[:, None] * n) * 0.5
        q = rng.normal(size=(E, 4))
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        qpos = np.concatenate([pos, q], axis=1)
        qvel = np.concatenate([rng.uniform(-1, 1, (E, 3)), rng.uniform(-3, 3, (E, 3))], axis=1)
        data = mj.MjData(model)
        data.set_state(qpos, qvel)
        qp, qv = qpos.copy(), qvel.copy()
        cnt = (np.zeros(E, np.uint32), np.zeros(E, np.uint32))
        kw = dict(geom="box", mass=model.body_mass[-1], inertia=model.body_inertia[-1], size=half, plane_pos=pp, plane_normal=n,
                  gravity=G, dt=0.009, restitution=0.2, friction=0.6, threshold=1e-4, counters=cnt)
        done = 0
        for upto in (4, 150):
            co.step_body_plane(qp, qv, upto - done, **kw)
            stepper.step_body_plane(model, data, -1, 0.009, 0.2, 0.6, 1e-4, substeps=upto - done, arith="fast")
            done = upto
            gq, gv = state_of(data)
            err = max(comp_rel_err(gq, qp, 1e-3), comp_rel_err(gv, qv, 1e-3))
            if upto == 4:
                assert err <= 4e-12, (euler, err)
        calls, imps = data.counters()
        mismatch = (calls[:, 0] != cnt[0]) | (imps[:, 0] != cnt[1])
        print("plane", euler, "envs with a different event count after 150 steps:", int(mismatch.sum()), "of", E)
        # cubes that come to rest sit at dist ~ -thr, where a last-bit difference decides whether a contact is skipped
        # (:79-80); a re-associated policy cannot promise every such decision.  Measured: a few envs in 10^4.
        assert mismatch.mean() <= 2e-3, (euler, mismatch.mean())
        assert cnt[0].sum() > E


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_multi_sphere_partner_lists_never_change_results(rb, arith):
    """The Verlet partner lists (PartnerLists in rbs_kernels.cuh) are a conservative superset of the contacts: every
    skin -- none (all partners scanned every substep), 5 %, the default 50 %, 400 % -- gives bit-identical states and
    event counts, for one-word (B = 64) and multi-word (B = 150) lists, uniform and per-body radii, fp64 and fp32."""
    from rigidbody_simulation_b200 import stepper, synth
    from rigidbody_simulation_b200.src.simulation import multi_sphere_bounce as ms
    for B, E, dtype, per_body in ((64, 600, np.float64, False), (150, 40, np.float64, False), (64, 300, np.float32, False),
                                  (27, 500, np.float64, True)):
        s = synth.multi_sphere(E, n_body=B, friction=0.3)
        results = []
        for skin in (-1, 5, 0, 400):
            model, data = ms.build(E, n_body=B, dtype=tdt(dtype))
            if per_body:        # radii 0.08 .. 0.12 by body index; mass / inertia stay those of the XML
                r = torch.tensor(0.08 + 0.04 * (np.arange(E * B) % B) / B, dtype=tdt(dtype), device="cuda")
                model.set_per_env(radius=r, mass=torch.full_like(r, float(model.body_mass[1])),
                                  inertia=torch.full((3, E * B), float(model.body_inertia[1][0]), dtype=tdt(dtype), device="cuda"))
            data.set_state(s["qpos"], s["qvel"])
            for k in (1, 7, 80):
                stepper.step_multi_sphere(model, data, 0.01, 1.0, 0.3, substeps=k, arith=arith, list_skin_percent=skin)
            calls, imps = data.counters()
            results.append((data.state.clone(), calls.copy(), imps.copy()))
        for st, calls, imps in results[1:]:
            assert torch.equal(st, results[0][0]), (B, dtype, per_body)
            assert (calls == results[0][1]).all() and (imps == results[0][2]).all()
        assert results[0][1].sum() > E * B          # pair and ground contacts are exercised
        assert torch.isfinite(results[0][0]).all()


def test_padded_stride_and_applied_wrench_per_env(rb):
    """stride > n_env (a window into a larger allocation) and a per-env applied wrench through the raw C ABI."""
    import ctypes
    from rigidbody_simulation_b200 import stepper, synth
    E, pad = 5000, 5120
    s = synth.sphere_incline(E)
    model, data = make_single(rb, "sphere", [0.2], 0.7, s["qpos"], s["qvel"])
    rng = np.random.default_rng(3)
    xf = np.concatenate([rng.uniform(-2, 2, (E, 3)), rng.uniform(-0.05, 0.05, (E, 3))], axis=1)
    data.set_xfrc(xf)
    big = torch.full((13, pad), float("nan"), dtype=torch.float64, device="cuda")
    big[:, :E] = data.state[:, 0, :]
    a = stepper.body_plane_args(model, data, -1, s["dt"], 0.8, 0.4, 0.0, rb._lib.RBS_SCHEME_A, 20)
    a.state, a.stride = ctypes.c_void_p(big.data_ptr()), pad
    a.stream = stepper.current_stream(model.device)
    rb._lib.check(rb._lib.load().rbs_step_body_plane(ctypes.byref(a)))
    qp, qv = s["qpos"].copy(), s["qvel"].copy()
    co.step_body_plane(qp, qv, 20, geom="sphere", mass=model.body_mass[-1], inertia=model.body_inertia[-1], size=0.2,
                       plane_pos=[0, 0, 0], plane_normal=model.plane_normal, gravity=G, dt=s["dt"], restitution=0.8,
                       friction=0.4, threshold=0.0, xfrc=xf)
    got = big[:, :E].cpu().numpy()
    assert np.isnan(big[:, E:].cpu().numpy()).all()                       # the padding is never touched
    assert comp_rel_err(got[:7].T, qp, 1e-3) <= 1e-11 and comp_rel_err(got[7:].T, qv, 1e-3) <= 1e-11


def test_headless_loop_batched_logger_and_cli(rb, capsys):
    """start_main_loop's contract with a batch: the device-side TrajectoryLog of env 0 equals the one-env run, and
    the headless CLI reports the same final state as the scenario module."""
    import json
    from rigidbody_simulation_b200.src import simulate
    from rigidbody_simulation_b200.src.simulation import single_sphere_bounce
    _, d1, log1 = single_sphere_bounce.run_headless(steps=150, nenv=1)
    _, d8, log8 = single_sphere_bounce.run_headless(steps=150, nenv=8)
    assert log8.buf.shape == (150, 4, 3)
    assert np.array_equal(np.array(log1.z_positions), np.array(log8.z_positions))
    assert np.array_equal(np.asarray(d1.qpos), d8.qpos.torch().cpu().numpy()[5])
    out = simulate.run_simulation("single_sphere", steps=150, envs=3)
    assert out["qpos_env0"] == np.asarray(d1.qpos).tolist()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["sim"] == "single_sphere" and line["envs"] == 3
    with pytest.raises(SystemExit):
        simulate.run_simulation("compare_builtin")


def test_env_windows_and_split_chains(rb):
    """A call restricted to a window of the batch (env_range) touches only that window and computes what the whole-batch
    call computes there, including per-env parameter rows ([3][E] sizes) and counters; SplitChains == one chain."""
    from rigidbody_simulation_b200 import stepper, synth
    E = 70_001
    s = synth.cube(E, kind="bounce")
    rng = np.random.default_rng(11)
    size = (0.4 + rng.uniform(-0.05, 0.05, (1, E))).repeat(3, axis=0)         # per-env cube half sizes, [3, E]

    def fresh():
        model, data = make_single(rb, "box", s["half"], 0.0, s["qpos"], s["qvel"])
        model.set_per_env(size=size)
        return model, data

    kw = dict(dt=0.009, restitution=0.2, friction_coeff=0.6, contact_threshold=1e-4, strict_inertia=False)
    model, whole = fresh()
    for _ in range(3):
        stepper.step_body_plane(model, whole, -1, substeps=40, **kw)
    model, data = fresh()
    before = data.state.clone()
    stepper.step_body_plane(model, data, -1, substeps=40, env_range=(1000, 5000), **kw)
    assert torch.equal(data.state[:, :, :1000], before[:, :, :1000]) and torch.equal(data.state[:, :, 6000:], before[:, :, 6000:])
    assert not torch.equal(data.state[:, :, 1000:6000], before[:, :, 1000:6000])
    model, data = fresh()
    chains = stepper.SplitChains(model, data, parts=3)
    assert sum(c for _, c in chains.ranges) == E
    chains.fork()
    for _ in range(3):
        chains.step(substeps=40, **kw)
    chains.join()
    torch.cuda.synchronize()
    assert torch.equal(data.state, whole.state)
    assert torch.equal(data.n_contacts, whole.n_contacts) and int(whole.n_contacts.sum()) > 0
    with pytest.raises(ValueError):
        stepper.step_body_plane(model, data, -1, substeps=1, env_range=(E - 10, 20), **kw)


def test_run_statistics_kernel(rb):
    """rbs_stats (energy / max height / event totals in one pass) against a NumPy evaluation of the same sums."""
    from rigidbody_simulation_b200 import shard, stepper, synth
    from rigidbody_simulation_b200.src.simulation import multi_sphere_bounce as ms
    E = 30_011
    s = synth.cube(E, kind="bounce")
    model, data = make_single(rb, "box", [0.3, 0.2, 0.1], 0.0, s["qpos"], s["qvel"])     # anisotropic: exercises R^T w
    stepper.step_body_plane(model, data, -1, 0.009, 0.2, 0.6, 1e-4, substeps=150)
    got = shard.local_stats(model, data).cpu().numpy()
    qp, qv = state_of(data)
    m, I = model.body_mass[-1], model.body_inertia[-1]
    from scipy.spatial.transform import Rotation
    R = Rotation.from_quat(qp[:, [4, 5, 6, 3]]).as_matrix()
    wb = np.einsum("eji,ej->ei", R, qv[:, 3:6])
    ke = 0.5 * m * (qv[:, :3] ** 2).sum() + 0.5 * (wb ** 2 * I).sum()
    pe = m * 9.8 * qp[:, 2].sum()
    calls, imps = data.counters()
    assert got[1] == calls.sum() and got[2] == imps.sum() and calls.sum() > 0
    assert got[3] == pytest.approx(ke + pe, rel=1e-10)
    assert got[4] == pytest.approx(qp[:, 2].max(), rel=1e-14)
    # thread-per-body layout
    s = synth.multi_sphere(500, n_body=27)
    model, data = ms.build(500, n_body=27)
    data.set_state(s["qpos"], s["qvel"])
    stepper.step_multi_sphere(model, data, 0.01, 1.0, 0.3, substeps=30)
    got = shard.local_stats(model, data).cpu().numpy()
    qp, qv = state_of(data)
    qp, qv = qp.reshape(-1, 7), qv.reshape(-1, 6)
    m, I = model.body_mass[1], model.body_inertia[1][0]
    ref = 0.5 * m * (qv[:, :3] ** 2).sum() + 0.5 * I * (qv[:, 3:] ** 2).sum() + m * 9.8 * qp[:, 2].sum()
    assert got[3] == pytest.approx(ref, rel=1e-10) and got[4] == pytest.approx(qp[:, 2].max(), rel=1e-14)
    assert got[1] == data.counters()[0].sum()


def test_host_buffer_drivers_two_ball_and_multi_sphere(rb):
    """rbs_run_two_ball_host / rbs_run_multi_sphere_host: reference-layout host arrays in and out == device path."""
    from rigidbody_simulation_b200 import stepper, synth
    from rigidbody_simulation_b200.src.simulation import ball_collision, multi_sphere_bounce as ms
    E = 30_000
    s = synth.two_ball(E)
    model, data = ball_collision.build(E)
    data.set_state(s["qpos"], s["qvel"])
    stepper.step_two_ball(model, data, 0.01, 1.0, 0.3, radius=0.1, substeps=130)
    qp_h, qv_h = s["qpos"].copy(), s["qvel"].copy()
    stepper.run_two_ball_host(model, qp_h, qv_h, 130, dt=0.01, restitution=1.0, friction=0.3, radius=0.1, substeps=50)
    gq, gv = state_of(data)
    assert (qp_h == gq).all() and (qv_h == gv).all()
    E, B = 700, 27
    s = synth.multi_sphere(E, n_body=B, friction=0.2)
    model, data = ms.build(E, n_body=B)
    data.set_state(s["qpos"], s["qvel"])
    stepper.step_multi_sphere(model, data, 0.01, 1.0, 0.2, substeps=45)
    qp_h, qv_h = s["qpos"].copy(), s["qvel"].copy()
    stepper.run_multi_sphere_host(model, qp_h, qv_h, 45, dt=0.01, restitution=1.0, friction=0.2, substeps=8)
    gq, gv = state_of(data)
    assert (qp_h == gq).all() and (qv_h == gv).all()
    with pytest.raises(ValueError):                       # wrong shape is refused before anything is launched
        stepper.run_multi_sphere_host(model, qp_h[:, :-1].copy(), qv_h, 1)


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_host_buffer_drivers_pipelined_chunks_two_ball_and_multi_sphere(rb, arith):
    """The two-ball and multi-sphere host-buffer drivers run the same chunk pipeline as the single-body one (whole waves
    per chunk, copies overlapped with stepping): sizes that need several chunks with a ragged last one, per-environment
    masses / radii (so the windowed parameter pointers and row strides are exercised), a step count that is not a multiple
    of the fuse factor -- the call must return exactly what the device-resident path computes with the same launches."""
    from rigidbody_simulation_b200 import stepper, synth
    from rigidbody_simulation_b200.src.simulation import ball_collision, multi_sphere_bounce as ms
    old = rb._lib.set_option("host_chunks", 7)
    try:
        E = 500_003
        s = synth.two_ball(E)
        model, data = ball_collision.build(E)
        rng = np.random.default_rng(5)
        model.set_per_env(mass=torch.tensor(rng.uniform(0.15, 0.3, (2, E)), dtype=torch.float64, device="cuda"),
                          radius=torch.tensor(rng.uniform(0.09, 0.11, E), dtype=torch.float64, device="cuda"))
        data.set_state(s["qpos"], s["qvel"])
        for k in (60, 60, 31):
            stepper.step_two_ball(model, data, 0.01, 1.0, 0.3, radius=0.1, substeps=k, arith=arith)
        qp_h, qv_h = s["qpos"].copy(), s["qvel"].copy()
        stepper.run_two_ball_host(model, qp_h, qv_h, 151, dt=0.01, restitution=1.0, friction=0.3, radius=0.1, substeps=60, arith=arith)
        gq, gv = state_of(data)
        assert (qp_h == gq).all() and (qv_h == gv).all()
        E, B = 5_001, 64
        s = synth.multi_sphere(E, n_body=B, friction=0.2)
        model, data = ms.build(E, n_body=B)
        r = torch.tensor(0.09 + 0.02 * (np.arange(E * B) % 7) / 7, dtype=torch.float64, device="cuda")
        model.set_per_env(radius=r, mass=torch.full_like(r, float(model.body_mass[1])),
                          inertia=torch.full((3, E * B), float(model.body_inertia[1][0]), dtype=torch.float64, device="cuda"))
        data.set_state(s["qpos"], s["qvel"])
        for k in (16, 16, 5):
            stepper.step_multi_sphere(model, data, 0.01, 1.0, 0.2, substeps=k, arith=arith)
        qp_h, qv_h = s["qpos"].copy(), s["qvel"].copy()
        stepper.run_multi_sphere_host(model, qp_h, qv_h, 37, dt=0.01, restitution=1.0, friction=0.2, substeps=16, arith=arith)
        gq, gv = state_of(data)
        assert (qp_h == gq).all() and (qv_h == gv).all()
    finally:
        rb._lib.set_option("host_chunks", old)


def test_free_functions_follow_input_precision(rb):
    """float32 CUDA tensors select the float kernels (and match the float oracle bit for bit); float64 the double ones."""
    rng = np.random.default_rng(9)
    n = 1000
    v, w, r = rng.uniform(-2, 2, (n, 3)), rng.uniform(-5, 5, (n, 3)), rng.uniform(-0.3, 0.3, (n, 3))
    nn = rng.normal(size=(n, 3))
    nn /= np.linalg.norm(nn, axis=1, keepdims=True)
    T32 = lambda a: torch.tensor(a, dtype=torch.float32, device="cuda")
    jn, jt = rb.compute_collision_impulse_friction(2.5, None, T32(v), T32(w), T32(r), T32(nn), 0.8, 0.4)
    assert jn.dtype == torch.float32 and jt.shape == (n, 3)
    ojn, ojt = co.impulse_friction(np.full(n, 2.5), v, w, r, nn, np.full(n, 0.8), np.full(n, 0.4), dtype=np.float32)
    assert (jn.cpu().numpy() == ojn).all() and (jt.cpu().numpy() == ojt).all()
    Iw = rb.compute_inertia_tensor_world(T32(rng.uniform(0.1, 2, (n, 3))), T32(rng.normal(size=(n, 4))))
    assert Iw.dtype == torch.float32 and Iw.shape == (n, 3, 3)
    sym = (Iw - Iw.transpose(1, 2)).abs().max()
    assert float(sym) < 1e-5


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_contact_threshold_semantics(rb, arith):
    """contacts with |dist| < contact_threshold are skipped (collision.py:79-80): the fast kernel folds
    `dist < 0 and not |dist| < thr` into one comparison -- same decisions as the oracle, including thr = exactly |dist|."""
    from rigidbody_simulation_b200 import stepper, synth
    E = 60_000
    s = synth.sphere_incline(E)
    for thr in (1e-3, 5e-2):
        model, data = make_single(rb, "sphere", [0.2], 0.7, s["qpos"], s["qvel"])
        model.set_per_env(restitution=s["restitution"], friction=s["friction"])
        qp, qv = s["qpos"].copy(), s["qvel"].copy()
        cnt = (np.zeros(E, np.uint32), np.zeros(E, np.uint32))
        co.step_body_plane(qp, qv, 250, geom="sphere", mass=model.body_mass[-1], inertia=model.body_inertia[-1], size=0.2,
                           plane_pos=[0, 0, 0], plane_normal=model.plane_normal, gravity=G, dt=s["dt"],
                           restitution=s["restitution"], friction=s["friction"], threshold=thr, counters=cnt)
        stepper.step_body_plane(model, data, -1, s["dt"], None, None, thr, substeps=250, arith=arith)
        calls, imps = data.counters()
        assert cnt[0].sum() > E
        assert (calls[:, 0] == cnt[0]).all() and (imps[:, 0] == cnt[1]).all(), thr
    # a sphere resting exactly thr deep: |dist| == thr is NOT below the threshold, so the contact is processed
    qpos = np.array([[0.0, 0.0, 0.2 - 0.015625, 1, 0, 0, 0]])          # dist = -2^-6 exactly
    model, data = make_single(rb, "sphere", [0.2], 0.0, qpos, np.zeros((1, 6)))
    stepper.step_body_plane(model, data, -1, 0.009, 1.0, 0.5, 0.015625, substeps=1, arith=arith)
    assert int(data.counters()[0].sum()) == 1
    model, data = make_single(rb, "sphere", [0.2], 0.0, qpos, np.zeros((1, 6)))
    stepper.step_body_plane(model, data, -1, 0.009, 1.0, 0.5, 0.015626, substeps=1, arith=arith)
    assert int(data.counters()[0].sum()) == 0


def test_checkpoint_resume_is_bit_exact(rb):
    """Stop after 70 substeps, checkpoint, resume in a fresh object: same bits as the uninterrupted run."""
    from rigidbody_simulation_b200 import stepper, synth
    E = 20_000
    s = synth.cube(E, kind="bounce")
    kw = dict(dt=0.009, restitution=0.2, friction_coeff=0.6, contact_threshold=1e-4)
    model, straight = make_single(rb, "box", s["half"], 0.0, s["qpos"], s["qvel"])
    stepper.step_body_plane(model, straight, -1, substeps=150, **kw)
    model, first = make_single(rb, "box", s["half"], 0.0, s["qpos"], s["qvel"])
    stepper.step_body_plane(model, first, -1, substeps=70, **kw)
    sd = first.state_dict()
    model, second = make_single(rb, "box", s["half"], 0.0, np.zeros_like(s["qpos"]) + [0, 0, 9, 1, 0, 0, 0], np.zeros_like(s["qvel"]))
    second.load_state_dict(sd)
    stepper.step_body_plane(model, second, -1, substeps=80, **kw)
    assert torch.equal(second.state, straight.state) and torch.equal(second.n_contacts, straight.n_contacts)


# fp64: 4 substeps x the per-step bar.  fp32: a contact multiplies the rounding of the tangential velocity by 1/I (up to
# ~110 for the smallest spheres here), so after a few contact steps two equally valid single-precision evaluations (the
# float oracle's order and the plane-frame order) are ~1e-4 apart in the spin; the bound only guards against a real bug.
@pytest.mark.parametrize("dtype,tol", [(np.float64, 4e-12), (np.float32, 2e-3)])
def test_plane_frame_kernel_arbitrary_plane(rb, dtype, tol):
    """The fused fast launches work in the plane frame (rotate in, step, rotate out).  Planes tilted about two axes
    and not through the origin, per-env radius / mass / inertia: 4 fused substeps stay within the per-step bar of the
    oracle, event counts over a longer horizon are exact."""
    import ctypes
    from rigidbody_simulation_b200 import scenes, stepper
    import rigidbody_simulation_b200.mj as mj
    rng = np.random.default_rng(21)
    E = 40_000
    for euler, ppos in (((0.3, -0.4, 0.0), (0.2, -0.1, 0.05)), ((-0.9, 0.2, 0.0), (0.0, 0.0, -0.3)), ((0.0, 0.0, 0.0), (0.0, 0.0, 0.0))):
        xml = scenes.single_body_xml("sphere", [0.2], plane_euler=euler)
        xml = xml.replace('<geom name="ground" type="plane"', f'<geom name="ground" pos="{ppos[0]} {ppos[1]} {ppos[2]}" type="plane"')
        model = mj.MjModel.from_xml_string(xml, nenv=E, dtype=tdt(dtype))
        n = np.array(model.plane_normal)
        pp = np.array(model.plane_point)
        assert np.allclose(pp, ppos) and abs(np.linalg.norm(n) - 1) < 1e-15
        h = rng.uniform(0.15, 0.6, E)                                  # some start in contact, most just above
        tang = rng.normal(size=(E, 3))
        pos = pp + h[:, None] * n + (tang - (tang @ n)[:, None] * n) * 0.5
        q = rng.normal(size=(E, 4))
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        qpos = np.concatenate([pos, q], axis=1)
        qvel = np.concatenate([rng.uniform(-2, 2, (E, 3)), rng.uniform(-5, 5, (E, 3))], axis=1)
        rad = rng.uniform(0.15, 0.25, E)
        mass = 50 * 4 / 3 * np.pi * rad ** 3
        inertia = np.tile(0.4 * mass * rad ** 2, (3, 1))
        e, mu = rng.uniform(0.3, 1.0, E), rng.uniform(0, 1, E)
        model.set_per_env(mass=mass, inertia=inertia, size=np.stack([rad, rad * 0, rad * 0]), restitution=e, friction=mu)
        data = mj.MjData(model)
        data.set_state(qpos, qvel)
        qp, qv = qpos.astype(dtype), qvel.astype(dtype)
        cnt = (np.zeros(E, np.uint32), np.zeros(E, np.uint32))
        kw = dict(geom="sphere", mass=mass, inertia=inertia.T.copy(), size=rad[:, None], plane_pos=pp, plane_normal=n, gravity=G,
                  dt=0.009, restitution=e, friction=mu, threshold=0.0, counters=cnt)
        done = 0
        for upto in (4, 260):
            co.step_body_plane(qp, qv, upto - done, **kw)
            # isotropic per-env inertia rows: the fast policy needs the isotropic mode, which per-env inertia arrays
            # do not select automatically -- go through the args builder and set it explicitly
            a = stepper.body_plane_args(model, data, -1, 0.009, None, None, 0.0, rb._lib.RBS_SCHEME_A, upto - done, arith="strict")
            a.arith, a.inertia_mode = rb._lib.RBS_ARITH_FAST, rb._lib.RBS_INERTIA_ISOTROPIC
            a.stream = stepper.current_stream(model.device)
            rb._lib.check(rb._lib.load().rbs_step_body_plane(ctypes.byref(a)))
            done = upto
            gq, gv = state_of(data)
            floor = 1e-3 if dtype == np.float64 else 1e-2
            err = max(comp_rel_err(gq, qp, floor), comp_rel_err(gv, qv, floor))
            if upto == 4:
                assert err <= tol, (euler, err)
        if dtype == np.float64:
            calls, imps = data.counters()
            assert (calls[:, 0] == cnt[0]).all() and (imps[:, 0] == cnt[1]).all(), euler
            assert cnt[0].sum() > E


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fused_launch_survives_fast_spin(rb, dtype):
    """The fused fast kernels carry the orientation unnormalised between rescales (kRenormMask in rbs_kernels.cuh).
    Spheres in free flight spinning at up to 2e4 rad/s (|0.5*dt*w| ~ 90 per substep; a 32-substep rescale interval
    overflows float at ~4): 256 fused substeps must return finite unit quaternions that agree with 256 single-substep
    launches of the same policy, which normalise after every substep like the reference (collision.py:94-95)."""
    from rigidbody_simulation_b200 import stepper
    rng = np.random.default_rng(5)
    E = 4096
    q = rng.normal(size=(E, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    qpos = np.concatenate([rng.uniform(-1, 1, (E, 2)), rng.uniform(500, 600, (E, 1)), q], axis=1)    # never reaches the plane
    spin = rng.normal(size=(E, 3))
    spin *= (np.geomspace(1.0, 2e4, E) / np.linalg.norm(spin, axis=1))[:, None]
    qvel = np.concatenate([rng.uniform(-1, 1, (E, 3)), spin], axis=1)
    out = {}
    for K in (256, 1):
        model, data = make_single(rb, "sphere", [0.2], 0.7, qpos, qvel, dtype=dtype)
        for _ in range(256 // K):
            stepper.step_body_plane(model, data, -1, 0.009, 0.9, 0.5, 0.0, substeps=K, arith="fast")
        out[K] = state_of(data)
    gq, gv = out[256]
    assert np.isfinite(gq).all() and np.isfinite(gv).all()
    tol = 1e-9 if dtype == np.float64 else 2e-3           # 256 steps of rounding at up to 90 rad per substep
    assert np.max(np.abs(np.linalg.norm(gq[:, 3:], axis=1) - 1)) < (1e-12 if dtype == np.float64 else 1e-5)
    # q and -q are the same orientation; compare up to sign
    dq = np.minimum(np.abs(gq[:, 3:] - out[1][0][:, 3:]).max(axis=1), np.abs(gq[:, 3:] + out[1][0][:, 3:]).max(axis=1))
    assert dq.max() < tol, dq.max()
    # positions ~550 m up: one float ulp is 6e-5 and both paths round 256 times (the plane-frame rotation mixes y and z)
    assert np.allclose(gq[:, :3], out[1][0][:, :3], rtol=0, atol=1e-9 if dtype == np.float64 else 5e-2)


def test_packed_float_kernel_matches_scalar_kernel(rb):
    """Float fused launches run two environments per thread on packed fp32x2 instructions with a branch-free contact
    path (step_sphere_plane_pf2_kernel).  Every environment goes through the scalar plane-frame kernel's operations
    in the same order, so states and event counters must be the same numbers (option pf_packed=0 selects the scalar
    kernel): ragged sizes (odd, smaller than one CTA, not a multiple of 256), with and without counters / threshold,
    over a horizon in which most environments bounce.  (The per-step bar against the oracle for this kernel is
    test_plane_frame_kernel_arbitrary_plane[float32], whose fused launches take this path by default.)"""
    from rigidbody_simulation_b200 import scenes, stepper, synth
    dev = torch.device("cuda:0")
    for E, thr, count in ((100_003, 0.0, True), (77, 1e-4, True), (65_536 + 129, 1e-4, False), (4096, 0.0, False)):
        s = synth.sphere_incline(E)
        res = {}
        for packed in ("0", "1"):
            old = rb._lib.set_option("pf_packed", int(packed))
            try:
                model = scenes.sphere_on_incline(E, device=dev, dtype=torch.float32)
                model.set_per_env(restitution=s["restitution"], friction=s["friction"])
                data = rb.BatchedData(model)
                data.set_state(s["qpos"], s["qvel"])
                for K in (4, 260, 37):
                    stepper.step_body_plane(model, data, -1, s["dt"], None, None, thr, substeps=K, count=count, arith="fast")
                res[packed] = state_of(data) + tuple(c.copy() for c in data.counters())
            finally:
                rb._lib.set_option("pf_packed", old)
        for a, b in zip(res["0"], res["1"]):
            assert np.array_equal(a, b), (E, thr, count, np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())
        if count:
            assert res["1"][2].sum() > E // 2                      # the horizon does exercise contacts


def test_packed_float_two_ball_kernel_matches_scalar_kernel(rb):
    """Float launches of the two-ball fast stepper run two environments per thread on packed fp32x2 instructions
    (step_two_ball_fast2_kernel); the rare event paths are the scalar kernel's statements on the unpacked slot, so state
    and both event counters must be the scalar kernel's numbers bit for bit (option tb_packed=0 selects it): ragged sizes
    (odd, smaller than one CTA, a tail of less than one CTA in the second slot), shipped and tilted gravity (the GZ and the
    general instantiation), over a horizon with ground hits and pair hits.  The bar against the oracle for the float
    kernel is test_two_ball_fast_policy_vs_oracle[float32], whose launches take this path by default."""
    from rigidbody_simulation_b200 import stepper, synth
    from rigidbody_simulation_b200.src.simulation import ball_collision
    dev = torch.device("cuda:0")
    for E, grav in ((100_003, None), (77, None), (65_536 + 129, (0.3, -0.2, -9.8)), (256, None)):
        s = synth.two_ball(E)
        res = {}
        for packed in (0, 1):
            old = rb._lib.set_option("tb_packed", packed)
            try:
                model, data = ball_collision.build(E, device=dev, dtype=torch.float32)
                if grav is not None:
                    model.opt.gravity[:] = grav
                data.set_state(s["qpos"], s["qvel"])
                for K in (4, 260, 37, 1, 150):
                    stepper.step_two_ball(model, data, 0.01, 1.0, 0.3, radius=0.1, substeps=K, count=True, arith="fast")
                res[packed] = state_of(data) + tuple(c.copy() for c in data.counters())
            finally:
                rb._lib.set_option("tb_packed", old)
        for a, b in zip(res[0], res[1]):
            assert np.array_equal(a, b), (E, grav, np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())
        assert res[1][2].sum() > E // 2 and res[1][3].sum() > E // 4    # ground hits and pair hits both happen


def test_two_ball_tilted_gravity_both_policies(rb):
    """The two-ball fast kernel has a gravity-along-z instantiation (every shipped model) and a general one: a gravity
    vector with horizontal components goes through the general one and must meet the same bar against the oracle;
    the strict policy stays bit-for-bit."""
    from rigidbody_simulation_b200 import stepper, synth
    from rigidbody_simulation_b200.src.simulation import ball_collision
    E = 20_000
    s = synth.two_ball(E)
    g = [0.7, -0.4, -9.8]
    for arith in ("strict", "fast"):
        model, data = ball_collision.build(E)
        model.opt.gravity[:] = g
        data.set_state(s["qpos"], s["qvel"])
        qp, qv = s["qpos"].copy(), s["qvel"].copy()
        hits = (np.zeros(E, np.uint32), np.zeros(E, np.uint32))
        m = float(model.body_mass[1])
        done = 0
        for upto in (1, 120):
            co.step_two_ball(qp, qv, upto - done, mass=[m, m], radius=0.1, gravity=g, dt=0.01, restitution=1.0, friction=0.3, counters=hits)
            stepper.step_two_ball(model, data, 0.01, 1.0, 0.3, radius=0.1, substeps=upto - done, arith=arith)
            done = upto
            gq, gv = state_of(data)
            err = max(comp_rel_err(gq, qp, 1e-3), comp_rel_err(gv, qv, 1e-3))
            assert err <= (F64_STEP if upto == 1 else 1e-7), (arith, upto, err)
            if arith == "strict":
                assert err == 0.0
        same = (data.n_contacts[:E].cpu().numpy() == hits[0]) & (data.n_impulses[:E].cpu().numpy() == hits[1])
        assert same.mean() >= (1.0 if arith == "strict" else 0.9999), (arith, same.mean())
        assert hits[0].sum() > 0 and hits[1].sum() > E // 2


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fused_fast_kernels_degenerate_inputs(rb, dtype):
    """The fused fast sphere kernels decide `dist < 0`, `u_n < 0` and `|u_t| > 1e-6` with integer tests on the bit
    patterns (double) or on packed pairs of environments (float).  Degenerate inputs must behave like the reference:
    a sphere exactly touching at rest (dist == 0) is not in contact on the first substep, one resting exactly on the
    plane with zero approach speed gets no impulse, a NaN environment stays NaN and does not disturb the environment
    that shares its thread / warp, and the event counters equal the oracle's."""
    from rigidbody_simulation_b200 import stepper
    E = 300                                # not a multiple of 256: the packed kernel's second slot idles in the tail
    qpos = np.tile(np.array([0, 0, 1.0, 1, 0, 0, 0.0]), (E, 1))
    qvel = np.zeros((E, 6))
    qpos[:, 2] = np.linspace(0.15, 0.6, E)                  # some start penetrating, some just above, most in the air
    qvel[:, 2] = np.linspace(-1.0, 0.5, E)
    qvel[:, 3] = 3.0
    qpos[9, 2], qvel[9, 2] = 0.2, 0.0                       # exactly touching, at rest
    qpos[10, 2], qvel[10, 2] = 0.2, -0.5                    # exactly touching, approaching: still dist == 0, no contact
    qpos[11, 2], qvel[11, 2] = 0.1999, 0.0                  # penetrating, v_z becomes g*dt < 0 on the first substep
    qpos[7, 2] = np.nan
    qpos[7 + 128, 5] = np.nan                               # shares a thread with env 7 in the packed float kernel
    model, data = make_single(rb, "sphere", [0.2], 0.0, qpos, qvel, dtype=dtype)
    ok = np.ones(E, bool)
    ok[[7, 7 + 128]] = False
    qp, qv = qpos.astype(dtype), qvel.astype(dtype)
    cnt = (np.zeros(E, np.uint32), np.zeros(E, np.uint32))
    kw = dict(geom="sphere", mass=model.body_mass[-1], inertia=model.body_inertia[-1], size=0.2, plane_pos=[0, 0, 0],
              plane_normal=[0, 0, 1], gravity=G, dt=0.009, restitution=0.8, friction=0.5, threshold=0.0, counters=cnt)
    for K in (1, 4, 60):                   # 1: world-frame kernel; 4, 60: plane-frame kernels (packed in float)
        co.step_body_plane(qp, qv, K, **kw)
        stepper.step_body_plane(model, data, -1, 0.009, 0.8, 0.5, 0.0, substeps=K, arith="fast")
        gq, gv = state_of(data)
        if K == 1:
            calls, _ = data.counters()
            assert calls[9, 0] == 0 and calls[10, 0] == 0 and calls[11, 0] == 1
        assert np.isnan(gq[7, 2]) and np.isnan(gq[7 + 128, 3:]).all()
        assert np.isfinite(gq[ok]).all() and np.isfinite(gv[ok]).all()
        if dtype == np.float64:
            assert comp_rel_err(gq[ok], qp[ok], 1e-3) <= 1e-10 and comp_rel_err(gv[ok], qv[ok], 1e-3) <= 1e-10
    calls, imps = data.counters()
    if dtype == np.float64:
        assert (calls[ok, 0] == cnt[0][ok]).all() and (imps[ok, 0] == cnt[1][ok]).all()
    assert cnt[0][ok].sum() > 50


# ------------------------------------------------------------------------------------ N4: spheres and boxes in one scene
MIXED_BODIES = ([{"type": "box", "size": [0.4, 0.4, 0.4]}, {"type": "sphere", "size": [0.2]}, {"type": "box", "size": [0.3, 0.2, 0.25]},
                 {"type": "sphere", "size": [0.25]}, {"type": "box", "size": [0.35, 0.35, 0.15]}, {"type": "sphere", "size": [0.15]},
                 {"type": "box", "size": [0.2, 0.45, 0.3]}, {"type": "sphere", "size": [0.3]}])


def _multi_body_case(bodies, E, dtype, plane_euler=(0.0, 0.0, 0.0), pitch=0.8):
    import rigidbody_simulation_b200.mj as mj
    from rigidbody_simulation_b200 import scenes, stepper, synth
    B = len(bodies)
    s = synth.multi_body(E, bodies, pitch=pitch)
    model = mj.MjModel.from_xml_string(scenes.multi_body_xml(bodies, plane_euler=plane_euler), nenv=E, dtype=tdt(dtype))
    data = mj.MjData(model, layout="body")
    data.set_state(s["qpos"], s["qvel"])
    tab = stepper.body_table(model)
    okw = dict(gtype=tab[:, 0].astype(np.int32), mass=tab[:, 4], inertia=tab[:, 5:8], size=tab[:, 1:4], plane_pos=model.plane_point,
               plane_normal=model.plane_normal, gravity=G, geom_pos=tab[:, 8:11] if model.has_offset_geoms else None,
               geom_quat=tab[:, 11:15] if model.has_offset_geoms else None)
    return model, data, s["qpos"].astype(dtype).reshape(E, B, 7).copy(), s["qvel"].astype(dtype).reshape(E, B, 6).copy(), okw


@pytest.mark.parametrize("dtype,tol", [(np.float64, F64_STEP), (np.float32, F32_STEP)])
@pytest.mark.parametrize("offsets", [False, True])
def test_multi_body_spheres_and_boxes_vs_oracle(rb, dtype, tol, offsets):
    """N4: eight bodies (four boxes, one of them anisotropic in every axis, four spheres of different radii) tumbling on
    a lattice over a tilted ground -- plane-sphere, plane-box, sphere-sphere, sphere-box and box-box contacts all occur --
    against the C oracle's restatement of the same loop: double is bit for bit over 400 steps with every per-body counter
    equal; float meets the per-step bar.  ``offsets``: every geom placed off its body's origin and axes (N1 remainder)."""
    from rigidbody_simulation_b200 import stepper
    bodies = [dict(b) for b in MIXED_BODIES]
    if offsets:
        for i, b in enumerate(bodies):
            b["geom_pos"] = [0.05 * ((i % 3) - 1), 0.03 * (i % 2), -0.04 * (i % 4 == 1)]
            b["geom_euler"] = [0.1 * i, -0.2, 0.05 * i]
    E, B = 3000, len(bodies)
    model, data, qp, qv, okw = _multi_body_case(bodies, E, dtype, plane_euler=(0.15, -0.1, 0.0))
    assert model.has_offset_geoms == offsets
    cnt = (np.zeros((E, B), np.uint32), np.zeros((E, B), np.uint32))
    done = 0
    for upto in (1, 10, 400):
        co.step_multi_body(qp, qv, upto - done, dt=0.005, restitution=0.2, friction=0.6, counters=cnt, **okw)
        stepper.step_multi_body(model, data, 0.005, 0.2, 0.6, substeps=upto - done)
        done = upto
        gq, gv = state_of(data)
        if dtype == np.float64:
            assert np.array_equal(gq, qp.reshape(E, -1)) and np.array_equal(gv, qv.reshape(E, -1)), upto
        elif upto == 1:
            assert max(comp_rel_err(gq.ravel(), qp.ravel(), 1e-3), comp_rel_err(gv.ravel(), qv.ravel(), 1e-3)) <= tol
    calls, imps = data.counters()
    if dtype == np.float64:
        assert np.array_equal(calls, cnt[0]) and np.array_equal(imps, cnt[1])
    assert cnt[0].sum() > 20 * E * B                                  # contacts are exercised ...
    assert (cnt[0][:, 1::2].sum(axis=0) > 0).all() and (cnt[0][:, 0::2].sum(axis=0) > 0).all()   # ... by every body


def test_multi_body_pair_types_all_occur_and_identities(rb):
    """(1) Every pair type of N4 is exercised: scenes of two bodies where the only possible non-ground contact is the
    pair itself (sphere on box, box on sphere, box on box) produce more calls than the ground alone could, bit for bit the
    oracle.  (2) Identities that pin the new kernel to reference-pinned ones: all spheres == step_multi_sphere (strict),
    one box == step_body_plane (strict box kernel) -- states and counters bit for bit."""
    import rigidbody_simulation_b200.mj as mj
    from rigidbody_simulation_b200 import scenes, stepper, synth
    from rigidbody_simulation_b200.src.simulation import multi_sphere_bounce as ms
    E = 2048
    big, small, ball = {"type": "box", "size": [0.8, 0.8, 0.4]}, {"type": "box", "size": [0.4, 0.4, 0.4]}, {"type": "sphere", "size": [0.2]}
    for bodies in ([big, ball], [ball, big], [big, small], [small, big]):
        B = 2
        model = mj.MjModel.from_xml_string(scenes.multi_body_xml(bodies), nenv=E, dtype=torch.float64)
        data = mj.MjData(model, layout="body")
        f = synth._Fields(synth.SEED, 0, E, 900)
        qpos = np.zeros((E, B, 7)); qvel = np.zeros((E, B, 6))
        low, high = (0, 1) if bodies[0]["size"][0] == 0.8 else (1, 0)        # the big box rests on the ground, the other falls on it
        qpos[:, low, :3] = [0, 0, 0.4]; qpos[:, low, 3] = 1
        qpos[:, high, :2] = f.u(-0.5, 0.5, (2,)); qpos[:, high, 2] = f.u(1.3, 2.0); qpos[:, high, 3:] = f.unit_quat()
        qvel[:, high, 3:] = f.u(-1, 1, (3,))
        data.set_state(qpos.reshape(E, -1), qvel.reshape(E, -1))
        tab = stepper.body_table(model)
        cnt = (np.zeros((E, B), np.uint32), np.zeros((E, B), np.uint32))
        qp, qv = qpos.copy(), qvel.copy()
        co.step_multi_body(qp, qv, 300, gtype=tab[:, 0].astype(np.int32), mass=tab[:, 4], inertia=tab[:, 5:8], size=tab[:, 1:4],
                           plane_pos=[0, 0, 0], plane_normal=[0, 0, 1], gravity=G, dt=0.005, restitution=0.2, friction=0.6, counters=cnt)
        stepper.step_multi_body(model, data, 0.005, 0.2, 0.6, substeps=300)
        gq, gv = state_of(data)
        assert np.array_equal(gq, qp.reshape(E, -1)) and np.array_equal(gv, qv.reshape(E, -1)), bodies
        calls, imps = data.counters()
        assert np.array_equal(calls, cnt[0]) and np.array_equal(imps, cnt[1])
        # The A9 loop never flips the normal (geom1 -> geom2, geom1 = the lower index), so only the HIGHER-index body of a
        # pair responds to an approach: falling with the higher index it is caught by the box and
        # never reaches the ground within 300 steps -- all its calls are pair contacts; falling with the lower index it
        # registers the pair contacts (calls) but gets no impulse from them and ends on the ground.
        above = qp[:, high, 2] > 0.8 + 0.15
        if high > low:
            assert above.mean() > 0.5 and (cnt[0][above, high] > 0).mean() > 0.9, (bodies, above.mean())
        else:
            assert above.mean() < 0.1 and (cnt[0][:, high] > cnt[1][:, high]).mean() > 0.9, (bodies, above.mean())
    # all spheres == the strict multi-sphere kernel
    B, E = 27, 1500
    s = synth.multi_sphere(E, n_body=B, friction=0.3)
    m1, d1 = ms.build(E, n_body=B, dtype=torch.float64)
    d1.set_state(s["qpos"], s["qvel"])
    stepper.step_multi_sphere(m1, d1, 0.01, 1.0, 0.3, substeps=80, arith="strict")
    m2 = mj.MjModel.from_xml_string(scenes.multi_body_xml([{"type": "sphere", "size": [0.1]}] * B, timestep=0.01), nenv=E, dtype=torch.float64)
    d2 = mj.MjData(m2, layout="body")
    d2.set_state(s["qpos"], s["qvel"])
    stepper.step_multi_body(m2, d2, 0.01, 1.0, 0.3, substeps=80)
    for a, b in zip(state_of(d1) + d1.counters(), state_of(d2) + d2.counters()):
        assert np.array_equal(a, b)
    assert d1.counters()[0].sum() > E * B
    # one box == the strict single-body box kernel (threshold 0: the multi-body loop has none, like multi_sphere_bounce.py)
    E = 4096
    sc = synth.cube(E, kind="bounce")
    m3 = scenes.cube_on_plane(E, theta=0.0, dtype=torch.float64)
    d3 = rb.BatchedData(m3)
    d3.set_state(sc["qpos"], sc["qvel"])
    stepper.step_body_plane(m3, d3, -1, 0.009, 0.2, 0.6, 0.0, substeps=300, arith="strict")
    m4 = mj.MjModel.from_xml_string(scenes.multi_body_xml([{"type": "box", "size": [0.4, 0.4, 0.4]}], timestep=0.009), nenv=E, dtype=torch.float64)
    d4 = mj.MjData(m4, layout="body")
    d4.set_state(sc["qpos"], sc["qvel"])
    stepper.step_multi_body(m4, d4, 0.009, 0.2, 0.6, substeps=300)
    for a, b in zip(state_of(d3) + d3.counters(), state_of(d4) + d4.counters()):
        assert np.array_equal(a, b)
    assert d3.counters()[0].sum() > E


def test_mixed_pile_cli_and_shard_invariance(rb, capsys):
    """The N4 scenario through the CLI: `--sim mixed_pile` prints the one JSON line with finite statistics and contacts,
    and its environments are a pure function of the global index -- environments [3, 7) of a 10-env run computed as
    their own shard (start = 3) are bit for bit the same states."""
    import json
    from rigidbody_simulation_b200.src import simulate
    from rigidbody_simulation_b200.src.simulation import mixed_pile
    out = simulate.run_simulation("mixed_pile", steps=120, envs=10, substeps=40, bodies=8)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["sim"] == "mixed_pile" and line["config"] == "random" and line["contacts"] > 0 and line["impulses"] > 0
    assert np.isfinite(line["stats"]["energy_sum"]) and 0.0 < line["stats"]["max_height"] < 20.0
    _, whole, _ = mixed_pile.run_headless(120, 10, substeps=40)
    _, part, _ = mixed_pile.run_headless(120, 4, substeps=40, start=3)
    assert out["qpos_env0"] == whole.qpos.torch().cpu().numpy()[0].tolist()
    assert np.array_equal(whole.qpos.torch().cpu().numpy()[3:7], part.qpos.torch().cpu().numpy())
    with pytest.raises(SystemExit):
        simulate.run_simulation("mixed_pile", steps=10, envs=2, arith="fast")


@pytest.mark.parametrize("B,E", [(70, 37), (256, 5), (3, 1001), (1, 130)])
def test_multi_body_shapes_vs_oracle(rb, B, E):
    """Launch shapes of the multi-body stepper: more than 64 bodies (several words of the broad-phase bitmask, B = 70 is
    also not a divisor of the CTA size), the ABI maximum of 256 bodies (one environment per CTA, more than 48 KB of
    shared memory), odd environment counts (ragged last CTA), a single body; bit for bit the oracle, counters included."""
    from rigidbody_simulation_b200 import stepper
    bodies = [dict(MIXED_BODIES[i % len(MIXED_BODIES)]) for i in range(B)]
    model, data, qp, qv, okw = _multi_body_case(bodies, E, np.float64, pitch=0.8)
    cnt = (np.zeros((E, B), np.uint32), np.zeros((E, B), np.uint32))
    steps = 60 if B >= 70 else 200
    co.step_multi_body(qp, qv, steps, dt=0.005, restitution=0.2, friction=0.6, counters=cnt, **okw)
    for k in (steps // 2, steps - steps // 2):
        stepper.step_multi_body(model, data, 0.005, 0.2, 0.6, substeps=k)
    gq, gv = state_of(data)
    assert np.array_equal(gq, qp.reshape(E, -1)) and np.array_equal(gv, qv.reshape(E, -1))
    calls, imps = data.counters()
    assert np.array_equal(calls, cnt[0]) and np.array_equal(imps, cnt[1])
    assert cnt[0].sum() > 0
    with pytest.raises(ValueError):
        stepper.step_multi_body(model, data, 0.005, 0.2, 0.6, substeps=0)


def test_multi_body_host_buffer_driver_matches_device_path(rb):
    """rbs_run_multi_body_host (reference-layout host arrays in and out, chunk pipeline inside) computes what the
    device-resident stepper computes, bit for bit, on a batch large enough for several pipeline chunks and an odd tail."""
    import rigidbody_simulation_b200.mj as mj
    from rigidbody_simulation_b200 import scenes, stepper, synth
    E, B = 40_003, len(MIXED_BODIES)
    model, data, _, _, _ = _multi_body_case(MIXED_BODIES, E, np.float64, pitch=0.8)
    s = synth.multi_body(E, MIXED_BODIES, pitch=0.8)
    qp, qv = s["qpos"].copy(), s["qvel"].copy()
    stepper.run_multi_body_host(model, qp, qv, 150, dt=0.005, restitution=0.2, friction=0.6, substeps=64)
    for k in (64, 64, 22):
        stepper.step_multi_body(model, data, 0.005, 0.2, 0.6, substeps=k)
    gq, gv = state_of(data)
    assert np.array_equal(gq, qp) and np.array_equal(gv, qv)
    with pytest.raises(ValueError):
        stepper.run_multi_body_host(model, qp[:10], qv, 1)


def test_multi_body_box_edges_through_boxes(rb):
    """Box pairs that touch without any vertex of one inside the other -- planks crossed at random angles, one dropped on
    the other -- are held by the edge-through-box contacts: bit for bit the oracle, the upper plank ends on the lower one
    (without those contacts it falls through to the ground)."""
    import rigidbody_simulation_b200.mj as mj
    from rigidbody_simulation_b200 import scenes, stepper, synth
    E, B = 1500, 2
    planks = [{"type": "box", "size": [1.0, 0.2, 0.15]}, {"type": "box", "size": [0.2, 1.0, 0.15]}]
    model = mj.MjModel.from_xml_string(scenes.multi_body_xml(planks, density=200.0), nenv=E, dtype=torch.float64)
    data = mj.MjData(model, layout="body")
    f = synth._Fields(synth.SEED, 0, E, 950)
    qpos = np.zeros((E, B, 7)); qvel = np.zeros((E, B, 6))
    qpos[:, 0, :3] = [0, 0, 0.15]; qpos[:, 0, 3] = 1
    yaw = f.u(-0.6, 0.6)
    qpos[:, 1, 0] = f.u(-0.1, 0.1); qpos[:, 1, 1] = f.u(-0.1, 0.1); qpos[:, 1, 2] = f.u(0.5, 0.9)
    qpos[:, 1, 3] = np.cos(yaw / 2); qpos[:, 1, 6] = np.sin(yaw / 2)
    data.set_state(qpos.reshape(E, -1), qvel.reshape(E, -1))
    tab = stepper.body_table(model)
    cnt = (np.zeros((E, B), np.uint32), np.zeros((E, B), np.uint32))
    qp, qv = qpos.copy(), qvel.copy()
    co.step_multi_body(qp, qv, 300, gtype=tab[:, 0].astype(np.int32), mass=tab[:, 4], inertia=tab[:, 5:8], size=tab[:, 1:4],
                       plane_pos=[0, 0, 0], plane_normal=[0, 0, 1], gravity=G, dt=0.005, restitution=0.2, friction=0.6, counters=cnt)
    stepper.step_multi_body(model, data, 0.005, 0.2, 0.6, substeps=300)
    gq, gv = state_of(data)
    assert np.array_equal(gq, qp.reshape(E, -1)) and np.array_equal(gv, qv.reshape(E, -1))
    calls, imps = data.counters()
    assert np.array_equal(calls, cnt[0]) and np.array_equal(imps, cnt[1])
    assert (qp[:, 1, 2] > 0.3 + 0.15 - 0.06).mean() > 0.95, (qp[:, 1, 2] > 0.39).mean()     # resting on the lower plank (top face at 0.3)


def test_multi_body_follows_model_edits(rb):
    """The cached body table is rebuilt when the model's mass / inertia arrays are edited between steps (MuJoCo-style):
    the second half of a run with a doubled mass of body 1 equals a fresh scene built with that mass."""
    import rigidbody_simulation_b200.mj as mj
    from rigidbody_simulation_b200 import scenes, stepper, synth
    E = 512
    model, data, _, _, _ = _multi_body_case(MIXED_BODIES, E, np.float64)
    stepper.step_multi_body(model, data, 0.005, 0.2, 0.6, substeps=40)
    half = data.state.clone()
    model.body_mass[model.free_ids[0]] *= 2.0
    stepper.step_multi_body(model, data, 0.005, 0.2, 0.6, substeps=40)
    model2, data2, _, _, _ = _multi_body_case(MIXED_BODIES, E, np.float64)
    model2.body_mass[model2.free_ids[0]] *= 2.0
    data2.state.copy_(half)
    stepper.step_multi_body(model2, data2, 0.005, 0.2, 0.6, substeps=40)
    assert torch.equal(data.state, data2.state)
    model3, data3, _, _, _ = _multi_body_case(MIXED_BODIES, E, np.float64)
    data3.state.copy_(half)
    stepper.step_multi_body(model3, data3, 0.005, 0.2, 0.6, substeps=40)
    assert not torch.equal(data.state, data3.state)            # the edit did change the dynamics
