"""CPU tests of the host-side logic: XML subset parser against the values the (fake) MuJoCo compiler gives
the reference, the index-keyed synthetic generator, sharding, config mirror, and the no-fallback rule."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parser_matches_reference_model_values(golden):
    from rigidbody_simulation_b200 import mjcf, scenes
    mv = golden("model_values")
    for name in ("sphere", "cube", "ball_collision", "multi_sphere"):
        sc = mjcf.parse_file(scenes.model_path(name))
        ref = mv[name]
        assert [b.name for b in sc.bodies] == ref["body_names"]
        assert [b.mass for b in sc.bodies] == ref["body_mass"]                  # bit-for-bit
        assert [b.inertia for b in sc.bodies] == ref["body_inertia"]
        assert sc.gravity == ref["gravity"] and sc.timestep == ref["timestep"]
        qpos0 = []
        for i in sc.free_bodies:
            qpos0 += sc.bodies[i].pos + sc.bodies[i].quat
        assert qpos0 == ref["qpos0"]
        _, normal = mjcf.plane_frame(sc, sc.planes()[0])
        assert normal == ref["plane_normal"]
    sc = mjcf.parse_file(scenes.model_path("sphere"))
    assert sc.body_id("sphere") == -1 and sc.body_id("ball") == 2           # the reference relies on the -1
    # the reference's own XML files parse to the same numbers (only where the checkout exists)
    ref_models = "/root/reference/models"
    if os.path.isdir(ref_models):
        for name in ("sphere", "cube", "ball_collision", "multi_sphere"):
            a, b = mjcf.parse_file(os.path.join(ref_models, name + ".xml")), mjcf.parse_file(scenes.model_path(name))
            assert [x.mass for x in a.bodies] == [x.mass for x in b.bodies]
            assert [x.pos + x.quat for x in a.bodies] == [x.pos + x.quat for x in b.bodies]
            assert a.timestep == b.timestep and a.gravity == b.gravity


def test_parser_templating_and_errors():
    from rigidbody_simulation_b200 import mjcf
    txt = ('<mujoco><compiler angle="radian"/><option gravity="0 0 -9.8" timestep="{TIMESTEP}"/><worldbody>'
           '<geom type="plane" size="1 1 1" euler="{INCLINE_ANGLE} 0 0"/>'
           '<body name="b" pos="0 0 1"><freejoint/><geom type="box" size="0.3 0.2 0.1" density="50"/></body>'
           '</worldbody></mujoco>')
    sc = mjcf.parse_string(mjcf.render_template(txt, incline_angle=0.25, timestep=0.004))
    assert sc.timestep == 0.004
    _, n = mjcf.plane_frame(sc, sc.planes()[0])
    assert n == pytest.approx([0, -math.sin(0.25), math.cos(0.25)], abs=1e-15)
    m = 50 * 8 * 0.3 * 0.2 * 0.1
    assert sc.bodies[1].mass == pytest.approx(m, rel=1e-15)
    assert sc.bodies[1].inertia == pytest.approx([m / 3 * (0.04 + 0.01), m / 3 * (0.09 + 0.01), m / 3 * (0.09 + 0.04)], rel=1e-14)
    with pytest.raises(ValueError):
        mjcf.parse_string('<mujoco><worldbody><geom type="capsule" size="1 1"/></worldbody></mujoco>')
    with pytest.raises(ValueError):
        mjcf.parse_string('<mujoco><compiler angle="gradian"/><worldbody/></mujoco>')
    with pytest.raises(ValueError):
        mjcf.parse_string("<notmujoco/>")


def test_synth_is_keyed_by_global_index():
    from rigidbody_simulation_b200 import synth
    from rigidbody_simulation_b200.shard import shard_range
    whole = synth.sphere_incline(1001)
    parts = [synth.sphere_incline(c, start=s) for s, c in (shard_range(1001, r, 4) for r in range(4))]
    assert (np.concatenate([p["qpos"] for p in parts]) == whole["qpos"]).all()
    assert (np.concatenate([p["friction"] for p in parts]) == whole["friction"]).all()
    q = whole["qpos"][:, 3:7]
    assert np.allclose(np.linalg.norm(q, axis=1), 1.0, atol=1e-15)
    h = whole["qpos"][:, :3] @ whole["plane_normal"]
    assert h.min() >= 0.25 - 1e-12 and h.max() <= 2.5 + 1e-12              # starts above the plane
    assert 0.5 <= whole["restitution"].min() and whole["restitution"].max() < 1.0
    assert synth.sphere_incline(10, seed=1)["qpos"].tolist() != whole["qpos"][:10].tolist()
    ms = synth.multi_sphere(3, n_body=64)
    assert ms["qpos"].shape == (3, 448) and ms["qvel"].shape == (3, 384)
    tb = synth.two_ball(5)
    assert tb["qpos"].shape == (5, 14) and (tb["qpos"][:, 3] == 1).all()
    for kind in ("bounce", "incline"):
        c = synth.cube(7, kind=kind)
        assert np.allclose(np.linalg.norm(c["qpos"][:, 3:7], axis=1), 1.0, atol=1e-15)
    u = synth.uniform01(1, 2, np.arange(100000))
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 5e-3


def test_shard_ranges_partition():
    from rigidbody_simulation_b200.shard import shard_range
    for n, w in ((1 << 20, 8), (65536, 8), (10, 4), (3, 8), (0, 2)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        for (s0, c0), (s1, _) in zip(spans, spans[1:]):
            assert s0 + c0 == s1
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 4, 4)


def test_config_mirror_matches_reference_values():
    sys.path.insert(0, os.path.join(ROOT, "rigidbody-simulation_b200"))
    try:
        from rigidbody_simulation_b200.src.config import load_sim_config
    finally:
        sys.path.pop(0)
    c = load_sim_config("cube_incline")
    assert (c["FRICTION_COEFFICIENT"], c["RESTITUTION"], c["TIMESTEP"], c["INCLINE_ANGLE_RAD"]) == (0.6, 0.2, 0.009, 0.7)
    assert load_sim_config("single_sphere_bounce")["RESTITUTION"] == 1.0
    assert load_sim_config("ball_collision")["FRICTION_COEFFICIENT"] == 0.3
    assert load_sim_config("multi_sphere_bounce")["FRICTION_COEFFICIENT"] == 0.0
    d = load_sim_config("unknown")
    assert (d["FRICTION_COEFFICIENT"], d["RESTITUTION"], d["TIMESTEP"], d["INCLINE_ANGLE_RAD"]) == (0.5, 0.9, 0.01, 0.0)
    assert set(d) == {"FRICTION_COEFFICIENT", "RESTITUTION", "TIMESTEP", "INCLINE_ANGLE_RAD", "RECORD_VIDEO", "CAMERA", "RECORDING_PATH"}


def test_reference_module_paths_importable():
    """`src.physics.collision` etc. resolve to the drop-in when rigidbody-simulation_b200/ is on sys.path."""
    code = ("import sys; sys.path.insert(0, 'rigidbody-simulation_b200');"
            "from src.physics.collision import compute_collision_impulse_friction, compute_inertia_tensor_world, "
            "custom_step_with_impulse_collision_friction, apply_impulse, apply_impulse_friction;"
            "from src.physics.physics_utils import apply_impulse, apply_impulse_friction;"
            "from src.physics.time_integeration import timestep_integration, general, compute_inertia_tensor_world;"
            "from src.simulation.ball_collision import compute_inverse_inertia, compute_collision_impulse, step_with_custom_collisions;"
            "from src.simulation.multi_sphere_bounce import custom_step_multi_sphere;"
            "from src.simulate import run_simulation, main;"
            "from src.config import load_sim_config;"
            "import inspect;"
            "s = inspect.signature(custom_step_with_impulse_collision_friction);"
            "assert list(s.parameters)[:7] == ['model','obj','data','dt','restitution','friction_coeff','contact_threshold'];"
            "assert [s.parameters[k].default for k in ('dt','restitution','friction_coeff','contact_threshold')] == [0.01,1.0,1.0,0];"
            "s = inspect.signature(timestep_integration);"
            "assert [s.parameters[k].default for k in ('dt','restitution','friction_coeff','contact_threshold')] == [0.01,1.0,0.5,1e-4];"
            "s = inspect.signature(general);"
            "assert [s.parameters[k].default for k in ('dt','restitution','friction_coeff','contact_threshold')] == [0.01,1.0,0.5,1e-4];"
            "print('ok')")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr


def test_cli_unknown_name_exits_1():
    code = ("import sys; sys.path.insert(0, 'rigidbody-simulation_b200');"
            "from src.simulate import main; main(['--sim', 'nope'])")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 1 and "Unknown simulation name" in r.stdout


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for GPU-less hosts")
    import rigidbody_simulation_b200 as rb
    from rigidbody_simulation_b200 import scenes, stepper
    with pytest.raises(rb.RbsError):
        rb.compute_collision_impulse_friction(1.0, None, np.zeros(3), np.zeros(3), np.zeros(3), np.array([0, 0, 1.0]), 1.0, 0.5)
    model = rb.BatchedModel.from_xml_path(scenes.model_path("sphere"), nenv=4, device="cpu")
    data = rb.BatchedData(model)
    with pytest.raises(rb.RbsError):
        stepper.step_body_plane(model, data, -1, 0.009, 1.0, 0.5, 0.0)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the product package may reference it."""
    pkg = os.path.join(ROOT, "rigidbody-simulation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "c_oracle" not in text and "pyport" not in text and "rb_oracle" not in text, f
                assert "import oracle" not in text and "from oracle" not in text, f


def test_batched_data_views_on_cpu():
    """AoS views over the SoA state (device-independent logic, exercised on CPU tensors)."""
    import torch
    import rigidbody_simulation_b200 as rb
    from rigidbody_simulation_b200 import scenes
    m1 = rb.BatchedModel.from_xml_path(scenes.model_path("sphere"), nenv=1, device="cpu")
    d1 = rb.BatchedData(m1)
    assert d1.qpos.shape == (7,) and np.asarray(d1.qpos).tolist() == [0, 0, 2.0, 1, 0, 0, 0]
    d1.qvel[3:6] = np.array([2.0, 2.0, 0.0])
    assert d1.qvel[3:6].tolist() == [2.0, 2.0, 0.0] and isinstance(d1.qvel[:3], np.ndarray)
    for layout, name in (("env", "ball_collision"), ("body", "multi_sphere")):
        m = rb.BatchedModel.from_xml_path(scenes.model_path(name), nenv=5, device="cpu")
        d = rb.BatchedData(m, layout=layout)
        B = m.nfree
        qp = np.arange(5 * 7 * B, dtype=np.float64).reshape(5, 7 * B)
        qv = -np.arange(5 * 6 * B, dtype=np.float64).reshape(5, 6 * B)
        d.set_state(qp, qv)
        assert (d.qpos.torch().numpy() == qp).all() and (d.qvel.torch().numpy() == qv).all()
        assert d.qpos[2, 7:10].tolist() == qp[2, 7:10].tolist()
        d.qvel[:, 0:3] = 0.0
        assert (d.qvel.torch().numpy()[:, 0:3] == 0).all() and (d.qvel.torch().numpy()[:, 3:] == qv[:, 3:]).all()
        mask = torch.tensor([True, False, False, True, False])
        with pytest.raises(rb.RbsError):          # reset(env_mask) is a CUDA kernel (rbs_reset_envs): no CPU path
            d.reset(mask)


def test_parser_orientation_specifiers_units_and_defaults():
    """N1 front-end: every MuJoCo orientation specifier, degree / radian, eulerseq (lower = rotating frame, upper =
    fixed frame), top-level geom defaults, explicit geom mass, <inertial> under inertiafromgeom auto / false --
    checked against SciPy's Rotation and the closed-form mass rules."""
    from scipy.spatial.transform import Rotation
    from rigidbody_simulation_b200 import mjcf

    def body_quat(compiler, attr):
        sc = mjcf.parse_string(f'<mujoco>{compiler}<worldbody><body {attr}><freejoint/><geom type="sphere" size="0.1"/></body>'
                               '</worldbody></mujoco>')
        return np.array(sc.bodies[1].quat)

    def same_rotation(q_wxyz, rot):
        want = rot.as_quat()[[3, 0, 1, 2]]
        return min(np.abs(q_wxyz - want).max(), np.abs(q_wxyz + want).max()) < 1e-14

    rad = '<compiler angle="radian"/>'
    e = [0.3, -1.1, 2.0]
    assert same_rotation(body_quat(rad, f'euler="{e[0]} {e[1]} {e[2]}"'), Rotation.from_euler("XYZ", e))       # SciPy: upper = intrinsic
    assert same_rotation(body_quat('<compiler angle="radian" eulerseq="XYZ"/>', f'euler="{e[0]} {e[1]} {e[2]}"'),
                         Rotation.from_euler("xyz", e))                                                         # fixed frame
    assert same_rotation(body_quat('<compiler angle="radian" eulerseq="zyx"/>', f'euler="{e[0]} {e[1]} {e[2]}"'),
                         Rotation.from_euler("ZYX", e))
    d = np.degrees(e)
    assert same_rotation(body_quat("", f'euler="{d[0]} {d[1]} {d[2]}"'), Rotation.from_euler("XYZ", e))          # default unit: degree
    assert same_rotation(body_quat("", 'axisangle="0 3 4 90"'), Rotation.from_rotvec(np.array([0, 0.6, 0.8]) * np.pi / 2))
    assert same_rotation(body_quat(rad, 'axisangle="1 0 0 0.7"'), Rotation.from_euler("x", 0.7))
    assert same_rotation(body_quat(rad, 'quat="2 0 0 2"'), Rotation.from_euler("z", np.pi / 2))
    R = Rotation.from_euler("XYZ", e).as_matrix()
    x, y = R[:, 0], R[:, 1]
    assert same_rotation(body_quat(rad, 'xyaxes="%s"' % " ".join(repr(float(c)) for c in (*(2 * x), *(y + 0.3 * x)))), Rotation.from_matrix(R))
    z = np.array([0.2, -0.5, 0.7]); z /= np.linalg.norm(z)
    qz = body_quat(rad, 'zaxis="%s"' % " ".join(repr(float(c)) for c in 3 * z))
    assert np.abs(Rotation.from_quat(qz[[1, 2, 3, 0]]).apply([0, 0, 1]) - z).max() < 1e-15
    assert abs(qz[3]) < 1e-15                                              # minimal rotation: no twist about z
    assert body_quat(rad, 'zaxis="0 0 -2"').tolist() == [0.0, 1.0, 0.0, 0.0]
    with pytest.raises(ValueError):
        body_quat(rad, 'euler="0 0 1" quat="1 0 0 0"')
    with pytest.raises(ValueError):
        body_quat(rad, 'quat="0 0 0 0"')

    sc = mjcf.parse_string('<mujoco><compiler angle="radian" inertiafromgeom="true"/><default><geom density="50" type="box" size="0.1 0.2 0.3"/>'
                           '</default><worldbody><geom type="plane" size="1 1 1" zaxis="0 -1 1"/>'
                           '<body><freejoint/><geom/><inertial pos="0 0 0" mass="9" diaginertia="1 1 1"/></body>'
                           '<body><freejoint/><geom type="sphere" size="0.2" mass="3"/></body></worldbody></mujoco>')
    m = 50 * 8 * 0.1 * 0.2 * 0.3
    assert sc.bodies[1].mass == pytest.approx(m, rel=1e-15)                # inertiafromgeom="true" overrides <inertial>
    assert sc.bodies[1].inertia == pytest.approx([m / 3 * 0.13, m / 3 * 0.10, m / 3 * 0.05], rel=1e-14)
    assert sc.bodies[2].mass == 3.0 and sc.bodies[2].inertia == pytest.approx([0.4 * 3 * 0.04] * 3, rel=1e-15)
    _, n = mjcf.plane_frame(sc, sc.planes()[0])
    assert n == pytest.approx([0, -math.sqrt(0.5), math.sqrt(0.5)], abs=1e-15)
    auto = mjcf.parse_string('<mujoco><worldbody><body><freejoint/><geom type="sphere" size="0.2"/>'
                             '<inertial pos="0 0 0" mass="9" diaginertia="1 2 3"/></body></worldbody></mujoco>')
    assert auto.bodies[1].mass == 9.0 and auto.bodies[1].inertia == [1.0, 2.0, 3.0]
    for bad in ('<mujoco><compiler inertiafromgeom="false"/><worldbody><body><freejoint/><geom type="sphere" size="0.2"/></body></worldbody></mujoco>',
                '<mujoco><worldbody><body><freejoint/><geom type="sphere" size="0.2"/><inertial pos="0 0 0.1" mass="1" diaginertia="1 1 1"/></body></worldbody></mujoco>',
                '<mujoco><worldbody><body><joint type="hinge"/><geom type="sphere" size="0.2"/></body></worldbody></mujoco>',
                '<mujoco><worldbody><body><freejoint/><geom type="box" size="0.2"/></body></worldbody></mujoco>',
                '<mujoco><default><default class="a"/></default><worldbody/></mujoco>'):
        with pytest.raises(ValueError):
            mjcf.parse_string(bad)
    # a geom offset from its free body's origin parses (N1 remainder); only the multi-body stepper takes such a scene
    import rigidbody_simulation_b200 as rb
    from rigidbody_simulation_b200 import stepper
    off = ('<mujoco><worldbody><geom type="plane" size="1 1 1"/><body pos="0 0 1"><freejoint/>'
           '<geom type="box" size="0.1 0.2 0.3" pos="0 0 0.1" euler="0.2 0 0" density="50"/></body></worldbody></mujoco>')
    sc = mjcf.parse_string(off.replace("<mujoco>", '<mujoco><compiler angle="radian"/>'))
    g = sc.geoms[sc.bodies[1].geoms[0]]
    assert g.pos == [0.0, 0.0, 0.1] and g.quat == pytest.approx([math.cos(0.1), math.sin(0.1), 0, 0], abs=1e-15)
    model = rb.BatchedModel(sc, nenv=2, device="cpu")
    assert model.has_offset_geoms
    tab = stepper.body_table(model)
    assert tab.shape == (1, 16) and tab[0, 0] == 1.0 and list(tab[0, 1:4]) == [0.1, 0.2, 0.3] and list(tab[0, 8:11]) == [0.0, 0.0, 0.1]
    assert tab[0, 15] == pytest.approx(math.sqrt(0.14) * (1 + 1e-6), rel=1e-15)
    for layout, fn, args in (("env", stepper.body_plane_args, (-1, 0.01, 1.0, 0.5, 0.0, 0, 1)),
                             ("body", stepper.multi_sphere_args, (0.01, 1.0, 0.0, 1))):
        with pytest.raises(ValueError, match="offset"):
            fn(model, rb.BatchedData(model, layout=layout), *args)
    with pytest.raises(rb.RbsError):                                       # and no CPU fallback for the one that does
        stepper.multi_body_args(model, rb.BatchedData(model, layout="body"), 0.01, 1.0, 0.5, 1)


def test_cli_help_and_headless_flags():
    code = ("import sys; sys.path.insert(0, 'rigidbody-simulation_b200');"
            "from src.simulate import main; main(['--help'])")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0
    for flag in ("--sim", "--headless", "--steps", "--envs", "--dtype", "--substeps-per-launch", "--seed", "--gpus", "--arith",
                 "--config", "--bodies", "--log"):
        assert flag in r.stdout


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints one JSON line with the agreed keys:
    the installed reference (baseline/_ref, kind "reference") when present, else the NumPy port (kind "port")."""
    import json
    have_ref = os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "physics", "collision.py"))
    for config in ("sphere_incline", "two_ball"):
        r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0", "--config", config,
                            "--cpu-seconds", "0.2"], cwd=ROOT, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        lines = r.stdout.strip().splitlines()
        assert len(lines) == 1, r.stdout                  # exactly one JSON line on stdout
        line = json.loads(lines[0])
        assert line["impl"] == "reference" and line["metric"] == "env-substeps/s" and line["unit"] == "env-substeps/s"
        assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
        assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and line["cpu_baseline"]["cores"] >= 1
        assert line["e2e"] == {"value": line["value"], "unit": "env-substeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        assert "workload" in line["config"] and "model" not in line["config"] and line["config"]["config"] == config


def test_installed_reference_and_port_agree():
    """When baseline/_ref exists (pip install --target of the unmodified reference), its step functions and the NumPy
    port produce the same numbers on the same sample -- the port is what stands in when the copy is absent."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_baseline as cb
    import pyport
    from rigidbody_simulation_b200 import synth
    if not cb.reference_installed():
        pytest.skip("baseline/_ref not installed")
    mj, col, ti, pu = cb._reference_modules()
    s = synth.sphere_incline(3)
    model = mj.MjModel.from_xml_string(pyport.single_body_xml("sphere", [0.2], plane_euler=(0.7, 0, 0)))
    for i in range(3):
        a, b = mj.MjData(model), mj.MjData(model)
        for d in (a, b):
            d.qpos[:], d.qvel[:] = s["qpos"][i], s["qvel"][i]
        for _ in range(150):
            col.custom_step_with_impulse_collision_friction(model, "obj", a, dt=0.009, restitution=s["restitution"][i],
                                                            friction_coeff=s["friction"][i], contact_threshold=0.0)
            pyport.step_scheme_a(model, "obj", b, dt=0.009, restitution=s["restitution"][i], friction_coeff=s["friction"][i],
                                 contact_threshold=0.0)
        assert np.array_equal(a.qpos, b.qpos) and np.array_equal(a.qvel, b.qvel)
    assert "site-packages" not in col.__file__ and os.path.join("baseline", "_ref") in col.__file__


def test_cli_random_configs_are_index_keyed_and_flags_reach_the_run():
    """`--seed S` selects the randomised BASELINE config of the scenario; a shard [start, start+count) holds exactly the
    rows of the whole batch (so `--gpus N` sees the same environments), for every scenario.  Built on a CPU device:
    state allocation is plumbing, stepping is not available there."""
    sys.path.insert(0, os.path.join(ROOT, "rigidbody-simulation_b200"))
    import torch
    from rigidbody_simulation_b200.src import simulate
    for sim, width in (("single_sphere", 7), ("cube_incline", 7), ("ball_collision", 14), ("multi_sphere", 7 * 5)):
        _, whole, _ = simulate.build_random(sim, 11, 0, 7, "cpu", torch.float64, bodies=5)
        _, part, _ = simulate.build_random(sim, 4, 6, 7, "cpu", torch.float64, bodies=5)
        _, other, _ = simulate.build_random(sim, 4, 6, 8, "cpu", torch.float64, bodies=5)
        qw, qp, qo = (np.asarray(d.qpos.torch()) for d in (whole, part, other))
        assert qw.shape == (11, width)
        assert np.array_equal(qw[6:10], qp) and not np.array_equal(qp, qo)
    # flag plumbing down to run_simulation (no CUDA here: the run itself refuses after validating its arguments)
    seen = {}
    orig = simulate.run_simulation
    simulate.run_simulation = lambda *a, **k: seen.update(args=a, kw=k)
    try:
        simulate.main(["--sim", "single_sphere", "--envs", "1048576", "--seed", "1", "--arith", "fast", "--gpus", "1",
                       "--substeps-per-launch", "256", "--steps", "2048"])
    finally:
        simulate.run_simulation = orig
    assert seen["args"] == ("single_sphere", 2048, 1048576, "fp64", 256)
    assert seen["kw"]["arith"] == "fast" and seen["kw"]["seed"] == 1 and seen["kw"]["config"] is None
    if not torch.cuda.is_available():
        with pytest.raises(SystemExit) as e:
            simulate.run_simulation("single_sphere", envs=4, seed=1, arith="fast")
        assert e.value.code == 1
    with pytest.raises(SystemExit):
        simulate.run_simulation("single_sphere", arith="sloppy")


def test_cli_gpus_mismatch_with_launcher_is_refused():
    code = ("import sys; sys.path.insert(0, '.');"
            "from rigidbody_simulation_b200.src.simulate import main; main(['--sim', 'single_sphere', '--gpus', '4'])")
    env = dict(os.environ, WORLD_SIZE="2", RANK="0", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, env=env)
    assert r.returncode == 1 and "does not match" in r.stdout


def test_free_function_scalar_arrays_never_alias_one_element():
    """a one-element array for a per-item scalar (mass=np.array([2.0])) next to N batched vectors must be
    treated as a uniform value, never as a device pointer the kernel would index with i < N; a wrong length raises."""
    import torch
    from rigidbody_simulation_b200.free_functions import _Batch
    b = object.__new__(_Batch)
    b.device, b.dtype, b.n, b.batched, b.keep, b.as_torch = torch.device("cpu"), torch.float64, None, False, [], False
    b.vec(np.zeros((5, 3)), (3,))
    assert b.n == 5
    assert b.scalar(np.array([2.0])) == (None, 2.0)
    assert b.scalar(torch.tensor([3.5])) == (None, 3.5)
    assert b.scalar(1.25) == (None, 1.25)
    ptr, val = b.scalar(np.arange(5.0))
    assert ptr is not None and val == 0.0
    with pytest.raises(ValueError):
        b.scalar(np.arange(3.0))


def test_tuning_options_api():
    import rigidbody_simulation_b200 as rb
    lib = rb._lib
    old = lib.set_option("pf_min_substeps", 1)
    assert old == 4 or old >= 1
    assert lib.get_option("pf_min_substeps") == 1
    lib.set_option("pf_min_substeps", old)
    assert lib.get_option("pf_min_substeps") == old
    for name in ("minb", "pf_packed", "strict_minb", "strict_compact", "strict_tb_minb", "strict_ms_regs", "box_minb", "box_compact", "tb_minb", "ms_skin_percent", "ms_kernel", "ms_walk_cost", "ms_tight_span", "ms_regs", "probe_mode", "host_chunks"):
        lib.get_option(name)
    with pytest.raises(ValueError):
        lib.get_option("no_such_knob")
    with pytest.raises(ValueError):
        lib.set_option("no_such_knob", 1)


def test_numa_binding_is_best_effort():
    """shard.bind_to_gpu_numa never raises: without a GPU (or without sysfs locality) it reports why it did nothing."""
    from rigidbody_simulation_b200 import shard
    before = os.sched_getaffinity(0)
    info = shard.bind_to_gpu_numa(0)
    assert isinstance(info, dict) and "bound" in info
    if not info["bound"]:
        assert os.sched_getaffinity(0) == before and "note" in info
    else:
        os.sched_setaffinity(0, before)
    assert shard._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}


def test_missing_extension_fails_loudly():
    """If librbsim_b200.so has not been built the package refuses to work -- it never computes some other way."""
    code = ("import sys; sys.path.insert(0, '.');"
            "import rigidbody_simulation_b200 as rb;"
            "rb._lib.LIB_PATH = rb._lib.LIB_PATH + '.absent'; rb._lib._lib = None\n"
            "try:\n    rb._lib.load()\nexcept rb.RbsError as e:\n    print('raised', 'no CPU fallback' in str(e).lower() or 'fallback' in str(e))\n"
            "import numpy as np\n"
            "try:\n    rb.compute_inverse_inertia(1.0, 0.1); rb.compute_collision_impulse(1.0, 2500.0, np.zeros(3), np.zeros(3), np.zeros(3), np.array([0,0,1.0]), 1.0, 0.3)\n"
            "except rb.RbsError:\n    print('free function refused')\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "raised True" in r.stdout and "free function refused" in r.stdout


def test_checkpoint_round_trip_cpu(tmp_path):
    import torch
    import rigidbody_simulation_b200 as rb
    from rigidbody_simulation_b200 import scenes
    m = rb.BatchedModel.from_xml_path(scenes.model_path("ball_collision"), nenv=6, device="cpu")
    d = rb.BatchedData(m)
    d.set_state(np.random.default_rng(0).normal(size=(6, 14)), np.random.default_rng(1).normal(size=(6, 12)))
    d.n_contacts += 3
    torch.save(d.state_dict(), tmp_path / "ckpt.pt")
    d2 = rb.BatchedData(m)
    d2.load_state_dict(torch.load(tmp_path / "ckpt.pt"))
    assert torch.equal(d2.state, d.state) and torch.equal(d2.n_contacts, d.n_contacts)
    m3 = rb.BatchedModel.from_xml_path(scenes.model_path("sphere"), nenv=6, device="cpu")
    with pytest.raises(ValueError):
        rb.BatchedData(m3).load_state_dict(d.state_dict())


def test_bit_pattern_comparisons_decide_like_the_fp_tests():
    """rbs_kernels.cuh moves the always-executed comparisons of the fused double kernels off the FP64 pipe by comparing
    bit patterns (below_nonneg, above_positive, clamp_to_minus_one, sign_bit).  NumPy restatement of those integer
    tests against the floating-point tests they replace, over random and edge operands: they must agree everywhere
    except the documented cases (x = -0.0 against y = +0.0; -0.0 under sign_bit; NaN)."""
    rng = np.random.default_rng(11)
    edge = np.array([0.0, -0.0, 5e-324, -5e-324, 2.2250738585072014e-308, -2.2250738585072014e-308, 1e-12, 1.0000000000000002e-12,
                     9.999999999999998e-13, 1.0, -1.0, np.nextafter(-1.0, 0.0), np.nextafter(-1.0, -2.0), 0.2, np.nextafter(0.2, 1.0),
                     np.nextafter(0.2, 0.0), 1e308, -1e308, np.inf, -np.inf])
    x = np.concatenate([edge, rng.normal(size=20000) * 10.0 ** rng.integers(-300, 300, 20000), rng.uniform(-2, 2, 20000)])
    bits = lambda a: np.asarray(a, np.float64).view(np.int64)
    # below_nonneg(x, y) = x < y for y >= +0
    for y in (0.0, 5e-324, 0.2, 1.0, 1e300, np.inf):
        got = bits(x) < bits(np.float64(y))
        want = x < y
        differ = got != want
        assert (differ <= ((x == 0) & np.signbit(x) & (y == 0.0))).all(), y      # only -0.0 against +0.0
    # above_positive(x, c) = x > c for x >= 0, c > 0
    xp = np.abs(x)
    for c in (1e-12, 1e-16, 5e-324, 1.0):
        assert ((bits(xp) > bits(np.float64(c))) == (xp > c)).all(), c
    # clamp_to_minus_one(c) = c > -1 ? c : -1 for c <= 0: |c| < 1 read off the high word
    xn = -np.abs(x)
    hi = (bits(xn) >> 32) & 0x7FFFFFFF
    got = np.where(hi < 0x3FF00000, xn, -1.0)
    want = np.where(xn > -1.0, xn, -1.0)
    assert np.array_equal(got, want)
    nan_hi = (bits(np.float64(np.nan)) >> 32) & 0x7FFFFFFF
    assert not (nan_hi < 0x3FF00000)                                             # NaN -> -1, like the FP test
    # sign_bit(x) against !(x >= 0): differ only for -0.0
    got = (bits(x) >> 32).astype(np.int32) < 0
    want = ~(x >= 0)
    assert ((got != want) <= ((x == 0) & np.signbit(x))).all()
    # the float kernels keep the FP forms; the same identities hold for binary32 (used by nobody, checked for the record)
    xf = x[np.abs(x) < 1e38].astype(np.float32)
    fb = lambda a: np.asarray(a, np.float32).view(np.int32)
    got, want = fb(xf) < fb(np.float32(0.2)), xf < np.float32(0.2)
    assert np.array_equal(got, want)
