"""Generate tests/golden/*.json by running the UNMODIFIED reference under the fake MuJoCo backend.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs BASELINE.json:reference_path):

    python oracle/make_golden.py            # rewrites every fixture (deterministic, seeded)

Each fixture stores inputs AND outputs as JSON doubles (Python repr round-trips float64 exactly),
so the tests on the GPU box need neither the reference nor this script.
"""
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_runner as rr  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
SEED = 20261018
SNAP = (1, 10, 100)


def dump(name, obj):
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name + ".json")
    with open(path, "w") as f:
        json.dump(obj, f, separators=(",", ":"))
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


def unit_quat(rng, n):
    q = rng.normal(size=(n, 4))
    return q / np.linalg.norm(q, axis=1, keepdims=True)


# ------------------------------------------------------------------------------- free functions
def free_function_vectors():
    rng = np.random.default_rng(SEED)
    m, I = 1.675516081914557, 0.026808257310632914
    Iw = I * np.eye(3)
    kats = [
        dict(name="stick", e=1.0, mu=0.5, v=[0.3, -0.2, -1.5], w=[2, 2, 0.5], r=[0, 0, -0.19], n=[0, 0, 1]),
        dict(name="low_mu", e=0.9, mu=0.05, v=[0.3, -0.2, -1.5], w=[2, 2, 0.5], r=[0, 0, -0.19], n=[0, 0, 1]),
        dict(name="separating", e=1.0, mu=0.5, v=[0, 0, 0.1], w=[2, 2, 0.5], r=[0, 0, -0.19], n=[0, 0, 1]),
        dict(name="no_tangential", e=0.2, mu=0.6, v=[0, 0, -1], w=[0, 0, 0], r=[0, 0, -0.19], n=[0, 0, 1]),
        dict(name="coulomb_clamp", e=0.9, mu=0.01, v=[0.3, -0.2, -1.5], w=[2, 2, 0.5], r=[0, 0, -0.19], n=[0, 0, 1]),
        dict(name="inclined", e=0.9, mu=0.5, v=[0.3, -0.2, -1.5], w=[2, 2, 0.5],
             r=(-0.195 * np.array([0, -math.sin(0.7), math.cos(0.7)])).tolist(),
             n=[0, -math.sin(0.7), math.cos(0.7)]),
    ]
    out = {"mass": m, "inertia": I, "kats": [], "random": []}
    with rr.reference_imports():
        F = rr.free_functions()
        for k in kats:
            v, w, r, n = (np.array(k[x], dtype=float) for x in "vwrn")
            jn, jt = F["A1"](m, Iw, v, w, r, n, k["e"], k["mu"])
            v2, w2 = F["A2"](v, w, m, Iw, r, n, jn, jt)
            out["kats"].append(dict(k, v=v.tolist(), w=w.tolist(), jn=float(jn), jt=np.asarray(jt).tolist(),
                                    v_out=v2.tolist(), w_out=w2.tolist()))
        v, w, r, n = (np.array(x, dtype=float) for x in ([0.3, -0.2, -1.5], [2, 2, 0.5], [0, 0, -0.19], [0, 0, 1]))
        v3, w3 = F["A3"](v, w, m, Iw, r, n, 0.7)
        out["A3"] = dict(v=v.tolist(), w=w.tolist(), r=r.tolist(), n=n.tolist(), impulse=0.7,
                         v_out=v3.tolist(), w_out=w3.tolist())
        idg, q = np.array([1.0, 2.0, 3.0]), np.array([0.9, 0.1, -0.3, 0.2])
        out["A4"] = dict(inertia_diag=idg.tolist(), q=q.tolist(), out=F["A4"](idg, q).tolist())
        # randomized batch: anisotropic inertia, arbitrary normals, both branches of every test
        for i in range(256):
            mass = float(rng.uniform(0.1, 30.0))
            idg = rng.uniform(0.01, 3.0, size=3)
            q = rng.normal(size=4) * rng.uniform(0.5, 2.0)       # un-normalised on purpose
            Iwr = F["A4"](idg, q)
            n = rng.normal(size=3)
            n /= np.linalg.norm(n)
            v = rng.uniform(-2, 2, size=3)
            w = rng.uniform(-5, 5, size=3)
            r = rng.uniform(-0.5, 0.5, size=3)
            if i % 8 == 0:                                      # pure normal approach: |u_t| <= 1e-6 branch
                w = np.zeros(3)
                v = -abs(rng.uniform(0.1, 2)) * n
            e, mu = float(rng.uniform(0, 1)), float(rng.uniform(0, 1.0))
            jn, jt = F["A1"](mass, Iwr, v, w, r, n, e, mu)
            v2, w2 = F["A2"](v, w, mass, Iwr, r, n, jn, jt)
            imp = float(rng.uniform(0, 3))
            v3, w3 = F["A3"](v, w, mass, Iwr, r, n, imp)
            out["random"].append(dict(mass=mass, inertia_diag=idg.tolist(), q=q.tolist(), Iw=Iwr.tolist(),
                                      n=n.tolist(), v=v.tolist(), w=w.tolist(), r=r.tolist(), e=e, mu=mu,
                                      jn=float(jn), jt=np.asarray(jt).tolist(), v_out=v2.tolist(),
                                      w_out=w2.tolist(), impulse=imp, v3=v3.tolist(), w3=w3.tolist()))
    # A10 lives in a script namespace: fetch it by running ball_collision.py with a zero step budget
    import runpy, shutil, tempfile
    tmp = tempfile.mkdtemp()
    cwd = os.getcwd()
    try:
        shutil.copytree(os.path.join(rr.reference_path(), "models"), os.path.join(tmp, "models"))
        os.chdir(tmp)
        with rr.reference_imports():
            import glfw
            glfw.reset(0)
            g = runpy.run_path(os.path.join(rr.reference_path(), "src", "simulation", "ball_collision.py"))
            f10 = g["compute_collision_impulse"]
            mb, ii = 0.20943951023931962, 1193.6620731892144
            J = f10(mb, ii * np.eye(3), np.array([1, 0.2, -0.5]), np.array([0.5, -1, 2.0]), np.array([0.1, 0, 0]),
                    np.array([1.0, 0, 0]), 1.0, 0.3)
            out["A10_kat"] = dict(mass=mb, inv_inertia=ii, v=[1, 0.2, -0.5], w=[0.5, -1, 2.0], r=[0.1, 0, 0],
                                  n=[1.0, 0, 0], e=1.0, mu=0.3, J=J.tolist())
            out["A10_inv_inertia_01"] = float(g["compute_inverse_inertia"](mb, 0.1)[0, 0])
            out["A10_random"] = []
            for i in range(128):
                mass = float(rng.uniform(0.05, 5))
                iinv = float(rng.uniform(10, 2000))
                n = rng.normal(size=3)
                n /= np.linalg.norm(n)
                v, w, r = rng.uniform(-2, 2, 3), rng.uniform(-5, 5, 3), rng.uniform(-0.2, 0.2, 3)
                if i % 8 == 0:
                    w = np.zeros(3)
                    v = rng.uniform(-2, 2) * n                  # t_norm <= 1e-8 branch
                e, mu = float(rng.uniform(0, 1)), float(rng.uniform(0, 1))
                J = f10(mass, iinv * np.eye(3), v, w, r, n, e, mu)
                out["A10_random"].append(dict(mass=mass, inv_inertia=iinv, v=v.tolist(), w=w.tolist(), r=r.tolist(),
                                              n=n.tolist(), e=e, mu=mu, J=J.tolist()))
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)
    dump("free_functions", out)


# ------------------------------------------------------------------------------- shipped scripts
def shipped_scripts():
    o = rr.run_script("single_sphere_bounce", 2000)
    z, t = np.array(o["log_z"]), np.array(o["log_t"])
    peaks = [[float(t[i]), float(z[i])] for i in range(1, len(z) - 1) if z[i] > z[i - 1] and z[i] >= z[i + 1]]
    dump("script_single_sphere_2000", dict(
        steps=2000, dt=0.009, e=1.0, mu=0.5, thr=0.0, radius=0.2, qpos0=[0, 0, 2.0, 1, 0, 0, 0],
        qvel0=[0, 0, 0, 2.0, 2.0, 0], qpos=o["qpos"], qvel=o["qvel"], calls=o["calls"], impulses=o["impulses"],
        peaks=peaks[:8], z_every_50=z[49::50].tolist()))
    o = rr.run_script("cube_incline", 240)
    z, t = np.array(o["log_z"]), np.array(o["log_t"])
    dump("script_cube_incline_240", dict(
        steps=240, dt=0.009, e=0.2, mu=0.6, thr=1e-4, half=[0.4, 0.4, 0.4], incline=0.7, qpos=o["qpos"],
        qvel=o["qvel"], calls=o["calls"], impulses=o["impulses"], log_z=z.tolist(), log_t=t.tolist(),
        log_x=o["log_x"], log_y=o["log_y"]))
    o = rr.run_script("ball_collision", 500, press_space=True)
    dump("script_ball_collision_500", dict(
        steps=500, dt=0.01, e=1.0, mu=0.3, radius=0.1, qpos=o["qpos"], qvel=o["qvel"],
        ball1_z=o["logger_ball1_z"][::10], ball2_z=o["logger_ball2_z"][::10], ball1_x=o["logger_ball1_x"][::10]))


def model_values():
    """XML -> mass / inertia / qpos0 / plane normal as the fake compiler produces them (Appendix A.1)."""
    out = {}
    with rr.reference_imports():
        import mujoco as mj
        for name in ("sphere", "cube", "ball_collision", "multi_sphere"):
            m = mj.MjModel.from_xml_path(os.path.join(rr.reference_path(), "models", name + ".xml"))
            d = mj.MjData(m)
            mj.mj_forward(m, d)
            planes = [g for g in m.geoms if g.type == "plane"]
            pR = mj._quat_to_mat(planes[0].quat)
            out[name] = dict(body_names=m.body_names, body_mass=m.body_mass.tolist(),
                             body_inertia=m.body_inertia.tolist(), qpos0=m.qpos0.tolist(),
                             gravity=m.opt.gravity.tolist(), timestep=m.opt.timestep,
                             plane_normal=pR[:, 2].tolist(), ncon0=d.ncon,
                             contact0=[dict(dist=c.dist, pos=c.pos.tolist(), n=c.frame[:3].tolist()) for c in d.contact])
    dump("model_values", out)


# ------------------------------------------------------------------------------- randomized envs
def sphere_incline_random():
    """cfg2-style: models/sphere.xml body on a plane tilted 0.7 rad about x, A5, randomized ICs."""
    rng = np.random.default_rng(SEED + 2)
    n_env, steps, theta = 48, 400, 0.7
    nrm = np.array([0.0, -math.sin(theta), math.cos(theta)])
    xml = rr.single_body_xml("sphere", [0.2], plane_euler=(theta, 0, 0))
    envs = []
    for i in range(n_env):
        h = rng.uniform(0.25, 2.5)
        off = rng.uniform(-1, 1, size=2)
        # in-plane basis (x, and n x x)
        tx = np.array([1.0, 0, 0])
        ty = np.cross(nrm, tx)
        p = h * nrm + off[0] * tx + off[1] * ty
        q = unit_quat(rng, 1)[0]
        v, w = rng.uniform(-2, 2, 3), rng.uniform(-5, 5, 3)
        e, mu = float(rng.uniform(0.5, 1.0)), float(rng.uniform(0, 1))
        qpos0, qvel0 = np.concatenate([p, q]), np.concatenate([v, w])
        o = rr.run_single_body("custom", xml, qpos0, qvel0, steps, 0.009, e, mu, 0.0, snapshots=SNAP)
        envs.append(dict(qpos0=qpos0.tolist(), qvel0=qvel0.tolist(), e=e, mu=mu, qpos=o["qpos"], qvel=o["qvel"],
                         calls=o["calls"], impulses=o["impulses"], snapshots=o["snapshots"]))
    dump("sphere_incline_random", dict(steps=steps, dt=0.009, thr=0.0, radius=0.2, theta=theta,
                                       plane_normal=o["plane_normal"], mass=o["mass"], inertia=o["inertia"], envs=envs))


def cube_random():
    """cfg4-style: models/cube.xml body; 'bounce' (flat plane, random pose/velocity) and 'incline'
    (0.7 rad, shipped pose + perturbation, from rest); A6 with its default threshold 1e-4."""
    rng = np.random.default_rng(SEED + 4)
    out = {}
    for kind, theta, n_env, steps in (("bounce", 0.0, 24, 300), ("incline", 0.7, 16, 240)):
        nrm = np.array([0.0, -math.sin(theta), math.cos(theta)])
        xml = rr.single_body_xml("box", [0.4, 0.4, 0.4], plane_euler=(theta, 0, 0), body_euler=(theta, 0, 0))
        envs = []
        for i in range(n_env):
            if kind == "bounce":
                p = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(0.8, 2.0)])
                q = unit_quat(rng, 1)[0]
                v, w = rng.uniform(-1, 1, 3), rng.uniform(-3, 3, 3)
            else:
                p = np.array([0, 0, 0.4]) + rng.uniform(-0.02, 0.02, 3)
                q = np.array([math.cos(0.35), math.sin(0.35), 0, 0]) + rng.uniform(-0.01, 0.01, 4)
                q /= np.linalg.norm(q)
                v, w = np.zeros(3), np.zeros(3)
            qpos0, qvel0 = np.concatenate([p, q]), np.concatenate([v, w])
            o = rr.run_single_body("timestep", xml, qpos0, qvel0, steps, 0.009, 0.2, 0.6, 1e-4, snapshots=SNAP)
            envs.append(dict(qpos0=qpos0.tolist(), qvel0=qvel0.tolist(), qpos=o["qpos"], qvel=o["qvel"],
                             calls=o["calls"], impulses=o["impulses"], snapshots=o["snapshots"]))
        out[kind] = dict(steps=steps, dt=0.009, thr=1e-4, e=0.2, mu=0.6, half=[0.4, 0.4, 0.4], theta=theta,
                         plane_normal=o["plane_normal"], mass=o["mass"], inertia=o["inertia"], envs=envs)
    dump("cube_random", out)


def general_and_xfrc():
    """A7 (scheme B) on sphere and an anisotropic box, and A5 with a non-zero applied wrench."""
    rng = np.random.default_rng(SEED + 7)
    out = {"general": [], "xfrc": []}
    for geom, size in (("sphere", [0.2]), ("box", [0.3, 0.2, 0.1])):
        xml = rr.single_body_xml(geom, size, plane_euler=(0.3, 0, 0))
        nrm = [0.0, -math.sin(0.3), math.cos(0.3)]
        for i in range(6):
            p = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(0.3, 1.0)])
            qpos0 = np.concatenate([p, unit_quat(rng, 1)[0]])
            qvel0 = np.concatenate([rng.uniform(-1, 1, 3), rng.uniform(-3, 3, 3)])
            o = rr.run_single_body("general", xml, qpos0, qvel0, 150, 0.01, 0.7, 0.4, 1e-4, snapshots=SNAP)
            out["general"].append(dict(geom=geom, size=size, plane_normal=o["plane_normal"], qpos0=qpos0.tolist(),
                                       qvel0=qvel0.tolist(), steps=150, dt=0.01, e=0.7, mu=0.4, thr=1e-4,
                                       mass=o["mass"], inertia=o["inertia"], qpos=o["qpos"], qvel=o["qvel"],
                                       calls=o["calls"], impulses=o["impulses"], snapshots=o["snapshots"]))
            xf = np.concatenate([rng.uniform(-3, 3, 3), rng.uniform(-0.2, 0.2, 3)])
            o = rr.run_single_body("custom", xml, qpos0, qvel0, 150, 0.01, 0.7, 0.4, 0.0, snapshots=SNAP, xfrc=xf)
            out["xfrc"].append(dict(geom=geom, size=size, plane_normal=o["plane_normal"], qpos0=qpos0.tolist(),
                                    qvel0=qvel0.tolist(), xfrc=xf.tolist(), steps=150, dt=0.01, e=0.7, mu=0.4,
                                    thr=0.0, mass=o["mass"], inertia=o["inertia"], qpos=o["qpos"], qvel=o["qvel"],
                                    calls=o["calls"], impulses=o["impulses"], snapshots=o["snapshots"]))
    dump("general_and_xfrc", out)


def two_ball_random():
    rng = np.random.default_rng(SEED + 3)
    envs = []
    for i in range(32):
        d = rng.uniform(-0.05, 0.05, size=(2, 3))
        qpos0 = np.array([-1 + d[0, 0], d[0, 1], 1 + d[0, 2], 1, 0, 0, 0, 1 + d[1, 0], d[1, 1], 1 + d[1, 2], 1, 0, 0, 0])
        v1 = np.array([1, 0, 0.5]) + rng.uniform(-0.2, 0.2, 3)
        v2 = np.array([-1, 0, 0.5]) + rng.uniform(-0.2, 0.2, 3)
        qvel0 = np.concatenate([v1, rng.uniform(-2, 2, 3), v2, rng.uniform(-2, 2, 3)])
        o = rr.run_two_ball(qpos0, qvel0, 400, 0.01, snapshots=SNAP)
        envs.append(dict(qpos0=qpos0.tolist(), qvel0=qvel0.tolist(), qpos=o["qpos"], qvel=o["qvel"],
                         snapshots=o["snapshots"]))
    dump("two_ball_random", dict(steps=400, dt=0.01, e=1.0, mu=0.3, radius=0.1, mass=o["mass"],
                                 inv_inertia=o["inv_inertia"], envs=envs))


def multi_sphere_cases():
    out = {}
    # shipped 4-ball scene (models/multi_sphere.xml ICs): purely vertical bouncing
    xml = rr.multi_sphere_xml(4)
    qpos0 = np.zeros((4, 7))
    qpos0[:, :3] = [[-1.5, -1.5, 2], [1.5, -1.5, 2], [-1.5, 1.5, 2], [1.5, 1.5, 2]]
    qpos0[:, 3] = 1
    o = rr.run_multi_sphere(xml, qpos0.ravel(), np.zeros(24), 300, 0.01, 1.0, 0.0, snapshots=SNAP)
    out["shipped4"] = dict(B=4, steps=300, dt=0.01, e=1.0, mu=0.0, radius=0.1, qpos0=qpos0.ravel().tolist(),
                           qvel0=[0.0] * 24, **o)
    # dense synthetic scenes (cfg5-style jittered lattice) so that ball-ball contacts happen
    rng = np.random.default_rng(SEED + 5)
    dense = []
    for B, grid, mu in ((8, (2, 2, 2), 0.0), (8, (2, 2, 2), 0.3), (27, (3, 3, 3), 0.3)):
        xml = rr.multi_sphere_xml(B)
        for rep in range(3):
            idx = np.stack(np.meshgrid(*[np.arange(g) for g in grid], indexing="ij"), -1).reshape(-1, 3)
            pos = idx * 0.3 + rng.uniform(-0.04, 0.04, size=(B, 3))
            pos[:, 2] += 0.3
            q0 = np.zeros((B, 7))
            q0[:, :3] = pos
            q0[:, 3] = 1
            v0 = np.zeros((B, 6))
            v0[:, :3] = rng.uniform(-1, 1, size=(B, 3))
            o = rr.run_multi_sphere(xml, q0.ravel(), v0.ravel(), 200, 0.01, 1.0, mu, snapshots=SNAP)
            dense.append(dict(B=B, steps=200, dt=0.01, e=1.0, mu=mu, radius=0.1, qpos0=q0.ravel().tolist(),
                              qvel0=v0.ravel().tolist(), **o))
    out["dense"] = dense
    dump("multi_sphere", out)


def main():
    if not rr.reference_available():
        sys.exit("reference checkout not found at " + rr.reference_path())
    free_function_vectors()
    shipped_scripts()
    model_values()
    sphere_incline_random()
    cube_random()
    general_and_xfrc()
    two_ball_random()
    multi_sphere_cases()


if __name__ == "__main__":
    main()
