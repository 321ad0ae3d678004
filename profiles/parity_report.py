"""Parity report: GPU (both arithmetic policies) vs the CPU oracle on the same seeded inputs, per BASELINE config:
max relative deviation (|gpu-ref| / max(|ref|, 1e-3)) as a function of step count, and the number of environments
whose contact-event counters differ from the oracle's.  Writes a markdown table to stdout.
    python profiles/parity_report.py > gpurun_out/parity_report.md
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

import c_oracle as co
import rigidbody_simulation_b200 as rb
import rigidbody_simulation_b200.mj as mj
from helpers import comp_rel_err
from rigidbody_simulation_b200 import scenes, stepper, synth
from rigidbody_simulation_b200.src.simulation import ball_collision, multi_sphere_bounce

G = [0.0, 0.0, -9.8]
CHECK = (1, 10, 100, 1000)


def state_of(data):
    return data.qpos.torch().cpu().numpy(), data.qvel.torch().cpu().numpy()


def row(name, policy, errs, mism, E):
    cells = " | ".join(f"{errs[s]:.1e}" if s in errs else "—" for s in CHECK)
    print(f"| {name} | {policy} | {E} | {cells} | {mism} |", flush=True)


print("| config | policy | envs | step 1 | step 10 | step 100 | step 1000 | envs with differing event counts (whole horizon) |")
print("|---|---|---|---|---|---|---|---|")

# config 2 / config 1-like sphere, config 4 cubes ------------------------------------------------------------------
for name, geom, gen, E in (("cfg2 sphere on incline", "sphere", lambda n: synth.sphere_incline(n), 100_000),
                           ("cfg4 cube bounce", "box", lambda n: synth.cube(n, kind="bounce"), 50_000),
                           ("cfg4 cube incline", "box", lambda n: synth.cube(n, kind="incline"), 50_000)):
    s = gen(E)
    size = [s["radius"]] if geom == "sphere" else s["half"]
    e = s["restitution"] if np.ndim(s["restitution"]) else np.full(E, s["restitution"])
    mu = s["friction"] if np.ndim(s["friction"]) else np.full(E, s["friction"])
    for policy in ("strict, isotropic shortcut", "strict (default: literal inertia)", "fast"):
        model = mj.MjModel.from_xml_string(scenes.single_body_xml(geom, size, plane_euler=(s.get("theta", 0.7), 0, 0)), nenv=E)
        model.set_per_env(restitution=e, friction=mu)
        data = mj.MjData(model)
        data.set_state(s["qpos"], s["qvel"])
        qp, qv = s["qpos"].copy(), s["qvel"].copy()
        cnt = (np.zeros(E, np.uint32), np.zeros(E, np.uint32))
        kw = dict(geom=geom, mass=model.body_mass[-1], inertia=model.body_inertia[-1], size=size if geom == "box" else size[0],
                  plane_pos=[0, 0, 0], plane_normal=model.plane_normal, gravity=G, dt=s["dt"], restitution=e, friction=mu,
                  threshold=s["threshold"], counters=cnt)
        errs, done = {}, 0
        for upto in CHECK:
            co.step_body_plane(qp, qv, upto - done, **kw)
            stepper.step_body_plane(model, data, -1, s["dt"], None, None, s["threshold"], substeps=upto - done,
                                    strict_inertia="literal" in policy, arith="fast" if policy == "fast" else "strict")
            done = upto
            gq, gv = state_of(data)
            errs[upto] = max(comp_rel_err(gq, qp, 1e-3), comp_rel_err(gv, qv, 1e-3))
        calls, imps = data.counters()
        row(name, policy, errs, int(((calls[:, 0] != cnt[0]) | (imps[:, 0] != cnt[1])).sum()), E)

# config 3 two balls ---------------------------------------------------------------------------------------------
E = 100_000
s = synth.two_ball(E)
model, data = ball_collision.build(E)
data.set_state(s["qpos"], s["qvel"])
qp, qv = s["qpos"].copy(), s["qvel"].copy()
hits = (np.zeros(E, np.uint32), np.zeros(E, np.uint32))
m = float(model.body_mass[1])
errs, done = {}, 0
for upto in CHECK:
    co.step_two_ball(qp, qv, upto - done, mass=[m, m], radius=0.1, gravity=G, dt=0.01, restitution=1.0, friction=0.3, counters=hits)
    stepper.step_two_ball(model, data, 0.01, 1.0, 0.3, radius=0.1, substeps=upto - done)
    done = upto
    gq, gv = state_of(data)
    errs[upto] = max(comp_rel_err(gq, qp, 1e-3), comp_rel_err(gv, qv, 1e-3))
mism = int(((data.n_contacts[:E].cpu().numpy() != hits[0]) | (data.n_impulses[:E].cpu().numpy() != hits[1])).sum())
row("cfg3 two balls", "strict (default)", errs, mism, E)
model, data = ball_collision.build(E)
data.set_state(s["qpos"], s["qvel"])
qp, qv = s["qpos"].copy(), s["qvel"].copy()
hits = (np.zeros(E, np.uint32), np.zeros(E, np.uint32))
errs, done = {}, 0
for upto in CHECK:
    co.step_two_ball(qp, qv, upto - done, mass=[m, m], radius=0.1, gravity=G, dt=0.01, restitution=1.0, friction=0.3, counters=hits)
    stepper.step_two_ball(model, data, 0.01, 1.0, 0.3, radius=0.1, substeps=upto - done, arith="fast")
    done = upto
    gq, gv = state_of(data)
    errs[upto] = max(comp_rel_err(gq, qp, 1e-3), comp_rel_err(gv, qv, 1e-3))
mism = int(((data.n_contacts[:E].cpu().numpy() != hits[0]) | (data.n_impulses[:E].cpu().numpy() != hits[1])).sum())
row("cfg3 two balls", "fast", errs, mism, E)

# config 5 multi sphere ------------------------------------------------------------------------------------------
# (spin components are measured against a floor of 0.1 rad/s: where a spin component is analytically zero the oracle's
# literal inv(R diag(I) R^T) @ (arm x J) leaves ~3e-14 of rounding noise, 1/I = 1194, which the fast kernels do not have)
E, B = 2000, 64
for mu in (0.3, 0.0):
    s = synth.multi_sphere(E, n_body=B, friction=mu)
    for policy in ("strict", "fast"):                       # strict = default = literal inertia
        model, data = multi_sphere_bounce.build(E, n_body=B)
        data.set_state(s["qpos"], s["qvel"])
        qp, qv = s["qpos"].reshape(E, B, 7).copy(), s["qvel"].reshape(E, B, 6).copy()
        cnt = (np.zeros((E, B), np.uint32), np.zeros((E, B), np.uint32))
        errs, done = {}, 0
        for upto in (1, 10, 100):
            co.step_multi_sphere(qp, qv, upto - done, mass=model.body_mass[1], inertia=model.body_inertia[1], radius=0.1,
                                 plane_pos=[0, 0, 0], plane_normal=[0, 0, 1], gravity=G, dt=0.01, restitution=1.0, friction=mu, counters=cnt)
            stepper.step_multi_sphere(model, data, 0.01, 1.0, mu, substeps=upto - done, arith=policy)
            done = upto
            gq, gv = state_of(data)
            g6, r6 = gv.reshape(E, B, 6), qv
            errs[upto] = max(comp_rel_err(gq.ravel(), qp.ravel(), 1e-3), comp_rel_err(g6[:, :, :3].ravel(), r6[:, :, :3].ravel(), 1e-3),
                             comp_rel_err(g6[:, :, 3:].ravel(), r6[:, :, 3:].ravel(), 1e-1))
        calls, imps = data.counters()
        row(f"cfg5 64 spheres (mu={mu}), horizon 100", "strict (default: literal inertia)" if policy == "strict" else policy, errs,
            int(((calls != cnt[0]).any(axis=1) | (imps != cnt[1]).any(axis=1)).sum()), E)
