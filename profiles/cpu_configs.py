"""CPU baseline (Python/NumPy port of the reference step under the fake MuJoCo, one process per host core) for every
BASELINE config, to sit beside profiles/bench_configs.py.
    python profiles/cpu_configs.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import cpu_baseline
    from rigidbody_simulation_b200 import synth
    cores = os.cpu_count() or 1
    out = []
    r = cpu_baseline.python_port_sphere_incline(synth.sphere_incline(cores * 16), cores, 16, 400)
    out.append(("cfg2 sphere on incline (A5)", r))
    r = cpu_baseline.python_port_config("two_ball", synth.two_ball(cores * 16), cores, 16, 400)
    out.append(("cfg3 two balls (A11)", r))
    for kind in ("bounce", "incline"):
        s = synth.cube(cores * 8, kind=kind)
        r = cpu_baseline.python_port_config("cube", s, cores, 8, 150, extra={"theta": s["theta"]})
        out.append((f"cfg4 cube {kind} (A6)", r))
    r = cpu_baseline.python_port_config("multi_sphere", synth.multi_sphere(cores * 2, n_body=64), cores, 2, 25,
                                        extra={"n_body": 64, "friction": 0.0})
    out.append(("cfg5 64 spheres (A9, repaired)", r))
    for name, r in out:
        print(json.dumps({"config": name, "env_substeps_per_s": r["value"], "cores": r["cores"], "sample": r["sample"]}), flush=True)


if __name__ == "__main__":
    main()
