"""Headless counterpart of the reference's CLI ``src/simulate.py``.

The reference maps ``--sim NAME`` to a script and runs it in a subprocess (:13-19, :37); it exits with status 1 on
an unknown name (:21-26).  Here the same names select the same scenarios, run in-process on the GPU with no
window.  New flags: --headless (implied), --steps, --envs, --dtype, --substeps-per-launch, --seed, --gpus (use
``torchrun --nproc-per-node N`` for N > 1: one process per GPU, environments sharded, no collective on the step
path).  ``compare_builtin`` exercises MuJoCo's own constraint solver and is not available."""
import argparse
import json
import sys
import time

SIMULATIONS = ["cube_incline", "ball_collision", "single_sphere", "compare_builtin", "multi_sphere"]


def run_simulation(sim_name, steps=None, envs=1, dtype="fp64", substeps=1, device=None, log_path=None):
    if sim_name not in SIMULATIONS:
        print(f"Unknown simulation name: '{sim_name}'")
        print("Available simulations:")
        for sim in SIMULATIONS:
            print(f"  {sim}")
        sys.exit(1)
    if sim_name == "compare_builtin":
        print("compare_builtin runs MuJoCo's own soft-contact solver, which is outside the accelerated path")
        sys.exit(1)
    import numpy as np
    import torch

    if not torch.cuda.is_available():
        print("No CUDA device: the stepping path runs as sm_100a kernels only (there is no CPU fallback)")
        sys.exit(1)
    from .simulation import ball_collision, cube_incline, multi_sphere_bounce, single_sphere_bounce
    tdtype = {"fp64": torch.float64, "fp32": torch.float32}[dtype]
    t0 = time.time()
    if sim_name == "single_sphere":
        model, data, logger = single_sphere_bounce.run_headless(steps or 2000, envs, device, tdtype, log=log_path is not None,
                                                                substeps_per_launch=substeps)
    elif sim_name == "cube_incline":
        model, data, logger = cube_incline.run_headless(steps or 240, envs, device, tdtype, log=log_path is not None,
                                                        substeps_per_launch=substeps)
    elif sim_name == "ball_collision":
        model, data, logger = ball_collision.run_headless(steps or 500, envs, device, tdtype, substeps)
    else:
        model, data, logger = multi_sphere_bounce.run_headless(steps or 300, envs, device, tdtype, substeps)
    torch.cuda.synchronize()
    if log_path is not None:
        if not hasattr(logger, "save_npz"):
            print(f"--log is available for single_sphere and cube_incline (the single-body steppers), not {sim_name}")
            sys.exit(1)
        logger.save_npz(log_path)
    contacts, impulses = data.counters()
    out = {"sim": sim_name, "envs": envs, "dtype": dtype, "wall_s": round(time.time() - t0, 4),
           "qpos_env0": np.asarray(data.qpos).reshape(envs, -1)[0].tolist(),
           "contacts": int(contacts.sum()), "impulses": int(impulses.sum())}
    print(json.dumps(out))
    return out


def main(argv=None):
    parser = argparse.ArgumentParser(description="Headless rigid-body simulation runner (B200)")
    parser.add_argument("--sim", type=str, required=True, help="one of: " + ", ".join(SIMULATIONS))
    parser.add_argument("--headless", action="store_true", help="accepted for clarity; this runner is always headless")
    parser.add_argument("--steps", type=int, default=None)
    parser.add_argument("--envs", type=int, default=1)
    parser.add_argument("--dtype", choices=["fp64", "fp32"], default="fp64")
    parser.add_argument("--substeps-per-launch", type=int, default=1)
    parser.add_argument("--seed", type=int, default=20261018)
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--log", type=str, default=None, metavar="PATH.npz",
                        help="save times [n] and positions [n, sampled envs, 3] of every step (recorded on the device, also "
                             "inside fused launches) -- the data behind the reference's height-vs-time plots")
    args = parser.parse_args(argv)
    run_simulation(args.sim, args.steps, args.envs, args.dtype, args.substeps_per_launch, log_path=args.log)


if __name__ == "__main__":
    main()
