"""Headless counterpart of the reference's CLI ``src/simulate.py``.

The reference maps ``--sim NAME`` to a script and runs it in a subprocess (:13-19, :37); it exits with status 1 on
an unknown name (:21-26).  Here the same names select the same scenarios, run in-process on the GPU with no
window.  ``compare_builtin`` exercises MuJoCo's own constraint solver and is not available.

New flags: ``--headless`` (implied), ``--steps``, ``--envs``, ``--dtype``,
``--substeps-per-launch``, ``--arith strict|fast``, ``--config shipped|random``, ``--seed``, ``--bodies``, ``--gpus``,
``--log``.

* ``--config shipped`` (default without ``--seed``): every environment starts from the script's own initial
  condition (e.g. single_sphere_bounce.py:40-41).
* ``--config random`` (default with ``--seed S``): the randomised BASELINE config of that scenario from
  ``rigidbody_simulation_b200.synth`` -- single_sphere -> sphere on a 0.7 rad incline with per-env restitution and
  friction (configs[1]); ball_collision -> perturbed two-ball ICs (configs[2]); cube_incline -> perturbed cube on the
  incline (configs[3]); multi_sphere -> ``--bodies`` (default 64) spheres on a jittered lattice (configs[4]).  Values
  are a pure function of (seed, GLOBAL environment index), so any sharding sees the same environments.
* ``--gpus N``: one process per GPU (re-launched under ``torch.distributed.run`` when started bare); rank r steps the
  contiguous shard ``shard_range(envs, r, N)``; no collective on the step path, one end-of-run all-reduce of the
  statistics (NCCL), rank 0 prints.
"""
import argparse
import json
import os
import sys
import time

SIMULATIONS = ["cube_incline", "ball_collision", "single_sphere", "compare_builtin", "multi_sphere"]   # src/simulate.py:13-19
NEW_SIMULATIONS = ["mixed_pile"]          # not in the reference: spheres and boxes in one scene
DEFAULT_STEPS = {"single_sphere": 2000, "cube_incline": 240, "ball_collision": 500, "multi_sphere": 300, "mixed_pile": 400}


def _rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def build_random(sim_name, count, start, seed, device, tdtype, bodies=64):
    """(model, data, advance(k, arith)) for the randomised BASELINE config of ``sim_name``: environments
    [start, start+count) of the global index space."""
    import rigidbody_simulation_b200 as rb
    from rigidbody_simulation_b200 import scenes, stepper, synth
    from .simulation import ball_collision, multi_sphere_bounce
    if sim_name == "single_sphere":
        s = synth.sphere_incline(count, start=start, seed=seed)
        model = scenes.sphere_on_incline(count, device=device, dtype=tdtype)
        model.set_per_env(restitution=s["restitution"], friction=s["friction"])
        data = rb.BatchedData(model)
        advance = lambda k, arith: stepper.step_body_plane(model, data, -1, s["dt"], None, None, 0.0, substeps=k, arith=arith)
    elif sim_name == "cube_incline":
        s = synth.cube(count, start=start, seed=seed, kind="incline")
        model = scenes.cube_on_plane(count, theta=s["theta"], device=device, dtype=tdtype)
        data = rb.BatchedData(model)
        advance = lambda k, arith: stepper.step_body_plane(model, data, -1, s["dt"], s["restitution"], s["friction"],
                                                           s["threshold"], substeps=k, arith=arith)
    elif sim_name == "ball_collision":
        s = synth.two_ball(count, start=start, seed=seed)
        model, data = ball_collision.build(count, device=device, dtype=tdtype)
        advance = lambda k, arith: stepper.step_two_ball(model, data, s["dt"], s["restitution"], s["friction"],
                                                         radius=s["radius"], substeps=k, arith=arith)
    else:
        s = synth.multi_sphere(count, n_body=bodies, start=start, seed=seed)
        model, data = multi_sphere_bounce.build(count, device=device, dtype=tdtype, n_body=bodies)
        advance = lambda k, arith: stepper.step_multi_sphere(model, data, s["dt"], s["restitution"], s["friction"],
                                                             substeps=k, arith=arith)
    data.set_state(s["qpos"], s["qvel"])
    return model, data, advance


def run_simulation(sim_name, steps=None, envs=1, dtype="fp64", substeps=1, device=None, log_path=None, arith="strict",
                   config=None, seed=None, bodies=64):
    if sim_name not in SIMULATIONS + NEW_SIMULATIONS:
        print(f"Unknown simulation name: '{sim_name}'")
        print("Available simulations:")
        for sim in SIMULATIONS + NEW_SIMULATIONS:
            print(f"  {sim}")
        sys.exit(1)
    if sim_name == "compare_builtin":
        print("compare_builtin runs MuJoCo's own soft-contact solver, which is outside the accelerated path")
        sys.exit(1)
    if arith not in ("strict", "fast"):
        print(f"Unknown arithmetic policy: '{arith}' (strict | fast)")
        sys.exit(1)
    config = config or ("random" if seed is not None else "shipped")
    if config not in ("shipped", "random"):
        print(f"Unknown config: '{config}' (shipped | random)")
        sys.exit(1)
    import numpy as np
    import torch

    if not torch.cuda.is_available():
        print("No CUDA device: the stepping path runs as sm_100a kernels only (there is no CPU fallback)")
        sys.exit(1)
    from rigidbody_simulation_b200 import shard, synth
    from .simulation import ball_collision, cube_incline, multi_sphere_bounce, single_sphere_bounce
    rank, world, local = _rank_world()
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        device = torch.device("cuda", local)
        if not dist.is_initialized():
            sys.stdout.flush()
            saved = os.dup(1)                     # NCCL's banner must not land on stdout (one JSON line)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=device)
            finally:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)
    start, count = shard.shard_range(envs, rank, world)
    tdtype = {"fp64": torch.float64, "fp32": torch.float32}[dtype]
    steps = steps or DEFAULT_STEPS[sim_name]
    K = max(1, int(substeps))
    t0 = time.time()
    logger = None
    if sim_name == "mixed_pile":
        if arith != "strict":
            print("mixed_pile runs the strict policy only (boxes of any shape need the literal world inertia)")
            sys.exit(1)
        config = "random"
    if count == 0:
        model = data = None
    elif sim_name == "mixed_pile":
        from .simulation import mixed_pile
        model, data, logger = mixed_pile.run_headless(steps, count, device, tdtype, K, n_body=min(bodies, 256), start=start,
                                                      seed=synth.SEED if seed is None else seed)
    elif config == "random":
        if log_path is not None:
            print("--log records the shipped scenarios (single_sphere, cube_incline); it is not available with --config random")
            sys.exit(1)
        model, data, advance = build_random(sim_name, count, start, synth.SEED if seed is None else seed, device, tdtype, bodies)
        done = 0
        while done < steps:
            k = min(K, steps - done)
            advance(k, arith)
            done += k
    elif sim_name == "single_sphere":
        model, data, logger = single_sphere_bounce.run_headless(steps, count, device, tdtype, log=log_path is not None,
                                                                substeps_per_launch=K, arith=arith)
    elif sim_name == "cube_incline":
        model, data, logger = cube_incline.run_headless(steps, count, device, tdtype, log=log_path is not None,
                                                        substeps_per_launch=K, arith=arith)
    elif sim_name == "ball_collision":
        model, data, logger = ball_collision.run_headless(steps, count, device, tdtype, K, arith=arith)
    else:
        model, data, logger = multi_sphere_bounce.run_headless(steps, count, device, tdtype, K, arith=arith)
    torch.cuda.synchronize()
    wall = time.time() - t0
    if log_path is not None and rank == 0:
        if not hasattr(logger, "save_npz"):
            print(f"--log is available for single_sphere and cube_incline (the single-body steppers), not {sim_name}")
            sys.exit(1)
        logger.save_npz(log_path)
    # end-of-run statistics: one pass of the statistics kernel per rank, then the only collective of the job
    if data is not None:
        local_stats = shard.local_stats(model, data)
    else:
        local_stats = torch.tensor([0.0, 0.0, 0.0, 0.0, -1.0e300], dtype=torch.float64, device=device or "cuda")
    stats = shard.gather_stats(local_stats, env_substeps=count * steps)
    out = {"sim": sim_name, "config": config, "envs": envs, "gpus": world, "steps": steps, "dtype": dtype, "arith": arith,
           "substeps_per_launch": K, "wall_s": round(wall, 4), "env_substeps_per_s_wall": envs * steps / max(wall, 1e-9),
           "stats": stats, "contacts": int(stats["contacts"]), "impulses": int(stats["impulses"])}
    if config == "random":
        out["seed"] = synth.SEED if seed is None else seed
    if rank == 0:
        if data is not None:
            out["qpos_env0"] = np.asarray(data.qpos).reshape(count, -1)[0].tolist()
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    return out


def main(argv=None):
    parser = argparse.ArgumentParser(description="Headless rigid-body simulation runner (B200)")
    parser.add_argument("--sim", type=str, required=True, help="one of: " + ", ".join(SIMULATIONS + NEW_SIMULATIONS))
    parser.add_argument("--headless", action="store_true", help="accepted for clarity; this runner is always headless")
    parser.add_argument("--steps", type=int, default=None)
    parser.add_argument("--envs", type=int, default=1, help="environments in total (sharded over --gpus)")
    parser.add_argument("--dtype", choices=["fp64", "fp32"], default="fp64")
    parser.add_argument("--substeps-per-launch", type=int, default=1)
    parser.add_argument("--arith", choices=["strict", "fast"], default="strict",
                        help="strict: the reference's rounding sequence (bit-faithful); fast: re-associated, <= 1e-12 per step")
    parser.add_argument("--config", choices=["shipped", "random"], default=None,
                        help="initial conditions: the script's own (default) or the randomised BASELINE config (default with --seed)")
    parser.add_argument("--seed", type=int, default=None, help="seed of the randomised initial states (implies --config random)")
    parser.add_argument("--bodies", type=int, default=None,
                        help="bodies per environment: spheres of the randomised multi_sphere config (default 64), spheres and boxes of mixed_pile (default 8)")
    parser.add_argument("--gpus", type=int, default=1, help="one process per GPU; environments are sharded, no collective on the step path")
    parser.add_argument("--log", type=str, default=None, metavar="PATH.npz",
                        help="save times [n] and positions [n, sampled envs, 3] of every step (recorded on the device, also "
                             "inside fused launches) -- the data behind the reference's height-vs-time plots")
    args = parser.parse_args(argv)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # started bare: re-launch under torchrun, one rank per GPU (rendezvous on 127.0.0.1)
        import socket
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        cli = list(sys.argv[1:] if argv is None else argv)
        os.execv(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                                  "--master-addr", "127.0.0.1", "--master-port", str(port), "-m",
                                  "rigidbody_simulation_b200.src.simulate"] + cli)
    if "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) != args.gpus:
        print(f"--gpus {args.gpus} does not match the launcher's WORLD_SIZE={os.environ['WORLD_SIZE']}")
        sys.exit(1)
    run_simulation(args.sim, args.steps, args.envs, args.dtype, args.substeps_per_launch, log_path=args.log, arith=args.arith,
                   config=args.config, seed=args.seed, bodies=args.bodies if args.bodies is not None else (8 if args.sim == "mixed_pile" else 64))


if __name__ == "__main__":
    main()
