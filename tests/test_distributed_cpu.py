"""World-size-2 gloo test of the N>1 host logic: shards partition the env range, the index-keyed generator
gives each rank its block, and the end-of-run statistics gather sums / maxes over ranks."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_total, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from rigidbody_simulation_b200 import shard, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, count = shard.shard_range(n_total, rank, world)
    s = synth.sphere_incline(count, start=start)
    stats = torch.tensor([0.0, float(count), float(2 * count), float(s["qpos"][:, 2].sum()), float(s["qpos"][:, 2].max())],
                         dtype=torch.float64)
    red = shard.gather_stats(stats, env_substeps=count * 10)
    t = shard.max_over_ranks(1.0 + rank, "cpu")
    np.save(os.path.join(out_dir, f"qpos_{rank}.npy"), s["qpos"])
    if rank == 0:
        np.save(os.path.join(out_dir, "red.npy"), np.array([red["env_substeps"], red["contacts"], red["impulses"],
                                                            red["energy_sum"], red["max_height"], t]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather(tmp_path):
    import torch.multiprocessing as mp
    from rigidbody_simulation_b200 import synth
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    n_total = 1001
    mp.spawn(_worker, args=(2, port, n_total, str(tmp_path)), nprocs=2, join=True)
    whole = synth.sphere_incline(n_total)
    got = np.concatenate([np.load(tmp_path / "qpos_0.npy"), np.load(tmp_path / "qpos_1.npy")])
    assert (got == whole["qpos"]).all()
    red = np.load(tmp_path / "red.npy")
    assert red[0] == n_total * 10 and red[1] == n_total and red[2] == 2 * n_total
    assert red[3] == pytest.approx(whole["qpos"][:, 2].sum(), rel=1e-12)
    assert red[4] == whole["qpos"][:, 2].max() and red[5] == 2.0
