"""CPU, world_size=2 over gloo: the read-sharding rule partitions the batch, and the ordered merge reproduces input order."""
import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from minimap2_b200 import dist as mdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    qlens = rng.integers(1, 20000, 1000)
    cut = mdist.shard_bounds(qlens, world)
    mine = [(i, int(qlens[i]) * 3) for i in range(cut[rank], cut[rank + 1])]  # stand-in for per-read results
    merged = mdist.gather_in_order(mine, cut, rank, world, dist)
    if rank == 0:
        ok = merged == [(i, int(qlens[i]) * 3) for i in range(len(qlens))]
        loads = [int(qlens[cut[r]:cut[r + 1]].sum()) for r in range(world)]
        q.put((ok, cut, loads))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_ordered_merge_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    ok, cut, loads = q.get(timeout=120)
    for p in ps:
        p.join(timeout=60)
    assert ok
    assert cut[0] == 0 and cut[-1] == 1000 and cut[1] > 0
    assert abs(loads[0] - loads[1]) < 0.02 * sum(loads)


def test_shard_bounds_properties():
    sys.path.insert(0, ROOT)
    from minimap2_b200 import dist as mdist
    rng = np.random.default_rng(1)
    for world in (1, 2, 4, 8):
        for n in (0, 1, 7, 1000):
            qlens = rng.integers(0, 5000, n)
            cut = mdist.shard_bounds(qlens, world)
            assert cut[0] == 0 and cut[-1] == n and all(cut[i] <= cut[i + 1] for i in range(world))
