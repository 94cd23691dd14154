"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path (contiguous env-id shards, the
optional episode-statistics reduction).  The step path itself has no collective."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dynenv_b200.sharding import all_reduce_episode_stats, shard_range


def test_shard_range_partitions():
    for n, w in ((65536, 8), (10, 3), (7, 8), (1048576, 4)):
        r = [shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = shard_range(101, rank, world)
    stats = torch.tensor([float(e - b), rank + 1.0, 0, 0, 0, 1.0], dtype=torch.float64)
    red = all_reduce_episode_stats(stats.clone())
    q.put((rank, b, e, red.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, b0, e0, s0), (r1, b1, e1, s1) = res
    assert (b0, e1) == (0, 101) and e0 == b1
    assert s0 == s1 == [101.0, 3.0, 0.0, 0.0, 0.0, 2.0]
