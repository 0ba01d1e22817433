"""CPU tier, world size 2 over gloo: env sharding covers every env exactly once with GPU-count-independent seeds,
and the episode-statistics reduction is sum/sum/sum/max across ranks."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from toybox_b200 import distributed as D


def test_shard_partitions_every_env_once():
    for total in (1, 7, 64, 65536, 1048577):
        for world in (1, 2, 3, 4, 8):
            covered = 0
            for r in range(world):
                env0, n = D.shard(total, r, world)
                assert env0 == covered and n >= total // world
                covered += n
            assert covered == total
    a = np.concatenate([D.global_seeds(1234, *D.shard(1000, r, 4)) for r in range(4)])
    b = D.global_seeds(1234, 0, 1000)
    assert np.array_equal(a, b)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local = [[3, 100, 4000, 70], [5, 50, 1000, 90]][rank]
    out = D.reduce_episode_stats(local)
    env0, n = D.shard(101, rank, world)
    t = torch.tensor([n], dtype=torch.int64)
    dist.all_reduce(t)
    q.put((rank, out, int(t[0])))
    dist.destroy_process_group()


def test_episode_stats_reduce_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out, total in res:
        assert out == [8, 150, 5000, 90]
        assert total == 101
