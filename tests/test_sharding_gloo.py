"""World-size-2 gloo test of the view sharding + throughput reduction used by bench.py's N>1 path (CPU only)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from damvsnet_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, num_views, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.shard_views(num_views, rank, world)
    local = [(i, torch.tensor([float(i) * 2.0])) for i in mine]
    total, ms = sharding.reduce_throughput(len(mine), 10.0 * (rank + 1))
    gathered = sharding.gather_results(local, num_views)
    ok = total == num_views and abs(ms - 10.0 * world) < 1e-9 and all(
        g is not None and g.item() == 2.0 * i for i, g in enumerate(gathered))
    ret[rank] = (ok, mine)
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_partition_is_exact_cover():
    for n in (0, 1, 7, 49):
        for w in (1, 2, 4, 8):
            parts = [sharding.shard_views(n, r, w) for r in range(w)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(n))
            assert [len(p) for p in parts] == sharding.shard_sizes(n, w)
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_two_rank_gloo_reduction_and_gather():
    world, num_views = 2, 7
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, num_views, ret), nprocs=world, join=True)
        assert ret[0][0] and ret[1][0]
        assert ret[0][1] == [0, 2, 4, 6] and ret[1][1] == [1, 3, 5]


def test_single_process_passthrough():
    assert sharding.reduce_throughput(5, 12.5) == (5, 12.5)
