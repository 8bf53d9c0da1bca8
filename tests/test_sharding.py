"""Multi-GPU host logic on CPU: shard arithmetic and the optional final gather over a 2-rank gloo group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fss_b200.sharding import gather_point_outputs, key_shard, leaf_shard


def test_key_shards_partition():
    for n in (0, 1, 7, 8, 1 << 22, (1 << 22) + 5):
        for world in (1, 2, 3, 4, 8):
            spans = [key_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        key_shard(10, 2, 2)


def test_leaf_shards_partition():
    for n, gran in ((28, 1 << 17), (20, 1 << 17), (12, 1 << 12), (24, 1 << 17)):
        for world in (1, 2, 4, 8):
            spans = [leaf_shard(n, gran, r, world) for r in range(world)]
            pos = 0
            for b, c in spans:
                assert b == pos or c == 0
                assert b % gran == 0 and c % gran == 0
                pos = b + c if c else pos
            assert pos == 1 << n
    with pytest.raises(ValueError):
        leaf_shard(10, 3, 0, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, nkeys):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = key_shard(nkeys, rank, world)
        ys_local = torch.arange(b * 4, e * 4, dtype=torch.int32).reshape(e - b, 4)  # stand-in for eval outputs
        full = gather_point_outputs(ys_local, nkeys)
        assert torch.equal(full, torch.arange(nkeys * 4, dtype=torch.int32).reshape(nkeys, 4))
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == float(world)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nkeys", [8, 9])
def test_gather_world2_gloo(nkeys):
    mp.spawn(_worker, args=(2, _free_port(), nkeys), nprocs=2, join=True)
