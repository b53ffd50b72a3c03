"""CPU, world_size 2 over gloo: the host-side multi-GPU logic (sharding, weight broadcast, counter
reduction, max-over-ranks timing).  No GPU and no compute calls."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from takzero_b200 import distributed as tzd
from takzero_b200 import weights


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        base, n = tzd.shard(rank, world, 64)
        # every rank initialises differently; after the broadcast all hold rank 0's weights
        mine = weights.random_init(4, seed=100 + rank, blocks=1)
        got = tzd.broadcast_weights(mine, src=0)
        want = weights.random_init(4, seed=100, blocks=1)
        same = all(np.array_equal(got[k], want[k]) for k in want) and list(got) == list(want)
        totals = tzd.sum_counters([10.0 * (rank + 1), float(n)])
        slowest = tzd.max_over_ranks([5.0 + rank])
        out[rank] = (base, n, same, totals, slowest)
    finally:
        dist.destroy_process_group()


def test_two_rank_host_logic():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out[0][:2] == (0, 64) and out[1][:2] == (64, 64)  # contiguous, disjoint game ranges
    for r in range(world):
        assert out[r][2], "weights differ from rank 0 after the broadcast"
        assert out[r][3] == [30.0, 128.0]
        assert out[r][4] == [6.0]


def test_pack_unpack_round_trip():
    t = weights.random_init(4, seed=1, blocks=1)
    names, flat = tzd.pack(t)
    back = tzd.unpack(names, {k: v.shape for k, v in t.items()}, flat)
    assert all(np.array_equal(back[k], t[k]) for k in t)
    assert flat.dtype == np.float32


def test_flops_per_position_matches_baseline_md():
    assert weights.flops_per_position(6) == 2 * 703_300_680
    assert weights.flops_per_position(4) == 2 * 305_205_280
    assert weights.flops_per_position(5) == 2 * 598_764_850
