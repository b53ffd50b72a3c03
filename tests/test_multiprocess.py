"""CPU, world_size 2 over gloo: the host-side multi-GPU logic (sharding of games and replay-buffer positions, the
hand-over of the communicator's unique id, counter reduction, max-over-ranks timing).  No GPU and no compute calls; the
weight broadcast and the set swap themselves run in the library (tests/test_gpu_weights.py, test_gpu_multirank.py)."""
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
        lo, hi = tzd.shard_range(1_000_001, rank, world)
        # the NCCL unique id of the library's communicator travels from rank 0 like this (128 bytes)
        uid = bytes(range(128)) if rank == 0 else None
        got = tzd.exchange_bytes(uid, 128, src=0)
        totals = tzd.sum_counters([10.0 * (rank + 1), float(n)])
        slowest = tzd.max_over_ranks([5.0 + rank])
        out[rank] = (base, n, got == bytes(range(128)), totals, slowest, lo, hi)
    finally:
        dist.destroy_process_group()


def test_two_rank_host_logic():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out[0][:2] == (0, 64) and out[1][:2] == (64, 64)  # contiguous, disjoint game ranges
    assert (out[0][5], out[0][6], out[1][5], out[1][6]) == (0, 500_001, 500_001, 1_000_001)
    for r in range(world):
        assert out[r][2], "the unique id differs from rank 0's after the exchange"
        assert out[r][3] == [30.0, 128.0]
        assert out[r][4] == [6.0]


def test_shard_range_covers_everything_once():
    for total, world in ((1_000_000, 8), (10, 3), (7, 8), (0, 2)):
        spans = [tzd.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_flops_per_position_matches_baseline_md():
    assert weights.flops_per_position(6) == 2 * 703_300_680
    assert weights.flops_per_position(4) == 2 * 305_205_280
    assert weights.flops_per_position(5) == 2 * 598_764_850
