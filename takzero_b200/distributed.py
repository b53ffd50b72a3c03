"""Multi-GPU plumbing: one process per GPU.

The reference scales by running independent `selfplay` / `reanalyze` processes that share files
(README.md:128-130; `model_latest.ot` re-read before every move at selfplay/src/main.rs:107, `buffer_lengths.txt`
counters at learn/src/main.rs:195-209).  Here games (or replay-buffer positions) shard by contiguous global index and
the two exchanges are NCCL collectives INSIDE the library (csrc/comm.cu): `tz_broadcast_weights` per generation and
`tz_allreduce_sum` of the counters.  There is no collective inside a simulation.

`torch.distributed` is only the rendezvous here: it carries the 128-byte NCCL unique id from rank 0 to the others
(and the max-over-ranks of the timings, a float the library has no business with)."""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np


def shard(rank: int, world: int, games_per_rank: int) -> Tuple[int, int]:
    """(game_base, n_games) of this rank: contiguous global game ids, so RNG streams keyed by the global
    id make a sharded run equal to an unsharded one."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    return rank * games_per_rank, games_per_rank


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of `total` items (replay-buffer positions) owned by this rank: contiguous, disjoint, covering,
    sizes differing by at most one."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def exchange_bytes(payload: bytes | None, size: int, src: int = 0, device=None) -> bytes:
    """`size` bytes from rank `src` to every rank over the default process group (gloo on CPU, NCCL with `device`)."""
    import torch
    import torch.distributed as dist

    t = torch.zeros(size, dtype=torch.uint8)
    if dist.get_rank() == src:
        if payload is None or len(payload) != size:
            raise ValueError("the source rank passes exactly `size` bytes")
        t = torch.frombuffer(bytearray(payload), dtype=torch.uint8).clone()
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def init_comm(mcts, rank: int, world: int, device=None) -> None:
    """The library's own NCCL communicator for `mcts` (tz_comm_init): rank 0 makes the unique id, the process group
    hands it round."""
    from . import network

    if world == 1:
        network.comm_init(mcts, None, 1, 0)
        return
    uid = network.comm_unique_id() if rank == 0 else None
    uid = exchange_bytes(uid, 128, src=0, device=device)
    network.comm_init(mcts, uid, world, rank)


def sum_counters(values: Sequence[float], device=None) -> list:
    """Whole-job totals over the default process group (host-side fallback of tz_allreduce_sum)."""
    import torch
    import torch.distributed as dist

    t = torch.tensor(list(values), dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().tolist()


def max_over_ranks(values: Sequence[float], device=None) -> list:
    """Device times are reported as the max over ranks."""
    import torch
    import torch.distributed as dist

    t = torch.tensor(list(values), dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.cpu().tolist()
