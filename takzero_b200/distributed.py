"""Multi-GPU plumbing: one process per GPU, `torch.distributed` for the only exchanges the path has.

The reference scales by running independent `selfplay` processes that share files
(README.md:128-130; `model_latest.ot` read at selfplay/src/main.rs:107, `buffer_lengths.txt` counters at
learn/src/main.rs:195-209).  Here games shard by contiguous global id and the two exchanges become
collectives: a broadcast of the weight blob per generation and a sum of the counters.  There is no
collective inside a simulation."""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import numpy as np


def shard(rank: int, world: int, games_per_rank: int) -> Tuple[int, int]:
    """(game_base, n_games) of this rank: contiguous global game ids, so RNG streams keyed by the global
    id make a sharded run equal to an unsharded one."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    return rank * games_per_rank, games_per_rank


def pack(tensors: Dict[str, np.ndarray]) -> Tuple[list, np.ndarray]:
    names = list(tensors)
    flat = np.concatenate([np.asarray(tensors[n], dtype=np.float32).ravel() for n in names])
    return names, flat


def unpack(names: Sequence[str], shapes: Dict[str, tuple], flat: np.ndarray) -> Dict[str, np.ndarray]:
    out, off = {}, 0
    for n in names:
        size = int(np.prod(shapes[n]))
        out[n] = flat[off:off + size].reshape(shapes[n]).copy()
        off += size
    if off != flat.size:
        raise ValueError("weight blob size mismatch")
    return out


def broadcast_weights(tensors: Dict[str, np.ndarray], src: int = 0, device=None) -> Dict[str, np.ndarray]:
    """`Net::load(model_latest.ot)` on every rank, as one broadcast of the flattened f32 blob from `src`
    (NCCL over NVLink when `device` is a CUDA device, gloo on CPU).  Every rank passes tensors of the same
    names / shapes (its own initialisation); the result holds `src`'s values."""
    import torch
    import torch.distributed as dist

    names, flat = pack(tensors)
    shapes = {n: tuple(np.asarray(tensors[n]).shape) for n in names}
    t = torch.from_numpy(flat)
    if device is not None:
        t = t.to(device)
    if dist.get_rank() != src:
        t.zero_()
    dist.broadcast(t, src=src)
    return unpack(names, shapes, t.cpu().numpy())


def sum_counters(values: Sequence[float], device=None) -> list:
    """Whole-job totals of per-rank counters (simulations, evaluations, positions, ...)."""
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
