"""CPU: the work-item schedule of the fused network launch (takzero_b200/csrc/conv_tcgen05.cuh `Schedule`, read
back through tz_debug_schedule -- the very code the kernel runs, compiled for the host).

The persistent kernel numbers its items chunk-major, then layer-major, and sizes everything from the device-side
position count; the host sizes the activation sets and the progress counters from upper bounds.  Checked here:
every (chunk, layer, pair tile) appears exactly once and in dependency order, chunks are balanced and never smaller
than the minimum (unless there is only one), and no count <= count_max exceeds the host's bounds."""
import ctypes as C

import numpy as np
import pytest

from takzero_b200 import capi


def schedule(count, count_max, n, min_tiles, layers, cap=0):
    lib = capi.lib()
    fn = lib.tz_debug_schedule
    fn.argtypes = [C.c_int] * 5 + [C.c_void_p, C.c_void_p, C.c_int]
    fn.restype = C.c_int
    out = np.zeros(8, dtype=np.int64)
    items = np.zeros((max(cap, 1), 3), dtype=np.int32)
    capi._check(fn(count, count_max, n, min_tiles, layers, capi._ptr(out), capi._ptr(items), cap))
    return out, items[:cap]


@pytest.mark.parametrize("n,count,min_tiles,layers", [(6, 8192, 150, 34), (6, 8000, 150, 34), (4, 8192, 150, 34),
                                                      (5, 3000, 40, 42), (6, 333, 3, 8), (4, 701, 3, 8), (6, 5, 150, 34)])
def test_items_cover_every_tile_once_in_dependency_order(n, count, min_tiles, layers):
    out, _ = schedule(count, count, n, min_tiles, layers)
    n_items, chunks, chunk_tiles, chunk_rows = (int(x) for x in out[:4])
    _, items = schedule(count, count, n, min_tiles, layers, cap=n_items)
    rows = count * n * n
    assert chunks == -(-rows // chunk_rows)
    seen = set()
    last = (-1, -1, -1)
    for c, l, t in items.tolist():
        assert (c, l, t) > last  # chunk-major, then layer, then tile: dependencies always have a smaller index
        last = (c, l, t)
        seen.add((c, l, t))
    want = set()
    for c in range(chunks):
        r = min(chunk_rows, rows - c * chunk_rows)
        for l in range(layers):
            for t in range(-(-r // 256)):
                want.add((c, l, t))
    assert seen == want and len(items) == len(want)
    # balanced: every chunk but the last is full, the last one is not a sliver
    if chunks > 1:
        assert chunk_tiles >= min_tiles
        assert rows - (chunks - 1) * chunk_rows > chunk_rows - chunks * n * n


@pytest.mark.parametrize("n,count_max,min_tiles", [(6, 8192, 150), (4, 8192, 150), (6, 2400, 150), (5, 9000, 40),
                                                   (3, 5000, 7), (6, 700, 3)])
def test_no_count_exceeds_the_host_bounds(n, count_max, min_tiles):
    rng = np.random.default_rng(n * count_max)
    counts = set(rng.integers(1, count_max + 1, size=300).tolist()) | {1, 2, count_max - 1, count_max}
    for count in sorted(counts):
        out, _ = schedule(count, count_max, n, min_tiles, 34)
        _, chunks, chunk_tiles, chunk_rows, b_chunks, b_tiles, b_rows_set = (int(x) for x in out[:7])
        assert chunks <= b_chunks and chunk_tiles <= b_tiles, (count, out)
        assert chunks * chunk_tiles + chunks <= b_chunks * b_tiles + b_chunks
        # a chunk's halo tile reads up to HALO rows past its last tile
        assert 8 + chunk_tiles * 256 + 8 <= b_rows_set
