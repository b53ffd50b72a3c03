"""Times the tower convolution kernel in isolation (tuning aid, not a bench):
python tools/conv_tune.py [n] [positions]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from takzero_b200 import capi, network, weights  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
count = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
m = capi.BatchedMCTS(n, 4, count, arena_slots=4096)
network.set_weights(m, weights.random_init(n, blocks=1))
# fill the activation buffers with real data: evaluate positions a few random plies into real games
import numpy as np  # noqa: E402
m.new_openings(seed=1)
rng = np.random.default_rng(0)
for _ in range(10):
    st = m.positions()
    mv, cnt = m.legal_moves(st)
    pick = mv[np.arange(count), rng.integers(0, 1 << 30, size=count) % np.maximum(cnt, 1)]
    m.step(pick.astype(np.uint16))
    m.restart_terminal_envs(seed=2)
st = m.positions()
mv, cnt = m.legal_moves(st)
network.evaluate(m, st, [mv[i, : cnt[i]] for i in range(count)])
for c in sorted({count, count // 2, count // 8}):
    ms = network.time_tower(m, c, reps)
    flops = 2.0 * c * n * n * 9 * 256 * 256
    print(f"n={n} positions={c}: {ms*1000:.1f} us/conv, {flops/ms/1e9:.1f} TFLOP/s (dense rows, no padding)")
m.close()
