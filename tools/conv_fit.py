"""Fits the tower convolution's launch time to  T = a + b * (tiles per CTA pair)  (tuning aid, not a bench):
positions are chosen so that the number of 256-row pair tiles is an exact multiple of the 74 CTA pairs, plus the
bench's own 8192 (15.57 tiles per pair -> 16 waves).  `a` is the per-launch fixed cost (launch gap, prologue,
pipeline fill, last epilogue); python tools/conv_fit.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from takzero_b200 import capi, network, weights  # noqa: E402

n, games = 6, 8448
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
m = capi.BatchedMCTS(n, 4, games, arena_slots=4096)
network.set_weights(m, weights.random_init(n, blocks=1))
m.new_openings(seed=1)
rng = np.random.default_rng(0)
for _ in range(10):
    st = m.positions()
    mv, cnt = m.legal_moves(st)
    pick = mv[np.arange(games), rng.integers(0, 1 << 30, size=games) % np.maximum(cnt, 1)]
    m.step(pick.astype(np.uint16))
    m.restart_terminal_envs(seed=2)
st = m.positions()
mv, cnt = m.legal_moves(st)
network.evaluate(m, st, [mv[i, : cnt[i]] for i in range(games)])
network.time_tower(m, 8192, 400)  # settle under the power cap
rows = []
for k in (1, 2, 4, 8, 12, 15, 16):
    c = 74 * k * 256 // 36
    tiles = -(-c * 36 // 256)
    ms = network.time_tower(m, c, reps)
    rows.append((tiles / 74.0, ms * 1000))
    print(f"positions={c} pair_tiles={tiles} ({tiles/74:.2f}/pair): {ms*1000:.1f} us/conv, "
          f"{2.0*c*36*9*256*256/ms/1e9:.0f} TFLOP/s")
ms = network.time_tower(m, 8192, reps)
print(f"positions=8192 pair_tiles=1152 (15.57/pair): {ms*1000:.1f} us/conv, {2.0*8192*36*9*256*256/ms/1e9:.0f} TFLOP/s")
x = np.array([r[0] for r in rows]); y = np.array([r[1] for r in rows])
b, a = np.polyfit(x, y, 1)
print(f"fit: T = {a:.1f} us + {b:.2f} us * tiles_per_pair")
m.close()
