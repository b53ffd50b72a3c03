"""Long self-play soak on one GPU: python tools/soak.py [n] [games] [moves] [k] [budget]
Reports throughput, finished games, known-leaf fraction and the arena high-water mark every 10 moves."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from takzero_b200 import capi, network, weights  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
G = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
moves = int(sys.argv[3]) if len(sys.argv) > 3 else 100
k = int(sys.argv[4]) if len(sys.argv) > 4 else 16
budget = int(sys.argv[5]) if len(sys.argv) > 5 else 256
m = capi.BatchedMCTS(n, 4, G)
network.set_weights(m, weights.random_init(n, seed=123))
m.set_agent(capi.AGENT_NETWORK)
m.new_openings(seed=1000)
steps = k.bit_length() - 1
p = capi.SelfplayParams(k, budget, 0.0, 10, 32, 0.5, float(budget // steps // k * (k - 1)), 0.25, 7)
print(f"n={n} games={G} arena_slots={m.arena_slots} k={k} budget={budget}", flush=True)
c_prev, t_prev = m.counters(), time.perf_counter()
for mv in range(1, moves + 1):
    m.selfplay_move(p)
    if mv % 10 == 0 or mv == moves:
        m.sync()
        st = m.status()
        c, t = m.counters(), time.perf_counter()
        roots = m.root_stats()
        plies = m.positions()["ply"]
        sims = c.simulations - c_prev.simulations
        print(f"move {mv:4d}: {sims / (t - t_prev):12,.0f} sims/s  known {100.0 * (c.known - c_prev.known) / sims:5.1f}%  "
              f"arena max {int(roots['arena_used'].max()):7d} / {m.arena_slots}  ply mean {plies.mean():5.1f} max {plies.max()}  "
              f"status {st}", flush=True)
        if st:
            break
        c_prev, t_prev = c, t
m.close()
