"""Long self-play soak on one GPU: python tools/soak.py [n] [games] [moves] [k] [budget]
A weight generation (the reference's Net::load before every move) precedes every move, alternating between two models.
Reports throughput, known-leaf fraction and the arena high-water mark every 10 moves."""
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
models = [weights.random_init(n, seed=123), weights.random_init(n, seed=124)]
network.broadcast_weights(m, models[0])
m.set_agent(capi.AGENT_NETWORK)
m.new_openings(seed=1000)
if len(sys.argv) > 6:  # start from positions this many random plies into the games (late-game behaviour without the wait)
    m.random_steps(int(sys.argv[6]), seed=5)
steps = k.bit_length() - 1
p = capi.SelfplayParams(k, budget, 0.0, 10, 32, 0.5, float(budget // steps // k * (k - 1)), 0.25, 7)
print(f"n={n} games={G} arena_slots={m.arena_slots} k={k} budget={budget}", flush=True)
c_prev, t_prev = m.counters(), time.perf_counter()
cat = dict(zip(capi.PROFILE_CATEGORIES, range(8)))
for mv in range(1, moves + 1):
    network.broadcast_weights(m, models[mv & 1])
    sampled = mv % 10 == 0 or mv == moves
    if sampled:
        m.profile_begin(8)
    m.selfplay_move(p)
    if sampled:
        prof = m.profile_end()
        per = {name: 1000.0 * prof.ms[i] / max(1, prof.locksteps) for name, i in cat.items() if prof.launches[i]}
        print("           per lock-step (us): " + ", ".join(f"{k} {v:.0f}" for k, v in per.items()), flush=True)
    if mv % 10 == 0 or mv == moves:
        m.sync()
        st = m.status()
        c, t = m.counters(), time.perf_counter()
        roots = m.root_stats()
        plies = m.positions()["ply"]
        sims = c.simulations - c_prev.simulations
        print(f"move {mv:4d}: {sims / (t - t_prev):12,.0f} sims/s  known {100.0 * (c.known - c_prev.known) / sims:5.1f}%  "
              f"arena max {int(roots['arena_used'].max()):7d} / {m.arena_slots}  ply mean {plies.mean():5.1f} max {plies.max()}  "
              f"generation {network.weight_generation(m)[0]}  status {st}", flush=True)
        if st:
            break
        c_prev, t_prev = c, t
m.close()
