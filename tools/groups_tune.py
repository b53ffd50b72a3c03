"""Resident self-play throughput with the 8192 games split over K independent lock-step groups (handles /
streams) on one GPU: python tools/groups_tune.py [groups] [moves]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from takzero_b200 import capi, network, weights  # noqa: E402

groups = int(sys.argv[1]) if len(sys.argv) > 1 else 2
moves = int(sys.argv[2]) if len(sys.argv) > 2 else 3
total = 8192
G = total // groups
tensors = weights.random_init(6, seed=123)
hs = []
for i in range(groups):
    m = capi.BatchedMCTS(6, 4, G, game_base=i * G, arena_slots=160000)
    network.set_weights(m, tensors)
    m.set_agent(capi.AGENT_NETWORK)
    m.new_openings(seed=1000)
    hs.append(m)
p = capi.SelfplayParams(16, 256, 0.0, 10, 32, 0.5, 60.0, 0.25, 1)
for _ in range(2):
    for m in hs:
        m.selfplay_move(p)
for m in hs:
    m.sync()
c0 = sum(m.counters().simulations for m in hs)
t0 = time.perf_counter()
for _ in range(moves):
    for m in hs:
        m.selfplay_move(p)
for m in hs:
    m.sync()
dt = time.perf_counter() - t0
sims = sum(m.counters().simulations for m in hs) - c0
print(f"groups={groups} games/group={G}: {sims / dt:,.0f} sims/s ({dt / moves * 1000:.0f} ms per move of all {total} games)")
