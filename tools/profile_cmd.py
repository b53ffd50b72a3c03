"""The short command the ncu captures under profiles/ come from: the headline workload (8192 games of 6x6, full
network) advanced by a few lock-step simulations.  python tools/profile_cmd.py [locksteps] [games]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from takzero_b200 import capi, network, weights  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
games = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
m = capi.BatchedMCTS(6, 4, games)
network.set_weights(m, weights.random_init(6, seed=123))
m.set_agent(capi.AGENT_NETWORK)
m.new_openings(seed=1000)
betas = np.zeros(games, dtype=np.float32)
for _ in range(steps):
    m.simulate(betas)
m.sync()
assert m.status() == 0
c = m.counters()
print(f"{c.simulations} simulations, {c.evaluations} evaluations, {m.launch_count()} launches")
m.close()
