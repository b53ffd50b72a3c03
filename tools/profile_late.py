"""Like profile_cmd.py but from late-game positions (random playouts of `plies` plies first): the tree kernels' cost
grows with the stacks.  python tools/profile_late.py [plies] [locksteps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from takzero_b200 import capi, network, weights  # noqa: E402

plies = int(sys.argv[1]) if len(sys.argv) > 1 else 90
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
m = capi.BatchedMCTS(6, 4, 8192)
network.set_weights(m, weights.random_init(6, seed=123))
m.set_agent(capi.AGENT_NETWORK)
m.new_openings(seed=1000)
m.random_steps(plies, seed=5)
live = (m.result(m.positions()) == 0)
m.new_openings(seed=7, mask=(~live).astype(np.uint8))  # finished playouts start over
betas = np.zeros(8192, dtype=np.float32)
for _ in range(steps):
    m.simulate(betas)
m.sync()
print("status", m.status(), "launches", m.launch_count())
m.close()
