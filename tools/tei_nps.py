"""Single-tree search speed (the `tei` / `analysis` path: Node::simulate_batch, mcts.rs:268-328, 128 leaves per network
batch like tei/src/main.rs): nodes per second on one tree with the device network.
  python tools/tei_nps.py [n] [batch] [batches]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from takzero_b200 import capi, network, weights  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
batches = int(sys.argv[3]) if len(sys.argv) > 3 else 200
m = capi.BatchedMCTS(n, 4, 1, tree_batch=batch, arena_slots=1 << 25)
network.set_weights(m, weights.random_init(n, seed=123))
m.set_agent(capi.AGENT_NETWORK)
m.new_openings(seed=1)
for _ in range(10):
    m.tree_simulate_batch(0.0, batch)
m.sync()
c0 = m.counters()
t0 = time.perf_counter()
for _ in range(batches):
    m.tree_simulate_batch(0.0, batch)
m.sync()
dt = time.perf_counter() - t0
c1 = m.counters()
assert m.status() == 0
sims = c1.simulations - c0.simulations
print(f"{n}x{n} single tree, {batch} leaves per batch: {sims / dt:,.0f} nodes/s ({dt / batches * 1e3:.2f} ms per batch, "
      f"{(c1.evaluations - c0.evaluations) / batches:.1f} network positions per batch, root visits {m.root_stats()['visit_count'][0]})")
m.close()
