"""BASELINE.json configs[4] on one GPU: `reanalyze` fresh-target search over a synthetic replay buffer.

The buffer holds the positions of random-playout games (uniform random legal moves from `new_opening`, every ply
kept, like `Replay::states`, target.rs:205-212); every batch takes 8192 distinct positions as FRESH roots
(`*node = Node::default()`, reanalyze/src/main.rs:159-165), searches them with Gumbel sequential halving (k = 16,
256 simulations, beta = 0) and reads back the reanalyze targets (value rule of main.rs:184-195, improved policy with
most_visited_count(), UBE target).  Host buffers every batch: positions H2D, moves / root table / targets D2H.
Not the bench line (that is self-play, bench.py); a recorded secondary measurement.
  python tools/reanalyze_bench.py [positions] [batches] [games]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from takzero_b200 import capi, network, weights  # noqa: E402

want_positions = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
batches = int(sys.argv[2]) if len(sys.argv) > 2 else 4
G = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
n, hk, k, budget = 6, 4, 16, 256
m = capi.BatchedMCTS(n, hk, G)
network.set_weights(m, weights.random_init(n, seed=123))
m.set_agent(capi.AGENT_NETWORK)

# ---- synthetic replay buffer: all plies of random playouts (seeds 2000 + i) ------------------------------
t0 = time.perf_counter()
m.new_openings(seed=2000)
chunks, have, it = [], 0, 0
while have < want_positions:
    pos = m.positions()
    live = m.result(pos) == 0
    chunks.append(pos[live])
    have += int(live.sum())
    m.random_steps(1, seed=2000 + it)
    # finished games start over from a new opening
    done = (m.result(m.positions()) != 0).astype(np.uint8)
    if done.any():
        m.new_openings(seed=3000 + it, mask=done)
    it += 1
buffer = np.concatenate(chunks)[:want_positions]
print(f"buffer: {len(buffer):,} positions from {it} plies of {G} random playouts in {time.perf_counter() - t0:.1f} s "
      f"(mean ply {buffer['ply'].mean():.1f})", flush=True)

rng = np.random.default_rng(1)
betas = np.zeros(G, dtype=np.float32)


def batch(i):
    idx = rng.choice(len(buffer), size=G, replace=False)
    m.set_positions(buffer[idx])  # fresh roots
    selected = m.gumbel_sequential_halving(betas, k, budget, None, seed=100 + i)
    roots, ch = m.root_stats(), m.root_children()
    pol, ube, cnt = m.targets(-1.0, 0.25)
    # value = root evaluation if solved, else -evaluation of the selected child (main.rs:184-195)
    sel = (ch["moves"] == selected[:, None]).argmax(axis=1)
    child_bits = ch["eval_bits"][np.arange(G), sel]
    solved = roots["eval_tag"] != capi.E_VALUE
    return int(solved.sum()), float(pol[0, 0]) + float(ube[0]) + float(child_bits[0] & 1)


batch(-1)  # warm-up
m.sync()
c0 = m.counters()
t0 = time.perf_counter()
solved = 0
for i in range(batches):
    solved += batch(i)[0]
m.sync()
dt = time.perf_counter() - t0
c1 = m.counters()
assert m.status() == 0
sims = c1.simulations - c0.simulations
print(f"reanalyze: {batches} batches of {G} fresh roots, k={k}, {budget} sims: {sims / dt:,.0f} simulations/s, "
      f"{batches * G / dt:,.0f} targets/s, {dt / batches * 1000:.0f} ms per batch (wall clock incl. host buffers); "
      f"{solved} solved roots, known leaves {100.0 * (c1.known - c0.known) / sims:.1f} %")
m.close()
