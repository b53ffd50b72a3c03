"""Times one whole network evaluation through the C ABI (tz_evaluate host round trip excluded: uses the sampled
profile of a short search) -- tuning aid: python tools/conv_tower.py [n] [games] [sims]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from takzero_b200 import capi, network, weights  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
games = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
sims = int(sys.argv[3]) if len(sys.argv) > 3 else 64
m = capi.BatchedMCTS(n, 4, games)
network.set_weights(m, weights.random_init(n))
m.set_agent(capi.AGENT_NETWORK)
m.new_openings(seed=1)
m.gumbel_sequential_halving(None, 16, sims, None, seed=1)
m.reset_roots()
m.sync()
m.profile_begin(1)
m.timer_start()
m.gumbel_sequential_halving(None, 16, 2 * sims, None, seed=2)
ms = m.timer_stop()
prof = m.profile_end()
assert m.status() == 0
cat = dict(zip(capi.PROFILE_CATEGORIES, range(8)))
conv = prof.ms[cat["conv_input"]] + prof.ms[cat["conv_tower"]] + prof.ms[cat["conv_policy"]]
flops = weights.flops_per_position(n) * prof.positions
print(f"n={n} games={games}: {games * (2 * sims + 1) / ms * 1000:.0f} sims/s, conv {flops / conv / 1e9:.1f} TFLOP/s, "
      f"per lock-step: " + ", ".join(f"{k} {prof.ms[i] / prof.locksteps * 1000:.0f} us" for k, i in cat.items() if prof.launches[i]))
m.close()
