"""GPU: the BASELINE.json configurations at FULL size with the device network, checked through
size-independent properties (the oracle cannot run these sizes in seconds):
  * the sequential-halving visit schedule (batched.rs:246-339; golden multiset as in runs/*.txt),
  * root visit_count = sum of child visits + 1 (batched.rs:374-380), improved policy sums to 1,
  * selected move = the survivor among the root children, replays replay to the current position,
  * no device-side invariant violation (tz_status), sharded == unsharded on a sub-batch."""
import numpy as np
import pytest

from takzero_b200 import capi, network, weights

pytestmark = pytest.mark.gpu


def schedule(k, budget):
    steps = k.bit_length() - 1
    per_step = budget // steps
    visits = np.zeros(k, dtype=np.int64)
    remaining = k
    for _ in range(steps):
        visits[:remaining] += per_step // remaining
        remaining //= 2
    return np.sort(visits)[::-1]


def run_config(n, hk, G, k, budget, moves, blocks=None):
    m = capi.BatchedMCTS(n, hk, G)
    network.set_weights(m, weights.random_init(n, seed=123, blocks=blocks))
    m.set_agent(capi.AGENT_NETWORK)
    m.new_openings(seed=1000)
    want = schedule(k, budget)
    for mv in range(moves):
        fresh = mv == 0
        if not fresh:
            m.reset_roots()  # reanalyze-style fresh roots so that the schedule is exact on every move
        selected = m.gumbel_sequential_halving(None, k, budget, None, seed=7)
        assert m.status() == 0
        tbl, st = m.root_children(), m.root_stats()
        vis = tbl["visits"].astype(np.int64)
        nchild = tbl["n"]
        assert (nchild >= k).all(), "6x6/4x4 openings have far more than k legal moves"
        assert np.array_equal(vis.sum(axis=1) + 1, st["visit_count"].astype(np.int64))
        top = -np.sort(-vis, axis=1)[:, :k]
        unsolved = st["eval_tag"] == capi.E_VALUE
        assert np.array_equal(top[unsolved], np.broadcast_to(want, top[unsolved].shape)), "halving schedule"
        assert (vis.sum(axis=1)[unsolved] == budget).all()
        # the selected move is the most visited child (the survivor got the last step's visits)
        rows = np.arange(G)
        best = vis.argmax(axis=1)
        assert (vis[rows, best] == want[0]).all()
        sel_idx = (tbl["moves"] == selected[:, None]).argmax(axis=1)
        assert (vis[rows, sel_idx] == want[0])[unsolved].all()
        pol, ube, cnt = m.targets(float(budget // (k.bit_length() - 1) // k * (k - 1)), 0.25)
        assert np.allclose(pol.sum(axis=1), 1.0, atol=1e-4) and (ube >= 0).all() and (ube <= 4.0).all()
        m.step(selected)
        term = m.restart_terminal_envs(seed=9)
        assert m.status() == 0
    c = m.counters()
    assert c.simulations == G * moves * (1 + budget)
    assert c.evaluations + c.known == c.simulations
    pos = m.positions()
    assert (pos["ply"] >= 2).all()
    m.close()
    return c


def test_config_4x4_1024_games_128_sims():
    """BASELINE.json configs[1]: 4x4, 1024 concurrent games, 128 sims/move, full 16-block net."""
    run_config(4, 4, 1024, 16, 128, 6)


def test_config_6x6_8192_games_256_sims():
    """BASELINE.json configs[2]: 6x6 half komi, 8192 concurrent games, 256 sims/move, full-size net."""
    c = run_config(6, 4, 8192, 16, 256, 1)
    assert c.evaluations > 0.9 * c.simulations


def test_config_5x5_20_blocks():
    run_config(5, 4, 512, 16, 128, 2)


@pytest.mark.parametrize("n,hk,G", [(3, 0, 1), (3, 0, 3), (4, 4, 5)])
def test_tiny_batches_play_whole_games_with_the_device_network(n, hk, G):
    """A handful of games played to the end (and restarted) with the device network: leaf batches of 0..G positions,
    including lock-steps where every leaf is already known and the fused network launch has nothing to do."""
    m = capi.BatchedMCTS(n, hk, G, arena_slots=1 << 14)
    network.set_weights(m, weights.random_init(n, seed=5, blocks=1))
    m.set_agent(capi.AGENT_NETWORK)
    m.new_openings(seed=3)
    finished = 0
    for mv in range(80):
        selected = m.gumbel_sequential_halving(None, 4, 16, None, seed=mv)
        assert m.status() == 0
        m.step(selected)
        finished += int((m.restart_terminal_envs(seed=100 + mv) != 0).sum())
    c = m.counters()
    assert c.simulations == G * 80 * 17 and c.evaluations + c.known == c.simulations
    assert finished >= 1 and c.known > 0
    assert m.status() == 0
    m.close()
