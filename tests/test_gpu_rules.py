"""GPU parity: Tak rules through the C ABI vs the CPU oracle, bit-exact.

Reference boundary: fast-tak `Game::{possible_moves, play, result}` as called from
takzero/src/search/env.rs:39-59."""
import numpy as np
import pytest

from oracle import oracle as O
from takzero_b200 import capi

from helpers import games_to_states, random_playout_states, state_to_game, states_equal, states_equal_bulk

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handles():
    hs = {}
    for n, hk in ((3, 0), (4, 4), (5, 4), (6, 4)):
        hs[n] = capi.BatchedMCTS(n, hk, 4, arena_slots=4096)
    yield hs
    for h in hs.values():
        h.close()


@pytest.mark.parametrize("n,half_komi,games", [(3, 0, 60), (4, 4, 60), (5, 4, 40), (6, 4, 40)])
def test_rules_random_playouts(handles, n, half_komi, games):
    h = handles[n]
    positions = []
    for seed in range(games):
        positions += random_playout_states(n, half_komi, 10_000 * n + seed)
    states = games_to_states(positions)
    moves, counts = h.legal_moves(states)
    terms = h.result(states)
    n_terminal = 0
    chosen = np.zeros(len(positions), dtype=np.uint16)
    expect_next = []
    rng = np.random.default_rng(n)
    for i, g in enumerate(positions):
        t = O.terminal(g)
        assert terms[i] == t, f"terminal mismatch at {O.to_tps(g)}"
        want = O.possible_moves(g)
        assert counts[i] == len(want), f"move count mismatch at {O.to_tps(g)}"
        assert list(moves[i, : counts[i]]) == want, f"move list/order mismatch at {O.to_tps(g)}"
        n_terminal += t != O.T_NONE
        m = want[int(rng.integers(len(want)))]
        chosen[i] = m
        g2 = g.copy()
        O.play(g2, m)
        expect_next.append(g2)
    after, ok = h.apply(states, chosen)
    assert ok.all()
    expect = games_to_states(expect_next)
    for i in range(len(positions)):
        assert states_equal(after[i], expect[i]), f"apply mismatch at {O.to_tps(positions[i])} {O.move_str(int(chosen[i]))}"
    assert n_terminal >= games // 2  # playouts really reach finished games (roads / flat wins)


@pytest.mark.parametrize("n,half_komi", [(3, 0), (4, 4), (5, 4), (6, 4)])
def test_rules_bulk_playouts(n, half_komi):
    """>= 10^5 positions per board size (SURVEY 7's plan), generated and answered by the oracle's batched C entry
    (oracle/tak_batch.c) and compared in bulk: ordered legal-move lists, terminal codes, absolute results and the
    position after the played move, bit for bit.  The playout policies make every way a game can end show up in
    quantity: roads, flat wins on a full board, flat wins by empty reserves, the reversible-ply draw."""
    per_policy = 24_000
    ends = {"road": 0, "flat_full": 0, "flat_out": 0, "rev_draw": 0}
    total = 0
    for policy, rev_limit in ((0, 100), (1, 100), (2, 100), (3, 12), (4, 100)):
        h = capi.BatchedMCTS(n, half_komi, 4, arena_slots=4096, reversible_limit=rev_limit)
        d = O.playout_positions(n, half_komi, 77 * n + policy, policy, per_policy, reversible_limit=rev_limit)
        states = d["states"].view(capi.STATE_DTYPE).reshape(-1)
        stride = int(d["n_moves"].max())
        for lo in range(0, per_policy, 8192):
            sl = slice(lo, min(lo + 8192, per_policy))
            st = states[sl]
            moves, counts = h.legal_moves(st, stride=stride)
            assert np.array_equal(counts, d["n_moves"][sl]), f"N={n} policy {policy}: legal-move counts"
            cols = np.arange(stride)[None, :] < counts[:, None]
            assert np.array_equal(np.where(cols, moves, 0), np.where(cols, d["moves"][sl, :stride], 0)), \
                f"N={n} policy {policy}: legal-move lists / order"
            assert np.array_equal(h.result(st), d["terminal"][sl]), f"N={n} policy {policy}: terminal"
            assert np.array_equal(h.game_result(st), d["result"][sl]), f"N={n} policy {policy}: result"
            live = d["chosen"][sl] != 0xFFFF
            after, ok = h.apply(st[live], d["chosen"][sl][live])
            assert ok.all()
            want = d["next_states"][sl].view(capi.STATE_DTYPE).reshape(-1)[live]
            assert states_equal_bulk(after, want).all(), f"N={n} policy {policy}: play"
        term = d["terminal"] != 0
        full = (states["height"][:, : n * n] > 0).all(axis=1)
        out = ((states["stones"].astype(int) + states["caps"]) == 0).any(axis=1)
        res = d["result"]
        ends["road"] += int(((res == 1) | (res == 2)).sum())
        ends["flat_full"] += int((term & full & (res >= 3)).sum())
        ends["flat_out"] += int((term & out & ~full & (res >= 3)).sum())
        ends["rev_draw"] += int((term & ~full & ~out & (res == 5)).sum())
        total += per_policy
        h.close()
    print(f"N={n}: {total} positions, finished games by kind {ends}")
    assert total >= 100_000
    assert ends["road"] >= 100 and ends["flat_full"] >= 100 and ends["rev_draw"] >= 100
    assert ends["flat_out"] >= (5 if n == 3 else 50)


def test_rules_golden_position(handles):
    """repr.rs:411-499: the 18 legal moves of `2,1,x/1S,221,x/x,2S,2 1 6` (3x3)."""
    g = O.from_tps(3, 0, "2,1,x/1S,221,x/x,2S,2 1 6")
    moves, counts = handles[3].legal_moves(games_to_states([g]))
    assert counts[0] == 18
    assert list(moves[0, :18]) == O.possible_moves(g)


def test_rules_tinue_positions(handles):
    """mcts.rs:345-411 positions: road detection after the winning replies."""
    g = O.from_ptn_moves(3, 0, ["a3", "c1", "c2", "c3", "b3", "c3-", "b1"])
    states = games_to_states([g])
    assert handles[3].result(states)[0] == O.terminal(g)


def test_empty_and_invalid(handles):
    h = handles[4]
    moves, counts = h.legal_moves(np.zeros(0, dtype=capi.STATE_DTYPE))
    assert moves.shape[0] == 0 and counts.shape[0] == 0
    g = O.new_game(4, 4)
    st = games_to_states([g])
    # a spread on ply 0 and a placement on an occupied square are rejected like Game::play
    bad = np.array([O.parse_move("a1+")], dtype=np.uint16)
    _, ok = h.apply(st, bad)
    assert ok[0] == 0


def test_random_steps_are_legal_playouts(handles):
    """new_opening_with_random_steps (env.rs:81-96): every position reached is what the oracle reaches by
    replaying some legal move from the previous one, and plies advance by `steps` unless the game ended."""
    n, hk = 5, 4
    m = capi.BatchedMCTS(n, hk, 64, arena_slots=4096)
    m.new_openings(seed=3)
    prev = m.positions()
    for step in range(12):
        m.random_steps(1, seed=11)
        cur = m.positions()
        for g in range(64):
            before = state_to_game(prev[g], n, hk)
            if O.terminal(before) != O.T_NONE:
                assert states_equal(cur[g], prev[g])
                continue
            nexts = []
            for mv in O.possible_moves(before):
                g2 = before.copy()
                O.play(g2, mv)
                nexts.append(games_to_states([g2])[0])
            assert any(states_equal(cur[g], x) for x in nexts), f"step {step} game {g}"
        prev = cur
    assert len({bytes(p["height"]) for p in prev}) > 20  # the games diverged
    m.close()


def test_impossible_host_states_are_refused():
    """A `tz_state_t` can describe what the reference's `Game` cannot (a C struct has no invariants): stacks taller
    than the 64-bit colour mask, unknown piece types, pieces off the board.  Refused at the ABI with TZ_EINVAL."""
    n = 5
    m = capi.BatchedMCTS(n, 4, 4, arena_slots=4096)
    good = m.positions()
    assert m.legal_moves(good)[1].min() > 0
    for field, sq, value in (("height", 3, 65), ("top", 3, 7), ("height", 30, 1), ("to_move", None, 2), ("caps", 0, 9)):
        bad = good.copy()
        if field in ("height", "top"):
            bad["height"][1, sq] = max(1, int(bad["height"][1, sq]))
            bad[field][1, sq] = value
        elif field == "caps":
            bad["caps"][1, 0] = value
        else:
            bad[field][1] = value
        with pytest.raises(capi.TakzeroError, match="state 1 is not a possible position"):
            m.legal_moves(bad)
        with pytest.raises(capi.TakzeroError, match="state 1 is not a possible position"):
            m.set_positions(bad)
    m.set_positions(good)
    assert m.status() == 0
    m.close()


def _perft(m, states, depth, stride, chunk=8192):
    """Move sequences of length `depth` from `states`, depth-first over chunks (the frontier of the last level never
    exists as a whole): tz_result filters finished games, tz_legal_moves counts, tz_apply expands."""
    total = 0
    for lo in range(0, len(states), chunk):
        part = states[lo:lo + chunk]
        part = part[m.result(part) == 0]
        if len(part) == 0:
            continue
        moves, cnt = m.legal_moves(part, stride)
        assert (cnt >= 0).all() and (cnt <= stride).all()
        if depth == 1:
            total += int(cnt.sum())
            continue
        parent = np.repeat(np.arange(len(part)), cnt)
        played = moves[np.arange(stride)[None, :] < cnt[:, None]]  # row-major = parent order
        children, ok = m.apply(part[parent], played)
        assert ok.all()
        total += _perft(m, children, depth - 1, stride, chunk)
    return total


@pytest.mark.parametrize("n,depth", [(3, 7), (4, 5), (5, 5), (6, 4)])
def test_perft_through_the_c_abi(n, depth):
    """The published perft counts of Tak (tests/test_oracle_golden.py::PERFT, where their provenance is stated) from
    the CUDA rules alone: tz_result / tz_legal_moves / tz_apply, no oracle in the loop.  3x3 to depth 7 (52 M move
    sequences, most lines finished on the way and must not be continued), 5x5 to depth 5 (188 M, capstones flatten
    walls), 6x6 to depth 4 (13.6 M)."""
    from test_oracle_golden import PERFT

    m = capi.BatchedMCTS(n, 0, 4, arena_slots=4096)
    root = games_to_states([O.new_game(n, 0)])
    for d in range(1, depth + 1):
        if d < depth - 1 and d > 2:
            continue  # the deepest two levels say it all; the shallow ones are cheap
        assert _perft(m, root, d, stride=256) == PERFT[n][d - 1], f"{n}x{n} perft({d})"
    m.close()


def test_a_move_that_completes_both_roads_wins_for_the_mover():
    """The rule of tests/test_oracle_golden.py::test_a_move_that_completes_both_roads_wins_for_the_mover through
    tz_apply / tz_game_result / tz_result."""
    m = capi.BatchedMCTS(3, 0, 2, arena_slots=4096)
    games = [O.from_tps(3, 0, "x2,21/2,2,x/1,1,x 1 10"), O.from_tps(3, 0, "x2,12/1,1,x/2,2,x 2 10")]
    move = O.parse_move("2c3-11")
    states, ok = m.apply(games_to_states(games), [move, move])
    assert ok.all()
    assert list(m.game_result(states)) == [1, 2]  # R-0, 0-R: the mover's road counts
    assert list(m.result(states)) == [2, 2]  # a loss for the side to move, both times
    m.close()
