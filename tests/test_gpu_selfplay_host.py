"""GPU: the C++ `selfplay` host (host/selfplay.cpp over include/takzero_b200.hpp) writes exactly the
`targets-selfplay.txt` / `replays.txt` an independent Python replay of the same loop produces, formatted with
the ORACLE's TPS / move notation and numpy's shortest round-trip floats (= Rust's `{}` of an f32):
reference formats takzero/src/target.rs:56-73,215-232, loop selfplay/src/main.rs:138-153,238-329."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from takzero_b200 import build as tz_build
from takzero_b200 import capi

from helpers import state_to_game

pytestmark = pytest.mark.gpu

RESULT = {1: "R-0", 2: "0-R", 3: "F-0", 4: "0-F", 5: "1/2-1/2"}


def f32(x) -> str:
    return np.format_float_positional(np.float32(x), unique=True, trim="-")


def python_selfplay(n, hk, G, k, budget, moves, seed, plies=10, beta=0.25):
    m = capi.BatchedMCTS(n, hk, G, arena_slots=1 << 15)
    m.new_openings(seed=seed)
    steps = k.bit_length() - 1
    vis = float(budget // steps // k * ((1 << steps) - 1))
    betas = np.zeros(G, dtype=np.float32)
    pending = [[] for _ in range(G)]
    targets_txt, replays_txt = "", ""
    cur = m.positions()
    for _ in range(moves):
        selected = m.gumbel_sequential_halving(betas, k, budget, None, seed=seed)
        sampled = m.select_actions_in_selfplay(plies, 32, 0.5, None, seed=seed)
        selected = np.where(cur["ply"] < plies, sampled, selected).astype(np.uint16)
        pol, ube, cnt, mv = m.targets(vis, beta, with_moves=True)
        for g in range(G):
            pending[g].append((cur[g].copy(), [(int(mv[g, i]), pol[g, i]) for i in range(cnt[g])], ube[g]))
        m.step(selected)
        after = m.positions()
        term = m.restart_terminal_envs(seed=seed)
        fin = [g for g in range(G) if term[g]]
        results = m.game_result(after[fin]) if fin else []
        for j, g in enumerate(fin):
            start, actions = m.finished_replay(g)
            line = '[TPS "%s"]' % O.to_tps(state_to_game(start, n, hk))
            line += "".join(" " + O.move_str(int(a)) for a in actions)
            # the oracle agrees on how the game ended
            final = state_to_game(after[g], n, hk)
            assert O.terminal(final) == term[g]
            replays_txt += line + " " + RESULT[int(results[j])] + "\n"
            value = O.make_eval(int(term[g]), 0)
            for env, policy, u in reversed(pending[g]):
                value = O.lib().tk_eval_negate(value)
                v = O.lib().tk_eval_to_f32(value)
                targets_txt += "%s;%s;%s;%s\n" % (
                    O.to_tps(state_to_game(env, n, hk)), f32(v), f32(u),
                    ",".join("%s:%s" % (O.move_str(a), f32(p)) for a, p in policy))
            pending[g] = []
        cur = m.positions()
    m.close()
    return targets_txt, replays_txt


def test_cpp_selfplay_host_writes_reference_formats(tmp_path):
    tz_build.build()
    exe = os.path.join(os.path.dirname(capi.LIB_PATH), "bin", "selfplay")
    n, hk, G, k, budget, moves, seed = 4, 4, 32, 8, 48, 70, 5
    out = subprocess.run(
        [exe, "--directory", str(tmp_path), "--board", str(n), "--half-komi", str(hk), "--games", str(G),
         "--sampled-actions", str(k), "--budget", str(budget), "--moves", str(moves), "--seed", str(seed),
         "--arena-slots", str(1 << 15)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    want_targets, want_replays = python_selfplay(n, hk, G, k, budget, moves, seed)
    got_replays = open(tmp_path / "replays.txt").read()
    got_targets = open(tmp_path / "targets-selfplay.txt").read()
    assert got_replays.count("\n") >= 5, "the run must finish some games"
    assert got_replays == want_replays
    assert got_targets == want_targets
    # every line parses back: replays re-play to the recorded result through the oracle
    for line in got_replays.splitlines():
        tps_part, rest = line.split('"]')
        g = O.from_tps(n, hk, tps_part.split('"')[1])
        toks = rest.split()
        for mvs in toks[:-1]:
            O.play(g, O.parse_move(mvs))
        assert O.terminal(g) != O.T_NONE and toks[-1] in RESULT.values()
    first = got_targets.splitlines()[0].split(";")
    assert len(first) == 4 and abs(sum(float(x.split(":")[1]) for x in first[3].split(",")) - 1.0) < 1e-3


def _mix64(x):
    m = (1 << 64) - 1
    x &= m
    x ^= x >> 33
    x = (x * 0xFF51AFD7ED558CCD) & m
    x ^= x >> 33
    x = (x * 0xC4CEB9FE1A85EC53) & m
    x ^= x >> 33
    return x


def _sample_distinct(seed, b, count, n):
    """host/reanalyze.cpp `sample_distinct`: slot g takes the first value of its hash sequence not taken before."""
    out, taken = [], set()
    for g in range(count):
        attempt = 0
        while True:
            i = _mix64(seed * 0x9E3779B97F4A7C15 + b * 1000003 + g + attempt * 0x632BE59BD9B4E019) % n
            attempt += 1
            if i not in taken:
                taken.add(i)
                out.append(i)
                break
    return out


def test_cpp_reanalyze_host(tmp_path):
    """host/reanalyze.cpp (reanalyze/src/main.rs:147-235): fresh-root search over positions expanded from
    replays.txt; value / policy / ube rules checked against an independent Python + oracle replay."""
    from helpers import games_to_states

    tz_build.build()
    bindir = os.path.join(os.path.dirname(capi.LIB_PATH), "bin")
    n, hk, G, k, budget, seed = 4, 4, 32, 8, 48, 5
    common = ["--directory", str(tmp_path), "--board", str(n), "--half-komi", str(hk), "--games", str(G),
              "--sampled-actions", str(k), "--budget", str(budget), "--seed", str(seed), "--arena-slots", str(1 << 15)]
    out = subprocess.run([os.path.join(bindir, "selfplay"), *common, "--moves", "60"], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr
    out = subprocess.run([os.path.join(bindir, "reanalyze"), *common, "--batches", "2"], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    # positions = Replay::states of every replay line, in file order (oracle parser + oracle rules)
    positions = []
    for line in open(tmp_path / "replays.txt").read().splitlines():
        g = O.from_tps(n, hk, line.split('"')[1])
        for tok in line.split('"]')[1].split():
            try:
                mv = O.parse_move(tok)
            except ValueError:
                continue
            positions.append(g.copy())
            O.play(g, mv)
    assert len(positions) >= G
    m = capi.BatchedMCTS(n, hk, G, arena_slots=1 << 15)
    want = ""
    for b in range(2):
        idx = _sample_distinct(seed, b, G, len(positions))
        batch = [positions[i] for i in idx]
        m.set_positions(games_to_states(batch))
        selected = m.gumbel_sequential_halving(np.zeros(G, np.float32), k, budget, None, seed=seed + b)
        roots, ch = m.root_stats(), m.root_children()
        pol, ube, cnt, mv = m.targets(-1.0, 0.25, with_moves=True)
        for g in range(G):
            if roots["eval_tag"][g] != 0:
                value = O.make_eval(int(roots["eval_tag"][g]), int(roots["eval_bits"][g]))
            else:
                i = list(ch["moves"][g, : ch["n"][g]]).index(selected[g])
                tag, bits = int(ch["eval_tag"][g, i]), int(ch["eval_bits"][g, i])
                child = O.make_eval(tag, float(np.uint32(bits).view(np.float32)) if tag == 0 else bits)
                value = O.lib().tk_eval_negate(child)
            want += "%s;%s;%s;%s\n" % (
                O.to_tps(batch[g]), f32(O.lib().tk_eval_to_f32(value)), f32(ube[g]),
                ",".join("%s:%s" % (O.move_str(int(mv[g, j])), f32(pol[g, j])) for j in range(cnt[g])))
    m.close()
    got = open(tmp_path / "targets-reanalyze.txt").read()
    assert got.count("\n") == 2 * G
    assert got == want


@pytest.mark.parametrize("fmt", ["ot", "tzw"])
def test_cpp_selfplay_host_with_network_weights_and_buffer_file(tmp_path, fmt):
    """The C++ host finds the model in --directory -- `model_latest.ot`, the tch archive `learn` writes
    (selfplay/src/main.rs:107), or a TZW1 file -- honours a valid buffer_lengths.txt (main.rs:93-104,371-387)
    and searches with the device network."""
    from takzero_b200 import weights

    tz_build.build()
    exe = os.path.join(os.path.dirname(capi.LIB_PATH), "bin", "selfplay")
    w = weights.random_init(4, seed=1, blocks=2)
    if fmt == "ot":
        weights.save_ot(str(tmp_path / "model_latest.ot"), w)
    else:
        weights.save_tzw(str(tmp_path / "model_latest.tzw"), w)
    (tmp_path / "buffer_lengths.txt").write_text("100,50,150")  # selfplay, reanalyze, checksum: below the limit
    out = subprocess.run([exe, "--directory", str(tmp_path), "--board", "4", "--half-komi", "4", "--games", "64",
                          "--sampled-actions", "8", "--budget", "48", "--moves", "30", "--seed", "3",
                          "--arena-slots", str(1 << 15)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    assert "30 moves of 64 games" in out.stdout and f"{64 * 30 * 49} simulations" in out.stdout
    lines = open(tmp_path / "targets-selfplay.txt").read().splitlines()
    assert lines and all(len(l.split(";")) == 4 for l in lines)
