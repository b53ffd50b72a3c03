"""GPU: the C++ `evaluation` host (host/evaluation.cpp; reference evaluation/src/main.rs:139-318: two networks, each
searching its own trees, play a batch of games with colours swapped) prints exactly the result lines an independent
Python replay of the same loop produces through the ctypes binding."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from takzero_b200 import build as tz_build
from takzero_b200 import capi

from helpers import state_to_game

pytestmark = pytest.mark.gpu

M64 = (1 << 64) - 1


def _mix64(x):
    x &= M64
    x ^= x >> 33
    x = (x * 0xFF51AFD7ED558CCD) & M64
    x ^= x >> 33
    x = (x * 0xC4CEB9FE1A85EC53) & M64
    x ^= x >> 33
    return x


def compete(white, black, games, G, k, budget, max_moves, seed, rnd):
    wins = losses = draws = 0
    white.set_positions(games)
    black.set_positions(games)
    betas = np.zeros(G, dtype=np.float32)
    done = np.zeros(G, dtype=bool)
    ply = 0
    for _ in range(max_moves):
        for is_white in (True, False):
            if done.all():
                return wins, losses, draws
            current, other = (white, black) if is_white else (black, white)
            s = (seed + 1000003 * rnd + ply) & M64
            top = current.gumbel_sequential_halving(betas, k, budget, None, seed=s)
            current.step(top)
            other.step(top)
            term = current.restart_terminal_envs(seed=s)
            for g in range(G):
                if done[g] or term[g] == 0:
                    continue
                done[g] = True
                # evaluation/src/main.rs:300-306: the terminal belongs to the player to move after the move
                if term[g] == capi.T_DRAW:
                    draws += 1
                elif (term[g] == capi.T_LOSS and is_white) or (term[g] == capi.T_WIN and not is_white):
                    wins += 1
                else:
                    losses += 1
            if done.any():
                other.set_positions(current.positions(), mask=done.astype(np.uint8))
            ply += 1
    return wins, losses, draws


def line(a, b, w, l, d):
    total = w + l + d
    rate = "NaN" if total == 0 else "%.1f" % (100.0 * w / total)
    return f"{a} vs. {b}: Evaluation {{ wins: {w}, losses: {l}, draws: {d} }} {rate}%\n"


@pytest.mark.parametrize("book", [False, True])
def test_cpp_evaluation_host_matches_python_replay(tmp_path, book):
    tz_build.build()
    exe = os.path.join(os.path.dirname(capi.LIB_PATH), "bin", "evaluation")
    n, hk, G, k, budget, max_moves, seed, rounds = 4, 4, 16, 8, 24, 40, 11, 2
    args = [exe, "--board", str(n), "--half-komi", str(hk), "--games", str(G), "--sampled-actions", str(k),
            "--budget", str(budget), "--max-moves", str(max_moves), "--seed", str(seed), "--rounds", str(rounds),
            "--arena-slots", str(1 << 14)]
    first = capi.BatchedMCTS(n, hk, G, arena_slots=1 << 14)
    second = capi.BatchedMCTS(n, hk, G, arena_slots=1 << 14)
    book_states = None
    if book:
        # an opening book: TPS of positions a few random plies into 40 games, written with the oracle's notation
        src = capi.BatchedMCTS(n, hk, 40, arena_slots=4096)
        src.new_openings(seed=5)
        src.random_steps(3, seed=6)
        book_states = src.positions()
        src.close()
        with open(tmp_path / "book.txt", "w") as f:
            for st in book_states:
                f.write(O.to_tps(state_to_game(st, n, hk)) + "\n")
        args += ["--opening-book", str(tmp_path / "book.txt")]
    out = subprocess.run(args, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    want = ""
    for rnd in range(rounds):
        if book:
            taken, idx = set(), []
            for g in range(G):
                attempt = 0
                while True:
                    i = _mix64(seed * 0x9E3779B97F4A7C15 + rnd * 1000003 + g + attempt * 0x632BE59BD9B4E019) % len(book_states)
                    attempt += 1
                    if i not in taken:
                        taken.add(i)
                        idx.append(i)
                        break
            games = book_states[idx].copy()
            games["reversible_plies"] = 0  # TPS does not carry it
        else:
            s = seed + 7919 * rnd
            first.new_openings(seed=s)
            first.random_steps(2, seed=s)
            third = np.array([_mix64(s * 0x9E3779B97F4A7C15 + g) & 1 for g in range(G)], dtype=np.uint8)
            first.random_steps(1, seed=s + 1, mask=third)
            games = first.positions()
        w, l, d = compete(first, second, games, G, k, budget, max_moves, seed, 2 * rnd)
        want += line("synthetic-a", "synthetic-b", w, l, d)
        assert w + l + d > 0
        w, l, d = compete(second, first, games, G, k, budget, max_moves, seed, 2 * rnd + 1)
        want += line("synthetic-b", "synthetic-a", w, l, d)
    assert first.status() == 0 and second.status() == 0
    assert out.stdout == want
    for h in (first, second):
        h.close()


def test_cpp_evaluation_host_pits_models_of_a_directory(tmp_path):
    """--model-path: a match-up of two `model_NNNNNNN.ot` files of the directory (model_latest.ot is skipped,
    evaluation/src/main.rs:163-185), both loaded by the library's libtorch-free reader; the output lines match the
    pattern the reference's python/get_match_results.py extracts results with."""
    import re

    from takzero_b200 import weights

    tz_build.build()
    exe = os.path.join(os.path.dirname(capi.LIB_PATH), "bin", "evaluation")
    for i, name in enumerate(["model_0000000.ot", "model_0050000.ot", "model_0100000.ot", "model_latest.ot"]):
        weights.save_ot(str(tmp_path / name), weights.random_init(4, seed=10 + i, blocks=1))
    out = subprocess.run([exe, "--model-path", str(tmp_path), "--board", "4", "--half-komi", "4", "--games", "16",
                          "--sampled-actions", "8", "--budget", "24", "--max-moves", "40", "--seed", "3", "--rounds", "2",
                          "--arena-slots", str(1 << 14)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    pattern = re.compile(r"([\w\_\-]+)[\_\-](\d+)\.ot vs\. ([\w\d_\-]+)[\_\-](\d+)\.ot: "
                         r"Evaluation { wins: (\d*), losses: (\d*), draws: (\d*) }")  # python/get_match_results.py:8
    lines = out.stdout.splitlines()
    assert len(lines) == 4
    for a_line, b_line in zip(lines[::2], lines[1::2]):
        a, b = pattern.match(a_line), pattern.match(b_line)
        assert a and b, (a_line, b_line)
        assert "latest" not in a_line
        assert (a.group(2), a.group(4)) == (b.group(4), b.group(2)) and a.group(2) != a.group(4)  # colours swapped
        assert sum(int(x) for x in a.groups()[4:]) <= 16
