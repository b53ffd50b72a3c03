"""GPU: the C++ `analysis` REPL (host/analysis.cpp; reference analysis/src/main.rs + `impl Display for Node`,
search/node/debug.rs:11-95) against a Python replay of the same session through the ctypes binding, formatted
independently (oracle TPS / move notation, Rust-style centring and `{:+.4}`)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from takzero_b200 import build as tz_build
from takzero_b200 import capi

from helpers import state_to_game

pytestmark = pytest.mark.gpu

LOGF = ctypes.CDLL("libm.so.6").logf
LOGF.restype = ctypes.c_float
LOGF.argtypes = [ctypes.c_float]


def center(s, width):
    pad = max(0, width - len(s))
    return " " * (pad // 2) + s + " " * (pad - pad // 2)


def eval_display(tag, bits):
    if tag == capi.E_VALUE:
        return "%+.4f" % float(np.uint32(bits).view(np.float32))
    return "%s(%d)" % ({capi.E_WIN: "Win", capi.E_LOSS: "Loss", capi.E_DRAW: "Draw"}[tag], bits)


def node_display(m):
    ch, root = m.root_children(), m.root_stats()[0]
    pol, _, _ = m.targets(-1.0, 0.0)
    n = int(ch["n"][0])
    out = ""
    if n == 0 and root["eval_tag"] == capi.E_VALUE:
        out += "--- This node still needs to be initialized! ---\n"
    else:
        order = sorted(range(n), key=lambda i: int(ch["visits"][0, i]))  # stable
        parent = np.float32(root["visit_count"])
        rate = np.float32(LOGF((np.float32(1.0) + parent + np.float32(500.0)) / np.float32(500.0))) + np.float32(4.0)
        for i in order:
            puct = rate * ch["prob"][0, i] * np.sqrt(parent) / (np.float32(1.0) + np.float32(ch["visits"][0, i]))
            out += " ".join([
                center(O.move_str(int(ch["moves"][0, i])), 10), center(str(int(ch["visits"][0, i])), 9),
                center("%+.4f" % float(ch["logit"][0, i]), 9), center("%.4f" % float(ch["prob"][0, i]), 9),
                center("%.4f" % float(pol[0, i]), 9), center("%.4f" % float(np.float32(puct)), 8),
                center("%.4f" % float(ch["std_dev"][0, i]), 9),
                center(eval_display(int(ch["eval_tag"][0, i]), int(ch["eval_bits"][0, i])), 14)]) + "\n"
        out += "[ action ] [ count ] [ logit ] [ proba ] [ impol ] [ puct ] [ stdev ] [ evaluation ]\n"
    out += "((node))  [count: %d]  [std_dev: %.4f]  [eval: %s]\n" % (
        int(root["visit_count"]), float(np.uint32(root["std_dev_bits"]).view(np.float32)),
        eval_display(int(root["eval_tag"]), int(root["eval_bits"])))
    return out


def test_analysis_repl_matches_python_replay():
    tz_build.build()
    exe = os.path.join(os.path.dirname(capi.LIB_PATH), "bin", "analysis")
    n, hk, batch = 4, 4, 32
    m = capi.BatchedMCTS(n, hk, 1, arena_slots=1 << 16, tree_batch=batch)
    start = m.positions()
    m.set_positions(start)

    def tps():
        return O.to_tps(state_to_game(m.positions()[0], n, hk))

    # the session: search twice, play the best move, search, a line that is no move (searches), an illegal move
    # (reported on stderr, position unchanged), then a reply and one more search
    want, script = "", []
    want += "tps: %s\n>>> " % tps()
    for step in ("", "", "best", "", "hello", "Ca1", "reply", ""):
        if step in ("best", "reply"):
            mv = int(m.select_best_actions()[0]) if step == "best" else int(m.legal_moves(m.positions())[0][0, 3])
            script.append(O.move_str(mv))
            m.tree_descend(mv)
        elif step == "Ca1":  # 4x4 has no capstones: illegal, nothing is printed on stdout but the next prompt
            script.append(step)
            want += "tps: %s\n>>> " % tps()
            continue
        else:
            script.append(step)
            m.tree_simulate_batch(0.0, batch)
        want += node_display(m) + "\n" + "tps: %s\n>>> " % tps()
    assert m.status() == 0
    m.close()
    out = subprocess.run([exe, "--board", str(n), "--half-komi", str(hk), "--batch-size", str(batch),
                          "--arena-slots", str(1 << 16)], input="\n".join(script) + "\n", capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stderr
    assert "illegal move Ca1" in out.stderr
    assert out.stdout == want
    assert out.stdout.count("[ action ] [ count ]") == 7


def test_analysis_example_plays_a_whole_game():
    tz_build.build()
    exe = os.path.join(os.path.dirname(capi.LIB_PATH), "bin", "analysis")
    out = subprocess.run([exe, "--board", "3", "--half-komi", "0", "--batch-size", "16", "--example",
                          "--arena-slots", str(1 << 16)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.splitlines()
    assert lines[0] == "tps: x3/x3/x3 1 1" and lines[1].startswith(">>> ")
    # replay the printed moves with the oracle: every one is legal and the game ends exactly at the end
    g = O.from_tps(3, 0, "x3/x3/x3 1 1")
    for t, mv in zip(lines[::2], lines[1::2]):
        assert t == "tps: " + O.to_tps(g)
        assert O.terminal(g) == O.T_NONE
        O.play(g, O.parse_move(mv[4:]))
    assert O.terminal(g) != O.T_NONE
