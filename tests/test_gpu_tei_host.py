"""GPU: the C++ TEI front-end (host/tei.cpp; reference tei/src/main.rs, tei/src/protocol.rs) against an
independent Python replay of the same command script through the C ABI: same best moves, same node counts,
same principal variations, info lines in the protocol's format."""
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from takzero_b200 import build as tz_build
from takzero_b200 import capi

from helpers import games_to_states

pytestmark = pytest.mark.gpu


def python_go(m, nodes):
    start = int(m.root_stats()[0]["visit_count"])
    while True:
        m.tree_simulate_batch(0.0, 128)
        visits = int(m.root_stats()[0]["visit_count"]) - start
        if visits >= nodes:
            break
    return visits, list(m.tree_principal_variation())


def test_tei_session_matches_python_replay():
    tz_build.build()
    exe = os.path.join(os.path.dirname(capi.LIB_PATH), "bin", "tei")
    n, hk = 5, 4
    # python side first, to know the engine's replies and extend the position with them (tree reuse)
    m = capi.BatchedMCTS(n, hk, 1, arena_slots=1 << 20, tree_batch=128)
    m.tree_simulate_simple(0.0)
    env = O.from_ptn_moves(n, hk, ["a1", "e5"])
    m.set_positions(games_to_states([env]))
    v1, pv1 = python_go(m, 2000)
    reply = O.possible_moves(_after(env, pv1[0]))[3]
    m.tree_descend(pv1[0])
    m.tree_descend(reply)
    v2, pv2 = python_go(m, 1000)
    m.close()

    proc = subprocess.Popen([exe, "--board", str(n), "--half-komi", str(hk), "--arena-slots", str(1 << 20)],
                            stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, bufsize=1)
    lines = []

    def send(cmd):
        proc.stdin.write(cmd + "\n")
        proc.stdin.flush()

    def read_until(prefix):
        while True:
            line = proc.stdout.readline()
            assert line, "engine exited: " + proc.stderr.read()
            lines.append(line.rstrip("\n"))
            if lines[-1].startswith(prefix):
                return

    send("tei")
    read_until("teiok")
    send("isready")
    read_until("readyok")
    send(f"teinewgame {n}")
    send("position startpos moves a1 e5")
    send("go nodes 2000")
    read_until("bestmove")
    send(f"position startpos moves a1 e5 {O.move_str(pv1[0])} {O.move_str(reply)}")  # extends: tree reuse
    send("go nodes 1000")
    read_until("bestmove")
    send("quit")
    assert proc.wait(timeout=60) == 0
    assert "teiok" in lines and "readyok" in lines
    best = [l.split()[1] for l in lines if l.startswith("bestmove")]
    assert best == [O.move_str(pv1[0]), O.move_str(pv2[0])]
    infos = [l for l in lines if l.startswith("info")]
    pat = re.compile(r"^info time \d+ nodes (\d+) nps \d+ wdl \d+ \d+ \d+( score mate -?\d+)? score cp -?\d+ pv( \S+)+$")
    assert infos and all(pat.match(l) for l in infos), infos
    # the last info line before each bestmove carries the final node count and PV of that search
    finals = [lines[i - 1] for i, l in enumerate(lines) if l.startswith("bestmove")]
    for line, visits, pv in zip(finals, (v1, v2), (pv1, pv2)):
        assert int(pat.match(line).group(1)) == visits
        assert line.split(" pv ")[1].split() == [O.move_str(x) for x in pv]


def _after(env, move):
    g = env.copy()
    O.play(g, move)
    return g
