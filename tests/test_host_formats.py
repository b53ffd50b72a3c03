"""CPU: the host-side text layer of include/takzero_b200.hpp (TPS / PTN move notation, Rust-style f32
printing, Replay parsing, Eval walk-back) against the oracle's independent implementations and numpy.
Reference formats: takzero/src/target.rs:56-73,215-272, takzero/src/search/eval.rs:40-47,95-105."""
import os
import subprocess

import numpy as np

from oracle import oracle as O
from takzero_b200 import build as tz_build

from helpers import random_playout_states


def run(lines):
    tz_build.build()
    exe = os.path.join(os.path.dirname(tz_build.LIB), "bin", "format_check")
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    return out.stdout.splitlines()


def test_tps_and_moves_round_trip_against_oracle():
    cmds, want = [], []
    for n, hk in ((3, 0), (4, 4), (5, 4), (6, 4)):
        for seed in range(6):
            for g in random_playout_states(n, hk, 77 * n + seed)[::3]:
                t = O.to_tps(g)
                cmds.append(f"tps {n} {t}")
                want.append(f"{t} | {g.stones[0]} {g.stones[1]} {g.caps[0]} {g.caps[1]} {g.ply}")
                for m in O.possible_moves(g)[:40]:
                    cmds.append(f"move {O.move_str(m)}")
                    want.append(f"{m} {O.move_str(m)}")
    got = run(cmds)
    assert len(got) == len(want) > 2000
    assert got == want


def test_f32_display_is_shortest_round_trip():
    rng = np.random.default_rng(0)
    vals = np.concatenate([
        rng.random(2000, dtype=np.float32), (rng.random(500, dtype=np.float32) - 0.5) * 2e-4,
        np.float32(0.997) ** np.arange(0, 200, dtype=np.float32),
        np.array([0.0, 1.0, -1.0, 0.5, 4.0, 1e-7, 123456.78, 3.4e38], dtype=np.float32)]).astype(np.float32)
    got = run([f"f32 {int(v.view(np.uint32)):08x}" for v in vals])
    for v, text in zip(vals, got):
        assert text == np.format_float_positional(v, unique=True, trim="-"), (v, text)
        assert np.float32(text) == v and "e" not in text


def test_replay_parse_and_eval_walk_back():
    line = '[TPS "x3,1/x4/x4/2,x3 1 2"] b2 Sc3 a1+ 2a2>11 b2< R-0'
    got = run([f"replay 4 {line}"])
    assert got == ['[TPS "x3,1/x4/x4/2,x3 1 2"] b2 Sc3 a1+ 2a2>11 b2<']
    cmds, want = [], []
    for tag in (1, 2, 3):
        for k in (0, 1, 2, 7, 30):
            e = O.make_eval(tag, 0)
            for _ in range(k):
                e = O.lib().tk_eval_negate(e)
            cmds.append(f"eval {tag} 0 {k}")
            want.append("%u %u %08x" % (e.tag, e.u.ply, int(np.float32(O.lib().tk_eval_to_f32(e)).view(np.uint32))))
    assert run(cmds) == want
