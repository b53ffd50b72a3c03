"""ctypes binding of the CPU ORACLE (test infrastructure, NOT product code).

Loads ``oracle/libtakoracle.so`` (built from tak_rules.c + tak_search.c by
``oracle/Makefile``).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module; ``takzero_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Callable, List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtakoracle.so")

MAX_SQ = 36
MAX_MOVES = 1024

T_NONE, T_WIN, T_LOSS, T_DRAW = 0, 1, 2, 3
E_VALUE, E_WIN, E_LOSS, E_DRAW = 0, 1, 2, 3


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("tak_rules.c", "tak_search.c", "tak_batch.c", "tak_oracle.h")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libtakoracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class Game(C.Structure):
    _fields_ = [
        ("stack", C.c_uint64 * MAX_SQ),
        ("height", C.c_uint8 * MAX_SQ),
        ("top", C.c_uint8 * MAX_SQ),
        ("n", C.c_uint8),
        ("half_komi", C.c_int8),
        ("to_move", C.c_uint8),
        ("stones", C.c_uint8 * 2),
        ("caps", C.c_uint8 * 2),
        ("ply", C.c_uint16),
        ("reversible_plies", C.c_uint16),
        ("reversible_limit", C.c_uint16),
    ]

    def copy(self) -> "Game":
        g = Game()
        C.memmove(C.byref(g), C.byref(self), C.sizeof(Game))
        return g


class EvalU(C.Union):
    _fields_ = [("value", C.c_float), ("ply", C.c_uint32)]


class Eval(C.Structure):
    _fields_ = [("tag", C.c_uint32), ("u", EvalU)]

    def key(self):
        return (self.tag, self.u.ply if self.tag else np.float32(self.u.value).view(np.uint32).item())


class Node(C.Structure):
    pass


Node._fields_ = [
    ("evaluation", Eval),
    ("visit_count", C.c_uint32),
    ("logit", C.c_float),
    ("probability", C.c_float),
    ("std_dev", C.c_float),
    ("n_children", C.c_uint32),
    ("actions", C.POINTER(C.c_uint16)),
    ("children", C.POINTER(Node)),
]


class Counters(C.Structure):
    _fields_ = [("simulations", C.c_uint64), ("evaluations", C.c_uint64), ("known", C.c_uint64)]


AGENT_FN = C.CFUNCTYPE(
    None,
    C.c_void_p,
    C.c_int,
    C.POINTER(Game),
    C.POINTER(C.c_uint16),
    C.POINTER(C.c_int),
    C.c_int,
    C.POINTER(C.c_float),
    C.POINTER(C.c_float),
    C.POINTER(C.c_float),
)

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    P = C.POINTER
    L.tk_game_init.argtypes = [P(Game), C.c_int, C.c_int]
    L.tk_game_from_tps.argtypes = [P(Game), C.c_int, C.c_int, C.c_char_p]
    L.tk_game_from_tps.restype = C.c_int
    L.tk_game_to_tps.argtypes = [P(Game), C.c_char_p, C.c_int]
    L.tk_game_to_tps.restype = C.c_int
    L.tk_possible_moves.argtypes = [P(Game), P(C.c_uint16)]
    L.tk_possible_moves.restype = C.c_int
    L.tk_play.argtypes = [P(Game), C.c_uint16]
    L.tk_play.restype = C.c_int
    L.tk_result.argtypes = [P(Game)]
    L.tk_result.restype = C.c_int
    L.tk_terminal.argtypes = [P(Game)]
    L.tk_terminal.restype = C.c_int
    L.tk_flat_diff.argtypes = [P(Game)]
    L.tk_flat_diff.restype = C.c_int
    L.tk_state_hash.argtypes = [P(Game)]
    L.tk_state_hash.restype = C.c_uint64
    L.tk_new_opening.argtypes = [P(Game), C.c_int, C.c_int, C.c_int, C.c_int]
    L.tk_move_to_str.argtypes = [C.c_uint16, C.c_char_p]
    L.tk_move_to_str.restype = C.c_int
    L.tk_move_from_str.argtypes = [C.c_char_p, P(C.c_uint16)]
    L.tk_move_from_str.restype = C.c_int
    L.tk_move_order_key.argtypes = [C.c_uint16, C.c_int]
    L.tk_move_order_key.restype = C.c_int
    L.tk_input_channels.argtypes = [C.c_int]
    L.tk_input_channels.restype = C.c_int
    L.tk_output_channels.argtypes = [C.c_int]
    L.tk_output_channels.restype = C.c_int
    L.tk_move_index.argtypes = [C.c_int, C.c_uint16]
    L.tk_move_index.restype = C.c_int
    L.tk_game_repr.argtypes = [P(Game), P(C.c_float)]
    L.tk_eval_negate.argtypes = [Eval]
    L.tk_eval_negate.restype = Eval
    L.tk_eval_cmp.argtypes = [Eval, Eval]
    L.tk_eval_cmp.restype = C.c_int
    L.tk_eval_to_f32.argtypes = [Eval]
    L.tk_eval_to_f32.restype = C.c_float
    L.tk_softmax.argtypes = [P(C.c_float), C.c_int, P(C.c_float)]
    L.tk_set_exact_math.argtypes = [C.c_int]
    L.tk_expf_restated.argtypes = [C.c_float]
    L.tk_expf_restated.restype = C.c_float
    L.tk_node_new.restype = P(Node)
    L.tk_node_free.argtypes = [P(Node)]
    L.tk_node_reset.argtypes = [P(Node)]
    L.tk_node_simulate_simple.argtypes = [P(Node), P(Game), C.c_float, C.c_void_p, C.c_void_p]
    L.tk_node_simulate_simple.restype = C.c_int
    L.tk_node_simulate_batch.argtypes = [P(Node), P(Game), C.c_float, C.c_int, C.c_void_p, C.c_void_p]
    L.tk_node_descend.argtypes = [P(Node), C.c_uint16]
    L.tk_node_select_best_action.argtypes = [P(Node)]
    L.tk_node_select_best_action.restype = C.c_uint16
    L.tk_node_select_selfplay_action.argtypes = [P(Node), C.c_int, C.c_uint32, C.c_float, C.c_uint64]
    L.tk_node_select_selfplay_action.restype = C.c_uint16
    L.tk_node_ube_target.argtypes = [P(Node), C.c_float]
    L.tk_node_ube_target.restype = C.c_float
    L.tk_node_most_visited_count.argtypes = [P(Node)]
    L.tk_node_most_visited_count.restype = C.c_float
    L.tk_node_improved_policy.argtypes = [P(Node), C.c_float, P(C.c_float)]
    L.tk_node_principal_variation.argtypes = [P(Node), P(C.c_uint16), C.c_int]
    L.tk_node_principal_variation.restype = C.c_int
    L.tk_node_count.argtypes = [P(Node)]
    L.tk_node_count.restype = C.c_uint64
    L.tk_batched_from_envs.argtypes = [P(Game), C.c_int]
    L.tk_batched_from_envs.restype = C.c_void_p
    L.tk_batched_free.argtypes = [C.c_void_p]
    L.tk_batched_node.argtypes = [C.c_void_p, C.c_int]
    L.tk_batched_node.restype = P(Node)
    L.tk_batched_env.argtypes = [C.c_void_p, C.c_int]
    L.tk_batched_env.restype = P(Game)
    L.tk_batched_counters.argtypes = [C.c_void_p, P(Counters)]
    L.tk_batched_simulate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, P(C.c_float)]
    L.tk_batched_gumbel_sequential_halving.argtypes = [
        C.c_void_p, C.c_void_p, C.c_void_p, P(C.c_float), C.c_int, C.c_uint32, P(C.c_float), C.c_int,
        P(C.c_uint16),
    ]
    L.tk_batched_step.argtypes = [C.c_void_p, P(C.c_uint16)]
    L.tk_batched_restart_terminal_envs.argtypes = [C.c_void_p, P(C.c_int), P(C.c_int), P(C.c_int)]
    L.tk_batched_replay_len.argtypes = [C.c_void_p, C.c_int]
    L.tk_batched_replay_len.restype = C.c_int
    L.tk_batched_replay_actions.argtypes = [C.c_void_p, C.c_int]
    L.tk_batched_replay_actions.restype = P(C.c_uint16)
    L.tk_batched_select_best_actions.argtypes = [C.c_void_p, P(C.c_uint16)]
    L.tk_games_pack.argtypes = [P(Game), C.c_int, C.c_void_p]
    L.tk_games_unpack.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, P(Game)]
    L.tk_game_repr_batch.argtypes = [P(Game), C.c_int, C.c_void_p]
    L.tk_perft.argtypes = [P(Game), C.c_int]
    L.tk_perft.restype = C.c_ulonglong
    L.tk_move_index_batch.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.tk_game_result5.argtypes = [P(Game)]
    L.tk_game_result5.restype = C.c_int
    L.tk_has_road.argtypes = [P(Game), C.c_int]
    L.tk_has_road.restype = C.c_int
    L.tk_playout_positions.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, P(Game),
                                       C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, P(Game)]
    L.tk_playout_positions.restype = C.c_int
    L.tk_expf_compare.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, P(C.c_longlong), P(C.c_float)]
    L.tk_expf_compare.restype = C.c_longlong
    L.tk_expf_restated_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    _lib = L
    return L


# ---------------------------------------------------------------- rules helpers


def new_game(n: int, half_komi: int = 0) -> Game:
    g = Game()
    lib().tk_game_init(C.byref(g), n, half_komi)
    return g


def from_tps(n: int, half_komi: int, tps: str) -> Game:
    g = Game()
    rc = lib().tk_game_from_tps(C.byref(g), n, half_komi, tps.encode())
    if rc != 0:
        raise ValueError(f"bad TPS ({rc}): {tps}")
    return g


def to_tps(g: Game) -> str:
    buf = C.create_string_buffer(4096)
    lib().tk_game_to_tps(C.byref(g), buf, 4096)
    return buf.value.decode()


def new_opening(n: int, half_komi: int, symmetry: int, adjacent: int) -> Game:
    g = Game()
    lib().tk_new_opening(C.byref(g), n, half_komi, symmetry, adjacent)
    return g


def possible_moves(g: Game) -> List[int]:
    buf = (C.c_uint16 * MAX_MOVES)()
    k = lib().tk_possible_moves(C.byref(g), buf)
    return list(buf[:k])


def play(g: Game, m: int) -> None:
    rc = lib().tk_play(C.byref(g), m)
    if rc != 0:
        raise ValueError(f"illegal move {move_str(m)} ({rc})")


def move_str(m: int) -> str:
    buf = C.create_string_buffer(16)
    lib().tk_move_to_str(m, buf)
    return buf.value.decode()


def parse_move(s: str) -> int:
    out = C.c_uint16()
    if lib().tk_move_from_str(s.encode(), C.byref(out)) != 0:
        raise ValueError(f"bad move {s}")
    return out.value


def from_ptn_moves(n: int, half_komi: int, moves: Sequence[str]) -> Game:
    g = new_game(n, half_komi)
    for s in moves:
        play(g, parse_move(s))
    return g


def terminal(g: Game) -> int:
    return lib().tk_terminal(C.byref(g))


def result(g: Game) -> int:
    return lib().tk_result(C.byref(g))


def game_repr(g: Game) -> np.ndarray:
    n = g.n
    out = np.zeros(lib().tk_input_channels(n) * n * n, dtype=np.float32)
    lib().tk_game_repr(C.byref(g), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def move_index(n: int, m: int) -> int:
    return lib().tk_move_index(n, m)


def softmax(logits: Sequence[float]) -> np.ndarray:
    a = np.ascontiguousarray(logits, dtype=np.float32)
    out = np.zeros_like(a)
    lib().tk_softmax(a.ctypes.data_as(C.POINTER(C.c_float)), len(a), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def make_eval(tag: int, payload) -> Eval:
    e = Eval()
    e.tag = tag
    if tag == E_VALUE:
        e.u.value = payload
    else:
        e.u.ply = payload
    return e


# ---------------------------------------------------------------- bulk helpers (tak_batch.c)


def pack_games(games, count: Optional[int] = None) -> np.ndarray:
    """`count` oracle Games (a ctypes array or a list) as tz_state_t records: uint8 [count, 384]."""
    if not isinstance(games, C.Array):
        games = (Game * len(games))(*games)
    count = len(games) if count is None else count
    out = np.zeros((count, 384), dtype=np.uint8)
    lib().tk_games_pack(games, count, out.ctypes.data_as(C.c_void_p))
    return out


def unpack_games(states: np.ndarray, n: int, half_komi: int, reversible_limit: int = 100):
    """tz_state_t records (any dtype of itemsize 384, or uint8 [count, 384]) -> ctypes array of Games."""
    raw = np.ascontiguousarray(states).view(np.uint8).reshape(-1, 384)
    games = (Game * len(raw))()
    lib().tk_games_unpack(raw.ctypes.data_as(C.c_void_p), len(raw), n, half_komi, reversible_limit, games)
    return games


def game_repr_batch(games, count: Optional[int] = None) -> np.ndarray:
    if not isinstance(games, C.Array):
        games = (Game * len(games))(*games)
    count = len(games) if count is None else count
    n = games[0].n
    out = np.zeros((count, lib().tk_input_channels(n), n, n), dtype=np.float32)
    lib().tk_game_repr_batch(games, count, out.ctypes.data_as(C.c_void_p))
    return out


def playout_positions(n: int, half_komi: int, seed: int, policy: int, max_positions: int, reversible_limit: int = 100,
                      max_plies: int = 1000, stride: int = MAX_MOVES) -> dict:
    """Seeded playouts (tak_batch.c `tk_playout_positions`): every visited position with its ordered legal moves,
    terminal code, absolute result, the move played and the position after it."""
    games = (Game * max_positions)()
    nexts = (Game * max_positions)()
    n_moves = np.zeros(max_positions, dtype=np.int32)
    moves = np.zeros((max_positions, stride), dtype=np.uint16)
    terminal_ = np.zeros(max_positions, dtype=np.int32)
    result_ = np.zeros(max_positions, dtype=np.int32)
    chosen = np.zeros(max_positions, dtype=np.uint16)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    k = lib().tk_playout_positions(n, half_komi, reversible_limit, seed, policy, max_positions, max_plies, games,
                                   p(n_moves), p(moves), stride, p(terminal_), p(result_), p(chosen), nexts)
    assert k == max_positions
    return {"games": games, "states": pack_games(games), "next_states": pack_games(nexts), "n_moves": n_moves,
            "moves": moves, "terminal": terminal_, "result": result_, "chosen": chosen}


# ---------------------------------------------------------------- agents


def c_agent(name: str):
    """Address of a built-in C agent: 'dummy', 'simple' or 'synthetic'."""
    return C.cast(getattr(lib(), f"tk_agent_{name}"), C.c_void_p)


PyAgent = Callable[[List[Game], List[List[int]]], tuple]


def py_agent(fn: PyAgent):
    """Wrap fn(envs, actions) -> (list of logit arrays, values, variances)."""

    def tramp(_ctx, batch, envs, actions, n_actions, stride, logits, values, variances):
        env_list = [envs[i] for i in range(batch)]
        act = [[actions[i * stride + j] for j in range(n_actions[i])] for i in range(batch)]
        lg, v, u = fn(env_list, act)
        for i in range(batch):
            row = np.asarray(lg[i], dtype=np.float32)
            C.memmove(
                C.addressof(logits.contents) + 4 * i * stride, row.ctypes.data, 4 * n_actions[i]
            )
            values[i] = float(v[i])
            variances[i] = float(u[i])

    cb = AGENT_FN(tramp)
    return cb


def array_agent(fn):
    """Like py_agent without a Python loop per position: fn(games: ctypes Game array view, batch, actions uint16
    [batch, stride], n_actions int32 [batch]) -> (logits float32 [batch, stride], values [batch], variances [batch])."""

    def tramp(_ctx, batch, envs, actions, n_actions, stride, logits, values, variances):
        games = C.cast(envs, C.POINTER(Game * batch)).contents
        act = np.ctypeslib.as_array(actions, shape=(batch, stride))
        na = np.ctypeslib.as_array(n_actions, shape=(batch,))
        lg, v, u = fn(games, batch, act, na)
        np.ctypeslib.as_array(logits, shape=(batch, stride))[:] = lg
        np.ctypeslib.as_array(values, shape=(batch,))[:] = v
        np.ctypeslib.as_array(variances, shape=(batch,))[:] = u

    return AGENT_FN(tramp)


def _agent_ptr(agent):
    if isinstance(agent, str):
        return c_agent(agent)
    if isinstance(agent, C.c_void_p):
        return agent
    return C.cast(agent, C.c_void_p)


# ---------------------------------------------------------------- search wrappers


class Tree:
    """Single search tree (reference `Node<E>`, search/node/mod.rs:14-23)."""

    def __init__(self):
        self.ptr = lib().tk_node_new()

    def __del__(self):
        if getattr(self, "ptr", None):
            lib().tk_node_free(self.ptr)
            self.ptr = None

    @property
    def node(self) -> Node:
        return self.ptr.contents

    def simulate_simple(self, agent, env: Game, beta: float) -> int:
        return lib().tk_node_simulate_simple(self.ptr, C.byref(env), beta, _agent_ptr(agent), None)

    def simulate_batch(self, agent, env: Game, beta: float, batch_size: int) -> None:
        lib().tk_node_simulate_batch(self.ptr, C.byref(env), beta, batch_size, _agent_ptr(agent), None)

    def descend(self, m: int) -> None:
        lib().tk_node_descend(self.ptr, m)


def node_children(node: Node):
    return [(node.actions[i], node.children[i]) for i in range(node.n_children)]


def improved_policy(node_ptr, visitations: float) -> np.ndarray:
    node = node_ptr.contents if hasattr(node_ptr, "contents") else node_ptr
    out = np.zeros(max(1, node.n_children), dtype=np.float32)
    lib().tk_node_improved_policy(C.byref(node), visitations, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out[: node.n_children]


class Batched:
    """Reference `BatchedMCTS<BATCH_SIZE, E>` (search/node/batched.rs:24-409)."""

    def __init__(self, envs: Sequence[Game]):
        arr = (Game * len(envs))(*envs)
        self.batch = len(envs)
        self.ptr = C.c_void_p(lib().tk_batched_from_envs(arr, len(envs)))

    def __del__(self):
        if getattr(self, "ptr", None):
            lib().tk_batched_free(self.ptr)
            self.ptr = None

    def node(self, i: int) -> Node:
        return lib().tk_batched_node(self.ptr, i).contents

    def node_ptr(self, i: int):
        return lib().tk_batched_node(self.ptr, i)

    def env(self, i: int) -> Game:
        return lib().tk_batched_env(self.ptr, i).contents

    def counters(self) -> Counters:
        c = Counters()
        lib().tk_batched_counters(self.ptr, C.byref(c))
        return c

    def simulate(self, agent, betas: Sequence[float]) -> None:
        b = (C.c_float * self.batch)(*betas)
        lib().tk_batched_simulate(self.ptr, _agent_ptr(agent), None, b)

    def gumbel_sequential_halving(self, agent, betas, sampled_actions: int, search_budget: int,
                                  gumbel: np.ndarray) -> List[int]:
        b = (C.c_float * self.batch)(*betas)
        gum = np.ascontiguousarray(gumbel, dtype=np.float32)
        assert gum.ndim == 2 and gum.shape[0] == self.batch
        out = (C.c_uint16 * self.batch)()
        lib().tk_batched_gumbel_sequential_halving(
            self.ptr, _agent_ptr(agent), None, b, sampled_actions, search_budget,
            gum.ctypes.data_as(C.POINTER(C.c_float)), gum.shape[1], out,
        )
        return list(out)

    def step(self, actions: Sequence[int]) -> None:
        a = (C.c_uint16 * self.batch)(*actions)
        lib().tk_batched_step(self.ptr, a)

    def restart_terminal_envs(self, sym: Sequence[int], adj: Sequence[int]) -> List[int]:
        s = (C.c_int * self.batch)(*sym)
        a = (C.c_int * self.batch)(*adj)
        out = (C.c_int * self.batch)()
        lib().tk_batched_restart_terminal_envs(self.ptr, s, a, out)
        return list(out)

    def replay(self, i: int) -> List[int]:
        k = lib().tk_batched_replay_len(self.ptr, i)
        p = lib().tk_batched_replay_actions(self.ptr, i)
        return [p[j] for j in range(k)]

    def select_best_actions(self) -> List[int]:
        out = (C.c_uint16 * self.batch)()
        lib().tk_batched_select_best_actions(self.ptr, out)
        return list(out)


def perft(g: Game, depth: int) -> int:
    """Move sequences of length `depth` from `g` (none continues past a finished game)."""
    return int(lib().tk_perft(C.byref(g), depth))
