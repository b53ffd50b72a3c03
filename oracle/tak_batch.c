/*
 * tak_batch.c -- CPU ORACLE batch helpers (test infrastructure, NOT product code).
 *
 * Nothing here has semantics of its own: every function loops over the restated rules /
 * encoding of tak_rules.c (fast-tak `Game::{possible_moves, play, result}` as called from
 * takzero/src/search/env.rs:39-59, `game_repr` / `move_index` of network/repr.rs:49-71,169-228)
 * so that the parity tests can compare 10^5..10^6 positions per board size against the CUDA
 * library without a Python loop per position.
 */
#include <math.h>
#include <string.h>

#include "tak_oracle.h"

/* tz_state_t (include/takzero_b200.h): 384 bytes per position */
void tk_games_pack(const tk_game* games, int count, uint8_t* out) {
    for (int i = 0; i < count; i++) {
        const tk_game* g = &games[i];
        uint8_t* o = out + (size_t)i * 384;
        memset(o, 0, 384);
        memcpy(o, g->stack, 8 * TK_MAX_SQ);
        memcpy(o + 288, g->height, TK_MAX_SQ);
        memcpy(o + 324, g->top, TK_MAX_SQ);
        o[360] = g->to_move;
        o[361] = g->stones[0];
        o[362] = g->stones[1];
        o[363] = g->caps[0];
        o[364] = g->caps[1];
        memcpy(o + 366, &g->ply, 2);
        memcpy(o + 368, &g->reversible_plies, 2);
    }
}

void tk_games_unpack(const uint8_t* in, int count, int n, int half_komi, int reversible_limit, tk_game* games) {
    for (int i = 0; i < count; i++) {
        tk_game* g = &games[i];
        const uint8_t* s = in + (size_t)i * 384;
        memset(g, 0, sizeof(*g));
        memcpy(g->stack, s, 8 * TK_MAX_SQ);
        memcpy(g->height, s + 288, TK_MAX_SQ);
        memcpy(g->top, s + 324, TK_MAX_SQ);
        g->n = (uint8_t)n;
        g->half_komi = (int8_t)half_komi;
        g->to_move = s[360];
        g->stones[0] = s[361];
        g->stones[1] = s[362];
        g->caps[0] = s[363];
        g->caps[1] = s[364];
        memcpy(&g->ply, s + 366, 2);
        memcpy(&g->reversible_plies, s + 368, 2);
        g->reversible_limit = (uint16_t)reversible_limit;
    }
}

/* game_repr of `count` positions: out[count][C][N][N] */
void tk_game_repr_batch(const tk_game* games, int count, float* out) {
    if (count <= 0) return;
    const int n = games[0].n, len = tk_input_channels(n) * n * n;
    for (int i = 0; i < count; i++) tk_game_repr(&games[i], out + (size_t)i * len);
}

/* move_index of every listed action: out[count][stride] (cells past n_actions[i] are 0) */
void tk_move_index_batch(int n, const tk_move* actions, const int* n_actions, int stride, int count, int32_t* out) {
    for (int i = 0; i < count; i++)
        for (int j = 0; j < stride; j++)
            out[(size_t)i * stride + j] = j < n_actions[i] ? tk_move_index(n, actions[(size_t)i * stride + j]) : 0;
}

/* Game::result in the absolute coding of tz_game_result (the suffix of a Replay line, target.rs:226-230):
 * 0 ongoing, 1 "R-0", 2 "0-R", 3 "F-0", 4 "0-F", 5 "1/2-1/2".  Same order of checks as tk_result: the road of
 * the player who just moved, then the other road, then the flat count. */
int tk_game_result5(const tk_game* g) {
    const int r = tk_result(g);
    if (r == TK_ONGOING) return 0;
    if (r == TK_DRAW) return 5;
    const int mover = g->to_move ^ 1;
    if (tk_has_road(g, mover)) return mover == TK_WHITE ? 1 : 2;
    if (tk_has_road(g, g->to_move)) return g->to_move == TK_WHITE ? 1 : 2;
    return r == TK_WHITE_WIN ? 3 : 4;
}

static inline uint64_t splitmix(uint64_t* s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

/* Seeded playouts for the rules parity tests.  Plays games from random openings until `max_positions` positions
 * (terminal ones included) have been recorded; for each: the position, its ordered legal moves, terminal code,
 * absolute result, the move the playout chose (0xffff at a terminal position) and the position after it.
 * `policy` biases the choice so that every way a game can end shows up in bulk:
 *   0 uniform over the legal moves (mostly road wins)
 *   1 flat placements preferred (boards fill up: flat wins on a full board)
 *   2 walls preferred (no roads; reserves run out / boards fill with walls)
 *   3 spreads preferred (long games: reversible-ply draws with a small limit)
 *   4 per move either walls or spreads preferred (stacks grow while placements go on: reserves run out)
 * Returns the number of positions written. */
int tk_playout_positions(int n, int half_komi, int reversible_limit, uint64_t seed, int policy, int max_positions,
                         int max_plies, tk_game* out_games, int32_t* out_n_moves, tk_move* out_moves, int stride,
                         int32_t* out_terminal, int32_t* out_result, tk_move* out_chosen, tk_game* out_next) {
    uint64_t s = seed * 0x2545f4914f6cdd1dULL + 0x1234567ULL;
    int written = 0;
    tk_move moves[TK_MAX_MOVES];
    while (written < max_positions) {
        tk_game g;
        tk_new_opening(&g, n, half_komi, (int)(splitmix(&s) & 7), (int)(splitmix(&s) & 1));
        g.reversible_limit = (uint16_t)reversible_limit;
        for (int ply = 0; ply <= max_plies && written < max_positions; ply++) {
            const int term = tk_terminal(&g);
            const int cnt = tk_possible_moves(&g, moves);
            const int i = written++;
            out_games[i] = g;
            out_n_moves[i] = cnt;
            for (int j = 0; j < cnt && j < stride; j++) out_moves[(size_t)i * stride + j] = moves[j];
            out_terminal[i] = term;
            out_result[i] = tk_game_result5(&g);
            out_chosen[i] = 0xffff;
            out_next[i] = g;
            if (term != TK_T_NONE || cnt == 0 || ply == max_plies) break;
            /* pick: with probability 7/8 among the preferred kind (when there is one), else uniformly */
            int pick = (int)(splitmix(&s) % (uint64_t)cnt);
            if (policy != 0 && (splitmix(&s) & 7) != 0) {
                int pref[TK_MAX_MOVES], np = 0;
                const int pol = policy == 4 ? ((splitmix(&s) & 1) ? 2 : 3) : policy;
                for (int j = 0; j < cnt; j++) {
                    const int pat = (moves[j] >> 8) & 0xff, kind = (moves[j] >> 6) & 3;
                    const int is = pol == 1 ? (pat == 0 && kind == TK_FLAT)
                                   : pol == 2 ? (pat == 0 && kind == TK_WALL)
                                              : (pat != 0);
                    if (is) pref[np++] = j;
                }
                if (np > 0) pick = pref[splitmix(&s) % (uint64_t)np];
            }
            out_chosen[i] = moves[pick];
            tk_play_unchecked(&g, moves[pick]);
            out_next[i] = g;
        }
    }
    return written;
}

/* expf: host libm (mode 0 of tak_search.c = what Rust's f32::exp calls) against the restated glibc algorithm
 * (mode 1 = what the CUDA library executes) on every `step`-th f32 bit pattern in [lo_bits, hi_bits].  Returns the
 * number of inputs whose results differ in any bit; *first_bad receives the first such input. */
long long tk_expf_compare(uint32_t lo_bits, uint32_t hi_bits, uint32_t step, long long* tested, float* first_bad) {
    long long bad = 0, n = 0;
    for (uint64_t b = lo_bits; b <= hi_bits; b += step) {
        const uint32_t u = (uint32_t)b;
        float x;
        memcpy(&x, &u, 4);
        const float a = expf(x), r = tk_expf_restated(x);
        uint32_t ua, ur;
        memcpy(&ua, &a, 4);
        memcpy(&ur, &r, 4);
        if (ua != ur && !(a != a && r != r)) {
            if (bad == 0 && first_bad) *first_bad = x;
            bad++;
        }
        n++;
    }
    if (tested) *tested = n;
    return bad;
}

void tk_expf_restated_batch(const float* in, int count, float* out) {
    for (int i = 0; i < count; i++) out[i] = tk_expf_restated(in[i]);
}

/* perft: number of move sequences of length `depth` from `g`, none continuing past a finished game -- the usual
 * known-answer test of Tak move generators.  Loops tk_terminal / tk_possible_moves / tk_play only. */
unsigned long long tk_perft(const tk_game* g, int depth) {
    if (depth <= 0) return 1;
    if (tk_terminal(g) != TK_T_NONE) return 0;
    tk_move moves[TK_MAX_MOVES];
    const int n = tk_possible_moves(g, moves);
    if (depth == 1) return (unsigned long long)n;
    unsigned long long total = 0;
    for (int i = 0; i < n; i++) {
        tk_game child = *g;
        tk_play_unchecked(&child, moves[i]);
        total += tk_perft(&child, depth - 1);
    }
    return total;
}
