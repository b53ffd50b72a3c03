/*
 * tak_rules.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Tak rules + state as used by the reference through the un-vendored crate
 * fast-tak 0.4.1 (Cargo.lock:611-616) and takparse 0.6.0 (Cargo.lock:1564-1569).
 * The crate sources are absent from /root/reference, so this file restates the
 * published rules of Tak and is anchored on the reference's own call sites and
 * golden vectors:
 *   - call sites: takzero/src/search/env.rs:39-79 (possible_moves, play, result,
 *     new_opening), takzero/src/network/repr.rs:49-71,169-228 (move_index, game_repr)
 *   - legal moves + index mapping: repr.rs:411-499 (3x3, 18 legal moves)
 *   - encoding: repr.rs:260-409 (three full plane vectors)
 *   - opening swap / road detection: search/node/mcts.rs:345-411 (two tinue KATs)
 *   - move-generation ORDER: derived from runs/{*}.txt (20 x 1024 ordered 5x5 lists)
 * Parity unpinned (nothing in the reference fixes them): reversible-ply draw
 * threshold (kept as a field), symmetry index order of new_opening.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tak_oracle.h"

static const int DROW[4] = {+1, -1, 0, 0}; /* Up, Down, Left, Right */
static const int DCOL[4] = {0, 0, -1, +1};

static inline int mv_col(tk_move m) { return m & 7; }
static inline int mv_row(tk_move m) { return (m >> 3) & 7; }
static inline int mv_kind(tk_move m) { return (m >> 6) & 3; }
static inline int mv_pat(tk_move m) { return (m >> 8) & 0xff; }
static inline tk_move mk_move(int row, int col, int kind, int pat) {
    return (tk_move)(col | (row << 3) | (kind << 6) | (pat << 8));
}

static void initial_reserves(int n, int* stones, int* caps) {
    /* standard Tak piece counts; consistent with repr.rs:303-409 (3: 10/0, 5: 21/1) */
    static const int S[9] = {0, 0, 0, 10, 15, 21, 30, 40, 50};
    static const int C[9] = {0, 0, 0, 0, 0, 1, 1, 2, 2};
    *stones = S[n];
    *caps = C[n];
}

void tk_game_init(tk_game* g, int n, int half_komi) {
    memset(g, 0, sizeof(*g));
    g->n = (uint8_t)n;
    g->half_komi = (int8_t)half_komi;
    int s, c;
    initial_reserves(n, &s, &c);
    g->stones[0] = g->stones[1] = (uint8_t)s;
    g->caps[0] = g->caps[1] = (uint8_t)c;
    g->to_move = TK_WHITE;
    g->reversible_limit = 100;
}

static inline int top_color(const tk_game* g, int sq) {
    return (int)((g->stack[sq] >> (g->height[sq] - 1)) & 1);
}

/* fast-tak `Game::possible_moves` order, derived from runs/{*}.txt:
 * squares file-major (a1,a2,..,b1,..); empty -> flat, wall, cap; own stack ->
 * carry 1..min(h,N) x direction (+,-,<,>) x drop sequences in descending
 * lexicographic order. */
int tk_possible_moves(const tk_game* g, tk_move* out) {
    const int n = g->n, me = g->to_move;
    int cnt = 0;
    if (g->ply < 2) {
        for (int col = 0; col < n; col++)
            for (int row = 0; row < n; row++)
                if (g->height[row * n + col] == 0) out[cnt++] = mk_move(row, col, TK_FLAT, 0);
        return cnt;
    }
    for (int col = 0; col < n; col++) {
        for (int row = 0; row < n; row++) {
            const int sq = row * n + col;
            const int h = g->height[sq];
            if (h == 0) {
                if (g->stones[me] > 0) {
                    out[cnt++] = mk_move(row, col, TK_FLAT, 0);
                    out[cnt++] = mk_move(row, col, TK_WALL, 0);
                }
                if (g->caps[me] > 0) out[cnt++] = mk_move(row, col, TK_CAP, 0);
                continue;
            }
            if (top_color(g, sq) != me) continue;
            const int maxc = h < n ? h : n;
            for (int c = 1; c <= maxc; c++) {
                for (int d = 0; d < 4; d++) {
                    int reach = 0, smash = 0;
                    int r = row, cc = col;
                    for (int step = 1; step <= c; step++) {
                        r += DROW[d];
                        cc += DCOL[d];
                        if (r < 0 || r >= n || cc < 0 || cc >= n) break;
                        const int t = r * n + cc;
                        if (g->height[t] > 0 && g->top[t] != TK_FLAT) {
                            if (g->top[t] == TK_WALL && g->top[sq] == TK_CAP) smash = 1;
                            break;
                        }
                        reach = step;
                    }
                    /* rev: bit (c-1-i) = "piece i (from the bottom of the carried
                     * pieces) starts a new drop"; piece 0 always does. Ascending
                     * rev == descending lexicographic drop sequence. */
                    for (int rev = 1 << (c - 1); rev < (1 << c); rev++) {
                        const int parts = __builtin_popcount((unsigned)rev);
                        if (!(parts <= reach || (smash && parts == reach + 1 && (rev & 1))))
                            continue;
                        int pat = 0;
                        for (int i = 0; i < c; i++)
                            if ((rev >> (c - 1 - i)) & 1) pat |= 1 << (8 - c + i);
                        out[cnt++] = mk_move(row, col, d, pat);
                    }
                }
            }
        }
    }
    return cnt;
}

/* Applies `m`; returns 0, or a negative PlayError-like code and leaves *g untouched. */
int tk_play(tk_game* g, tk_move m) {
    const int n = g->n, me = g->to_move;
    const int row = mv_row(m), col = mv_col(m), kind = mv_kind(m), pat = mv_pat(m);
    if (row >= n || col >= n) return -1;
    const int sq = row * n + col;
    tk_game t = *g;
    if (pat == 0) {
        if (kind > TK_CAP) return -2;
        if (t.height[sq] != 0) return -3;
        int color = me;
        if (t.ply < 2) {
            if (kind != TK_FLAT) return -4;
            color = me ^ 1; /* opening swap (mcts.rs:352 needs it) */
        }
        if (kind == TK_CAP) {
            if (t.caps[color] == 0) return -5;
            t.caps[color]--;
        } else {
            if (t.stones[color] == 0) return -5;
            t.stones[color]--;
        }
        t.stack[sq] = (uint64_t)color;
        t.height[sq] = 1;
        t.top[sq] = (uint8_t)kind;
        t.reversible_plies = 0;
    } else {
        if (t.ply < 2) return -4;
        const int h = t.height[sq];
        if (h == 0 || top_color(&t, sq) != me) return -6;
        const int c = 8 - __builtin_ctz((unsigned)pat);
        if (c > h || c > n) return -7;
        const uint64_t carried = (t.stack[sq] >> (h - c)) & ((1ull << c) - 1);
        const int toptype = t.top[sq];
        t.height[sq] = (uint8_t)(h - c);
        t.stack[sq] &= (1ull << (h - c)) - 1;
        t.top[sq] = TK_FLAT;
        int r = row, cc = col, pos = sq, smashed = 0;
        for (int i = 0; i < c; i++) {
            if ((pat >> (8 - c + i)) & 1) {
                r += DROW[kind];
                cc += DCOL[kind];
                if (r < 0 || r >= n || cc < 0 || cc >= n) return -8;
                pos = r * n + cc;
                if (t.height[pos] > 0 && t.top[pos] != TK_FLAT) {
                    if (t.top[pos] == TK_WALL && toptype == TK_CAP && i == c - 1) {
                        t.top[pos] = TK_FLAT;
                        smashed = 1;
                    } else {
                        return -9;
                    }
                }
            }
            t.stack[pos] |= ((carried >> i) & 1ull) << t.height[pos];
            t.height[pos]++;
            t.top[pos] = TK_FLAT;
        }
        t.top[pos] = (uint8_t)toptype;
        t.reversible_plies = smashed ? 0 : (uint16_t)(t.reversible_plies + 1);
    }
    t.ply++;
    t.to_move ^= 1;
    *g = t;
    return 0;
}

void tk_play_unchecked(tk_game* g, tk_move m) {
    if (tk_play(g, m) != 0) {
        /* env.rs:44 `.expect("Action should be valid")` panics */
        char buf[16];
        tk_move_to_str(m, buf);
        fprintf(stderr, "tak_oracle: Action should be valid (%s)\n", buf);
        abort();
    }
}

int tk_has_road(const tk_game* g, int color) {
    const int n = g->n;
    uint8_t road[TK_MAX_SQ];
    for (int sq = 0; sq < n * n; sq++)
        road[sq] = g->height[sq] > 0 && g->top[sq] != TK_WALL && top_color(g, sq) == color;
    /* two searches: rank 1 -> rank N, and file a -> last file */
    for (int pass = 0; pass < 2; pass++) {
        uint8_t seen[TK_MAX_SQ] = {0};
        int queue[TK_MAX_SQ], qh = 0, qt = 0;
        for (int i = 0; i < n; i++) {
            const int sq = pass == 0 ? i : i * n; /* row 0 / col 0 */
            if (road[sq]) {
                seen[sq] = 1;
                queue[qt++] = sq;
            }
        }
        while (qh < qt) {
            const int sq = queue[qh++];
            const int r = sq / n, c = sq % n;
            if ((pass == 0 && r == n - 1) || (pass == 1 && c == n - 1)) return 1;
            for (int d = 0; d < 4; d++) {
                const int rr = r + DROW[d], cc = c + DCOL[d];
                if (rr < 0 || rr >= n || cc < 0 || cc >= n) continue;
                const int t = rr * n + cc;
                if (road[t] && !seen[t]) {
                    seen[t] = 1;
                    queue[qt++] = t;
                }
            }
        }
    }
    return 0;
}

/* white top flats - black top flats (walls/caps excluded: repr.rs:404-408) */
int tk_flat_diff(const tk_game* g) {
    int diff = 0;
    for (int sq = 0; sq < g->n * g->n; sq++)
        if (g->height[sq] > 0 && g->top[sq] == TK_FLAT) diff += top_color(g, sq) == TK_WHITE ? 1 : -1;
    return diff;
}

/* fast-tak `Game::result` as the reference consumes it (env.rs:47-59): the
 * player who just moved is checked for a road first. */
int tk_result(const tk_game* g) {
    const int mover = g->to_move ^ 1;
    if (tk_has_road(g, mover)) return mover == TK_WHITE ? TK_WHITE_WIN : TK_BLACK_WIN;
    if (tk_has_road(g, g->to_move)) return g->to_move == TK_WHITE ? TK_WHITE_WIN : TK_BLACK_WIN;
    int full = 1;
    for (int sq = 0; sq < g->n * g->n; sq++)
        if (g->height[sq] == 0) full = 0;
    const int w_out = g->stones[0] == 0 && g->caps[0] == 0;
    const int b_out = g->stones[1] == 0 && g->caps[1] == 0;
    if (full || w_out || b_out) {
        const int score2 = 2 * tk_flat_diff(g) - g->half_komi;
        if (score2 > 0) return TK_WHITE_WIN;
        if (score2 < 0) return TK_BLACK_WIN;
        return TK_DRAW;
    }
    if (g->reversible_plies >= g->reversible_limit) return TK_DRAW;
    return TK_ONGOING;
}

int tk_terminal(const tk_game* g) {
    const int r = tk_result(g);
    if (r == TK_ONGOING) return TK_T_NONE;
    if (r == TK_DRAW) return TK_T_DRAW;
    const int winner = r == TK_WHITE_WIN ? TK_WHITE : TK_BLACK;
    return winner == g->to_move ? TK_T_WIN : TK_T_LOSS;
}

/* env.rs:65-79: two flat placements on opposite (a1,xN) or adjacent (a1,aN)
 * corners under one of 8 symmetries.  The crate's symmetry index order is
 * unpinned; ours: bit0 = mirror columns, bit1 = mirror rows, bit2 = transpose. */
void tk_new_opening(tk_game* g, int n, int half_komi, int symmetry, int adjacent) {
    tk_game_init(g, n, half_komi);
    const int squares[2][2] = {{0, 0}, {adjacent ? 0 : n - 1, n - 1}}; /* {col,row} */
    for (int i = 0; i < 2; i++) {
        int col = squares[i][0], row = squares[i][1];
        if (symmetry & 1) col = n - 1 - col;
        if (symmetry & 2) row = n - 1 - row;
        if (symmetry & 4) {
            const int t = col;
            col = row;
            row = t;
        }
        tk_play_unchecked(g, mk_move(row, col, TK_FLAT, 0));
    }
}

/* 64-bit state hash used by the synthetic agent (shared definition with the
 * CUDA library: takzero_b200/csrc/state.cuh `state_hash`). */
static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

uint64_t tk_state_hash(const tk_game* g) {
    uint64_t h = 0x9e3779b97f4a7c15ULL ^ (uint64_t)g->to_move;
    for (int sq = 0; sq < g->n * g->n; sq++) {
        const uint64_t hh = g->height[sq];
        const uint64_t tt = hh ? g->top[sq] : 0;
        /* per-square contribution is order independent (sum), so a warp can
         * reduce it in any order */
        h += mix64(g->stack[sq] * 0x100000001b3ULL + (hh << 8) + (tt << 16) + ((uint64_t)(sq + 1) << 24));
    }
    return mix64(h);
}

/* ---- notation -------------------------------------------------------- */

int tk_move_to_str(tk_move m, char* buf) {
    char* p = buf;
    const int pat = mv_pat(m);
    if (pat == 0) {
        const int k = mv_kind(m);
        if (k == TK_WALL) *p++ = 'S';
        if (k == TK_CAP) *p++ = 'C';
        *p++ = (char)('a' + mv_col(m));
        *p++ = (char)('1' + mv_row(m));
    } else {
        const int c = 8 - __builtin_ctz((unsigned)pat);
        if (c > 1) *p++ = (char)('0' + c);
        *p++ = (char)('a' + mv_col(m));
        *p++ = (char)('1' + mv_row(m));
        *p++ = "+-<>"[mv_kind(m)];
        int drops[8], nd = 0;
        for (int i = 0; i < c; i++) {
            if ((pat >> (8 - c + i)) & 1) drops[nd++] = 0;
            drops[nd - 1]++;
        }
        if (nd > 1)
            for (int i = 0; i < nd; i++) *p++ = (char)('0' + drops[i]);
    }
    *p = 0;
    return (int)(p - buf);
}

int tk_move_from_str(const char* s, tk_move* out) {
    int kind = TK_FLAT, c = 0;
    if (*s == 'S') {
        kind = TK_WALL;
        s++;
    } else if (*s == 'C') {
        kind = TK_CAP;
        s++;
    } else if (*s == 'F') {
        s++;
    }
    if (*s >= '1' && *s <= '8') c = *s++ - '0';
    if (*s < 'a' || *s > 'h') return -1;
    const int col = *s++ - 'a';
    if (*s < '1' || *s > '8') return -1;
    const int row = *s++ - '1';
    if (*s == 0) {
        if (c != 0) return -1;
        *out = mk_move(row, col, kind, 0);
        return 0;
    }
    int dir;
    switch (*s++) {
        case '+': dir = TK_UP; break;
        case '-': dir = TK_DOWN; break;
        case '<': dir = TK_LEFT; break;
        case '>': dir = TK_RIGHT; break;
        default: return -1;
    }
    if (c == 0) c = 1;
    int drops[8], nd = 0, total = 0;
    while (*s >= '1' && *s <= '8' && nd < 8) {
        drops[nd] = *s++ - '0';
        total += drops[nd++];
    }
    while (*s == '*' || *s == '\'' || *s == '!' || *s == '?') s++;
    if (*s != 0) return -1;
    if (nd == 0) {
        drops[nd++] = c;
        total = c;
    }
    if (total != c || c > 8) return -1;
    int pat = 0, i = 0;
    for (int d = 0; d < nd; d++) {
        pat |= 1 << (8 - c + i);
        i += drops[d];
    }
    *out = mk_move(row, col, dir, pat);
    return 0;
}

/* Sort key that is strictly increasing along tk_possible_moves (for checking the
 * ordering rule against runs/{*}.txt without knowing the positions). */
int tk_move_order_key(tk_move m, int n) {
    const int sqkey = mv_col(m) * n + mv_row(m);
    const int pat = mv_pat(m);
    int inner;
    if (pat == 0) {
        inner = mv_kind(m);
    } else {
        const int c = 8 - __builtin_ctz((unsigned)pat);
        int rev = 0;
        for (int i = 0; i < c; i++)
            if ((pat >> (8 - c + i)) & 1) rev |= 1 << (c - 1 - i);
        inner = 3 + ((c * 4 + mv_kind(m)) << 8) + rev;
    }
    return (sqkey << 16) | inner;
}

int tk_game_from_tps(tk_game* g, int n, int half_komi, const char* tps) {
    tk_game_init(g, n, half_komi);
    const char* s = tps;
    int row = n - 1, col = 0;
    while (*s && *s != ' ') {
        if (*s == '/') {
            if (col != n) return -1;
            row--;
            col = 0;
            s++;
        } else if (*s == ',') {
            s++;
        } else if (*s == 'x') {
            s++;
            int k = 1;
            if (*s >= '1' && *s <= '8') k = *s++ - '0';
            col += k;
        } else if (*s == '1' || *s == '2') {
            if (row < 0 || col >= n) return -1;
            const int sq = row * n + col;
            int h = 0;
            uint64_t bits = 0;
            int color = 0;
            while (*s == '1' || *s == '2') {
                color = *s - '1';
                bits |= (uint64_t)color << h;
                h++;
                s++;
            }
            int type = TK_FLAT;
            if (*s == 'S') {
                type = TK_WALL;
                s++;
            } else if (*s == 'C') {
                type = TK_CAP;
                s++;
            }
            g->stack[sq] = bits;
            g->height[sq] = (uint8_t)h;
            g->top[sq] = (uint8_t)type;
            for (int i = 0; i < h; i++) {
                const int cl = (int)((bits >> i) & 1);
                if (i == h - 1 && type == TK_CAP) {
                    if (g->caps[cl] == 0) return -2;
                    g->caps[cl]--;
                } else {
                    if (g->stones[cl] == 0) return -2;
                    g->stones[cl]--;
                }
            }
            col++;
        } else {
            return -1;
        }
    }
    if (row != 0 || col != n) return -1;
    int player = 1, move_no = 1;
    if (sscanf(s, " %d %d", &player, &move_no) != 2) return -3;
    g->to_move = (uint8_t)(player - 1);
    g->ply = (uint16_t)((move_no - 1) * 2 + (player - 1));
    g->reversible_plies = 0;
    return 0;
}

int tk_game_to_tps(const tk_game* g, char* buf, int buflen) {
    const int n = g->n;
    char tmp[4096];
    char* p = tmp;
    for (int row = n - 1; row >= 0; row--) {
        int empties = 0, first = 1;
        for (int col = 0; col <= n; col++) {
            const int sq = row * n + col;
            if (col < n && g->height[sq] == 0) {
                empties++;
                continue;
            }
            if (empties) {
                if (!first) *p++ = ',';
                *p++ = 'x';
                if (empties > 1) *p++ = (char)('0' + empties);
                empties = 0;
                first = 0;
            }
            if (col == n) break;
            if (!first) *p++ = ',';
            first = 0;
            for (int i = 0; i < g->height[sq]; i++) *p++ = (char)('1' + ((g->stack[sq] >> i) & 1));
            if (g->top[sq] == TK_WALL) *p++ = 'S';
            if (g->top[sq] == TK_CAP) *p++ = 'C';
        }
        if (row > 0) *p++ = '/';
    }
    p += sprintf(p, " %d %d", g->to_move + 1, g->ply / 2 + 1);
    const int len = (int)(p - tmp);
    if (len + 1 > buflen) return -1;
    memcpy(buf, tmp, (size_t)len + 1);
    return len;
}

/* ---- network/repr.rs -------------------------------------------------- */

static int stack_size(int n) { return 3 + (n - 1) + (n + 1); } /* repr.rs:119-125 */

int tk_input_channels(int n) { return 2 * (stack_size(n) + 2) + 1 + 1; } /* repr.rs:133-139 */

int tk_output_channels(int n) { return 3 + 4 * ((1 << n) - 2); } /* repr.rs:103-108 */

/* repr.rs:49-71 */
int tk_move_index(int n, tk_move m) {
    const int row = mv_row(m), col = mv_col(m), pat = mv_pat(m);
    int channel;
    if (pat == 0) {
        channel = mv_kind(m); /* flat 0, wall 1, cap 2 */
    } else {
        static const int dir_off[4] = {0 /*Up*/, 2 /*Down*/, 3 /*Left*/, 1 /*Right*/};
        const int pattern_offset = (pat >> (8 - n)) - 1;
        channel = 3 + pattern_offset + ((1 << n) - 2) * dir_off[mv_kind(m)];
    }
    return channel * n * n + row * n + col;
}

/* repr.rs:169-228 */
void tk_game_repr(const tk_game* g, float* out) {
    const int n = g->n, nn = n * n, ss = stack_size(n);
    memset(out, 0, sizeof(float) * (size_t)(tk_input_channels(n) * nn));
    for (int row = 0; row < n; row++) {
        for (int col = 0; col < n; col++) {
            const int sq = row * n + col, h = g->height[sq];
            if (h == 0) continue;
            const int off_top = (top_color(g, sq) != g->to_move) * ss;
            out[nn * (g->top[sq] + off_top) + sq] = 1.0f;
            for (int i = 0; i < ss - 3 && h - 2 - i >= 0; i++) {
                const int color = (int)((g->stack[sq] >> (h - 2 - i)) & 1);
                const int off = (color != g->to_move) * ss;
                out[nn * (3 + off + i) + sq] = 1.0f;
            }
        }
    }
    int s0, c0;
    initial_reserves(n, &s0, &c0);
    const int me = g->to_move, other = me ^ 1;
    /* reserves_ratio: NotNan::new(x / y).unwrap_or_default() -> 0/0 = NaN -> 0.0 */
    const float my_stones = (float)g->stones[me] / (float)s0;
    const float my_caps = c0 ? (float)g->caps[me] / (float)c0 : 0.0f;
    const float op_stones = (float)g->stones[other] / (float)s0;
    const float op_caps = c0 ? (float)g->caps[other] / (float)c0 : 0.0f;
    const int base = 2 * ss * nn;
    const float fcd = (float)tk_flat_diff(g) - (float)g->half_komi / 2.0f;
    const float fcd_per_square = fcd / (float)nn;
    for (int i = 0; i < nn; i++) {
        out[base + i] = my_stones;
        out[base + nn + i] = my_caps;
        out[base + 2 * nn + i] = op_stones;
        out[base + 3 * nn + i] = op_caps;
        if (me == TK_BLACK) out[base + 4 * nn + i] = 1.0f;
        out[base + 5 * nn + i] = fcd_per_square;
    }
}
