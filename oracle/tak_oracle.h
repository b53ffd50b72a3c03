/*
 * tak_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the batched self-play search of ViliamVadocz/takzero
 * (paths below are relative to the reference checkout):
 *   takzero/src/search/{env,agent,eval,mod}.rs
 *   takzero/src/search/node/{mod,mcts,policy,batched}.rs
 *   takzero/src/network/repr.rs
 * plus the Tak rules of the un-vendored crate `fast-tak 0.4.1` / `takparse 0.6.0`
 * (Cargo.lock:611-616,1564-1569), restated from the published rules of Tak and
 * pinned by the reference's own golden vectors (repr.rs:260-499, mcts.rs:345-411,
 * runs/{*}.txt).  PARITY STATUS: rules/encoding/search are pinned by those vectors;
 * the reversible-ply draw threshold, the symmetry index order and the `rand`
 * streams are "parity unpinned" (nothing in the reference repo fixes them).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (takzero_b200/) never does.
 */
#ifndef TAK_ORACLE_H
#define TAK_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TK_MAX_N 6
#define TK_MAX_SQ 36
#define TK_MAX_MOVES 1024

/* Move encoding (u16), shared by documentation with include/takzero_b200.h:
 *   bits 0..2  column (file a = 0)
 *   bits 3..5  row    (rank 1 = 0)
 *   bits 6..7  placement: piece (0 flat, 1 wall, 2 cap)
 *              spread:    direction (0 Up '+', 1 Down '-', 2 Left '<', 3 Right '>')
 *   bits 8..15 spread pattern = takparse `Pattern::mask()` byte (MSB aligned,
 *              repr.rs:437-484); 0 for placements. */
typedef uint16_t tk_move;

enum { TK_FLAT = 0, TK_WALL = 1, TK_CAP = 2 };
enum { TK_UP = 0, TK_DOWN = 1, TK_LEFT = 2, TK_RIGHT = 3 };
enum { TK_WHITE = 0, TK_BLACK = 1 };

typedef struct tk_game {
    uint64_t stack[TK_MAX_SQ];  /* bit i = colour of the piece at height i (0 = bottom) */
    uint8_t height[TK_MAX_SQ];
    uint8_t top[TK_MAX_SQ];     /* piece type of the top piece, valid when height > 0 */
    uint8_t n;
    int8_t half_komi;
    uint8_t to_move;
    uint8_t stones[2];
    uint8_t caps[2];
    uint16_t ply;
    uint16_t reversible_plies;
    uint16_t reversible_limit;  /* parity unpinned (fast-tak source absent); default 100 */
} tk_game;

/* game result codes */
enum { TK_ONGOING = 0, TK_WHITE_WIN = 1, TK_BLACK_WIN = 2, TK_DRAW = 3 };
/* terminal relative to the side to move (env.rs:47-59) */
enum { TK_T_NONE = 0, TK_T_WIN = 1, TK_T_LOSS = 2, TK_T_DRAW = 3 };

void tk_game_init(tk_game* g, int n, int half_komi);
int tk_game_from_tps(tk_game* g, int n, int half_komi, const char* tps);
int tk_game_to_tps(const tk_game* g, char* buf, int buflen);
int tk_possible_moves(const tk_game* g, tk_move* out);
int tk_play(tk_game* g, tk_move m);
void tk_play_unchecked(tk_game* g, tk_move m);
int tk_result(const tk_game* g);
int tk_has_road(const tk_game* g, int color);
int tk_terminal(const tk_game* g);
int tk_flat_diff(const tk_game* g);
uint64_t tk_state_hash(const tk_game* g);
void tk_new_opening(tk_game* g, int n, int half_komi, int symmetry, int adjacent);

int tk_move_to_str(tk_move m, char* buf);
int tk_move_from_str(const char* s, tk_move* out);
int tk_move_order_key(tk_move m, int n); /* strictly increasing along possible_moves */

/* network/repr.rs */
int tk_input_channels(int n);
int tk_output_channels(int n);
int tk_move_index(int n, tk_move m);
void tk_game_repr(const tk_game* g, float* out);

/* search/eval.rs */
enum { TK_E_VALUE = 0, TK_E_WIN = 1, TK_E_LOSS = 2, TK_E_DRAW = 3 };
typedef struct tk_eval {
    uint32_t tag;
    union {
        float value;
        uint32_t ply;
    } u;
} tk_eval;

tk_eval tk_eval_negate(tk_eval e);
int tk_eval_cmp(tk_eval a, tk_eval b);
float tk_eval_to_f32(tk_eval e);
void tk_softmax(const float* logits, int n, float* out);
void tk_set_exact_math(int on); /* 1: expf = tk_expf_restated (what the CUDA library executes) */
float tk_expf_restated(float x);

/* search/node/mod.rs */
typedef struct tk_node {
    tk_eval evaluation;
    uint32_t visit_count;
    float logit;
    float probability;
    float std_dev;
    uint32_t n_children;
    tk_move* actions;
    struct tk_node* children;
} tk_node;

/* Agent (search/agent.rs:5-14).  Batched; logits are un-normalised and in the
 * order of the action lists. */
typedef void (*tk_agent_fn)(void* ctx, int batch, const tk_game* envs, const tk_move* actions,
                            const int* n_actions, int stride, float* logits, float* values,
                            float* variances);
void tk_agent_dummy(void*, int, const tk_game*, const tk_move*, const int*, int, float*, float*,
                    float*);
void tk_agent_simple(void*, int, const tk_game*, const tk_move*, const int*, int, float*, float*,
                     float*);
/* Deterministic integer-hash agent shared bit-for-bit with the CUDA library's
 * TZ_AGENT_SYNTHETIC (include/takzero_b200.h). */
void tk_agent_synthetic(void*, int, const tk_game*, const tk_move*, const int*, int, float*,
                        float*, float*);

/* The search is generic over `Environment` (env.rs:11-25).  0 (default): Tak.  1: the reference's test
 * environment SafeCrack (env.rs:108-209) with its agent SafeCracker, for mcts.rs:413-445. */
void tk_set_environment(int safecrack);
void tk_safecrack_new(tk_game* g, const uint8_t* key, int key_len);
void tk_agent_safecracker(void*, int, const tk_game*, const tk_move*, const int*, int, float*, float*, float*);

tk_node* tk_node_new(void);
void tk_node_free(tk_node* node);
void tk_node_reset(tk_node* node);
/* returns propagated eval tag (TK_E_*) */
int tk_node_simulate_simple(tk_node* root, const tk_game* env, float beta, tk_agent_fn agent,
                            void* ctx);
void tk_node_simulate_batch(tk_node* root, const tk_game* env, float beta, int batch_size,
                            tk_agent_fn agent, void* ctx);
void tk_node_descend(tk_node* root, tk_move action);
tk_move tk_node_select_best_action(const tk_node* node);
tk_move tk_node_select_selfplay_action(const tk_node* node, int use_threshold, uint32_t threshold,
                                       float allowed_eval_drop, uint64_t random);
float tk_node_ube_target(const tk_node* node, float beta);
float tk_node_most_visited_count(const tk_node* node);
void tk_node_improved_policy(const tk_node* node, float visitations, float* out);
int tk_node_principal_variation(const tk_node* node, tk_move* out, int max);
int tk_node_is_terminal(const tk_node* node);
int tk_node_needs_initialization(const tk_node* node);
const tk_node* tk_node_child(const tk_node* node, int i);
tk_move tk_node_action(const tk_node* node, int i);
uint64_t tk_node_count(const tk_node* node);

/* search/node/batched.rs */
typedef struct tk_batched tk_batched;
typedef struct tk_counters {
    uint64_t simulations; /* calls of Node::forward */
    uint64_t evaluations; /* positions sent to the agent */
    uint64_t known;       /* forwards that ended in Forward::Known */
} tk_counters;

tk_batched* tk_batched_from_envs(const tk_game* envs, int batch);
void tk_batched_free(tk_batched* b);
int tk_batched_size(const tk_batched* b);
tk_node* tk_batched_node(tk_batched* b, int i);
tk_game* tk_batched_env(tk_batched* b, int i);
void tk_batched_counters(const tk_batched* b, tk_counters* out);
void tk_batched_simulate(tk_batched* b, tk_agent_fn agent, void* ctx, const float* betas);
/* gumbel: [batch][gumbel_stride] injected Gumbel(0,1) draws, one per root child
 * in child order (batched.rs:226-239). */
void tk_batched_gumbel_sequential_halving(tk_batched* b, tk_agent_fn agent, void* ctx,
                                          const float* betas, int sampled_actions,
                                          uint32_t search_budget, const float* gumbel,
                                          int gumbel_stride, tk_move* out_moves);
void tk_batched_step(tk_batched* b, const tk_move* actions);
/* out_terminal[i] = TK_T_* of env i before the restart; openings: per game
 * {symmetry, adjacent} pairs consumed only for restarted games. */
void tk_batched_restart_terminal_envs(tk_batched* b, const int* opening_sym,
                                      const int* opening_adj, int* out_terminal);
int tk_batched_replay_len(const tk_batched* b, int i);
const tk_move* tk_batched_replay_actions(const tk_batched* b, int i);
void tk_batched_select_best_actions(tk_batched* b, tk_move* out);

/* tak_batch.c: loops of the functions above for the bulk parity tests */
void tk_games_pack(const tk_game* games, int count, uint8_t* out384);
void tk_games_unpack(const uint8_t* in384, int count, int n, int half_komi, int reversible_limit, tk_game* games);
void tk_game_repr_batch(const tk_game* games, int count, float* out);
unsigned long long tk_perft(const tk_game* g, int depth);
void tk_move_index_batch(int n, const tk_move* actions, const int* n_actions, int stride, int count, int32_t* out);
int tk_game_result5(const tk_game* g);
long long tk_expf_compare(uint32_t lo_bits, uint32_t hi_bits, uint32_t step, long long* tested, float* first_bad);
void tk_expf_restated_batch(const float* in, int count, float* out);
int tk_playout_positions(int n, int half_komi, int reversible_limit, uint64_t seed, int policy, int max_positions,
                         int max_plies, tk_game* out_games, int32_t* out_n_moves, tk_move* out_moves, int stride,
                         int32_t* out_terminal, int32_t* out_result, tk_move* out_chosen, tk_game* out_next);

#ifdef __cplusplus
}
#endif
#endif
