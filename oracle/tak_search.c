/*
 * tak_search.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Restates, function by function, the search of the reference:
 *   takzero/src/search/eval.rs            Eval: negate :40-47, f32 :95-105, Ord :138-163
 *   takzero/src/search/node/mod.rs        Node :14-38, descend :95-102, select_best_action
 *                                         :132-163, select_selfplay_action :170-207, ube_target :215-230
 *   takzero/src/search/node/policy.rs     softmax :10-19, improved_policy :29-48,
 *                                         select_with_puct :78-95, sigma_select :121-128, PUCT :140-156
 *   takzero/src/search/node/mcts.rs       update_mean_value :49-53, update_standard_deviation :56-61,
 *                                         node_solver :66-76, propagate_child_eval :78-102, forward :107-138,
 *                                         backward_known_eval :141-163, backward_network_eval :171-225,
 *                                         simulate_simple :235-266, simulate_batch :268-328
 *   takzero/src/search/node/batched.rs    simulate :63-128, step :131-144, restart_terminal_envs :185-203,
 *                                         gumbel_sequential_halving :207-409
 *   takzero/src/search/agent.rs           Dummy :16-42, Simple :44-87
 * Float semantics: every f32 expression is evaluated in the reference's order in
 * IEEE binary32 (compile with -ffp-contract=off, no fast-math); `powi` follows
 * compiler-rt __powisf2; `exp`/`ln` are libm expf/logf like Rust's f32::exp/ln on
 * Linux (tk_set_exact_math(1) switches expf to the restated glibc algorithm that the
 * CUDA path executes; tests assert both modes agree).
 * The `rand` streams of the reference are parity-unpinned, so Gumbel noise,
 * openings and sampling randomness are injected by the caller.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tak_oracle.h"

#define DISCOUNT_FACTOR 0.997f /* search/mod.rs:7 */
#define CONTEMPT (-0.05f)      /* eval.rs:128 */

/* Math modes.  0: the host libm's expf/logf, i.e. what Rust's f32::exp / f32::ln call on
 * Linux.  1: expf through `tk_expf_restated`, the algorithm of glibc's generic expf
 * (sysdeps/ieee754/flt-32/e_expf.c: exp2f_data table, N = 32, degree-3 polynomial in
 * double) written out without FMA contraction -- the same sequence of IEEE operations the
 * CUDA library executes (takzero_b200/csrc/tree.cuh `expf_libm`), so GPU-vs-oracle
 * comparisons are bit-exact by construction.  tests/ assert that mode 1 agrees with mode 0
 * on sampled inputs (a libm built with FMA may differ in ~1e-9 of inputs).  ln is always
 * the host logf (the CUDA library tabulates it with the host logf as well). */
static int g_exact_math = 0;
void tk_set_exact_math(int on) { g_exact_math = on; }

static const uint64_t exp2f_tab[32] = {
    0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL,
    0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL,
    0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL,
    0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL,
    0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL,
    0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL,
    0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL,
    0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL};

float tk_expf_restated(float x) {
    if (x != x) return x;
    if (x > 0x1.62e42ep6f) return INFINITY;
    if (x < -0x1.9fe368p6f) return 0.0f;
    const double N = 32.0;
    const double inv_ln2_n = 0x1.71547652b82fep+0 * N;
    const double c0 = 0x1.c6af84b912394p-5 / N / N / N;
    const double c1 = 0x1.ebfce50fac4f3p-3 / N / N;
    const double c2 = 0x1.62e42ff0c52d6p-1 / N;
    const double shift = 0x1.8p+52;
    const double xd = (double)x;
    double z = inv_ln2_n * xd;
    double kd = z + shift;
    uint64_t ki;
    memcpy(&ki, &kd, 8);
    kd = kd - shift;
    const double r = z - kd;
    uint64_t t = exp2f_tab[ki & 31];
    t += ki << (52 - 5);
    double s;
    memcpy(&s, &t, 8);
    z = c0 * r + c1;
    const double r2 = r * r;
    double y = c2 * r + 1.0;
    y = z * r2 + y;
    y = y * s;
    return (float)y;
}

/* ---- Environment (env.rs:11-25) ------------------------------------------------------
 * The reference's search is generic over `Environment`; Tak (env.rs:33-96) is the instance every caller uses and
 * the default here.  The second instance is the reference's own test environment `SafeCrack` (env.rs:108-209),
 * needed to reproduce `safe_cracker_value_propagation` (mcts.rs:413-445); it keeps its state in the bytes of a
 * tk_game: to_move = !active, ply = tried.len(), stack bytes = tried, top[] = key, caps[0] = key.len(). */
typedef struct {
    int (*terminal)(const tk_game*);
    int (*populate_actions)(const tk_game*, tk_move*);
    void (*step)(tk_game*, tk_move);
} env_ops;

static int sc_terminal(const tk_game* g) {
    (void)g;
    return TK_T_NONE; /* "The game never ends." */
}
static int sc_populate(const tk_game* g, tk_move* out) {
    if (g->to_move == 0) { /* active */
        for (int i = 0; i <= 9; i++) out[i] = (tk_move)i; /* Some(i) */
        return 10;
    }
    out[0] = 0xffff; /* None */
    return 1;
}
static void sc_step(tk_game* g, tk_move m) {
    if (g->to_move == 0) {
        if (m > 9 || g->ply >= 8 * TK_MAX_SQ) abort(); /* "All actions should be Some()" */
        ((uint8_t*)g->stack)[g->ply++] = (uint8_t)m;
    } else if (m != 0xffff) {
        abort();
    }
    g->to_move ^= 1;
}
static const env_ops TAK_ENV = {tk_terminal, tk_possible_moves, tk_play_unchecked};
static const env_ops SAFECRACK_ENV = {sc_terminal, sc_populate, sc_step};
static const env_ops* g_env = &TAK_ENV;
void tk_set_environment(int safecrack) { g_env = safecrack ? &SAFECRACK_ENV : &TAK_ENV; }

void tk_safecrack_new(tk_game* g, const uint8_t* key, int key_len) { /* SafeCrack::new */
    memset(g, 0, sizeof(*g));
    for (int i = 0; i < key_len && i < TK_MAX_SQ; i++) g->top[i] = key[i];
    g->caps[0] = (uint8_t)key_len;
}
static int sc_solved(const tk_game* g) { /* tried.starts_with(&key) */
    if (g->ply < g->caps[0]) return 0;
    return memcmp(g->stack, g->top, g->caps[0]) == 0;
}
/* `SafeCracker` (env.rs:190-209): logit 1.0 for every action, value = +-1 * solved, uncertainty 0 */
void tk_agent_safecracker(void* ctx, int batch, const tk_game* envs, const tk_move* actions, const int* n_actions,
                          int stride, float* logits, float* values, float* variances) {
    (void)ctx;
    (void)actions;
    for (int b = 0; b < batch; b++) {
        for (int i = 0; i < n_actions[b]; i++) logits[(size_t)b * stride + i] = 1.0f;
        values[b] = (envs[b].to_move == 0 ? 1.0f : -1.0f) * (float)sc_solved(&envs[b]);
        variances[b] = 0.0f;
    }
}

static inline float f_exp(float x) { return g_exact_math ? tk_expf_restated(x) : expf(x); }
static inline float f_ln(float x) { return logf(x); }

/* ---- Eval ------------------------------------------------------------ */

static inline tk_eval ev_value(float v) {
    tk_eval e;
    e.tag = TK_E_VALUE;
    e.u.value = v;
    return e;
}
static inline tk_eval ev_known(uint32_t tag, uint32_t ply) {
    tk_eval e;
    e.tag = tag;
    e.u.ply = ply;
    return e;
}
static inline int ev_is_known(tk_eval e) { return e.tag != TK_E_VALUE; }

tk_eval tk_eval_negate(tk_eval e) { /* eval.rs:40-47 */
    switch (e.tag) {
        case TK_E_VALUE: return ev_value(-e.u.value);
        case TK_E_WIN: return ev_known(TK_E_LOSS, e.u.ply + 1);
        case TK_E_DRAW: return ev_known(TK_E_DRAW, e.u.ply + 1);
        default: return ev_known(TK_E_WIN, e.u.ply + 1);
    }
}

static float powi_f32(float a, int b) { /* compiler-rt __powisf2 (what f32::powi lowers to) */
    const int recip = b < 0;
    float r = 1.0f;
    while (1) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0f / r : r;
}

float tk_eval_to_f32(tk_eval e) { /* eval.rs:95-105 */
    const int ply = e.tag == TK_E_VALUE ? 0 : (int)e.u.ply;
    float m;
    switch (e.tag) {
        case TK_E_VALUE: m = e.u.value; break;
        case TK_E_WIN: m = 1.0f; break;
        case TK_E_LOSS: m = -1.0f; break;
        default: m = 0.0f; break;
    }
    return powi_f32(DISCOUNT_FACTOR, ply) * m;
}

static inline float ev_notnan(tk_eval e) { /* eval.rs:107-116 */
    return e.tag == TK_E_VALUE ? e.u.value : tk_eval_to_f32(e);
}

static inline int cmp_f32(float a, float b) { return (a > b) - (a < b); }
static inline int cmp_u32(uint32_t a, uint32_t b) { return (a > b) - (a < b); }

int tk_eval_cmp(tk_eval a, tk_eval b) { /* eval.rs:138-163 */
    switch (a.tag) {
        case TK_E_VALUE:
            switch (b.tag) {
                case TK_E_VALUE: return cmp_f32(a.u.value, b.u.value);
                case TK_E_WIN: return -1;
                case TK_E_DRAW: return cmp_f32(a.u.value, CONTEMPT);
                default: return 1;
            }
        case TK_E_WIN: return b.tag == TK_E_WIN ? cmp_u32(b.u.ply, a.u.ply) : 1;
        case TK_E_DRAW:
            switch (b.tag) {
                case TK_E_VALUE: return cmp_f32(CONTEMPT, b.u.value);
                case TK_E_WIN: return -1;
                case TK_E_DRAW: return cmp_u32(b.u.ply, a.u.ply);
                default: return 1;
            }
        default: return b.tag == TK_E_LOSS ? cmp_u32(a.u.ply, b.u.ply) : -1;
    }
}

/* ---- policy.rs ------------------------------------------------------- */

void tk_softmax(const float* logits, int n, float* out) { /* policy.rs:10-19 */
    float max = 0.0f; /* unwrap_or_default */
    for (int i = 0; i < n; i++)
        if (i == 0 || !(logits[i] < max)) max = logits[i]; /* Iterator::max keeps the last max */
    float sum = 0.0f;
    for (int i = 0; i < n; i++) {
        out[i] = f_exp(logits[i] - max);
        sum += out[i];
    }
    for (int i = 0; i < n; i++) out[i] = out[i] / sum;
}

static inline float exploration_rate(float visit_count) { /* policy.rs:143-145 */
    return f_ln((1.0f + visit_count + 500.0f) / 500.0f) + 4.0f;
}

static inline float puct_term(float parent_visits, float visits, float probability) {
    /* policy.rs:148-156 */
    return exploration_rate(parent_visits) * probability * sqrtf(parent_visits) / (1.0f + visits);
}

static inline float sigma_select(float q, float std_dev, float beta, float visit_count) {
    return (q + std_dev * beta) * (50.0f + visit_count); /* policy.rs:121-128 */
}

static inline float sigma_improve(float q, float std_dev, float beta, float visit_count) {
    return (q + std_dev * beta) * sqrtf(visit_count); /* policy.rs:131-138 */
}

/* ---- Node ------------------------------------------------------------ */

static void node_default(tk_node* n) { memset(n, 0, sizeof(*n)); }

tk_node* tk_node_new(void) { return (tk_node*)calloc(1, sizeof(tk_node)); }

static void node_drop_children(tk_node* n) {
    for (uint32_t i = 0; i < n->n_children; i++) node_drop_children(&n->children[i]);
    free(n->children);
    free(n->actions);
    n->children = NULL;
    n->actions = NULL;
    n->n_children = 0;
}

void tk_node_reset(tk_node* n) {
    node_drop_children(n);
    node_default(n);
}

void tk_node_free(tk_node* n) {
    if (!n) return;
    node_drop_children(n);
    free(n);
}

int tk_node_needs_initialization(const tk_node* n) { /* mod.rs:83-85 */
    return n->n_children == 0 && !ev_is_known(n->evaluation);
}

int tk_node_is_terminal(const tk_node* n) { /* mod.rs:106-108 */
    return ev_is_known(n->evaluation) && n->evaluation.u.ply == 0;
}

const tk_node* tk_node_child(const tk_node* n, int i) { return &n->children[i]; }
tk_move tk_node_action(const tk_node* n, int i) { return n->actions[i]; }

uint64_t tk_node_count(const tk_node* n) {
    uint64_t c = 1;
    for (uint32_t i = 0; i < n->n_children; i++) c += tk_node_count(&n->children[i]);
    return c;
}

static inline float q_value(const tk_node* child) { /* mod.rs:114-124 */
    return ev_notnan(tk_eval_negate(child->evaluation));
}

static int select_with_puct(const tk_node* node, float beta) { /* policy.rs:78-95 */
    const float parent_visit_count = (float)node->visit_count;
    const int parent_is_loss = node->evaluation.tag == TK_E_LOSS;
    int best = -1;
    float best_key = 0.0f;
    for (uint32_t i = 0; i < node->n_children; i++) {
        const tk_node* child = &node->children[i];
        if (!(parent_is_loss || child->evaluation.tag != TK_E_WIN)) continue;
        const float q = q_value(child);
        const float puct = puct_term(parent_visit_count, (float)child->visit_count, child->probability);
        const float key = q + puct + child->std_dev * beta;
        if (best < 0 || !(key < best_key)) { /* max_by_key: last maximum wins */
            best = (int)i;
            best_key = key;
        }
    }
    if (best < 0) {
        fprintf(stderr, "tak_oracle: there should always be a child to simulate\n");
        abort();
    }
    return best;
}

static inline void update_mean_value(tk_node* n, float value) { /* mcts.rs:49-53 */
    if (n->evaluation.tag == TK_E_VALUE) {
        float m = n->evaluation.u.value;
        m += (-m + value) / (float)n->visit_count;
        n->evaluation.u.value = m;
    }
}

static inline void update_standard_deviation(tk_node* n, float variance) { /* mcts.rs:56-61 */
    if (ev_is_known(n->evaluation)) return;
    n->std_dev += (-n->std_dev + sqrtf(variance)) / (float)n->visit_count;
}

static tk_eval min_child_eval(const tk_node* n) { /* Iterator::min keeps the first minimum */
    tk_eval best = n->children[0].evaluation;
    for (uint32_t i = 1; i < n->n_children; i++)
        if (tk_eval_cmp(n->children[i].evaluation, best) < 0) best = n->children[i].evaluation;
    return best;
}

static void node_solver(tk_node* n, tk_eval child_eval) { /* mcts.rs:66-76 */
    int all_known = 1;
    for (uint32_t i = 0; i < n->n_children; i++)
        if (!ev_is_known(n->children[i].evaluation)) {
            all_known = 0;
            break;
        }
    if (child_eval.tag == TK_E_LOSS || all_known) {
        n->evaluation = tk_eval_negate(min_child_eval(n));
        n->std_dev = 0.0f;
    }
}

typedef struct {
    tk_eval eval;
    float variance;
} propagated;

static propagated propagate_child_eval(tk_node* n, tk_eval child_eval, float child_variance) {
    /* mcts.rs:78-102 */
    propagated p;
    node_solver(n, child_eval);
    if (ev_is_known(n->evaluation)) {
        p.eval = n->evaluation;
        p.variance = n->std_dev * n->std_dev;
        return p;
    }
    const float negated = ev_notnan(tk_eval_negate(child_eval));
    update_mean_value(n, negated);
    update_standard_deviation(n, child_variance);
    p.eval = ev_value(negated * DISCOUNT_FACTOR);
    p.variance = child_variance * DISCOUNT_FACTOR * DISCOUNT_FACTOR;
    return p;
}

#define MAX_DEPTH 512

typedef struct {
    int len;
    uint32_t idx[MAX_DEPTH];
} trajectory;

/* returns 1 = Forward::Known(*known), 0 = Forward::NeedsNetwork(*env) ; mcts.rs:107-138 */
static int node_forward(tk_node* root, trajectory* traj, tk_game* env, float beta, tk_eval* known) {
    tk_node* node = root;
    traj->len = 0;
    for (;;) {
        node->visit_count += 1;
        if (tk_node_is_terminal(node)) {
            *known = node->evaluation;
            return 1;
        }
        if (tk_node_needs_initialization(node)) {
            const int t = g_env->terminal(env);
            if (t != TK_T_NONE) {
                node->evaluation =
                    ev_known(t == TK_T_WIN ? TK_E_WIN : t == TK_T_LOSS ? TK_E_LOSS : TK_E_DRAW, 0);
                node->std_dev = 0.0f;
                *known = node->evaluation;
                return 1;
            }
            return 0;
        }
        const int index = select_with_puct(node, beta);
        if (traj->len >= MAX_DEPTH) {
            fprintf(stderr, "tak_oracle: trajectory too deep\n");
            abort();
        }
        traj->idx[traj->len++] = (uint32_t)index;
        g_env->step(env, node->actions[index]);
        node = &node->children[index];
    }
}

static propagated backward_known_eval(tk_node* node, const trajectory* traj, int depth, tk_eval eval) {
    /* mcts.rs:141-163 */
    if (depth < traj->len) {
        propagated c = backward_known_eval(&node->children[traj->idx[depth]], traj, depth + 1, eval);
        return propagate_child_eval(node, c.eval, c.variance);
    }
    propagated p;
    p.eval = eval;
    p.variance = 0.0f;
    return p;
}

static propagated backward_network_eval(tk_node* node, const trajectory* traj, int depth,
                                        const tk_move* actions, const float* logits,
                                        const float* probabilities, int n_actions, float value,
                                        float variance) { /* mcts.rs:171-225 */
    if (depth < traj->len) {
        propagated c = backward_network_eval(&node->children[traj->idx[depth]], traj, depth + 1,
                                             actions, logits, probabilities, n_actions, value, variance);
        return propagate_child_eval(node, c.eval, c.variance);
    }
    update_mean_value(node, value);
    update_standard_deviation(node, variance);
    node->n_children = (uint32_t)n_actions;
    node->children = (tk_node*)calloc((size_t)n_actions, sizeof(tk_node));
    node->actions = (tk_move*)malloc(sizeof(tk_move) * (size_t)n_actions);
    const float parent_value = ev_notnan(node->evaluation);
    for (int i = 0; i < n_actions; i++) {
        tk_node* c = &node->children[i];
        node->actions[i] = actions[i];
        c->evaluation = ev_value(-parent_value); /* mod.rs:66-79 */
        c->logit = logits[i];
        c->probability = probabilities[i];
        c->std_dev = node->std_dev;
    }
    propagated p;
    p.eval = ev_value(value * DISCOUNT_FACTOR);
    p.variance = variance * DISCOUNT_FACTOR * DISCOUNT_FACTOR;
    return p;
}

int tk_node_simulate_simple(tk_node* root, const tk_game* env0, float beta, tk_agent_fn agent,
                            void* ctx) { /* mcts.rs:235-266 */
    trajectory traj;
    tk_game env = *env0;
    tk_eval known;
    propagated p;
    if (node_forward(root, &traj, &env, beta, &known)) {
        p = backward_known_eval(root, &traj, 0, known);
    } else {
        static _Thread_local tk_move actions[TK_MAX_MOVES];
        static _Thread_local float logits[TK_MAX_MOVES], probs[TK_MAX_MOVES];
        const int n = g_env->populate_actions(&env, actions);
        float value, variance;
        agent(ctx, 1, &env, actions, &n, TK_MAX_MOVES, logits, &value, &variance);
        tk_softmax(logits, n, probs);
        p = backward_network_eval(root, &traj, 0, actions, logits, probs, n, value, variance);
    }
    return (int)p.eval.tag;
}

void tk_node_simulate_batch(tk_node* root, const tk_game* env0, float beta, int batch_size,
                            tk_agent_fn agent, void* ctx) { /* mcts.rs:268-328 */
    trajectory* trajs = (trajectory*)malloc(sizeof(trajectory) * (size_t)batch_size);
    tk_game* envs = (tk_game*)malloc(sizeof(tk_game) * (size_t)batch_size);
    tk_move* actions = (tk_move*)malloc(sizeof(tk_move) * TK_MAX_MOVES * (size_t)batch_size);
    int* n_actions = (int*)malloc(sizeof(int) * (size_t)batch_size);
    int filled = 0;
    for (int it = 0; it < batch_size * 4; it++) {
        tk_game env = *env0;
        tk_eval known;
        if (node_forward(root, &trajs[filled], &env, beta, &known)) {
            backward_known_eval(root, &trajs[filled], 0, known);
        } else {
            n_actions[filled] = g_env->populate_actions(&env, actions + (size_t)filled * TK_MAX_MOVES);
            envs[filled] = env;
            filled++;
        }
        if (filled == batch_size) break;
    }
    if (filled > 0) {
        float* logits = (float*)malloc(sizeof(float) * TK_MAX_MOVES * (size_t)filled);
        float* probs = (float*)malloc(sizeof(float) * TK_MAX_MOVES);
        float* values = (float*)malloc(sizeof(float) * (size_t)filled);
        float* variances = (float*)malloc(sizeof(float) * (size_t)filled);
        agent(ctx, filled, envs, actions, n_actions, TK_MAX_MOVES, logits, values, variances);
        for (int i = 0; i < filled; i++) {
            const float* lg = logits + (size_t)i * TK_MAX_MOVES;
            tk_softmax(lg, n_actions[i], probs);
            backward_network_eval(root, &trajs[i], 0, actions + (size_t)i * TK_MAX_MOVES, lg, probs,
                                  n_actions[i], values[i], variances[i]);
        }
        free(logits);
        free(probs);
        free(values);
        free(variances);
    }
    free(trajs);
    free(envs);
    free(actions);
    free(n_actions);
}

void tk_node_descend(tk_node* root, tk_move action) { /* mod.rs:95-102 */
    tk_node me = *root;
    node_default(root);
    int found = -1;
    for (uint32_t i = 0; i < me.n_children; i++)
        if (me.actions[i] == action) {
            found = (int)i;
            break;
        }
    if (found >= 0) {
        *root = me.children[found];
        node_default(&me.children[found]);
    }
    node_drop_children(&me);
}

tk_move tk_node_select_best_action(const tk_node* node) { /* mod.rs:132-163 */
    if (node->n_children == 0) {
        fprintf(stderr, "tak_oracle: There should be at least one child\n");
        abort();
    }
    if (ev_is_known(node->evaluation)) {
        int best = 0; /* min_by_key: first minimum */
        for (uint32_t i = 1; i < node->n_children; i++)
            if (tk_eval_cmp(node->children[i].evaluation, node->children[best].evaluation) < 0) best = (int)i;
        return node->actions[best];
    }
    int most = 0; /* max_by_key: last maximum */
    for (uint32_t i = 1; i < node->n_children; i++)
        if (node->children[i].visit_count >= node->children[most].visit_count) most = (int)i;
    if (node->children[most].visit_count == 0) {
        int bp = 0;
        for (uint32_t i = 1; i < node->n_children; i++)
            if (!(node->children[i].probability < node->children[bp].probability)) bp = (int)i;
        return node->actions[bp];
    }
    return node->actions[most];
}

/* mod.rs:170-207.  `random` replaces rand's choose_weighted draw: the sampled
 * index is the first child whose cumulative weight exceeds random % total. */
tk_move tk_node_select_selfplay_action(const tk_node* node, int use_threshold, uint32_t threshold,
                                       float allowed_eval_drop, uint64_t random) {
    if (ev_is_known(node->evaluation) || !use_threshold) return tk_node_select_best_action(node);
    tk_eval best_eval = min_child_eval(node);
    tk_eval limit = best_eval;
    if (limit.tag == TK_E_VALUE) limit.u.value = limit.u.value + allowed_eval_drop;
    uint64_t total = 0;
    for (uint32_t i = 0; i < node->n_children; i++) {
        const tk_node* c = &node->children[i];
        if (c->visit_count < threshold || c->evaluation.tag == TK_E_WIN ||
            tk_eval_cmp(c->evaluation, limit) > 0)
            continue;
        total += c->visit_count;
    }
    if (total == 0) return tk_node_select_best_action(node);
    uint64_t x = random % total;
    for (uint32_t i = 0; i < node->n_children; i++) {
        const tk_node* c = &node->children[i];
        if (c->visit_count < threshold || c->evaluation.tag == TK_E_WIN ||
            tk_eval_cmp(c->evaluation, limit) > 0)
            continue;
        if (x < c->visit_count) return node->actions[i];
        x -= c->visit_count;
    }
    return tk_node_select_best_action(node); /* unreachable */
}

float tk_node_ube_target(const tk_node* node, float beta) { /* mod.rs:215-230 */
    if (ev_is_known(node->evaluation) || tk_node_needs_initialization(node)) return 0.0f;
    int best = -1;
    float best_key = 0.0f;
    for (uint32_t i = 0; i < node->n_children; i++) {
        const tk_node* c = &node->children[i];
        const float key = ev_notnan(tk_eval_negate(c->evaluation)) + c->std_dev * beta;
        if (best < 0 || !(key < best_key)) {
            best = (int)i;
            best_key = key;
        }
    }
    const float s = node->children[best].std_dev;
    return s * s;
}

float tk_node_most_visited_count(const tk_node* node) { /* policy.rs:23-30 */
    uint32_t m = 0;
    for (uint32_t i = 0; i < node->n_children; i++)
        if (node->children[i].visit_count > m) m = node->children[i].visit_count;
    return (float)m;
}

void tk_node_improved_policy(const tk_node* node, float visitations, float* out) {
    /* policy.rs:36-48 */
    const int n = (int)node->n_children;
    float* p = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        const tk_node* c = &node->children[i];
        const float completed = tk_node_needs_initialization(c)
                                    ? ev_notnan(node->evaluation)
                                    : ev_notnan(tk_eval_negate(c->evaluation));
        p[i] = sigma_improve(completed, c->std_dev, 0.0f, visitations) + c->logit;
    }
    tk_softmax(p, n, out);
    free(p);
}

int tk_node_principal_variation(const tk_node* node, tk_move* out, int max) { /* mod.rs:44-62 */
    int len = 0;
    while (len < max && !tk_node_needs_initialization(node) && !tk_node_is_terminal(node)) {
        if (node->n_children == 0) break;
        const tk_move best = tk_node_select_best_action(node);
        const tk_node* next = NULL;
        for (uint32_t i = 0; i < node->n_children; i++)
            if (node->actions[i] == best) {
                next = &node->children[i];
                break;
            }
        out[len++] = best;
        node = next;
    }
    return len;
}

/* ---- agents (agent.rs) ------------------------------------------------ */

void tk_agent_dummy(void* ctx, int batch, const tk_game* envs, const tk_move* actions,
                    const int* n_actions, int stride, float* logits, float* values, float* variances) {
    (void)ctx;
    (void)envs;
    (void)actions;
    for (int b = 0; b < batch; b++) {
        for (int i = 0; i < n_actions[b]; i++) logits[(size_t)b * stride + i] = 1.0f;
        values[b] = 0.0f;
        variances[b] = 0.0f;
    }
}

void tk_agent_simple(void* ctx, int batch, const tk_game* envs, const tk_move* actions,
                     const int* n_actions, int stride, float* logits, float* values, float* variances) {
    (void)ctx;
    for (int b = 0; b < batch; b++) {
        const tk_game* env = &envs[b];
        /* agent.rs:65-68: integer division HALF_KOMI / 2 on i8 */
        float fcd = (float)(tk_flat_diff(env) - env->half_komi / 2) / (float)(env->n * env->n);
        if (env->to_move == TK_BLACK) fcd = -fcd;
        for (int i = 0; i < n_actions[b]; i++) {
            const tk_move m = actions[(size_t)b * stride + i];
            float p;
            if ((m >> 8) != 0) p = 1.0f;
            else if (((m >> 6) & 3) == TK_FLAT) p = 4.0f;
            else if (((m >> 6) & 3) == TK_CAP) p = 3.0f;
            else p = 2.0f;
            logits[(size_t)b * stride + i] = p;
        }
        values[b] = fcd;
        variances[b] = 0.0f;
    }
}

/* Synthetic agent: all outputs are exact binary fractions of integer hashes, so
 * the CUDA library reproduces them bit for bit (csrc/agent_synth.cuh). */
static inline uint64_t mix64s(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

void tk_agent_synthetic(void* ctx, int batch, const tk_game* envs, const tk_move* actions,
                        const int* n_actions, int stride, float* logits, float* values,
                        float* variances) {
    (void)ctx;
    for (int b = 0; b < batch; b++) {
        const uint64_t h = tk_state_hash(&envs[b]);
        for (int i = 0; i < n_actions[b]; i++) {
            const uint64_t hm = mix64s(h ^ ((uint64_t)actions[(size_t)b * stride + i] * 0x9e3779b97f4a7c15ULL));
            /* logits in [-4, 4) with 1/4096 resolution */
            logits[(size_t)b * stride + i] = (float)((int)(hm & 0x7fff) - 16384) * (1.0f / 4096.0f);
        }
        /* value in (-0.75, 0.75), variance in [0, 1) */
        values[b] = (float)((int)((h >> 20) & 0xffff) - 32768) * (0.75f / 32768.0f);
        variances[b] = (float)((h >> 40) & 0xffff) * (1.0f / 65536.0f);
    }
}

/* ---- BatchedMCTS (batched.rs) ------------------------------------------ */

struct tk_batched {
    int batch;
    tk_node* nodes;
    tk_game* envs;
    trajectory* trajectories;
    tk_game* start_envs; /* Replay.env */
    tk_move** replay_actions;
    int* replay_len;
    int* replay_cap;
    tk_counters counters;
    /* scratch for the agent call */
    tk_game* env_batch;
    tk_move* actions_batch;
    int* n_actions;
    int* game_of;
    float* logits;
    float* values;
    float* variances;
};

tk_batched* tk_batched_from_envs(const tk_game* envs, int batch) { /* batched.rs:39-47 */
    tk_batched* b = (tk_batched*)calloc(1, sizeof(tk_batched));
    b->batch = batch;
    b->nodes = (tk_node*)calloc((size_t)batch, sizeof(tk_node));
    b->envs = (tk_game*)malloc(sizeof(tk_game) * (size_t)batch);
    b->start_envs = (tk_game*)malloc(sizeof(tk_game) * (size_t)batch);
    memcpy(b->envs, envs, sizeof(tk_game) * (size_t)batch);
    memcpy(b->start_envs, envs, sizeof(tk_game) * (size_t)batch);
    b->trajectories = (trajectory*)calloc((size_t)batch, sizeof(trajectory));
    b->replay_actions = (tk_move**)calloc((size_t)batch, sizeof(tk_move*));
    b->replay_len = (int*)calloc((size_t)batch, sizeof(int));
    b->replay_cap = (int*)calloc((size_t)batch, sizeof(int));
    b->env_batch = (tk_game*)malloc(sizeof(tk_game) * (size_t)batch);
    b->actions_batch = (tk_move*)malloc(sizeof(tk_move) * TK_MAX_MOVES * (size_t)batch);
    b->n_actions = (int*)malloc(sizeof(int) * (size_t)batch);
    b->game_of = (int*)malloc(sizeof(int) * (size_t)batch);
    b->logits = (float*)malloc(sizeof(float) * TK_MAX_MOVES * (size_t)batch);
    b->values = (float*)malloc(sizeof(float) * (size_t)batch);
    b->variances = (float*)malloc(sizeof(float) * (size_t)batch);
    return b;
}

void tk_batched_free(tk_batched* b) {
    if (!b) return;
    for (int i = 0; i < b->batch; i++) {
        node_drop_children(&b->nodes[i]);
        free(b->replay_actions[i]);
    }
    free(b->nodes);
    free(b->envs);
    free(b->start_envs);
    free(b->trajectories);
    free(b->replay_actions);
    free(b->replay_len);
    free(b->replay_cap);
    free(b->env_batch);
    free(b->actions_batch);
    free(b->n_actions);
    free(b->game_of);
    free(b->logits);
    free(b->values);
    free(b->variances);
    free(b);
}

int tk_batched_size(const tk_batched* b) { return b->batch; }
tk_node* tk_batched_node(tk_batched* b, int i) { return &b->nodes[i]; }
tk_game* tk_batched_env(tk_batched* b, int i) { return &b->envs[i]; }
void tk_batched_counters(const tk_batched* b, tk_counters* out) { *out = b->counters; }
int tk_batched_replay_len(const tk_batched* b, int i) { return b->replay_len[i]; }
const tk_move* tk_batched_replay_actions(const tk_batched* b, int i) { return b->replay_actions[i]; }

/* One lock-step simulation over `sim_nodes[i]` / `sim_envs[i]`
 * (batched.rs:63-128 and the inlined copy at :266-333). */
static void lockstep_simulation(tk_batched* b, tk_node** sim_nodes, const tk_game* sim_envs,
                                const float* betas, int use_betas, tk_agent_fn agent, void* ctx) {
    int filled = 0;
    static _Thread_local float probs[TK_MAX_MOVES];
    for (int i = 0; i < b->batch; i++) {
        tk_game env = sim_envs[i];
        tk_eval known;
        b->counters.simulations++;
        if (node_forward(sim_nodes[i], &b->trajectories[i], &env, use_betas ? betas[i] : 0.0f, &known)) {
            backward_known_eval(sim_nodes[i], &b->trajectories[i], 0, known);
            b->counters.known++;
        } else {
            b->n_actions[filled] = g_env->populate_actions(&env, b->actions_batch + (size_t)filled * TK_MAX_MOVES);
            b->env_batch[filled] = env;
            b->game_of[filled] = i;
            filled++;
        }
    }
    if (filled == 0) return;
    b->counters.evaluations += (uint64_t)filled;
    agent(ctx, filled, b->env_batch, b->actions_batch, b->n_actions, TK_MAX_MOVES, b->logits,
          b->values, b->variances);
    for (int j = 0; j < filled; j++) {
        const int i = b->game_of[j];
        const float* lg = b->logits + (size_t)j * TK_MAX_MOVES;
        tk_softmax(lg, b->n_actions[j], probs);
        backward_network_eval(sim_nodes[i], &b->trajectories[i], 0,
                              b->actions_batch + (size_t)j * TK_MAX_MOVES, lg, probs, b->n_actions[j],
                              b->values[j], b->variances[j]);
    }
}

void tk_batched_simulate(tk_batched* b, tk_agent_fn agent, void* ctx, const float* betas) {
    tk_node** nodes = (tk_node**)malloc(sizeof(tk_node*) * (size_t)b->batch);
    for (int i = 0; i < b->batch; i++) nodes[i] = &b->nodes[i];
    lockstep_simulation(b, nodes, b->envs, betas, 1, agent, ctx);
    free(nodes);
}

void tk_batched_step(tk_batched* b, const tk_move* actions) { /* batched.rs:131-144 */
    for (int i = 0; i < b->batch; i++) {
        if (tk_node_is_terminal(&b->nodes[i])) continue;
        tk_node_descend(&b->nodes[i], actions[i]);
        if (b->replay_len[i] == b->replay_cap[i]) {
            b->replay_cap[i] = b->replay_cap[i] ? b->replay_cap[i] * 2 : 64;
            b->replay_actions[i] =
                (tk_move*)realloc(b->replay_actions[i], sizeof(tk_move) * (size_t)b->replay_cap[i]);
        }
        b->replay_actions[i][b->replay_len[i]++] = actions[i];
        g_env->step(&b->envs[i], actions[i]);
    }
}

void tk_batched_restart_terminal_envs(tk_batched* b, const int* opening_sym, const int* opening_adj,
                                      int* out_terminal) { /* batched.rs:185-203 */
    for (int i = 0; i < b->batch; i++) {
        const int t = g_env->terminal(&b->envs[i]);
        out_terminal[i] = t;
        if (t == TK_T_NONE) continue;
        const uint16_t limit = b->envs[i].reversible_limit;
        tk_new_opening(&b->envs[i], b->envs[i].n, b->envs[i].half_komi, opening_sym[i], opening_adj[i]);
        b->envs[i].reversible_limit = limit;
        tk_node_reset(&b->nodes[i]);
        b->start_envs[i] = b->envs[i];
        b->replay_len[i] = 0;
    }
}

void tk_batched_select_best_actions(tk_batched* b, tk_move* out) {
    for (int i = 0; i < b->batch; i++) out[i] = tk_node_select_best_action(&b->nodes[i]);
}

typedef struct {
    float key; /* logit + gumbel */
    int child;
} set_entry;

/* stable insertion sort, descending by `skey` (slice::sort_by_key(Reverse(..)) is stable) */
static void stable_sort_desc(set_entry* e, float* skey, int n) {
    for (int i = 1; i < n; i++) {
        set_entry cur = e[i];
        float ck = skey[i];
        int j = i - 1;
        while (j >= 0 && skey[j] < ck) {
            e[j + 1] = e[j];
            skey[j + 1] = skey[j];
            j--;
        }
        e[j + 1] = cur;
        skey[j + 1] = ck;
    }
}

static uint32_t ilog2_u32(uint32_t x) {
    uint32_t r = 0;
    while (x >>= 1) r++;
    return r;
}

void tk_batched_gumbel_sequential_halving(tk_batched* b, tk_agent_fn agent, void* ctx,
                                          const float* betas, int sampled_actions,
                                          uint32_t search_budget, const float* gumbel,
                                          int gumbel_stride, tk_move* out_moves) {
    /* batched.rs:207-409 */
    if (sampled_actions <= 0) {
        fprintf(stderr, "tak_oracle: At least one action must be sampled\n");
        abort();
    }
    const uint32_t steps = ilog2_u32((uint32_t)sampled_actions);
    if (steps == 0 || search_budget % (steps * (uint32_t)sampled_actions) != 0) {
        fprintf(stderr, "tak_oracle: The search budget should be a multiple of k*log2(k) for clean visits\n");
        abort();
    }
    const int B = b->batch;
    tk_batched_simulate(b, agent, ctx, betas); /* :223 */

    set_entry** sets = (set_entry**)malloc(sizeof(set_entry*) * (size_t)B);
    int* set_len = (int*)malloc(sizeof(int) * (size_t)B);
    float* skey = (float*)malloc(sizeof(float) * TK_MAX_MOVES);
    for (int g = 0; g < B; g++) { /* :230-244 */
        tk_node* node = &b->nodes[g];
        const int nc = (int)node->n_children;
        sets[g] = (set_entry*)malloc(sizeof(set_entry) * (size_t)(nc > 0 ? nc : 1));
        for (int i = 0; i < nc; i++) {
            sets[g][i].key = node->children[i].logit + gumbel[(size_t)g * gumbel_stride + i];
            sets[g][i].child = i;
            skey[i] = sets[g][i].key;
        }
        stable_sort_desc(sets[g], skey, nc);
        set_len[g] = nc < sampled_actions ? nc : sampled_actions;
    }

    const uint32_t visits_per_step = search_budget / steps;
    uint32_t visits_to_most_visited_action = 0;
    int remaining = sampled_actions;
    tk_node** sim_nodes = (tk_node**)malloc(sizeof(tk_node*) * (size_t)B);
    tk_game* sim_envs = (tk_game*)malloc(sizeof(tk_game) * (size_t)B);

    for (uint32_t step = 0; step < steps; step++) {
        const uint32_t visits_per_action = visits_per_step / (uint32_t)remaining;
        for (int i = 0; i < remaining; i++) {
            for (int g = 0; g < B; g++) { /* :255-264 */
                if (set_len[g] == 0) {
                    fprintf(stderr, "tak_oracle: root without children in sequential halving\n");
                    abort();
                }
                const int child = sets[g][i % set_len[g]].child;
                sim_envs[g] = b->envs[g];
                g_env->step(&sim_envs[g], b->nodes[g].actions[child]);
                sim_nodes[g] = &b->nodes[g].children[child];
            }
            for (uint32_t v = 0; v < visits_per_action; v++)
                lockstep_simulation(b, sim_nodes, sim_envs, betas, 0, agent, ctx);
        }
        visits_to_most_visited_action += visits_per_action;
        remaining /= 2;
        for (int g = 0; g < B; g++) { /* :342-355 */
            for (int j = 0; j < set_len[g]; j++) {
                const tk_node* child = &b->nodes[g].children[sets[g][j].child];
                skey[j] = sets[g][j].key + sigma_select(ev_notnan(tk_eval_negate(child->evaluation)),
                                                        child->std_dev, betas[g],
                                                        (float)visits_to_most_visited_action);
            }
            stable_sort_desc(sets[g], skey, set_len[g]);
            if (set_len[g] > remaining) set_len[g] = remaining;
        }
    }

    for (int g = 0; g < B; g++) { /* :358-370 */
        if (set_len[g] != 1) {
            fprintf(stderr, "tak_oracle: After sequential halving, every set should have exactly 1 action left\n");
            abort();
        }
        out_moves[g] = b->nodes[g].actions[sets[g][0].child];
    }

    for (int g = 0; g < B; g++) { /* :373-406 */
        tk_node* node = &b->nodes[g];
        uint32_t sum = 0;
        int any_loss = 0, all_known = 1;
        for (uint32_t i = 0; i < node->n_children; i++) {
            sum += node->children[i].visit_count;
            if (node->children[i].evaluation.tag == TK_E_LOSS) any_loss = 1;
            if (!ev_is_known(node->children[i].evaluation)) all_known = 0;
        }
        node->visit_count = sum + 1;
        if (any_loss || all_known) {
            node->evaluation = tk_eval_negate(min_child_eval(node));
            node->std_dev = 0.0f;
        } else {
            float sum_p = 0.0f, weighted_q = 0.0f;
            for (uint32_t i = 0; i < node->n_children; i++) {
                const tk_node* c = &node->children[i];
                if (c->visit_count == 0) continue;
                sum_p += c->probability;
            }
            for (uint32_t i = 0; i < node->n_children; i++) {
                const tk_node* c = &node->children[i];
                if (c->visit_count == 0) continue;
                weighted_q += c->probability * tk_eval_to_f32(tk_eval_negate(c->evaluation));
            }
            node->evaluation = ev_value(weighted_q / sum_p);
        }
    }

    for (int g = 0; g < B; g++) free(sets[g]);
    free(sets);
    free(set_len);
    free(skey);
    free(sim_nodes);
    free(sim_envs);
}
