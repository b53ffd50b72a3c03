// nn.cu -- the policy / value / uncertainty ResNet of the reference as a device agent.
//
// Replaces `impl Agent<Env> for Net` (takzero/src/network/net6_simhash.rs:259-324 and its
// N = 4 / N = 5 siblings): board -> input planes (network/repr.rs:169-228) -> conv tower
// (net6_simhash.rs:43-72, residual.rs) -> policy conv + legal-logit gather (:74-86,277-306),
// value / UBE heads (:88-119) and the uncertainty combine (:309-317), without libtorch.
// The 3x3 convolutions run on tcgen05 (conv_tcgen05.cuh); everything else here is small.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <type_traits>
#include <vector>

#include "conv_tcgen05.cuh"
#include "nn.cuh"
#include "rules.cuh"

#define WPB TZ_WARPS_PER_BLOCK
#define FILTERS 256
#define CIN_PAD 64

struct ConvLayer {
    __nv_bfloat16* w = nullptr;  // pre-arranged blocks
    float* bias = nullptr;
    int cin = 0;
};

struct NnState {
    int n = 0, blocks = 0;
    int in_channels = 0, out_channels = 0;
    int max_positions = 0;
    size_t rows = 0;      // rows per chunk plane of the buffers that hold all positions (planes)
    size_t rows_set = 0;  // rows per chunk plane of one activation set (one chunk of positions)
    int chunk_min_tiles = 0;   // least pair tiles the network runs through all layers at a time (conv_tcgen05.cuh)
    int chunk_positions = 0;   // most positions of one chunk = what one activation set holds
    int chunk_tiles = 0, max_chunks = 0;  // bounds: pair tiles of one chunk, chunks of one launch
    ConvLayer input, policy;
    std::vector<ConvLayer> tower;  // 2 per residual block
    float* head_w = nullptr;       // [2][256] conv1x1 weights (value, ube)
    float* head_misc = nullptr;    // [2] conv bias, [2][36] linear weights, [2] linear bias
    // chunk-planar activations (conv_tcgen05.cuh): [channels / 8][rows][8]
    __nv_bfloat16* planes = nullptr;  // [8][rows][8]   input planes, 64 channels (C real ones)
    __nv_bfloat16* act_x = nullptr;   // [2 sets][32][rows_set][8]  residual stream (even / odd chunks)
    __nv_bfloat16* act_t = nullptr;   // [2 sets][32][rows_set][8]  middle of a residual block
    __nv_bfloat16* tune_buf[3] = {nullptr, nullptr, nullptr};  // tuning hook only: [32][rows][8] each
    float* head_feat = nullptr;       // [max_positions * n*n][2] value / UBE 1x1 convolution outputs per row
    float* logits_full = nullptr;     // [64][max_positions * n*n][4] policy logits, 4-channel planes
    uint4* masks = nullptr;           // [n*n][9] disable-output-lane masks (conv_tcgen05.cuh)
    int novelty = 0;                  // which hash indexes the set: 0 none yet, 1 SimHash, 2 LCG hash (last one set)
    float* lcghash_init = nullptr;    // [C][N][N] (net4_lcghash.rs:131-137), optional
    float* simhash_matrix = nullptr;  // [C*N*N][32] (net6_simhash.rs:136-139), optional
    uint32_t* simhash_set = nullptr;  // 2^32-bit set (bitvec.bin), optional; absent = empty set
    uint32_t* simhash_idx = nullptr;  // [max_positions] hash index of each queued position
    size_t progress_len = 0;
    unsigned* progress = nullptr;     // [chunks][pair tiles] tile-completion counters of the fused launch, then
                                      // [chunks] chunk-completion counters
    int max_pairs = 74;               // CTA pairs that can be resident at once (cooperative launch bound)
    int fused = 1;                    // 1: the tower is one multi-layer launch; 0 (TZ_TOWER=layers): one launch per layer
    int layer_limit = -1;             // debug: stop the tower after this many convolutions
    int f16 = 0;                      // 16-bit type of weights / activations: 0 bf16 (default), 1 fp16
    std::vector<void*> allocs;
    int sm_count = 148;
};

bool nn_ready(const tz_handle* h) { return h->nn != nullptr; }

void nn_free(tz_handle* h) {
    if (!h->nn) return;
    for (void* p : h->nn->allocs) cudaFree(p);
    delete h->nn;
    h->nn = nullptr;
}

// ---- input planes (network/repr.rs:169-228) ----------------------------------------------------

// One warp per position.  out_f32: [count][C][N][N] exactly like `game_repr` (parity hook);
// out_bf16: chunk-planar [8][rows][8] (row = guard + position * N*N + square) feeding the first convolution.
__global__ void __launch_bounds__(32 * WPB) k_encode(const TzState* states, const int* count_ptr, int count_max, int n,
                                                      int half_komi, float* out_f32, __nv_bfloat16* out_bf16,
                                                      int guard, long long rows, int f16) {
    __shared__ TzState s_state[WPB];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * WPB + warp;
    const int count = count_ptr ? *count_ptr : count_max;
    if (q >= count) return;
    TzState* st = &s_state[warp];
    warp_load_state(st, &states[q], lane);
    const int nn = n * n, ss = 2 * n + 3, C = 2 * (ss + 2) + 2;
    const int me = st->to_move, other = me ^ 1;
    const TzBoards b = warp_boards(st, nn, lane);
    const int s0 = n == 3 ? 10 : n == 4 ? 15 : n == 5 ? 21 : 30;
    const int c0 = n >= 5 ? 1 : 0;
    const float my_stones = __fdiv_rn((float)st->stones[me], (float)s0);
    const float my_caps = c0 ? __fdiv_rn((float)st->caps[me], (float)c0) : 0.0f;
    const float op_stones = __fdiv_rn((float)st->stones[other], (float)s0);
    const float op_caps = c0 ? __fdiv_rn((float)st->caps[other], (float)c0) : 0.0f;
    const float fcd = __fsub_rn((float)(__popcll(b.flat[0]) - __popcll(b.flat[1])), __fdiv_rn((float)half_komi, 2.0f));
    const float fcd_sq = __fdiv_rn(fcd, (float)nn);
    const float side = me == 1 ? 1.0f : 0.0f;
    for (int sq = lane; sq < nn; sq += 32) {
        float v[CIN_PAD];
#pragma unroll
        for (int c = 0; c < CIN_PAD; c++) v[c] = 0.0f;
        const int h = st->height[sq];
        if (h > 0) {
            const uint64_t stack = st->stack[sq];
            const int top_col = (int)((stack >> (h - 1)) & 1ull);
            v[st->top[sq] + (top_col != me ? ss : 0)] = 1.0f;
            for (int i = 0; i < ss - 3 && h - 2 - i >= 0; i++) {
                const int col = (int)((stack >> (h - 2 - i)) & 1ull);
                v[3 + i + (col != me ? ss : 0)] = 1.0f;
            }
        }
        const int base = 2 * ss;
        v[base] = my_stones;
        v[base + 1] = my_caps;
        v[base + 2] = op_stones;
        v[base + 3] = op_caps;
        v[base + 4] = side;
        v[base + 5] = fcd_sq;
        if (out_f32) {
            float* o = out_f32 + (size_t)q * C * nn + sq;
            for (int c = 0; c < C; c++) o[(size_t)c * nn] = v[c];
        }
        if (out_bf16) {
            const size_t r = (size_t)guard + (size_t)q * nn + sq;
#pragma unroll
            for (int j = 0; j < CIN_PAD / 8; j++)
                *reinterpret_cast<uint4*>(out_bf16 + ((size_t)j * (size_t)rows + r) * 8) =
                    make_uint4(conv::pack16(v[j * 8], v[j * 8 + 1], f16), conv::pack16(v[j * 8 + 2], v[j * 8 + 3], f16),
                               conv::pack16(v[j * 8 + 4], v[j * 8 + 5], f16), conv::pack16(v[j * 8 + 6], v[j * 8 + 7], f16));
        }
    }
}

// ---- SimHash novelty index (net6_simhash.rs:203-234 `get_indices`) ---------------------------------

// One warp per position, lane = hash bit: dot of the f32 planes (side-to-move plane zeroed) with column
// `lane` of the [C*N*N][32] matrix; bit set when the dot is >= 0; index = sum of 2^bit.
// f32 input planes of one position into shared memory (x[plane * nn + square], `game_repr` order); the warp's
// lanes own the squares.  zero_colour leaves the "black to move" plane at 0 (SimHash, net6_simhash.rs:209-222).
__device__ __forceinline__ void warp_fill_planes(float* x, const TzState* st, int n, int half_komi, int lane,
                                                 bool zero_colour) {
    const int nn = n * n, ss = 2 * n + 3, C = 2 * (ss + 2) + 2;
    for (int i = lane; i < C * nn; i += 32) x[i] = 0.0f;
    __syncwarp();
    const int me = st->to_move, other = me ^ 1;
    const TzBoards b = warp_boards(st, nn, lane);
    const int s0 = n == 3 ? 10 : n == 4 ? 15 : n == 5 ? 21 : 30;
    const int c0 = n >= 5 ? 1 : 0;
    const float r0 = __fdiv_rn((float)st->stones[me], (float)s0);
    const float r1 = c0 ? __fdiv_rn((float)st->caps[me], (float)c0) : 0.0f;
    const float r2 = __fdiv_rn((float)st->stones[other], (float)s0);
    const float r3 = c0 ? __fdiv_rn((float)st->caps[other], (float)c0) : 0.0f;
    const float fcd = __fsub_rn((float)(__popcll(b.flat[0]) - __popcll(b.flat[1])), __fdiv_rn((float)half_komi, 2.0f));
    const float fcd_sq = __fdiv_rn(fcd, (float)nn);
    for (int sq = lane; sq < nn; sq += 32) {
        const int h = st->height[sq];
        if (h > 0) {
            const uint64_t stack = st->stack[sq];
            const int top_col = (int)((stack >> (h - 1)) & 1ull);
            x[(st->top[sq] + (top_col != me ? ss : 0)) * nn + sq] = 1.0f;
            for (int i = 0; i < ss - 3 && h - 2 - i >= 0; i++) {
                const int col = (int)((stack >> (h - 2 - i)) & 1ull);
                x[(3 + i + (col != me ? ss : 0)) * nn + sq] = 1.0f;
            }
        }
        const int base = 2 * ss;
        x[(base + 0) * nn + sq] = r0;
        x[(base + 1) * nn + sq] = r1;
        x[(base + 2) * nn + sq] = r2;
        x[(base + 3) * nn + sq] = r3;
        if (!zero_colour && me == 1) x[(base + 4) * nn + sq] = 1.0f;
        x[(base + 5) * nn + sq] = fcd_sq;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(32 * WPB) k_simhash(const TzState* states, const int* count_ptr, int count_max, int n,
                                                       int half_komi, const float* matrix, uint32_t* out_idx) {
    __shared__ TzState s_state[WPB];
    __shared__ float s_planes[WPB][36 * 36];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * WPB + warp;
    const int count = count_ptr ? *count_ptr : count_max;
    if (q >= count) return;
    TzState* st = &s_state[warp];
    warp_load_state(st, &states[q], lane);
    const int nn = n * n, C = 2 * (2 * n + 3 + 2) + 2;
    float* x = s_planes[warp];
    warp_fill_planes(x, st, n, half_komi, lane, true);
    float dot = 0.0f;
    const int total = C * nn;
    for (int j = 0; j < total; j++) dot = fmaf(x[j], matrix[(size_t)j * 32 + lane], dot);
    const uint32_t bits = __ballot_sync(0xffffffffu, !(dot < 0.0f));
    if (lane == 0) out_idx[q] = bits;
}

// ---- heads + legal-logit gather ------------------------------------------------------------------

// network/repr.rs:49-71 `move_index` split into (channel, square)
__device__ __forceinline__ int move_channel(int n, uint16_t m) {
    const int kind = (m >> 6) & 3, pat = m >> 8;
    if (pat == 0) return kind;  // flat 0, wall 1, cap 2
    const int dir_off = kind == 0 ? 0 : kind == 1 ? 2 : kind == 2 ? 3 : 1;  // Up, Right, Down, Left order of repr.rs:61-66
    return 3 + ((pat >> (8 - n)) - 1) + ((1 << n) - 2) * dir_off;
}

// ---- LCG-hash novelty index (net4_lcghash.rs:203-241 `get_indices`) -------------------------------------
// planes * lcghash_init (f32, elementwise) reinterpreted as i32, folded with the 64-bit LCG
// acc = acc * 6364136223846793005 + 1 + v  along the columns, then the rows, then the channels (wrapping i64);
// index = |acc| >> 31.  Integer work: bit-exact.  One warp per position.
__global__ void __launch_bounds__(32 * WPB) k_lcghash(const TzState* states, const int* count_ptr, int count_max, int n,
                                                       int half_komi, const float* init, uint32_t* out_idx) {
    __shared__ TzState s_state[WPB];
    __shared__ float s_planes[WPB][36 * 36];
    __shared__ unsigned long long s_rows[WPB][36 * 6];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * WPB + warp;
    const int count = count_ptr ? *count_ptr : count_max;
    if (q >= count) return;
    TzState* st = &s_state[warp];
    warp_load_state(st, &states[q], lane);
    const int C = 2 * (2 * n + 3 + 2) + 2;
    float* x = s_planes[warp];
    warp_fill_planes(x, st, n, half_komi, lane, false);
    const unsigned long long MUL = 6364136223846793005ull;
    unsigned long long* rows = s_rows[warp];
    for (int cr = lane; cr < C * n; cr += 32) {  // (channel, row): fold the columns
        unsigned long long acc = 0;
        for (int j = 0; j < n; j++) {
            const int at = cr * n + j;
            const int v = __float_as_int(__fmul_rn(x[at], init[at]));
            acc = acc * MUL + 1ull + (unsigned long long)(long long)v;
        }
        rows[cr] = acc;
    }
    __syncwarp();
    for (int c = lane; c < C; c += 32) {  // channel: fold the rows (in place: slot c * n)
        unsigned long long acc = 0;
        for (int r = 0; r < n; r++) acc = acc * MUL + 1ull + rows[c * n + r];
        __syncwarp();
        rows[c * n] = acc;
    }
    __syncwarp();
    if (lane == 0) {
        unsigned long long acc = 0;
        for (int c = 0; c < C; c++) acc = acc * MUL + 1ull + rows[c * n];
        long long sacc = (long long)acc;
        if (sacc < 0) sacc = -sacc;  // i64::MIN stays negative, as a wrapping abs does
        out_idx[q] = (uint32_t)((unsigned long long)sacc >> 31);
    }
}

// One warp per position: value head (conv1x1 + ReLU + Linear + tanh), UBE head (same, no tanh),
// uncertainty = clamp(max(exp(ube), local), 0, 4) with local = 4.0 (empty SimHash set, i.e. a
// freshly initialised reference network), and logits[i] = policy[move_index(action_i)].
__global__ void __launch_bounds__(32 * WPB) k_heads_gather(const float* head_feat, const float* logits_full,
                                                            long long f32_rows, const float* head_misc,
                                                            const uint16_t* actions, const int* n_actions,
                                                            const int* count_ptr, int count_max, int n, int M,
                                                            const uint32_t* simhash_set, const uint32_t* simhash_idx,
                                                            float* out_logits, float* out_value, float* out_variance) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * WPB + warp;
    const int count = count_ptr ? *count_ptr : count_max;
    if (q >= count) return;
    const int nn = n * n;
    const float bv = head_misc[0], bu = head_misc[1];
    const float* lin_v = head_misc + 2;
    const float* lin_u = head_misc + 2 + 36;
    // lane = square (two passes for N = 6); the 1x1 convolutions (per-square dot products over the 256 channels)
    // were computed by the epilogue of the last tower convolution: head_feat[row] = {value, ube} features
    float acc_v = 0.0f, acc_u = 0.0f;
    for (int sq = lane; sq < nn; sq += 32) {
        const float2 d = *reinterpret_cast<const float2*>(head_feat + ((size_t)q * nn + sq) * 2);
        acc_v += fmaxf(d.x + bv, 0.0f) * lin_v[sq];
        acc_u += fmaxf(d.y + bu, 0.0f) * lin_u[sq];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_v += __shfl_xor_sync(0xffffffffu, acc_v, o);
        acc_u += __shfl_xor_sync(0xffffffffu, acc_u, o);
    }
    if (lane == 0) {
        out_value[q] = tanhf(acc_v + head_misc[2 + 72]);
        const float ube = acc_u + head_misc[2 + 73];
        // forward_hash (net6_simhash.rs:243-256): 0 when the position's SimHash bit is set, else 4.0
        float local = 4.0f;
        if (simhash_set) {
            const uint32_t idx = simhash_idx[q];
            if ((simhash_set[idx >> 5] >> (idx & 31)) & 1u) local = 0.0f;
        }
        out_variance[q] = fminf(fmaxf(fmaxf(expf(ube), local), 0.0f), 4.0f);
    }
    const int cnt = n_actions[q];
    const uint16_t* a = actions + (size_t)q * M;
    for (int i = lane; i < cnt; i += 32) {
        const uint16_t m = a[i];
        const int sq = ((m >> 3) & 7) * n + (m & 7);
        const int ch = move_channel(n, m);
        out_logits[(size_t)q * M + i] =
            logits_full[((size_t)(ch >> 2) * (size_t)f32_rows + (size_t)q * nn + sq) * 4 + (ch & 3)];
    }
}

// ---- weights ------------------------------------------------------------------------------------------

struct HostTensor {
    std::string name;
    const float* data;
    std::vector<long long> shape;
    size_t numel() const {
        size_t k = 1;
        for (long long s : shape) k *= (size_t)s;
        return k;
    }
};

static const HostTensor* find(const std::vector<HostTensor>& ts, const std::string& name) {
    for (const HostTensor& t : ts)
        if (t.name == name) return &t;
    return nullptr;
}

// Upper bounds of what conv::Schedule::init derives from a device-side count <= count_max: number of chunks and
// pair tiles of one chunk (a chunk has fewer than 2 * chunk_min_tiles tiles, + rounding).  The activation sets and
// the progress counters are sized with these.
struct ChunkBounds {
    int chunks, chunk_tiles;
};
static ChunkBounds chunk_bounds(int count_max, int nn, int chunk_min_tiles) {
    const int all_tiles = (int)(((long long)count_max * nn + 2 * conv::TILE_M - 1) / (2 * conv::TILE_M));
    ChunkBounds b;
    b.chunks = all_tiles >= 2LL * chunk_min_tiles ? all_tiles / chunk_min_tiles : 1;
    b.chunk_tiles = b.chunks > 1 ? 2 * chunk_min_tiles + 2 : (all_tiles > 0 ? all_tiles : 1);
    return b;
}

static uint16_t f32_to_bf16(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

static uint16_t f32_to_f16(float f) {  // round to nearest even, overflow -> inf, subnormals kept
    uint32_t u;
    memcpy(&u, &f, 4);
    const uint32_t sign = (u >> 16) & 0x8000u;
    const int32_t exp = (int32_t)((u >> 23) & 0xff) - 127 + 15;
    uint32_t man = u & 0x7fffffu;
    if (((u >> 23) & 0xff) == 0xff) return (uint16_t)(sign | 0x7c00u | (man ? 0x200u : 0));
    if (exp >= 31) return (uint16_t)(sign | 0x7c00u);
    if (exp <= 0) {
        if (exp < -10) return (uint16_t)sign;
        man |= 0x800000u;
        const int shift = 14 - exp;
        uint32_t h = man >> shift;
        const uint32_t rem = man & ((1u << shift) - 1), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (h & 1))) h++;
        return (uint16_t)(sign | h);
    }
    uint32_t h = ((uint32_t)exp << 10) | (man >> 13);
    const uint32_t rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) h++;
    return (uint16_t)(sign | h);
}

// conv [cout][cin][3][3] (+ BN, folded) -> bf16 blocks [cin_pad/64][9][8][256][8] + f32 bias[256]
static int upload_conv(NnState* s, ConvLayer* L, const HostTensor* w, const HostTensor* conv_bias, const HostTensor* bn_w,
                       const HostTensor* bn_b, const HostTensor* bn_m, const HostTensor* bn_v, int cin_pad) {
    const int cout = (int)w->shape[0], cin = (int)w->shape[1];
    if (cout > FILTERS || cin > cin_pad || w->shape[2] != 3 || w->shape[3] != 3) return TZ_EINVAL;
    std::vector<float> scale(cout, 1.0f), bias(FILTERS, 0.0f);
    for (int co = 0; co < cout; co++) {
        if (bn_w) {
            const float inv = 1.0f / sqrtf(bn_v->data[co] + 1e-5f);  // tch BatchNormConfig::default eps
            scale[co] = bn_w->data[co] * inv;
            bias[co] = bn_b->data[co] - bn_m->data[co] * scale[co];
        }
        if (conv_bias) bias[co] += conv_bias->data[co] * scale[co];
    }
    const int kblocks = cin_pad / 64;
    std::vector<uint16_t> blk((size_t)kblocks * 9 * 8 * 256 * 8, 0);
    // blocks are stored in the order the kernel consumes the taps: centre first, then the rest
    for (int co = 0; co < cout; co++)
        for (int ci = 0; ci < cin; ci++)
            for (int ti = 0; ti < 9; ti++) {
                const int tap = ti == 0 ? 4 : (ti <= 4 ? ti - 1 : ti);  // tap = ky * 3 + kx
                const float v = w->data[((size_t)co * cin + ci) * 9 + tap] * scale[co];
                const int kb = ci / 64, kc = (ci % 64) / 8, e = ci % 8;
                // one 32 KB block per (kb, tap) holding the two N halves as separate contiguous 16 KB
                // shared-memory images (one per CTA of the pair): [2 halves][8 k-chunks][128 n][8]
                const size_t in_blk = (((size_t)(co / 128) * 8 + kc) * 128 + co % 128) * 8 + e;
                blk[((size_t)kb * 9 + ti) * (8 * 256 * 8) + in_blk] = s->f16 ? f32_to_f16(v) : f32_to_bf16(v);
            }
    void* dw = nullptr;
    void* db = nullptr;
    if (cudaMalloc(&dw, blk.size() * 2) != cudaSuccess) return TZ_ENOMEM;
    s->allocs.push_back(dw);
    if (cudaMalloc(&db, FILTERS * sizeof(float)) != cudaSuccess) return TZ_ENOMEM;
    s->allocs.push_back(db);
    if (cudaMemcpy(dw, blk.data(), blk.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) return TZ_ECUDA;
    if (cudaMemcpy(db, bias.data(), FILTERS * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return TZ_ECUDA;
    L->w = (__nv_bfloat16*)dw;
    L->bias = (float*)db;
    L->cin = cin_pad;
    return TZ_OK;
}

static thread_local char g_nn_err[256] = "";
const char* nn_last_error() { return g_nn_err; }
#define NN_FAIL(code, ...)                              \
    do {                                                \
        snprintf(g_nn_err, sizeof(g_nn_err), __VA_ARGS__); \
        return code;                                    \
    } while (0)

int nn_set_weights(tz_handle* h, const char* const* names, const float* const* data, const long long* const* shapes,
                   const int* ndims, int count) {
    cudaSetDevice(h->device);
    std::vector<HostTensor> ts;
    for (int i = 0; i < count; i++) {
        HostTensor t;
        t.name = names[i];
        t.data = data[i];
        t.shape.assign(shapes[i], shapes[i] + ndims[i]);
        ts.push_back(t);
    }
    // check names and shapes before touching the current model: a reload from a bad file (the reference keeps its
    // previous `net` when Net::load fails, selfplay/src/main.rs:107-119) must leave the handle usable
    {
        const long long n0 = h->d.n, cin0 = 2 * (2 * n0 + 3 + 2) + 2, cout0 = 3 + 4 * ((1ll << n0) - 2);
        std::vector<std::pair<std::string, std::vector<long long>>> req;
        auto bn = [&](const std::string& p) {
            for (const char* f : {"weight", "bias", "running_mean", "running_var"}) req.push_back({p + "." + f, {FILTERS}});
        };
        req.push_back({"core.input_conv2d.weight", {FILTERS, cin0, 3, 3}});
        bn("core.batch_norm");
        int nb = 0;
        while (find(ts, "core.res_block_" + std::to_string(nb) + ".0.conv2d.weight")) nb++;
        if (nb == 0) NN_FAIL(TZ_EINVAL, "no residual blocks (core.res_block_0.0.conv2d.weight) found");
        for (int b = 0; b < nb; b++)
            for (int j = 0; j < 2; j++) {
                const std::string p = "core.res_block_" + std::to_string(b) + "." + std::to_string(j);
                req.push_back({p + ".conv2d.weight", {FILTERS, FILTERS, 3, 3}});
                bn(p + ".batch_norm");
            }
        req.push_back({"policy.conv2d.weight", {cout0, FILTERS, 3, 3}});
        req.push_back({"policy.conv2d.bias", {cout0}});
        for (const char* head : {"value", "ube"}) {
            req.push_back({std::string(head) + ".conv2d.weight", {1, FILTERS, 1, 1}});
            req.push_back({std::string(head) + ".conv2d.bias", {1}});
            req.push_back({std::string(head) + ".linear.weight", {1, n0 * n0}});
            req.push_back({std::string(head) + ".linear.bias", {1}});
        }
        for (const auto& r : req) {
            const HostTensor* t = find(ts, r.first);
            if (!t) NN_FAIL(TZ_EINVAL, "missing tensor %s", r.first.c_str());
            if (t->shape != r.second) NN_FAIL(TZ_EINVAL, "tensor %s has the wrong shape", r.first.c_str());
        }
    }
    // a model reload (selfplay/src/main.rs:107 does one per move) keeps the activation buffers of the previous
    // state: only the weights are re-folded and re-uploaded
    struct OldState {
        NnState* p;
        ~OldState() {
            if (!p) return;
            for (void* a : p->allocs) cudaFree(a);
            delete p;
        }
        void* take(void* ptr) {  // hand one allocation over to the new state
            if (!ptr) return nullptr;
            for (size_t i = 0; i < p->allocs.size(); i++)
                if (p->allocs[i] == ptr) {
                    p->allocs.erase(p->allocs.begin() + (long)i);
                    return ptr;
                }
            return nullptr;
        }
    } old{h->nn};
    h->nn = nullptr;
    NnState* s = new NnState();
    h->nn = s;
    s->f16 = h->nn_f16;
    const int n = h->d.n, nn = n * n;
    s->n = n;
    s->in_channels = 2 * (2 * n + 3 + 2) + 2;
    s->out_channels = 3 + 4 * ((1 << n) - 2);
    cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, h->device);
    auto need = [&](const std::string& name, std::vector<long long> shape) -> const HostTensor* {
        const HostTensor* t = find(ts, name);
        if (!t) {
            snprintf(g_nn_err, sizeof(g_nn_err), "missing tensor %s", name.c_str());
            return nullptr;
        }
        if (t->shape != shape) {
            snprintf(g_nn_err, sizeof(g_nn_err), "tensor %s has the wrong shape", name.c_str());
            return nullptr;
        }
        return t;
    };
    int rc;
#define NEED(var, name, ...)                                  \
    const HostTensor* var = need(name, std::vector<long long>{__VA_ARGS__}); \
    if (!var) {                                               \
        nn_free(h);                                           \
        return TZ_EINVAL;                                     \
    }
    {
        NEED(w, "core.input_conv2d.weight", FILTERS, s->in_channels, 3, 3);
        NEED(bw, "core.batch_norm.weight", FILTERS);
        NEED(bb, "core.batch_norm.bias", FILTERS);
        NEED(bm, "core.batch_norm.running_mean", FILTERS);
        NEED(bv, "core.batch_norm.running_var", FILTERS);
        if ((rc = upload_conv(s, &s->input, w, nullptr, bw, bb, bm, bv, CIN_PAD)) != TZ_OK) {
            nn_free(h);
            NN_FAIL(rc, "input conv upload failed");
        }
    }
    int blocks = 0;
    while (find(ts, "core.res_block_" + std::to_string(blocks) + ".0.conv2d.weight")) blocks++;
    if (blocks == 0) {
        nn_free(h);
        NN_FAIL(TZ_EINVAL, "no residual blocks (core.res_block_0.0.conv2d.weight) found");
    }
    s->blocks = blocks;
    for (int b = 0; b < blocks; b++)
        for (int j = 0; j < 2; j++) {
            const std::string p = "core.res_block_" + std::to_string(b) + "." + std::to_string(j);
            NEED(w, p + ".conv2d.weight", FILTERS, FILTERS, 3, 3);
            NEED(bw, p + ".batch_norm.weight", FILTERS);
            NEED(bb, p + ".batch_norm.bias", FILTERS);
            NEED(bm, p + ".batch_norm.running_mean", FILTERS);
            NEED(bv, p + ".batch_norm.running_var", FILTERS);
            ConvLayer L;
            if ((rc = upload_conv(s, &L, w, nullptr, bw, bb, bm, bv, FILTERS)) != TZ_OK) {
                nn_free(h);
                NN_FAIL(rc, "tower conv upload failed");
            }
            s->tower.push_back(L);
        }
    {
        NEED(w, "policy.conv2d.weight", s->out_channels, FILTERS, 3, 3);
        NEED(b, "policy.conv2d.bias", s->out_channels);
        if ((rc = upload_conv(s, &s->policy, w, b, nullptr, nullptr, nullptr, nullptr, FILTERS)) != TZ_OK) {
            nn_free(h);
            NN_FAIL(rc, "policy conv upload failed");
        }
    }
    {
        NEED(vw, "value.conv2d.weight", 1, FILTERS, 1, 1);
        NEED(vb, "value.conv2d.bias", 1);
        NEED(vlw, "value.linear.weight", 1, nn);
        NEED(vlb, "value.linear.bias", 1);
        NEED(uw, "ube.conv2d.weight", 1, FILTERS, 1, 1);
        NEED(ub, "ube.conv2d.bias", 1);
        NEED(ulw, "ube.linear.weight", 1, nn);
        NEED(ulb, "ube.linear.bias", 1);
        std::vector<float> hw(2 * FILTERS), misc(2 + 72 + 2, 0.0f);
        memcpy(hw.data(), vw->data, FILTERS * 4);
        memcpy(hw.data() + FILTERS, uw->data, FILTERS * 4);
        misc[0] = vb->data[0];
        misc[1] = ub->data[0];
        memcpy(misc.data() + 2, vlw->data, nn * 4);
        memcpy(misc.data() + 2 + 36, ulw->data, nn * 4);
        misc[2 + 72] = vlb->data[0];
        misc[2 + 73] = ulb->data[0];
        void *d1 = nullptr, *d2 = nullptr;
        if (cudaMalloc(&d1, hw.size() * 4) != cudaSuccess || cudaMalloc(&d2, misc.size() * 4) != cudaSuccess) {
            nn_free(h);
            NN_FAIL(TZ_ENOMEM, "cudaMalloc heads");
        }
        s->allocs.push_back(d1);
        s->allocs.push_back(d2);
        cudaMemcpy(d1, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(d2, misc.data(), misc.size() * 4, cudaMemcpyHostToDevice);
        s->head_w = (float*)d1;
        s->head_misc = (float*)d2;
    }
#undef NEED
    // activation buffers: guard rows + all boards, rounded up to whole tiles, + halo
    s->max_positions = h->d.Q;
    const size_t used = (size_t)s->max_positions * nn;
    s->rows = conv::HALO + ((used + 2 * conv::TILE_M - 1) / (2 * conv::TILE_M)) * (2 * conv::TILE_M) + 2 * conv::HALO;
    auto dalloc = [&](void** p, size_t bytes) -> bool {
        if (cudaMalloc(p, bytes) != cudaSuccess) return false;
        s->allocs.push_back(*p);
        return cudaMemset(*p, 0, bytes) == cudaSuccess;
    };
    {
        // chunking: at least TZ_NN_CHUNK_TILES pair tiles (256 rows each) per chunk, default 150 = two per CTA
        // pair (measured plateau 144..192 on 8192 6x6 positions); 0 or the per-layer mode = one chunk
        // (a chain longer than MAX_LAYERS is cut into several launches and cannot keep chunks apart)
        const char* mode = getenv("TZ_TOWER");
        s->fused = !(mode && strcmp(mode, "layers") == 0);
        const char* ct = getenv("TZ_NN_CHUNK_TILES");
        const int tiles = ct ? atoi(ct) : 150;
        const int all_tiles = (int)((used + 2 * conv::TILE_M - 1) / (2 * conv::TILE_M));
        s->chunk_min_tiles = (!s->fused || tiles <= 0 || all_tiles < 2 * tiles || 2 * s->blocks + 2 > conv::MAX_LAYERS)
                                 ? (1 << 28) : tiles;
        const ChunkBounds cb = chunk_bounds(s->max_positions, nn, s->chunk_min_tiles);
        s->chunk_tiles = cb.chunk_tiles;
        s->max_chunks = cb.chunks;
        s->chunk_positions = cb.chunks > 1 ? (cb.chunk_tiles - 1) * 2 * conv::TILE_M / nn : s->max_positions;
        s->rows_set = conv::HALO + (size_t)s->chunk_tiles * (2 * conv::TILE_M) + 2 * conv::HALO;
    }
    const bool reuse = old.p && old.p->n == n && old.p->max_positions == h->d.Q && old.p->rows == s->rows &&
                       old.p->rows_set == s->rows_set && old.p->max_chunks == s->max_chunks;
    if (reuse) {
        auto steal = [&](auto*& dst, auto* src) {
            dst = static_cast<std::remove_reference_t<decltype(dst)>>(old.take(src));
            if (dst) s->allocs.push_back(dst);
        };
        steal(s->planes, old.p->planes);
        steal(s->act_x, old.p->act_x);
        steal(s->act_t, old.p->act_t);
        steal(s->head_feat, old.p->head_feat);
        for (int i = 0; i < 3; i++) steal(s->tune_buf[i], old.p->tune_buf[i]);
        steal(s->logits_full, old.p->logits_full);
        steal(s->masks, old.p->masks);
        steal(s->simhash_matrix, old.p->simhash_matrix);
        steal(s->lcghash_init, old.p->lcghash_init);
        s->novelty = old.p->novelty;
        steal(s->simhash_set, old.p->simhash_set);
        steal(s->simhash_idx, old.p->simhash_idx);
    }
    if (!reuse || !s->planes || !s->act_x || !s->act_t || !s->logits_full || !s->head_feat)
    {
        if (!dalloc((void**)&s->planes, s->rows * CIN_PAD * 2) || !dalloc((void**)&s->act_x, 2 * s->rows_set * FILTERS * 2) ||
            !dalloc((void**)&s->act_t, 2 * s->rows_set * FILTERS * 2) ||
            !dalloc((void**)&s->head_feat, (size_t)s->max_positions * nn * 2 * sizeof(float)) ||
            !dalloc((void**)&s->logits_full, (size_t)s->max_positions * nn * FILTERS * 4)) {
            nn_free(h);
            NN_FAIL(TZ_ENOMEM, "cudaMalloc activations");
        }
    }
    if (!s->masks) {
        // lane masks: bit i of mask[start][tap] is set when tile row i (board square (start + i) mod nn)
        // has no (dy,dx) neighbour on the board
        std::vector<uint32_t> mk((size_t)nn * 9 * 4, 0);
        for (int start = 0; start < nn; start++)
            for (int tap = 0; tap < 9; tap++) {
                const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                for (int i = 0; i < conv::TILE_M; i++) {
                    const int sq = (start + i) % nn, y = sq / n, x = sq % n;
                    const bool off = y + dy < 0 || y + dy >= n || x + dx < 0 || x + dx >= n;
                    if (off) mk[((size_t)start * 9 + tap) * 4 + i / 32] |= 1u << (i % 32);
                }
            }
        if (!dalloc((void**)&s->masks, mk.size() * 4) ||
            cudaMemcpy(s->masks, mk.data(), mk.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
            nn_free(h);
            NN_FAIL(TZ_ENOMEM, "cudaMalloc masks");
        }
    }
    if (cudaFuncSetAttribute(conv::k_conv3x3_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, conv::SMEM_BYTES) !=
        cudaSuccess) {
        nn_free(h);
        NN_FAIL(TZ_ECUDA, "cudaFuncSetAttribute(k_conv3x3_pair, %d B smem) failed", conv::SMEM_BYTES);
    }
    {
        // CTA pairs that fit on the device at once: the fused tower spins on other pairs' progress, so its grid
        // must never exceed this
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * (s->sm_count / 2));
        cfg.blockDim = dim3(conv::THREADS);
        cfg.dynamicSmemBytes = conv::SMEM_BYTES;
        int clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&clusters, conv::k_conv3x3_pair, &cfg) == cudaSuccess && clusters > 0)
            s->max_pairs = clusters < s->sm_count / 2 ? clusters : s->sm_count / 2;
        else
            s->max_pairs = s->sm_count / 2;
        const size_t tiles = (size_t)s->max_chunks * s->chunk_tiles * 4 + s->max_chunks + 16;
        s->progress_len = tiles;
        if (reuse) {
            s->progress = static_cast<unsigned*>(old.take(old.p->progress));
            if (s->progress) s->allocs.push_back(s->progress);
        }
        if (!s->progress && !dalloc((void**)&s->progress, tiles * sizeof(unsigned))) {
            nn_free(h);
            NN_FAIL(TZ_ENOMEM, "cudaMalloc progress");
        }
    }
    cudaDeviceSynchronize();
    return TZ_OK;
}

// test hook (no GPU needed): the kernel's own work-item schedule for `count` positions, and the bounds the host
// sizes the activation sets and progress counters with.  out[0..4] = items, chunks, chunk_tiles, chunk_rows of the
// schedule; out[4..7] = host bounds: chunks, chunk_tiles, rows per activation set; triples (chunk, layer, pair tile)
// of up to `cap` items go to out_items.
int nn_debug_schedule(int count, int count_max, int n, int chunk_min_tiles, int layers, long long* out, int* out_items,
                      int cap) {
    const int nn = n * n;
    conv::Schedule sc;
    sc.init(count, nn, chunk_min_tiles, layers);
    out[0] = sc.items;
    out[1] = sc.rows_used > 0 ? (sc.rows_used + sc.chunk_rows - 1) / sc.chunk_rows : 0;
    out[2] = sc.chunk_tiles;
    out[3] = sc.chunk_rows;
    const ChunkBounds cb = chunk_bounds(count_max, nn, chunk_min_tiles);
    out[4] = cb.chunks;
    out[5] = cb.chunk_tiles;
    out[6] = conv::HALO + (long long)cb.chunk_tiles * (2 * conv::TILE_M) + 2 * conv::HALO;
    for (int i = 0; i < sc.items && i < cap; i++) {
        const conv::Item it = sc.at(i);
        out_items[3 * i] = it.chunk;
        out_items[3 * i + 1] = it.layer;
        out_items[3 * i + 2] = it.pt;
    }
    return 0;
}

void nn_set_layer_limit(tz_handle* h, int limit) {
    if (h->nn) h->nn->layer_limit = limit;
}

// ---- forward ---------------------------------------------------------------------------------------------

static conv::Layer conv_layer(const ConvLayer& L, const __nv_bfloat16* in, const __nv_bfloat16* residual,
                              __nv_bfloat16* out_act, float* out_f32, int relu) {
    conv::Layer l;
    l.in = in;
    l.w = L.w;
    l.bias = L.bias;
    l.residual = residual;
    l.out_act = out_act;
    l.out_f32 = out_f32;
    l.head_w = nullptr;
    l.head_out = nullptr;
    l.cin = L.cin;
    l.relu = relu;
    l.in_global = 0;
    return l;
}

// Launches p.layers[0..n_layers) as ONE persistent kernel over activation sets of `rows_set` rows holding
// one chunk (at least `chunk_min_tiles` pair tiles) each.  With more than one layer the CTA pairs synchronise through s->progress
// inside the kernel, so the launch is cooperative (all pairs resident, or it fails loudly).
static cudaError_t launch_layers(tz_handle* h, conv::Params& p, const int* count_ptr, int count_max, size_t rows_set,
                                 int chunk_min_tiles) {
    const NnState* s = h->nn;
    const int nn = s->n * s->n;
    p.rows_global = (long long)s->rows;
    p.rows_set = (long long)rows_set;
    p.set_stride = (long long)rows_set * FILTERS;
    p.chunk_min_tiles = chunk_min_tiles;
    p.f32_rows = (long long)s->max_positions * nn;
    p.count_ptr = count_ptr;
    p.count_max = count_max;
    p.n = s->n;
    p.guard = conv::HALO;
    p.masks = s->masks;
    p.f16 = s->f16;
    // upper bounds of what the kernel derives from the device-side count (conv::Schedule::init)
    const int all_tiles = (count_max * nn + 2 * conv::TILE_M - 1) / (2 * conv::TILE_M);
    const ChunkBounds cb = chunk_bounds(count_max, nn, chunk_min_tiles);
    const int chunks = cb.chunks, chunk_tiles = cb.chunk_tiles;
    const size_t counters = (size_t)chunks * chunk_tiles * 4 + chunks;  // 4 channel blocks per tile, then the chunks
    p.progress = s->progress;
    p.chunk_done = s->progress + (size_t)chunks * chunk_tiles * 4;
    p.status = h->d.status;
    p.debug_drop_progress = getenv("TZ_EXP_DROP_PROGRESS") != nullptr;  // watchdog test only
    const long long items = (long long)all_tiles * p.n_layers;
    const int pairs = items < s->max_pairs ? (items > 0 ? (int)items : 1) : s->max_pairs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(conv::THREADS);
    cfg.dynamicSmemBytes = conv::SMEM_BYTES;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    if (p.n_layers > 1) {
        if (counters > s->progress_len || (chunks > 1 && (size_t)chunk_tiles * 2 * conv::TILE_M + 3 * conv::HALO > rows_set))
            return cudaErrorInvalidValue;
        cudaError_t e = cudaMemsetAsync(s->progress, 0, counters * sizeof(unsigned), h->stream);
        if (e != cudaSuccess) return e;
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, conv::k_conv3x3_pair, p);
}

// one convolution over `count_max` positions that all sit in ONE activation set of `rows_set` rows
static cudaError_t launch_conv(tz_handle* h, const conv::Layer& layer, const int* count_ptr, int count_max,
                               size_t rows_set) {
    conv::Params p;
    p.layers[0] = layer;
    p.n_layers = 1;
    return launch_layers(h, p, count_ptr, count_max, rows_set, 1 << 28);
}

// The whole network body: input conv (planes -> x), residual blocks (conv(x) -> t, conv(t) + x -> x; the last one
// also emits the value / UBE head features), policy conv (x -> f32 logits).  `upto` < 0: all of it; otherwise only
// the first `upto` convolutions (debug hook).  Fused: one launch (chunks of conv::MAX_LAYERS layers), else one
// launch per layer.  Returns the number of launches, or -1 when a launch was refused.
static int launch_network(tz_handle* h, const int* count_ptr, int count_max, int upto) {
    NnState* s = h->nn;
    std::vector<conv::Layer> all;
    all.push_back(conv_layer(s->input, s->planes, nullptr, s->act_x, nullptr, 1));
    all.back().in_global = 1;
    for (int l = 0; l < 2 * s->blocks; l++)
        all.push_back((l & 1) ? conv_layer(s->tower[l], s->act_t, s->act_x, s->act_x, nullptr, 1)
                              : conv_layer(s->tower[l], s->act_x, nullptr, s->act_t, nullptr, 1));
    all.back().head_w = s->head_w;
    all.back().head_out = s->head_feat;
    all.push_back(conv_layer(s->policy, s->act_x, nullptr, nullptr, s->logits_full, 0));
    if (upto >= 0 && (size_t)upto < all.size()) all.resize((size_t)upto);
    int launches = 0;
    for (size_t first = 0; first < all.size();) {
        const size_t left = all.size() - first;
        const size_t chunk = s->fused ? (left < (size_t)conv::MAX_LAYERS ? left : (size_t)conv::MAX_LAYERS) : 1;
        conv::Params p;
        for (size_t i = 0; i < chunk; i++) p.layers[i] = all[first + i];
        p.n_layers = (int)chunk;
        // a chain cut into several launches (more than MAX_LAYERS layers) cannot keep chunks in flight across the
        // cut: it then runs with one chunk (full-size sets are required, see nn_set_weights)
        if (launch_layers(h, p, count_ptr, count_max, s->rows_set, s->chunk_min_tiles) != cudaSuccess) return -1;
        first += chunk;
        launches++;
    }
    return launches;
}

// states[count] (device), actions/n_actions by position -> logits/value/variance by position.
// count_ptr (device) may be smaller than count_max; kernels size themselves from it.
int nn_forward(tz_handle* h, const TzState* states, const int* count_ptr, int count_max, const uint16_t* actions,
               const int* n_actions, float* logits, float* value, float* variance) {
    NnState* s = h->nn;
    if (!s) return TZ_ENOWEIGHTS;
    if (count_max > s->max_positions) return TZ_EINVAL;
    const TzDev& d = h->d;
    const int wblocks = (count_max + WPB - 1) / WPB;
    {
        ProfScope ps(h, TZ_PROF_ENCODE);
        k_encode<<<wblocks, 32 * WPB, 0, h->stream>>>(states, count_ptr, count_max, d.n, d.half_komi, nullptr, s->planes,
                                                      conv::HALO, (long long)s->rows, s->f16);
    }
    const int limit = s->layer_limit;
    // the debug read-back shows one chunk: refuse more positions than run as a single chunk
    if (limit >= 0 && (count_max * d.n * d.n + 2 * conv::TILE_M - 1) / (2 * conv::TILE_M) >= 2LL * s->chunk_min_tiles)
        return TZ_EINVAL;
    {
        // input, tower and policy convolutions are one launch, so the sampled profile books all of it here
        ProfScope ps(h, TZ_PROF_CONV_TOWER);
        const int launched = launch_network(h, count_ptr, count_max, limit);
        if (launched < 0) {
            cudaGetLastError();
            return TZ_ECUDA;
        }
        h->launches += 1 + launched;
    }
    if (limit >= 0) return cudaGetLastError() == cudaSuccess ? TZ_OK : TZ_ECUDA;
    if (s->simhash_set && s->novelty == 2)
        k_lcghash<<<wblocks, 32 * WPB, 0, h->stream>>>(states, count_ptr, count_max, d.n, d.half_komi, s->lcghash_init,
                                                       s->simhash_idx);
    else if (s->simhash_set)
        k_simhash<<<wblocks, 32 * WPB, 0, h->stream>>>(states, count_ptr, count_max, d.n, d.half_komi, s->simhash_matrix,
                                                       s->simhash_idx);
    {
        ProfScope ps(h, TZ_PROF_HEADS);
        k_heads_gather<<<wblocks, 32 * WPB, 0, h->stream>>>(
            s->head_feat, s->logits_full, (long long)s->max_positions * d.n * d.n, s->head_misc, actions, n_actions,
            count_ptr, count_max, d.n, d.M, s->simhash_set, s->simhash_idx, logits, value, variance);
    }
    h->launches += 1;
    return cudaGetLastError() == cudaSuccess ? TZ_OK : TZ_ECUDA;
}

int nn_forward_queue(tz_handle* h) {
    const TzDev& d = h->d;
    return nn_forward(h, d.leaf_state, d.nn_count, d.Q, d.actions, d.n_actions, d.logits, d.value, d.variance);
}

int nn_encode_planes(tz_handle* h, const TzState* states, int count, float* out_f32) {
    const TzDev& d = h->d;
    k_encode<<<(count + WPB - 1) / WPB, 32 * WPB, 0, h->stream>>>(states, nullptr, count, d.n, d.half_komi, out_f32,
                                                                  nullptr, 0, 0, 0);
    return cudaGetLastError() == cudaSuccess ? TZ_OK : TZ_ECUDA;
}

// debug read-back: which = 0 act_x, 1 act_t, 2 planes; f32 [count][n*n][channels]
__global__ void k_unpad(const __nv_bfloat16* buf, int channels, int count, int n, int guard, long long rows, int f16,
                        float* out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int nn = n * n;
    if (idx >= (size_t)count * nn * channels) return;
    const int c = (int)(idx % channels);
    const size_t cell = idx / channels;  // position * nn + square
    const size_t at = ((size_t)(c >> 3) * (size_t)rows + (size_t)guard + cell) * 8 + (c & 7);
    out[idx] = f16 ? __half2float(reinterpret_cast<const __half*>(buf)[at]) : __bfloat162float(buf[at]);
}

int nn_debug_read(tz_handle* h, int which, int count, float* out_dev) {
    NnState* s = h->nn;
    if (!s) return TZ_ENOWEIGHTS;
    const __nv_bfloat16* buf = which == 0 ? s->act_x : which == 1 ? s->act_t : s->planes;
    const int channels = which == 2 ? CIN_PAD : FILTERS;
    // the activation sets hold one chunk
    if (which != 2 && (count * s->n * s->n + 2 * conv::TILE_M - 1) / (2 * conv::TILE_M) >= 2LL * s->chunk_min_tiles)
        return TZ_EINVAL;
    const size_t rows = which == 2 ? s->rows : s->rows_set;
    const size_t total = (size_t)count * s->n * s->n * channels;
    k_unpad<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(buf, channels, count, s->n, conv::HALO,
                                                                    (long long)rows, s->f16, out_dev);
    return cudaGetLastError() == cudaSuccess ? TZ_OK : TZ_ECUDA;
}

// test / tuning hook: time `reps` repetitions of one residual block (2 tower convolutions) over
// `count` positions with CUDA events; returns the mean milliseconds per convolution launch.  The block
// stream X is only read (the second convolution writes to a scratch buffer), so the data stay whatever
// the last tz_evaluate left there -- realistic activations, which matters under the power cap.
// TZ_EXP_GAP_US (environment, tuning experiment): idle this many microseconds between the timed launches, to see
// whether the power-capped tensor clock absorbs idle gaps (it does: see profiles/r1_conv_timing.txt)
__global__ void k_idle(long long ns) {
    long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
        __nanosleep(1000);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    } while (t - t0 < ns);
}

int nn_time_tower(tz_handle* h, int count, int reps, double* ms_per_conv) {
    NnState* s = h->nn;
    if (!s) return TZ_ENOWEIGHTS;
    if (count <= 0 || count > s->max_positions || reps <= 0) return TZ_EINVAL;
    // own full-size buffers (the network's activation sets hold one chunk): X is filled by the input convolution
    // from whatever positions the last tz_evaluate encoded -- realistic activations, which matters under the
    // power cap -- and only read afterwards
    for (int i = 0; i < 3; i++)
        if (!s->tune_buf[i]) {
            if (cudaMalloc((void**)&s->tune_buf[i], s->rows * FILTERS * 2) != cudaSuccess) return TZ_ENOMEM;
            s->allocs.push_back(s->tune_buf[i]);
            cudaMemset(s->tune_buf[i], 0, s->rows * FILTERS * 2);
        }
    __nv_bfloat16 *x = s->tune_buf[0], *t = s->tune_buf[1], *scratch = s->tune_buf[2];
    conv::Layer first = conv_layer(s->input, s->planes, nullptr, x, nullptr, 1);
    first.in_global = 1;
    launch_conv(h, first, nullptr, count, s->rows);
    const conv::Layer c1 = conv_layer(s->tower[0], x, nullptr, t, nullptr, 1);
    const conv::Layer c2 = conv_layer(s->tower[1], t, x, scratch, nullptr, 1);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int i = 0; i < 2; i++) {
        launch_conv(h, c1, nullptr, count, s->rows);
        launch_conv(h, c2, nullptr, count, s->rows);
    }
    const char* gap_env = getenv("TZ_EXP_GAP_US");
    const long long gap_ns = gap_env ? 1000ll * atoll(gap_env) : 0;
    cudaEventRecord(a, h->stream);
    for (int i = 0; i < reps; i++) {
        launch_conv(h, c1, nullptr, count, s->rows);
        if (gap_ns) k_idle<<<1, 1, 0, h->stream>>>(gap_ns);
        launch_conv(h, c2, nullptr, count, s->rows);
        if (gap_ns) k_idle<<<1, 1, 0, h->stream>>>(gap_ns);
    }
    cudaEventRecord(b, h->stream);
    cudaEventSynchronize(b);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *ms_per_conv = (double)ms / (2.0 * reps);
    return cudaGetLastError() == cudaSuccess ? TZ_OK : TZ_ECUDA;
}

// SimHash matrix ([C*N*N][32] f32) and the optional 2^32-bit set (512 MiB, the reference's bitvec.bin)
int nn_set_simhash(tz_handle* h, const float* matrix, const unsigned char* bitset) {
    NnState* s = h->nn;
    if (!s) return TZ_ENOWEIGHTS;
    const size_t rows = (size_t)s->in_channels * s->n * s->n;
    if (!s->simhash_matrix) {
        if (cudaMalloc((void**)&s->simhash_matrix, rows * 32 * 4) != cudaSuccess) return TZ_ENOMEM;
        s->allocs.push_back(s->simhash_matrix);
    }
    if (!s->simhash_idx) {
        if (cudaMalloc((void**)&s->simhash_idx, (size_t)s->max_positions * 4) != cudaSuccess) return TZ_ENOMEM;
        s->allocs.push_back(s->simhash_idx);
    }
    if (cudaMemcpy(s->simhash_matrix, matrix, rows * 32 * 4, cudaMemcpyHostToDevice) != cudaSuccess) return TZ_ECUDA;
    s->novelty = 1;
    if (bitset) {
        const size_t bytes = (size_t)1 << 29;
        if (!s->simhash_set) {
            if (cudaMalloc((void**)&s->simhash_set, bytes) != cudaSuccess) return TZ_ENOMEM;
            s->allocs.push_back(s->simhash_set);
        }
        if (cudaMemcpy(s->simhash_set, bitset, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return TZ_ECUDA;
    } else {
        s->simhash_set = nullptr;  // empty set: local uncertainty is MAXIMUM_VARIANCE everywhere
    }
    return TZ_OK;
}

// LCG-hash novelty (net4_lcghash.rs): the per-cell multipliers and the optional 2^32-bit set; replaces SimHash
int nn_set_lcghash(tz_handle* h, const float* init, const unsigned char* bitset) {
    NnState* s = h->nn;
    if (!s) return TZ_ENOWEIGHTS;
    const size_t cells = (size_t)s->in_channels * s->n * s->n;
    if (!s->lcghash_init) {
        if (cudaMalloc((void**)&s->lcghash_init, cells * 4) != cudaSuccess) return TZ_ENOMEM;
        s->allocs.push_back(s->lcghash_init);
    }
    if (!s->simhash_idx) {
        if (cudaMalloc((void**)&s->simhash_idx, (size_t)s->max_positions * 4) != cudaSuccess) return TZ_ENOMEM;
        s->allocs.push_back(s->simhash_idx);
    }
    if (cudaMemcpy(s->lcghash_init, init, cells * 4, cudaMemcpyHostToDevice) != cudaSuccess) return TZ_ECUDA;
    s->novelty = 2;
    if (bitset) {
        const size_t bytes = (size_t)1 << 29;
        if (!s->simhash_set) {
            if (cudaMalloc((void**)&s->simhash_set, bytes) != cudaSuccess) return TZ_ENOMEM;
            s->allocs.push_back(s->simhash_set);
        }
        if (cudaMemcpy(s->simhash_set, bitset, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return TZ_ECUDA;
    } else {
        s->simhash_set = nullptr;
    }
    return TZ_OK;
}

int nn_lcghash_indices(tz_handle* h, const TzState* states, int count, uint32_t* out_dev) {
    NnState* s = h->nn;
    if (!s || !s->lcghash_init) return TZ_ENOWEIGHTS;
    k_lcghash<<<(count + WPB - 1) / WPB, 32 * WPB, 0, h->stream>>>(states, nullptr, count, h->d.n, h->d.half_komi,
                                                                   s->lcghash_init, out_dev);
    return cudaGetLastError() == cudaSuccess ? TZ_OK : TZ_ECUDA;
}

int nn_simhash_indices(tz_handle* h, const TzState* states, int count, uint32_t* out_dev) {
    NnState* s = h->nn;
    if (!s || !s->simhash_matrix) return TZ_ENOWEIGHTS;
    k_simhash<<<(count + WPB - 1) / WPB, 32 * WPB, 0, h->stream>>>(states, nullptr, count, h->d.n, h->d.half_komi,
                                                                   s->simhash_matrix, out_dev);
    return cudaGetLastError() == cudaSuccess ? TZ_OK : TZ_ECUDA;
}
