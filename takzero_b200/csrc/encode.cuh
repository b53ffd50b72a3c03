// encode.cuh -- per-square input planes and the value / UBE heads of the reference network, as device functions
// shared by the stand-alone kernels (nn.cu), the first layer's A producer of the fused network launch
// (conv_tcgen05.cuh) and k_expand (kernels.cu).
//
//   game_repr      takzero/src/network/repr.rs:169-228
//   value / ube    takzero/src/network/net6_simhash.rs:88-119,194-201, uncertainty combine :309-317
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace enc {

// two 16-bit values of the network's storage type (bf16 or IEEE fp16) from two floats
__device__ __forceinline__ uint32_t pack16(float a, float b, int f16) {
    if (f16) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// What a position contributes to every one of its squares: the six scalar planes (repr.rs:199-227), packed as three
// pairs of 16-bit values of the network's storage type.  Computed once per queued position (k_select / k_prepare_eval)
// and kept in the queued copy of the state: pad1[0] = white - black top flats, bytes 372..383 = these three words.
struct PositionScalars {
    uint32_t s01;    // my stones / initial, my caps / initial
    uint32_t s23;    // opponent's
    uint32_t s45;    // 1.0 if Black to move, (white flats - black flats - komi / 2) / N^2
};

__device__ __forceinline__ PositionScalars position_scalars(int me, const uint8_t stones[2], const uint8_t caps[2],
                                                            int flat_diff, int n, int half_komi, int f16) {
    PositionScalars p;
    const int s0 = n == 3 ? 10 : n == 4 ? 15 : n == 5 ? 21 : 30;
    const int c0 = n >= 5 ? 1 : 0;
    const int other = me ^ 1;
    const float my_stones = __fdiv_rn((float)stones[me], (float)s0);
    const float my_caps = c0 ? __fdiv_rn((float)caps[me], (float)c0) : 0.0f;
    const float op_stones = __fdiv_rn((float)stones[other], (float)s0);
    const float op_caps = c0 ? __fdiv_rn((float)caps[other], (float)c0) : 0.0f;
    const float fcd = __fsub_rn((float)flat_diff, __fdiv_rn((float)half_komi, 2.0f));
    const float fcd_sq = __fdiv_rn(fcd, (float)(n * n));
    p.s01 = pack16(my_stones, my_caps, f16);
    p.s23 = pack16(op_stones, op_caps, f16);
    p.s45 = pack16(me == 1 ? 1.0f : 0.0f, fcd_sq, f16);
    return p;
}

// lane 0 of the warp that owns `st` (shared memory): derive and store the per-position words described above
__device__ __forceinline__ void store_position_scalars(TzState* st, int flat_diff, int n, int half_komi, int f16) {
    const PositionScalars ps = position_scalars(st->to_move, st->stones, st->caps, flat_diff, n, half_komi, f16);
    st->pad1[0] = (uint8_t)(int8_t)flat_diff;
    uint32_t* w = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(st) + 372);
    w[0] = ps.s01;
    w[1] = ps.s23;
    w[2] = ps.s45;
}

// The 0 / 1 indicator channels of one square as a bit mask: channels [0, 2*ss), ss = 2N + 3: flat, wall, cap on top,
// then the colours of the (up to 2N) pieces below the top; mine (colour `me`) first, then the opponent's.
__device__ __forceinline__ uint64_t square_indicator_bits(uint64_t stack, int h, int top, int me, int n) {
    if (h <= 0) return 0;
    const int ss = 2 * n + 3;
    const int top_col = (int)((stack >> (h - 1)) & 1ull);
    uint64_t ones = 1ull << (top + (top_col != me ? ss : 0));
    const int below = h - 1 < ss - 3 ? h - 1 : ss - 3;
    if (below > 0) {
        // bit i of r = colour of the piece i + 1 below the top
        const uint64_t r = __brevll(stack << (65 - h)) & ((1ull << below) - 1);
        const uint64_t all = (1ull << below) - 1;
        const uint64_t black = r, white = ~r & all;
        ones |= (me ? black : white) << 3;
        ones |= (me ? white : black) << (3 + ss);
    }
    return ones;
}

// The 64 input channels (C real ones, the rest zero) of one square as eight uint4 = 8 x 8 16-bit values, the pieces
// the chunk-planar activation layout stores (conv_tcgen05.cuh); the scalar planes occupy the pairs ss, ss + 1, ss + 2.
__device__ __forceinline__ void encode_square16(uint64_t stack, int h, int top, int me, const PositionScalars& ps, int n,
                                                int f16, uint4 out[8]) {
    const uint64_t ones = square_indicator_bits(stack, h, top, me, n);
    const uint32_t one = f16 ? 0x3c00u : 0x3f80u;
    const int sp = 2 * n + 3;
    uint32_t w[32];
#pragma unroll
    for (int p = 0; p < 32; p++) {
        const uint32_t b = (uint32_t)(ones >> (2 * p)) & 3u;
        uint32_t v = (b & 1u) * one + (b >> 1) * (one << 16);
        v = p == sp ? ps.s01 : v;
        v = p == sp + 1 ? ps.s23 : v;
        v = p == sp + 2 ? ps.s45 : v;
        w[p] = v;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) out[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
}

// The same straight into the canonical UMMA tile of the first convolution: `row` points at this square's 16 bytes of
// chunk plane 0, consecutive chunk planes are `plane_bytes` apart.  Indicator words first (8 x 16 B), then the three
// scalar words over their (zero) places.
__device__ __forceinline__ void encode_square16_smem(uint8_t* row, int plane_bytes, uint64_t ones, uint32_t s01,
                                                     uint32_t s23, uint32_t s45, int n, int f16) {
    const uint32_t one = f16 ? 0x3c00u : 0x3f80u;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t b = (uint32_t)(ones >> (8 * j + 2 * k)) & 3u;
            w[k] = (b & 1u) * one + (b >> 1) * (one << 16);
        }
        *reinterpret_cast<uint4*>(row + j * plane_bytes) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    const int sp = 2 * n + 3;
    *reinterpret_cast<uint32_t*>(row + (sp >> 2) * plane_bytes + (sp & 3) * 4) = s01;
    *reinterpret_cast<uint32_t*>(row + ((sp + 1) >> 2) * plane_bytes + ((sp + 1) & 3) * 4) = s23;
    *reinterpret_cast<uint32_t*>(row + ((sp + 2) >> 2) * plane_bytes + ((sp + 2) & 3) * 4) = s45;
}

// Value head (conv1x1 + ReLU + Linear + tanh) and uncertainty = clamp(max(exp(ube), local), 0, 4) of queue slot q
// from the two per-row features the epilogue of the last tower convolution wrote (the 1x1 convolutions), with
// local = 0 when the position's novelty-hash bit is in the set, else MAXIMUM_VARIANCE = 4.0
// (net6_simhash.rs:243-256), or the RND estimate of the 5x5 network.  Warp-convergent; every lane returns the same values.
// head_misc: [2] conv biases, [2][36] linear weights, [2] linear biases.
__device__ __forceinline__ void warp_heads(const float* head_feat, const float* head_misc, const uint32_t* novelty_set,
                                           const uint32_t* novelty_idx, const float* rnd_unc, int q, int nn, int lane,
                                           float* value, float* variance) {
    const float bv = head_misc[0], bu = head_misc[1];
    const float* lin_v = head_misc + 2;
    const float* lin_u = head_misc + 2 + 36;
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
    *value = tanhf(acc_v + head_misc[2 + 72]);
    const float ube = acc_u + head_misc[2 + 73];
    float local = 4.0f;
    if (rnd_unc) {  // net5: normalized_rnd of the position (net5.rs:271-277), rnd.cu
        local = rnd_unc[q];
    } else if (novelty_set) {
        const uint32_t idx = novelty_idx[q];
        if ((novelty_set[idx >> 5] >> (idx & 31)) & 1u) local = 0.0f;
    }
    *variance = fminf(fmaxf(fmaxf(expf(ube), local), 0.0f), 4.0f);
}

// network/repr.rs:49-71 `move_index` split into (channel, square): the channel of a move
__device__ __forceinline__ int move_channel(int n, uint16_t m) {
    const int kind = (m >> 6) & 3, pat = m >> 8;
    if (pat == 0) return kind;  // flat 0, wall 1, cap 2
    const int dir_off = kind == 0 ? 0 : kind == 1 ? 2 : kind == 2 ? 3 : 1;  // Up, Right, Down, Left order of repr.rs:61-66
    return 3 + ((pat >> (8 - n)) - 1) + ((1 << n) - 2) * dir_off;
}

}  // namespace enc
