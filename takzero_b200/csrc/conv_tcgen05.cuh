// conv_tcgen05.cuh -- 3x3 convolution of the ResNet tower as a bf16 implicit GEMM on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), sm_100a only.
//
// Replaces libtorch's conv2d + batch_norm2d (+ add + relu) of the reference tower
// (takzero/src/network/residual.rs:13-63, net6_simhash.rs:43-86); BN is folded into the
// weights / bias on the host (nn.cu), inference mode like `forward_t(xs, false)`.
//
// Data layout: activations are channels-last bf16 [rows][C] with DENSE rows, row = position *
// N*N + square (no padding anywhere).  A 3x3 tap (dy,dx) of output row r is input row
// r + dy*N + dx, so the convolution is the GEMM
//   out[r, co] = sum_{tap, ci} act[r + off(tap), ci] * W[tap][co][ci]
// whose A operand for each tap is the SAME shared-memory tile read at a shifted row: the tile is
// stored K-chunk-major ([C/8][rows][8 ch], row pitch 16 B) in the no-swizzle canonical UMMA
// layout, so a row shift is a +16 B/row change of the descriptor's start address.  Rows whose
// neighbour (dy,dx) lies off the board (and would wrap to another board row / board) are
// excluded with the `disable-output-lane` mask of tcgen05.mma: for that tap their accumulator
// lanes are simply not updated, which is exactly "add zero padding".  The masks depend only on
// (first row of the tile) mod N*N and come from a small host-built table.  No tensor-core work
// is spent on padding.
//
// One CTA = 128 output rows x 256 output channels (one UMMA M128 N256 accumulator of 256
// TMEM columns, double buffered), persistent over row tiles.  Warp roles:
//   warps 0-3 epilogue    tcgen05.ld -> bias (+ residual) (+ ReLU) -> bf16 / f32 stores
//   warp 4   A producer   cp.async 16 B pieces of the 144-row halo tile, one 64-channel block per stage
//   warp 5   B producer   cp.async.bulk of pre-arranged 32 KB weight blocks (64 K x 256 N) from L2
//   warp 6   MMA issuer   one thread: 9 taps x 4 K-steps of tcgen05.mma per A block
//   warp 7   TMEM allocator
// This file is the ONE-CTA kernel (kept for A/B and as the base of the shared helpers); the product path is
// the CTA-pair kernel in conv_pair_tcgen05.cuh.  TZ_DEBUG_* macros are compile-time tuning experiments
// (tools/build_variant.sh); the numbers they produced are in profiles/r1_conv_timing.txt.
#pragma once
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdint.h>

namespace conv {

constexpr int TILE_M = 128;
constexpr int HALO = 8;                       // >= N+1 for N <= 6
constexpr int A_ROWS = TILE_M + 2 * HALO;     // 144
constexpr int A_KC_PITCH = (A_ROWS + 1) * 16; // 2320 B: +1 row keeps cp.async writes bank-conflict free
constexpr int A_STAGE_BYTES = 8 * A_KC_PITCH; // 18560 B = one 64-channel block
#ifndef TZ_DEBUG_SKIP_RES
#define TZ_DEBUG_SKIP_RES 0
#endif
#ifndef TZ_DEBUG_SKIP_STORE
#define TZ_DEBUG_SKIP_STORE 0
#endif
#ifndef TZ_A_STAGES
#define TZ_A_STAGES 4
#endif
#ifndef TZ_B_STAGES
#define TZ_B_STAGES 4
#endif
constexpr int A_STAGES = TZ_A_STAGES;
constexpr int B_KC_PITCH = 256 * 16;          // 4096 B
constexpr int B_STAGE_BYTES = 8 * B_KC_PITCH; // 32768 B = 64 K x 256 N
constexpr int B_STAGES = TZ_B_STAGES;
constexpr int N_OUT = 256;
constexpr int THREADS = 256;
// Warp roles.  The SM sub-partition arbiter favours the highest warp id, so the single-thread producer /
// MMA roles sit on warps 4-7 and the instruction-heavy epilogue on warps 0-3 (warp w may only touch TMEM
// lanes 32*(w%4)..+31, so any four warps with distinct w%4 can be the epilogue).
constexpr int W_EPI0 = 0, W_APROD = 4, W_BPROD = 5, W_MMA = 6, W_ALLOC = 7;
constexpr int MASK_BYTES = 36 * 9 * 16;         // [N*N][9 taps] 128-bit lane masks
constexpr int SMEM_BYTES =
    A_STAGES * A_STAGE_BYTES + B_STAGES * B_STAGE_BYTES + 1024 /*bias*/ + 256 /*barriers*/ + MASK_BYTES;

struct Params {
    const __nv_bfloat16* in;        // [rows][cin] activations, dense rows
    int cin;                        // channels of `in` (multiple of 64)
    const __nv_bfloat16* w;         // [cin/64][9][8][256][8] pre-arranged weight blocks
    const float* bias;              // [256]
    const __nv_bfloat16* residual;  // [rows][256] or null
    __nv_bfloat16* out_act;         // [rows][256] or null
    float* out_f32;                 // [positions * n*n][256] (no guard rows), or null
    int relu;
    const int* count_ptr;           // number of positions (device), or null: use count_max
    int count_max;
    int n;                          // board size
    int guard;                      // leading guard rows of the buffers (= HALO)
    const uint4* masks;             // [n*n][9] disable-output-lane masks by (first tile row) mod n*n
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
#ifdef TZ_DEBUG_TIMING
#define TWAIT(acc, stmt)                 \
    do {                                 \
        const long long t0_ = clock64(); \
        stmt;                            \
        acc += clock64() - t0_;          \
    } while (0)
#else
#define TWAIT(acc, stmt) stmt
#endif
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// `mask`: disable-output-lane, bit i of the 128-bit vector = do not update TMEM lane (= tile row) i
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc,
                                       const uint4& mask) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(mask.x), "r"(mask.y), "r"(mask.z), "r"(mask.w)
        : "memory");
}
// K-major, no swizzle: 8x8 core matrices of 128 contiguous bytes; LBO = byte distance of the
// two K-chunks of one MMA, SBO = byte distance of consecutive 8-row groups (128: rows packed)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(THREADS, 1) k_conv3x3(const Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* a_smem = smem;
    uint8_t* b_smem = smem + A_STAGES * A_STAGE_BYTES;
    float* s_bias = reinterpret_cast<float*>(b_smem + B_STAGES * B_STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_bias) + 1024);
    // barrier map: a_full[6] a_empty[6] b_full[3] b_empty[3] t_full[2] t_empty[2], then the TMEM base
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t a_full = bar0, a_empty = bar0 + 8 * A_STAGES;
    const uint32_t b_full = a_empty + 8 * A_STAGES, b_empty = b_full + 8 * B_STAGES;
    const uint32_t t_full = b_empty + 8 * B_STAGES, t_empty = t_full + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * A_STAGES + 2 * B_STAGES + 4);
    uint4* s_masks = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(bars) + 256);

    const int count = p.count_ptr ? *p.count_ptr : p.count_max;
    const int nn = p.n * p.n;
    const int rows_used = count * nn;
    const int tiles = (rows_used + TILE_M - 1) / TILE_M;
    const int kblocks = p.cin >> 6;

    if (threadIdx.x == 0) {
        for (int i = 0; i < A_STAGES; i++) {
            mbar_init(a_full + 8 * i, 1);
            mbar_init(a_empty + 8 * i, 1);
        }
        for (int i = 0; i < B_STAGES; i++) {
            mbar_init(b_full + 8 * i, 1);
            mbar_init(b_empty + 8 * i, 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(t_full + 8 * i, 1);
            mbar_init(t_empty + 8 * i, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_ALLOC) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < N_OUT; i += THREADS) s_bias[i] = p.bias[i];
    for (int i = threadIdx.x; i < nn * 9; i += THREADS) s_masks[i] = p.masks[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == W_APROD) {
        // ---- A producer: halo tile rows [guard + 128 t - 8, +144), one 64-channel block per stage
        int stage = 0, phase = 0, pending = -1;
        const size_t row_bytes = (size_t)p.cin * 2;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            const uint8_t* src_tile =
                reinterpret_cast<const uint8_t*>(p.in) + (size_t)(p.guard + t * TILE_M - HALO) * row_bytes;
            for (int kb = 0; kb < kblocks; kb++) {
                mbar_wait(a_empty + 8 * stage, phase ^ 1);
                const uint32_t dst = smem_u32(a_smem + stage * A_STAGE_BYTES);
                const uint8_t* src = src_tile + kb * 128;
#ifndef TZ_DEBUG_NO_A_LOAD  // tuning experiment
#pragma unroll 4
                for (int it = 0; it < A_ROWS * 8 / 32; it++) {
                    const int piece = it * 32 + lane;
                    const int row = piece >> 3, kc = piece & 7;
                    cp_async16(dst + kc * A_KC_PITCH + row * 16, src + (size_t)row * row_bytes + kc * 16);
                }
#endif
                cp_async_commit();
                if (pending >= 0) {
                    // the previous block has landed: publish it to the async proxy (tcgen05 reads)
                    cp_async_wait<1>();
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(a_full + 8 * pending);
                }
                pending = stage;
                if (++stage == A_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
        if (pending >= 0) {
            cp_async_wait<0>();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full + 8 * pending);
        }
    } else if (warp == W_BPROD) {
        // ---- B producer: weight blocks stream from L2 in exactly the order the MMA consumes them
        if (lane == 0) {
            int stage = 0, phase = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                const uint8_t* src = reinterpret_cast<const uint8_t*>(p.w);
                for (int blk = 0; blk < kblocks * 9; blk++) {
                    mbar_wait(b_empty + 8 * stage, phase ^ 1);
#ifdef TZ_DEBUG_NO_B_LOAD  // tuning experiment: no weight traffic (results are garbage)
                    mbar_arrive(b_full + 8 * stage);
#else
                    mbar_arrive_expect_tx(b_full + 8 * stage, B_STAGE_BYTES);
                    bulk_g2s(smem_u32(b_smem + stage * B_STAGE_BYTES), src + (size_t)blk * B_STAGE_BYTES, B_STAGE_BYTES,
                             b_full + 8 * stage);
#endif
                    if (++stage == B_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == W_MMA) {
        // ---- MMA issuer
        if (lane == 0) {
            // kind::f16: D = f32, A = B = bf16, both K-major, N = 256, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N_OUT >> 3) << 17) |
                                   ((uint32_t)(TILE_M >> 4) << 24);
            int a_stage = 0, a_phase = 0, b_stage = 0, b_phase = 0, it = 0;
#ifdef TZ_DEBUG_TIMING
            long long w_t = 0, w_a = 0, w_b = 0;
            const long long mma_start = clock64();
#endif
            for (int t = blockIdx.x; t < tiles; t += gridDim.x, it++) {
                const int acc = it & 1;
                TWAIT(w_t, mbar_wait(t_empty + 8 * acc, ((it >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * N_OUT;
                const uint4* tile_masks = s_masks + ((size_t)t * TILE_M % nn) * 9;
                for (int kb = 0; kb < kblocks; kb++) {
                    TWAIT(w_a, mbar_wait(a_full + 8 * a_stage, a_phase));
                    const uint32_t a_base = smem_u32(a_smem + a_stage * A_STAGE_BYTES);
                    // tap order: the centre tap first (no mask, it initialises every lane), then the rest
                    for (int ti = 0; ti < 9; ti++) {
                        const int tap = ti == 0 ? 4 : (ti <= 4 ? ti - 1 : ti);
                        TWAIT(w_b, mbar_wait(b_full + 8 * b_stage, b_phase));
                        tc_fence_after();
                        const int off = (tap / 3 - 1) * p.n + (tap % 3 - 1);
#ifdef TZ_DEBUG_NO_MASK
                        const uint4 mask = make_uint4(0, 0, 0, 0);
#else
                        const uint4 mask = tile_masks[tap];
#endif
                        const uint32_t a_tap = a_base + (HALO + off) * 16;
                        const uint32_t b_base = smem_u32(b_smem + b_stage * B_STAGE_BYTES);
#pragma unroll
                        for (int ks = 0; ks < 4; ks++) {
                            const uint64_t adesc = make_desc(a_tap + ks * 2 * A_KC_PITCH, A_KC_PITCH);
                            const uint64_t bdesc = make_desc(b_base + ks * 2 * B_KC_PITCH, B_KC_PITCH);
#ifndef TZ_DEBUG_NO_MMA  // tuning experiment: data movement only
                            tc_mma(tmem_d, adesc, bdesc, idesc, (kb | ti | ks) != 0, mask);
#endif
                        }
                        tc_commit(b_empty + 8 * b_stage);
                        if (++b_stage == B_STAGES) {
                            b_stage = 0;
                            b_phase ^= 1;
                        }
                    }
                    tc_commit(a_empty + 8 * a_stage);
                    if (++a_stage == A_STAGES) {
                        a_stage = 0;
                        a_phase ^= 1;
                    }
                }
                tc_commit(t_full + 8 * acc);
            }
#ifdef TZ_DEBUG_TIMING
            if (blockIdx.x == 0 || blockIdx.x == 77)
                printf("cta %d mma: tiles %d total %lld wait_tmem %lld wait_a %lld wait_b %lld\n", blockIdx.x, it,
                       clock64() - mma_start, w_t, w_a, w_b);
#endif
        }
    } else if (warp >= W_EPI0 && warp < W_EPI0 + 4) {
        // ---- epilogue: TMEM lane = tile row; warp w reads lanes 32*(w%4)..+31.
        // Each 32-column chunk goes TMEM -> registers (+bias) -> a per-warp f32 staging tile in shared
        // memory -> coalesced copy-out (4 lanes per row for bf16, 8 for f32), where the residual is
        // loaded with the same coalesced pattern, added in f32, and ReLU / rounding are applied.  A warp
        // store then touches 8 (or 4) 128-B lines instead of 32.
        const int wq = warp & 3;
        int it = 0;
#ifdef TZ_DEBUG_TIMING
        long long e_wait = 0, e_busy = 0;
#endif
        for (int t = blockIdx.x; t < tiles; t += gridDim.x, it++) {
            const int acc = it & 1;
            const int tr = wq * 32 + lane;
            const int rel = t * TILE_M + tr;  // row = position * n*n + square
            const bool valid = rel < rows_used;
            const size_t grow = (size_t)(p.guard + rel);
            const size_t crow = (size_t)rel;
            TWAIT(e_wait, mbar_wait(t_full + 8 * acc, (it >> 1) & 1));
            tc_fence_after();
#ifdef TZ_DEBUG_TIMING
            const long long e0 = clock64();
#endif
            const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + acc * N_OUT;
#ifndef TZ_DEBUG_NO_EPILOGUE
#pragma unroll 1
            for (int c0 = 0; c0 < N_OUT; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
                tmem_ld_wait();
                if (valid) {
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) f[j] = __uint_as_float(v[j]) + s_bias[c0 + j];
                    if (p.residual) {
                        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + grow * N_OUT + c0);
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const uint4 r = rp[j];
                            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                f[j * 8 + e * 2] += __uint_as_float(w[e] << 16);
                                f[j * 8 + e * 2 + 1] += __uint_as_float(w[e] & 0xffff0000u);
                            }
                        }
                    }
                    if (p.relu) {
#pragma unroll
                        for (int j = 0; j < 32; j++) f[j] = fmaxf(f[j], 0.0f);
                    }
                    if (p.out_act) {
                        uint4* op = reinterpret_cast<uint4*>(p.out_act + grow * N_OUT + c0);
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            op[j] = make_uint4(pack_bf16(f[j * 8], f[j * 8 + 1]), pack_bf16(f[j * 8 + 2], f[j * 8 + 3]),
                                               pack_bf16(f[j * 8 + 4], f[j * 8 + 5]), pack_bf16(f[j * 8 + 6], f[j * 8 + 7]));
                    }
                    if (p.out_f32) {
                        float4* op = reinterpret_cast<float4*>(p.out_f32 + crow * N_OUT + c0);
#pragma unroll
                        for (int j = 0; j < 8; j++) op[j] = make_float4(f[j * 4], f[j * 4 + 1], f[j * 4 + 2], f[j * 4 + 3]);
                    }
                }
            }
#endif
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty + 8 * acc);
#ifdef TZ_DEBUG_TIMING
            e_busy += clock64() - e0;
#endif
        }
#ifdef TZ_DEBUG_TIMING
        if ((blockIdx.x == 0 || blockIdx.x == 77) && lane == 0 && wq == 1)
            printf("cta %d epi warp %d: tiles %d wait_full %lld busy %lld\n", blockIdx.x, warp, it, e_wait, e_busy);
#endif
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_ALLOC) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

}  // namespace conv
