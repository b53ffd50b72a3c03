// conv_pair_tcgen05.cuh -- the tower convolution of conv_tcgen05.cuh on CTA PAIRS (tcgen05
// cta_group::2): two SMs of one TPC compute one 256-row x 256-channel tile (UMMA M256 N256 K16).
//
// Why: in the one-CTA kernel every MMA pulls 4 KB of A and 8 KB of B through the SM's shared
// memory read path, and the measured MMA cadence is ~160 cycles instead of 128 (shared-memory
// operand fetch bound, profiles/r1_conv_timing.txt).  In pair mode each CTA holds its own 128
// activation rows (A) but only HALF of every weight block (128 of the 256 N rows), so per CTA an MMA
// reads 4 KB + 4 KB and only 16 KB per weight block is copied into each SM.
//
// Protocol (rank 0 = leader issues all MMAs; every barrier lives at the same offset in both CTAs):
//   a_full[s]   leader, count 2   each CTA's A producer arrives (remote for rank 1) when its block landed
//   b_land[s]   local,  count 1+tx  the CTA's own cp.async.bulk of its weight half
//   b_full[s]   leader, count 2   each CTA's relay thread forwards b_land -> leader
//   a_empty/b_empty/t_full[s]  local, count 1, arrived on BOTH CTAs by tcgen05.commit ... multicast
//   t_empty[acc] leader, count 8  epilogue warps of both CTAs
// Data layout, row-shift taps and lane-mask padding are exactly those of conv_tcgen05.cuh.
#pragma once
#include "conv_tcgen05.cuh"

namespace conv {

constexpr int P_A_STAGES = 4;
constexpr int P_B_HALF_PITCH = 128 * 16;               // 2048 B between K-chunks of a weight half
constexpr int P_B_STAGE_BYTES = 8 * P_B_HALF_PITCH;    // 16384 B = 64 K x 128 N
constexpr int P_B_STAGES = 8;
constexpr int P_SMEM_BYTES =
    P_A_STAGES * A_STAGE_BYTES + P_B_STAGES * P_B_STAGE_BYTES + 1024 /*bias*/ + 512 /*barriers*/ + MASK_BYTES;

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `local` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // default semantics (release, cta scope) like CUTLASS ClusterBarrier::arrive: the cluster-scope form makes
    // ptxas emit MEMBAR.ALL.GPU + CGAERRBAR per arrive, which serialises the whole pipeline
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// waits on barriers that the other CTA arrives on use the plain try_wait too (an acquire.cluster wait costs a
// CCTL.IVALL = L1 invalidate per wait); the protected data is read by the tensor core, not by this thread
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"((uint16_t)3)
        : "memory");
}
// lane masks: m0 = rows of rank 0 (TMEM lanes of the leader), m1 = rows of rank 1
__device__ __forceinline__ void tc_mma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc,
                                            const uint4& m0, const uint4& m1) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8, %9, %10, %11, %12}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(m0.x), "r"(m0.y), "r"(m0.z), "r"(m0.w), "r"(m1.x), "r"(m1.y),
        "r"(m1.z), "r"(m1.w)
        : "memory");
}

// Params.w here points at weight blocks stored as [cin/64][9][2 halves][8][128][8] (nn.cu `upload_conv`)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) k_conv3x3_pair(const Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    uint8_t* a_smem = smem;
    uint8_t* b_smem = smem + P_A_STAGES * A_STAGE_BYTES;
    float* s_bias = reinterpret_cast<float*>(b_smem + P_B_STAGES * P_B_STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_bias) + 1024);
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t a_full = bar0, a_empty = a_full + 8 * P_A_STAGES;
    const uint32_t b_land = a_empty + 8 * P_A_STAGES, b_full = b_land + 8 * P_B_STAGES;
    const uint32_t b_empty = b_full + 8 * P_B_STAGES;
    const uint32_t t_full = b_empty + 8 * P_B_STAGES, t_empty = t_full + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * P_A_STAGES + 3 * P_B_STAGES + 4);
    uint4* s_masks = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(bars) + 512);

    const int count = p.count_ptr ? *p.count_ptr : p.count_max;
    const int nn = p.n * p.n;
    const int rows_used = count * nn;
    const int pair_tiles = (rows_used + 2 * TILE_M - 1) / (2 * TILE_M);
    const int kblocks = p.cin >> 6;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < P_A_STAGES; i++) {
            mbar_init(a_full + 8 * i, 2);
            mbar_init(a_empty + 8 * i, 1);
        }
        for (int i = 0; i < P_B_STAGES; i++) {
            mbar_init(b_land + 8 * i, 1);
            mbar_init(b_full + 8 * i, 2);
            mbar_init(b_empty + 8 * i, 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(t_full + 8 * i, 1);
            mbar_init(t_empty + 8 * i, 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_ALLOC) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < N_OUT; i += THREADS) s_bias[i] = p.bias[i];
    for (int i = threadIdx.x; i < nn * 9; i += THREADS) s_masks[i] = p.masks[i];
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // both CTAs' barriers are initialised before anyone arrives remotely
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == W_APROD) {
        // ---- A producer (both CTAs): this CTA's 144-row halo tile, published on the LEADER's a_full
        int stage = 0, phase = 0, pending = -1;
        const size_t row_bytes = (size_t)p.cin * 2;
        for (int pt = pair; pt < pair_tiles; pt += npairs) {
            const int t = pt * 2 + (int)rank;
            const uint8_t* src_tile =
                reinterpret_cast<const uint8_t*>(p.in) + (size_t)(p.guard + t * TILE_M - HALO) * row_bytes;
            for (int kb = 0; kb < kblocks; kb++) {
                mbar_wait(a_empty + 8 * stage, phase ^ 1);
                const uint32_t dst = smem_u32(a_smem + stage * A_STAGE_BYTES);
                const uint8_t* src = src_tile + kb * 128;
#pragma unroll 4
                for (int it = 0; it < A_ROWS * 8 / 32; it++) {
                    const int piece = it * 32 + lane;
                    const int row = piece >> 3, kc = piece & 7;
                    cp_async16(dst + kc * A_KC_PITCH + row * 16, src + (size_t)row * row_bytes + kc * 16);
                }
                cp_async_commit();
                if (pending >= 0) {
                    cp_async_wait<1>();
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(map_to_rank(a_full + 8 * pending, 0));
                }
                pending = stage;
                if (++stage == P_A_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
        if (pending >= 0) {
            cp_async_wait<0>();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(map_to_rank(a_full + 8 * pending, 0));
        }
    } else if (warp == W_BPROD) {
        // ---- B producer (both CTAs): this CTA's half (128 of the 256 N rows) of every weight block
        if (lane == 0) {
            int stage = 0, phase = 0;
            for (int pt = pair; pt < pair_tiles; pt += npairs) {
                const uint8_t* src = reinterpret_cast<const uint8_t*>(p.w) + (size_t)rank * P_B_STAGE_BYTES;
                for (int blk = 0; blk < kblocks * 9; blk++) {
                    mbar_wait(b_empty + 8 * stage, phase ^ 1);
                    mbar_arrive_expect_tx(b_land + 8 * stage, P_B_STAGE_BYTES);
                    bulk_g2s(smem_u32(b_smem + stage * P_B_STAGE_BYTES), src + (size_t)blk * 2 * P_B_STAGE_BYTES,
                             P_B_STAGE_BYTES, b_land + 8 * stage);
                    if (++stage == P_B_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == W_ALLOC) {
        // ---- relay (both CTAs): "my weight half has landed" -> the leader's b_full
        if (lane == 0) {
            int stage = 0, phase = 0;
            for (int pt = pair; pt < pair_tiles; pt += npairs)
                for (int blk = 0; blk < kblocks * 9; blk++) {
                    mbar_wait(b_land + 8 * stage, phase);
                    mbar_arrive_cluster(map_to_rank(b_full + 8 * stage, 0));
                    if (++stage == P_B_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
        }
    } else if (warp == W_MMA) {
        // ---- MMA issuer (leader only): M = 256 (128 rows per CTA), N = 256, K = 16
        if (rank == 0 && lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N_OUT >> 3) << 17) |
                                   ((uint32_t)((2 * TILE_M) >> 4) << 24);
            int a_stage = 0, a_phase = 0, b_stage = 0, b_phase = 0, it = 0;
#ifdef TZ_DEBUG_TIMING
            long long w_t = 0, w_a = 0, w_b = 0;
            const long long mma_start = clock64();
#endif
            for (int pt = pair; pt < pair_tiles; pt += npairs, it++) {
                const int acc = it & 1;
                TWAIT(w_t, mbar_wait_cluster(t_empty + 8 * acc, ((it >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * N_OUT;
                const uint4* masks0 = s_masks + ((size_t)(pt * 2) * TILE_M % nn) * 9;
                const uint4* masks1 = s_masks + ((size_t)(pt * 2 + 1) * TILE_M % nn) * 9;
                for (int kb = 0; kb < kblocks; kb++) {
                    TWAIT(w_a, mbar_wait_cluster(a_full + 8 * a_stage, a_phase));
                    const uint32_t a_base = smem_u32(a_smem + a_stage * A_STAGE_BYTES);
                    for (int ti = 0; ti < 9; ti++) {
                        const int tap = ti == 0 ? 4 : (ti <= 4 ? ti - 1 : ti);
                        TWAIT(w_b, mbar_wait_cluster(b_full + 8 * b_stage, b_phase));
                        tc_fence_after();
                        const int off = (tap / 3 - 1) * p.n + (tap % 3 - 1);
                        const uint4 m0 = masks0[tap], m1 = masks1[tap];
                        const uint32_t a_tap = a_base + (HALO + off) * 16;
                        const uint32_t b_base = smem_u32(b_smem + b_stage * P_B_STAGE_BYTES);
#pragma unroll
                        for (int ks = 0; ks < 4; ks++) {
                            const uint64_t adesc = make_desc(a_tap + ks * 2 * A_KC_PITCH, A_KC_PITCH);
                            const uint64_t bdesc = make_desc(b_base + ks * 2 * P_B_HALF_PITCH, P_B_HALF_PITCH);
                            tc_mma_pair(tmem_d, adesc, bdesc, idesc, (kb | ti | ks) != 0, m0, m1);
                        }
                        tc_commit_pair(b_empty + 8 * b_stage);
                        if (++b_stage == P_B_STAGES) {
                            b_stage = 0;
                            b_phase ^= 1;
                        }
                    }
                    tc_commit_pair(a_empty + 8 * a_stage);
                    if (++a_stage == P_A_STAGES) {
                        a_stage = 0;
                        a_phase ^= 1;
                    }
                }
                tc_commit_pair(t_full + 8 * acc);
            }
#ifdef TZ_DEBUG_TIMING
            if (pair == 0 || pair == 40)
                printf("pair %d mma: tiles %d total %lld wait_tmem %lld wait_a %lld wait_b %lld\n", pair, it,
                       clock64() - mma_start, w_t, w_a, w_b);
#endif
        }
    } else if (warp >= W_EPI0 && warp < W_EPI0 + 4) {
        // ---- epilogue (both CTAs): own 128 rows from own TMEM; t_empty lives in the leader
        const int wq = warp & 3;
        const uint32_t t_empty_leader0 = map_to_rank(t_empty, 0);
        int it = 0;
        for (int pt = pair; pt < pair_tiles; pt += npairs, it++) {
            const int acc = it & 1;
            const int t = pt * 2 + (int)rank;
            const int tr = wq * 32 + lane;
            const int rel = t * TILE_M + tr;
            const bool valid = rel < rows_used;
            const size_t grow = (size_t)(p.guard + rel);
            const size_t crow = (size_t)rel;
            mbar_wait(t_full + 8 * acc, (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + acc * N_OUT;
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
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(t_empty_leader0 + 8 * acc);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // nobody leaves (or frees TMEM) while the pair's MMAs may still touch its memory
    if (warp == W_ALLOC) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

}  // namespace conv
