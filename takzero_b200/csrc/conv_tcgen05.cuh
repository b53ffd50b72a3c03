// conv_tcgen05.cuh -- 3x3 convolution of the ResNet tower as a bf16 (optionally fp16) implicit GEMM on the
// 5th-generation tensor cores (tcgen05.mma cta_group::2, accumulators in TMEM), sm_100a only.
//
// Replaces libtorch's conv2d + batch_norm2d (+ add + relu) of the reference tower
// (takzero/src/network/residual.rs:13-63, net6_simhash.rs:43-86); BN is folded into the
// weights / bias on the host (nn.cu), inference mode like `forward_t(xs, false)`.
//
// Data layout ("chunk-planar"): activations are bf16 [C/8][rows][8]: one plane per 8-channel chunk, rows
// dense (row = guard + position * N*N + square, no padding anywhere; small 5x5 / 6x6 batches use PACKED rows instead,
// whole positions per 128-row CTA tile, see Params::pack).  A 3x3 tap (dy,dx) of output row r is
// input row r + dy*N + dx, so the convolution is the GEMM
//   out[r, co] = sum_{tap, ci} act[r + off(tap), ci] * W[tap][co][ci]
// * The A tile of a 64-channel block is 8 contiguous 2304-byte runs of global memory (144 halo rows x 16 B
//   per chunk plane) and lands with 8 cp.async.bulk copies directly in the no-swizzle K-major canonical UMMA
//   layout ([8 chunks][rows][8 ch], row pitch 16 B).  The operand of tap (dy,dx) is that SAME tile with the
//   descriptor start address moved by (dy*N+dx)*16 B: the tile is loaded once and used by all 9 taps.
// * Rows whose neighbour (dy,dx) lies off the board are excluded with the `disable-output-lane` mask of
//   tcgen05.mma (their accumulator lanes are not updated for that tap = zero padding).  The masks depend only
//   on (first row of the tile) mod N*N and come from a small host-built table.  No tensor-core work is spent
//   on padding.
// * In the epilogue thread = row, so for one chunk plane the 32 lanes of a warp write 32 consecutive 16-byte
//   pieces: every store / residual load of a warp is one contiguous 512-byte run.
//
// CTA pairs (cluster of 2): two SMs compute one 256-row x 256-channel tile (UMMA M256 N256 K16); each CTA
// holds its own 128 activation rows but only HALF of every weight block (128 of the 256 N rows), so per CTA
// an MMA reads 4 KB + 4 KB of shared memory and 16 KB per weight block is copied into each SM.
// Protocol (rank 0 = leader issues all MMAs; every barrier lives at the same offset in both CTAs):
//   a_full[s] / b_full[s]   leader, count 2: the leader's own copies complete_tx on them directly; the peer's
//                           copies land on its local a_land / b_land and a relay thread forwards the arrival
//   a_empty / b_empty / t_full   local, count 1, arrived on BOTH CTAs by tcgen05.commit ... multicast
//   t_empty[acc]            leader, count 8: epilogue warps of both CTAs
// Warp roles per CTA: warps 0-3 epilogue (tcgen05.ld -> bias, residual, ReLU -> stores), warp 4 A producer,
// warp 5 B producer, warp 6 MMA issuer (leader) / A relay (peer), warp 7 TMEM allocator / B relay (peer).
//
// Layer chain: one launch runs a CHAIN of layers (Params::layers[0..n_layers)) -- the whole network body: input
// convolution, residual tower, policy convolution.  The work items (layer, pair tile) are dealt round-robin to the
// persistent CTA pairs, so a pair gets 15.57 tiles per layer on average instead of a whole number per launch, and
// the launch gaps, prologues and pipeline refills between layers disappear.  A 3x3 tap reaches at most HALO rows
// into the neighbouring tiles, so item (L, t) depends only on items (L-1, t-1..t+1): the epilogue warps publish
// their tile with release increments of progress[t][0..3] (per 64 output channels when a layer has fewer than
// two tiles per pair, so that the next layer's first MMAs overlap the rest of the epilogue; else once per tile),
// the A producer of a dependent item spins on an acquire load before it fetches a 64-channel block and crosses to
// the async proxy with fence.proxy.async before its bulk copies.  The same waits cover the write-after-read
// hazards of the two ping-pong activation buffers (a layer's output buffer is the input buffer of the layer
// before it).  The launch is cooperative: every pair must be resident, or the spin would deadlock.
//
// Chunks: the positions of a launch are cut into equal chunks of at least Params::chunk_min_tiles pair tiles and
// the items are numbered chunk-major, so a chunk runs through ALL its layers before the next one starts.
// Activations live in two small ping-pong sets (even / odd chunks) that stay resident in the 126 MB L2: the
// layer-to-layer traffic never reaches HBM, which under the power cap is worth ~5 % of throughput.  Only the input
// planes, the policy logits and the two head features per row (the value / UBE 1x1 convolutions, folded into
// the last tower layer's epilogue) use rows of all positions.  A chunk's first layer waits until the chunk two
// before it has completely finished (chunk_done), because it overwrites that chunk's activation set.
//
// First and last layer talk to the search directly (BASELINE north star (3) / (4)):
// * Layer::enc_states: the A producer warp of the input convolution builds its halo tile from the queued TzState
//   records (game_repr, network/repr.rs:169-228: 64 16-bit channels per square, written straight into the canonical
//   UMMA layout with st.shared, then fence.proxy.async + a plain mbarrier arrive) -- the input planes never exist in
//   global memory.
// * Layer::g_out: the epilogue of the policy convolution keeps only the logits of the legal moves
//   (net6_simhash.rs:277-306 `gather`): thread = (position, square) looks up the moves that start on its square
//   (TzDev::sq_ranges, contiguous in possible_moves order), stages its 32 channels of the current block in shared
//   memory and stores logit[move_channel] to the queue's logit row.  The [positions][9036] f32 policy tensor
//   (302 MB per pass of 8192 positions) is never written.
//
// Local chain (small batches of 4x4 boards): when N*N divides the 128 rows of a CTA, no position straddles two CTAs, so
// a tile needs nothing from its neighbours; when moreover a layer has no more tiles than there are CTA pairs, pair p
// keeps tile p for ALL layers and the activations never leave the SM: the epilogue writes the 16-bit outputs of a
// layer straight into the A stages of the next one (the canonical UMMA layout, st.shared + fence.proxy.async + mbarrier
// arrive) 64 channels at a time, so the next layer's first MMAs start ~2 us after the last MMA of this one instead of
// after a round trip through L2 with gpu-scope release / acquire flags (measured chain: 6.5 us per layer, which bounded
// 1024 games of 4x4 at 16.5 us per layer against 10 us of MMAs).  Only the residual stream x still goes through
// global memory, written and read back by the same thread.
// 5x5 / 6x6 boards get the same chain on PACKED rows (Params::pack): 5 / 3 whole positions per CTA tile, the tile's last
// 3 / 20 rows dead (masked for every tap, nothing stored), whenever every CTA has at most one such tile (<= 740 / 444
// positions on 148 SMs).  The schedule then counts 128-row units; row_position() is the only place that knows.
//
// TZ_DEBUG_TIMING is a compile-time tuning experiment (tools/build_variant.sh); the numbers it produced for
// the earlier row-major one-CTA / pair kernels are in profiles/r1_conv_timing.txt.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include "encode.cuh"

namespace conv {

constexpr int TILE_M = 128;                    // rows per CTA (a pair computes 256)
constexpr int HALO = 8;                        // >= N+1 for N <= 6
constexpr int A_ROWS = TILE_M + 2 * HALO;      // 144
constexpr int A_KC_BYTES = A_ROWS * 16;        // 2304 B: one chunk plane of the halo tile
constexpr int A_STAGE_BYTES = 8 * A_KC_BYTES;  // 18432 B = one 64-channel block
#ifndef TZ_A_STAGES
#define TZ_A_STAGES 4
#endif
#ifndef TZ_B_STAGES
#define TZ_B_STAGES 8
#endif
constexpr int A_STAGES = TZ_A_STAGES;
constexpr int B_KC_BYTES = 128 * 16;           // 2048 B between K-chunks of a weight half
constexpr int B_STAGE_BYTES = 8 * B_KC_BYTES;  // 16384 B = 64 K x 128 N (one CTA's half)
constexpr int B_STAGES = TZ_B_STAGES;
constexpr int N_OUT = 256;
constexpr int THREADS = 256;
constexpr int W_EPI0 = 0, W_APROD = 4, W_BPROD = 5, W_MMA = 6, W_ALLOC = 7;
constexpr int MASK_BYTES = 37 * 9 * 16;        // [N*N + 1][9 taps] 128-bit lane masks (the last entry: packed tiles)
constexpr int GATHER_BYTES = 4 * 32 * 32 * 4;  // policy epilogue: 32 rows x 32 channels f32 per epilogue warp
constexpr int SMEM_BYTES =
    A_STAGES * A_STAGE_BYTES + B_STAGES * B_STAGE_BYTES + 4096 /*bias, one copy per epilogue warp*/ + 512 /*barriers*/ +
    MASK_BYTES + GATHER_BYTES;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget of one CTA");
constexpr int MAX_LAYERS = 48;                 // layers one launch can chain (net5: 40 tower convolutions)

struct Layer {
    const __nv_bfloat16* in;        // [cin/8][rows][8] activations of the chunk's set; null for the encoding layer
    const TzState* enc_states;      // input convolution: the queued positions (row = position * N*N + square), the A
                                    // tile is encoded from them (pad1[0] holds white - black top flats)
    const __nv_bfloat16* w;         // [cin/64][9 taps][2 halves][8][128][8] pre-arranged weight blocks
    const float* bias;              // [256]
    const __nv_bfloat16* residual;  // [32][rows_set][8] or null
    __nv_bfloat16* out_act;         // [32][rows_set][8] or null
    const float* head_w;            // [2][256] value / UBE 1x1 convolution weights, or null
    float* head_out;                // [positions * N*N][2]: the two head dot products of every (global) row
    // policy convolution: legal-logit gather in the epilogue (all by position = evaluation-queue slot)
    float* g_out;                   // [positions][g_stride] logits of the listed moves, or null
    const uint16_t* g_actions;      // [positions][g_stride] moves
    const uint32_t* g_ranges;       // [positions][36] per square: index of its first move | count << 16 (indices into
                                    // g_perm when that is set, else into g_actions directly)
    const uint16_t* g_perm;         // [positions][g_stride] move indices grouped by square, or null (already grouped)
    int g_stride;
    int cin;                        // input channels (multiple of 64)
    int relu;
    int out_is_residual;            // the output is read again as a residual two layers on (else, in the local chain,
                                    // it only ever lives in shared memory)
};

struct Params {
    Layer layers[MAX_LAYERS];       // run back to back; layer l reads what layer l-1 wrote
    const __nv_bfloat16* dead_after[MAX_LAYERS];  // buffer of the chunk's activation set that nobody reads any more
                                    // once layer l's MMAs are done (or null): its lines are discarded from L2
    int n_layers;
    long long rows_set;             // rows per chunk plane of one activation set
    long long set_stride;           // elements between activation set 0 (even chunks) and set 1 (odd chunks)
    int chunk_min_tiles;            // least pair tiles (256 rows) per chunk; huge: the whole launch is one chunk
    const int* count_ptr;           // number of positions (device), or null: use count_max
    int count_max;
    int n;                          // board size
    int guard;                      // leading guard rows of the activation planes (= HALO)
    const uint4* masks;             // [n*n][9] disable-output-lane masks by (first tile row) mod n*n
    int f16;                        // 16-bit storage / operand type: 0 = bf16, 1 = IEEE fp16 (same UMMA kind::f16)
    unsigned* progress;             // [chunks][pair tiles per chunk][4], zeroed before the launch: epilogue-warp
                                    // arrivals per tile and 64-channel block of the output (8 per finished
                                    // layer); may be null when n_layers == 1
    unsigned* chunk_done;           // [chunks]: epilogue-warp arrivals of the last layer (8 per tile)
    unsigned* status;               // the handle's sticky error word (watchdog of the dependency waits)
    int debug_drop_progress;        // test hook: CTA pair 0 never publishes its tiles (exercises the watchdog)
    int allow_local;                // 0: never use the local chain (debug read-backs of the activation buffers)
    int pack;                       // > 0: PACKED rows for small batches of boards whose N*N does not divide 128 (5x5,
                                    // 6x6): every CTA tile holds `pack` whole positions (row = tile * 128 + position
                                    // in tile * N*N + square) and its last 128 - pack * N*N rows are dead, so no
                                    // position straddles two CTAs and the local chain applies; masks[N*N] is the lane
                                    // mask table of such a tile.  0: dense rows.
};

// work item -> (chunk, layer, pair tile); every role of the kernel walks the same sequence
struct Item {
    int chunk, layer, pt;
    int tiles;   // pair tiles of this chunk
    int rows;    // valid rows of this chunk
};
struct Schedule {
    int chunk_rows;   // rows of a full chunk
    int chunk_tiles;  // pair tiles of a full chunk
    int rows_used;    // rows of the launch
    int n_layers;
    int items;
    __host__ __device__ __forceinline__ void init(int count, int nn, int chunk_min_tiles, int layers) {
        rows_used = count * nn;
        // balanced chunks of at least chunk_min_tiles pair tiles (fewer than about two tiles per CTA pair and the
        // pairs stall on each other's progress; a sliver of a last chunk would walk through the layers alone)
        const int c1 = count > 0 ? count : 1;
        const int all_tiles = (rows_used + 2 * TILE_M - 1) / (2 * TILE_M);
        const int n_chunks = all_tiles >= 2 * chunk_min_tiles ? all_tiles / chunk_min_tiles : 1;
        const int cp = (c1 + n_chunks - 1) / n_chunks;
        chunk_rows = cp * nn;
        chunk_tiles = (chunk_rows + 2 * TILE_M - 1) / (2 * TILE_M);
        n_layers = layers;
        const int full = rows_used / chunk_rows, rest = rows_used - full * chunk_rows;
        items = (full * chunk_tiles + (rest + 2 * TILE_M - 1) / (2 * TILE_M)) * layers;
    }
    __host__ __device__ __forceinline__ Item at(int item) const {
        Item it;
        const int per_chunk = chunk_tiles * n_layers;
        it.chunk = item / per_chunk;
        const int r = item - it.chunk * per_chunk;
        const int left = rows_used - it.chunk * chunk_rows;
        it.rows = left < chunk_rows ? left : chunk_rows;
        it.tiles = (it.rows + 2 * TILE_M - 1) / (2 * TILE_M);
        it.layer = r / it.tiles;
        it.pt = r - it.layer * it.tiles;
        return it;
    }
};

#ifdef TZ_DEBUG_CHAIN
// chain-latency experiment: globaltimer stamps of one CTA pair's roles around two mid-network layers, kept in global
// memory and printed by the pair when the launch ends (printing at the events themselves distorts them)
__device__ long long g_chain[64];
__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define CHAIN(cond, layer, kind, k)                                        \
    do {                                                                   \
        if (cond) g_chain[((layer) - 10) * 32 + (kind) * 4 + (k)] = gtime(); \
    } while (0)
#else
#define CHAIN(cond, layer, kind, k)
#endif

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

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// row of the launch -> (position, square); false for a dead row of a packed tile or a position past the count
__device__ __forceinline__ bool row_position(int pack, int nn, int count, unsigned row, int* q, int* sq) {
    if (pack == 0) {
        *q = (int)(row / (unsigned)nn);
        *sq = (int)(row - (unsigned)*q * (unsigned)nn);
        return true;  // the callers bound dense rows by the chunk's row count
    }
    const unsigned tile = row / (unsigned)TILE_M, w = row - tile * (unsigned)TILE_M, b = w / (unsigned)nn;
    *q = (int)(tile * (unsigned)pack + b);
    *sq = (int)(w - b * (unsigned)nn);
    return (int)b < pack && *q < count;
}

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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint32_t bar, uint32_t n) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n) : "memory");
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
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Waits until *counter >= need.  Watchdog: the launch is cooperative, so the awaited pair is resident and a wait
// lasts microseconds; should one ever exceed ~1 s (2^22 polls of >= 0.25 us each), the error bit is raised, every other
// wait of the launch gives up at once and the host sees TZ_STATUS_NETWORK_STALL instead of a hung GPU.
constexpr unsigned STALL_BIT = 256;  // TZ_ERR_NETWORK_STALL
__device__ __forceinline__ void wait_counter(const unsigned* counter, unsigned need, unsigned* status) {
    unsigned polls = 0;
    while (ld_acquire_gpu(counter) < need) {
        __nanosleep(32);
        if ((++polls & 1023u) == 0) {
            if (*reinterpret_cast<volatile unsigned*>(status) & STALL_BIT) return;
            if (polls >= (1u << 22)) {
                atomicOr(status, STALL_BIT);
                return;
            }
        }
    }
}
// generic-proxy writes (observed through the acquire above) -> async-proxy reads of this thread's bulk copies
// the 128-byte line at `p` is dead: drop it from L2 without writing it back
__device__ __forceinline__ void discard_l2_line(const void* p) {
    asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
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
// two 16-bit activations <-> f32, in the network's storage type (bf16 or fp16; warp-uniform flag)
__device__ __forceinline__ uint32_t pack16(float a, float b, int f16) {
    if (f16) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    return pack_bf16(a, b);
}
__device__ __forceinline__ float2 unpack16(uint32_t w, int f16) {
    if (f16) return __half22float2(*reinterpret_cast<const __half2*>(&w));
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}


// one lane of the (converged) warp, chosen by the hardware
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
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
__device__ __forceinline__ void mbar_arrive_cluster_n(uint32_t cluster_addr, uint32_t n) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(n) : "memory");
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


__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) k_conv3x3_pair(const __grid_constant__ Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    uint8_t* a_smem = smem;
    uint8_t* b_smem = smem + A_STAGES * A_STAGE_BYTES;
    float* s_bias = reinterpret_cast<float*>(b_smem + B_STAGES * B_STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_bias) + 4096);
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t a_full = bar0, a_land = a_full + 8 * A_STAGES, a_empty = a_land + 8 * A_STAGES;
    const uint32_t b_full = a_empty + 8 * A_STAGES, b_land = b_full + 8 * B_STAGES;
    const uint32_t b_empty = b_land + 8 * B_STAGES;
    const uint32_t t_full = b_empty + 8 * B_STAGES, t_empty = t_full + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * A_STAGES + 3 * B_STAGES + 4);
    uint4* s_masks = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(bars) + 512);
    float* s_gather = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s_masks) + MASK_BYTES);

    const int count = p.count_ptr ? *p.count_ptr : p.count_max;
    const int nn = p.n * p.n;
    // packed rows: the schedule sees "positions" of 128 rows (one CTA tile = p.pack real positions + dead rows)
    const int period = p.pack ? TILE_M : nn;
    Schedule sched;
    sched.init(p.pack ? (count + p.pack - 1) / p.pack : count, period, p.chunk_min_tiles, p.n_layers);
    const int items = sched.items;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    // With fewer than two tiles per pair and layer the launch is bound by the layer-to-layer dependency chain: the
    // epilogues then publish every 64 output channels separately so that the next layer's first MMAs overlap the
    // rest of the epilogue.  With more tiles that only costs (four device-wide fences per tile instead of one).
    const bool fine = sched.chunk_tiles < 2 * npairs;
    // Local chain (see the header): pair p owns tile p of every layer, activations stay in shared memory.
    const bool local = p.allow_local && p.n_layers > 1 && (TILE_M % period) == 0 && sched.chunk_tiles <= npairs &&
                       items == sched.chunk_tiles * p.n_layers;
    const int item0 = local ? (pair < sched.chunk_tiles ? pair : items) : pair;
    const int istep = local ? sched.chunk_tiles : npairs;
    // arrivals that complete an A stage: one per CTA (its producer), or in the local chain one per epilogue warp
    const uint32_t a_arrivals = local ? 4u : 1u;

    if (threadIdx.x == 0) {
        for (int i = 0; i < A_STAGES; i++) {
            mbar_init(a_full + 8 * i, 2 * a_arrivals);
            mbar_init(a_land + 8 * i, 1);
            mbar_init(a_empty + 8 * i, 1);
        }
        for (int i = 0; i < B_STAGES; i++) {
            mbar_init(b_full + 8 * i, 2);
            mbar_init(b_land + 8 * i, 1);
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
    for (int i = threadIdx.x; i < (nn + 1) * 9; i += THREADS) s_masks[i] = p.masks[i];
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // both CTAs' barriers are initialised before anyone arrives remotely
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // the leader's copies signal the pair-level "full" barrier directly; the peer's land locally and are relayed
    const uint32_t a_sig = rank == 0 ? a_full : a_land, b_sig = rank == 0 ? b_full : b_land;

    if (warp == W_APROD) {
        // ---- A producer (both CTAs): this CTA's 144-row halo tile, one 64-channel block per stage.  Every lane walks
        // the item sequence (stage / phase stay warp-uniform); lane 0 issues the bulk copies of ordinary layers, all
        // lanes build the tile of the encoding layer.
        int stage = 0, phase = 0;
        for (int item = item0; item < items; item += istep) {
            const Item it = sched.at(item);
            const int layer = it.layer, pt = it.pt;
            const Layer& L = p.layers[layer];
            const int t = pt * 2 + (int)rank;
            if (L.enc_states != nullptr) {
                // ---- input convolution: encode rows [t*128 - HALO, t*128 + 128 + HALO) of the chunk from the queued
                // positions (rows outside the chunk are never read unmasked: a tap only looks at squares of the same
                // position, and positions do not straddle chunks)
                if (lane == 0) {
                    if (it.chunk >= 2 && p.n_layers > 1)  // overwrites nothing itself, but its epilogue does: see below
                        wait_counter(p.chunk_done + it.chunk - 2, 8u * (unsigned)sched.chunk_tiles, p.status);
                    mbar_wait(a_empty + 8 * stage, phase ^ 1);
                }
                __syncwarp();
#ifdef TZ_DEBUG_TIMING
                const long long enc_t0 = clock64();
#endif
                uint8_t* dst = a_smem + stage * A_STAGE_BYTES;
                const int first = t * TILE_M - HALO;  // chunk-relative row of tile row 0 (rows of a launch fit 31 bits)
                // a lane owns rows lane, lane + 32, ...: all loads first (they are independent), then the encoding
                constexpr int RPL = (A_ROWS + 31) / 32;
                uint64_t stack[RPL];
                uint4 tail[RPL];     // bytes 368..383 of the state: pad1[0], then the three scalar words
                uint32_t hts[RPL];   // height | top << 8 | to_move << 16, or ~0u for a row outside the chunk
#pragma unroll
                for (int k = 0; k < RPL; k++) {
                    const int r = lane + 32 * k;
                    const int rel = first + r;
                    hts[k] = 0xffffffffu;
                    stack[k] = 0;
                    tail[k] = make_uint4(0u, 0u, 0u, 0u);
                    if (r < A_ROWS && rel >= 0 && rel < it.rows) {
                        const unsigned grow = (unsigned)(it.chunk * sched.chunk_rows + rel);
                        int q, sq;
                        if (row_position(p.pack, nn, count, grow, &q, &sq)) {
                            const uint8_t* st = reinterpret_cast<const uint8_t*>(L.enc_states + q);
                            stack[k] = *reinterpret_cast<const uint64_t*>(st + 8 * sq);
                            hts[k] = (uint32_t)st[288 + sq] | ((uint32_t)st[324 + sq] << 8) | ((uint32_t)st[360] << 16);
                            tail[k] = *reinterpret_cast<const uint4*>(st + 368);
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < RPL; k++) {
                    const int r = lane + 32 * k;
                    if (r >= A_ROWS) continue;
                    uint8_t* row = dst + r * 16;
                    if (hts[k] != 0xffffffffu) {
                        const uint64_t ones = enc::square_indicator_bits(stack[k], (int)(hts[k] & 0xff), (int)((hts[k] >> 8) & 0xff),
                                                                         (int)((hts[k] >> 16) & 0xff), p.n);
                        enc::encode_square16_smem(row, A_KC_BYTES, ones, tail[k].y, tail[k].z, tail[k].w, p.n, p.f16);
                    } else {
#pragma unroll
                        for (int kc = 0; kc < 8; kc++) *reinterpret_cast<uint4*>(row + kc * A_KC_BYTES) = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
                // generic-proxy writes -> the tensor core's (async proxy) reads, then publish the stage
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    if (rank == 0) mbar_arrive_n(a_sig + 8 * stage, a_arrivals);
                    else mbar_arrive(a_sig + 8 * stage);  // a_land; the relay forwards a_arrivals
                }
#ifdef TZ_DEBUG_TIMING
                if (pair == 0 && rank == 0 && lane == 0 && item < 2 * npairs) printf("encode of one tile: %lld cycles\n", clock64() - enc_t0);
#endif
                if (++stage == A_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
                continue;
            }
            const int kblocks = L.cin >> 6;
            if (local) continue;  // the previous layer's epilogue fills the stages of this item
            if (lane == 0) {
                // rows [t*128 - HALO, t*128 + 128 + HALO) of the previous layer's output: CTA tiles t-1, t, t+1, i.e.
                // pair tiles {pt-1, pt} for rank 0 and {pt, pt+1} for rank 1; waited for per 64-channel block below
                const unsigned need = 8u * (unsigned)layer;
                const unsigned* prog = p.progress + ((size_t)it.chunk * sched.chunk_tiles) * 4;
                const int lo = pt - 1 + (int)rank;
                const __nv_bfloat16* in_base = L.in + (size_t)(it.chunk & 1) * p.set_stride;
                const uint8_t* src_tile =
                    reinterpret_cast<const uint8_t*>(in_base) + (size_t)(p.guard + t * TILE_M - HALO) * 16;
                int st = stage, ph = phase;
                for (int kb = 0; kb < kblocks; kb++) {
                    if (layer > 0 && (fine || kb == 0)) {
                        // fine: wait for this 64-channel block only; else for the whole tile (its last block)
                        const int blk = fine ? kb : 3;
                        for (int q = lo; q <= lo + 1; q++)
                            if (q >= 0 && q < it.tiles) wait_counter(prog + q * 4 + blk, need, p.status);
                        fence_proxy_async();
                        CHAIN(pair == 10 && rank == 0 && (layer == 10 || layer == 11), layer, 0, kb);
                    }
                    mbar_wait(a_empty + 8 * st, ph ^ 1);
                    mbar_arrive_expect_tx(a_sig + 8 * st, A_STAGE_BYTES);
                    const uint32_t dst = smem_u32(a_smem + st * A_STAGE_BYTES);
#pragma unroll
                    for (int kc = 0; kc < 8; kc++)
                        bulk_g2s(dst + kc * A_KC_BYTES, src_tile + (size_t)(kb * 8 + kc) * (size_t)p.rows_set * 16,
                                 A_KC_BYTES, a_sig + 8 * st);
                    if (++st == A_STAGES) {
                        st = 0;
                        ph ^= 1;
                    }
                }
            }
            for (int kb = 0; kb < kblocks; kb++)
                if (++stage == A_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
        }
    } else if (warp == W_BPROD) {
        // ---- B producer (both CTAs): this CTA's half (128 of the 256 N rows) of every weight block
        if (lane == 0) {
            int stage = 0, phase = 0;
            for (int item = item0; item < items; item += istep) {
                const Layer& L = p.layers[sched.at(item).layer];
                const uint8_t* src = reinterpret_cast<const uint8_t*>(L.w) + (size_t)rank * B_STAGE_BYTES;
                const int blocks = (L.cin >> 6) * 9;
                for (int blk = 0; blk < blocks; blk++) {
                    mbar_wait(b_empty + 8 * stage, phase ^ 1);
                    mbar_arrive_expect_tx(b_sig + 8 * stage, B_STAGE_BYTES);
                    bulk_g2s(smem_u32(b_smem + stage * B_STAGE_BYTES), src + (size_t)blk * 2 * B_STAGE_BYTES,
                             B_STAGE_BYTES, b_sig + 8 * stage);
                    if (++stage == B_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == W_ALLOC) {
        // ---- B relay (peer only): "my weight half has landed" -> the leader's b_full
        if (rank != 0 && lane == 0) {
            int stage = 0, phase = 0;
            const uint32_t b_full_leader = map_to_rank(b_full, 0);
            for (int item = item0; item < items; item += istep) {
                const int blocks = (p.layers[sched.at(item).layer].cin >> 6) * 9;
                for (int blk = 0; blk < blocks; blk++) {
                    mbar_wait(b_land + 8 * stage, phase);
                    mbar_arrive_cluster(b_full_leader + 8 * stage);
                    if (++stage == B_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == W_MMA) {
        if (rank != 0) {
            // ---- A relay (peer only)
            if (lane == 0) {
                int stage = 0, phase = 0;
                const uint32_t a_full_leader = map_to_rank(a_full, 0);
                for (int item = item0; item < items; item += istep) {
                    const Layer& RL = p.layers[sched.at(item).layer];
                    if (local && RL.enc_states == nullptr) break;  // only the encoded first layer lands here
                    const int kblocks = RL.cin >> 6;
                    for (int kb = 0; kb < kblocks; kb++) {
                        mbar_wait(a_land + 8 * stage, phase);
                        mbar_arrive_cluster_n(a_full_leader + 8 * stage, a_arrivals);
                        if (++stage == A_STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        } else {
            // ---- MMA issuer (leader only): kind::f16, D = f32, A = B = bf16 K-major, M = 256, N = 256, K = 16.
            // The whole warp runs the loop in uniform control flow (so descriptors, masks and barrier
            // addresses live in uniform registers) and one elected lane issues; a `lane == 0` branch around
            // the loop makes ptxas convert ~15 registers to uniform ones before EVERY mma, which made the
            // issuing thread, not the tensor pipe, the limit.
            // a_format / b_format (bits 7..9 / 10..12): 0 = F16, 1 = BF16
            const uint32_t fmt = p.f16 ? 0u : 1u;
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N_OUT >> 3) << 17) |
                                   ((uint32_t)((2 * TILE_M) >> 4) << 24);
            int a_stage = 0, a_phase = 0, b_stage = 0, b_phase = 0, it = 0;
#ifdef TZ_DEBUG_TIMING
            long long w_t = 0, w_a = 0, w_b = 0, w_a0 = 0;
            const long long mma_start = clock64();
            long long gt_start;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_start));
#endif
            for (int item = item0; item < items; item += istep, it++) {
                const Item wi = sched.at(item);
                const int pt = wi.pt;
                const int kblocks = p.layers[wi.layer].cin >> 6;
                const int acc = it & 1;
                TWAIT(w_t, mbar_wait(t_empty + 8 * acc, ((it >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * N_OUT;
                const uint4* masks0 = s_masks + (p.pack ? (size_t)nn : (size_t)(pt * 2) * TILE_M % nn) * 9;
                const uint4* masks1 = s_masks + (p.pack ? (size_t)nn : (size_t)(pt * 2 + 1) * TILE_M % nn) * 9;
                for (int kb = 0; kb < kblocks; kb++) {
#ifdef TZ_DEBUG_TIMING
                    if (p.layers[wi.layer].enc_states != nullptr)
                        TWAIT(w_a0, mbar_wait(a_full + 8 * a_stage, a_phase));
                    else
#endif
                        TWAIT(w_a, mbar_wait(a_full + 8 * a_stage, a_phase));
                    const uint32_t a_base = smem_u32(a_smem + a_stage * A_STAGE_BYTES);
                    // tap order: the centre tap first (no mask, it initialises every lane), then the rest
                    for (int ti = 0; ti < 9; ti++) {
                        const int tap = ti == 0 ? 4 : (ti <= 4 ? ti - 1 : ti);
                        TWAIT(w_b, mbar_wait(b_full + 8 * b_stage, b_phase));
                        tc_fence_after();
                        CHAIN(pair == 10 && lane == 0 && ti == 0 && (wi.layer == 10 || wi.layer == 11), wi.layer, 1, kb);
                        const int off = (tap / 3 - 1) * p.n + (tap % 3 - 1);
                        const uint4 m0 = masks0[tap], m1 = masks1[tap];
                        const uint32_t a_tap = a_base + (HALO + off) * 16;
                        const uint32_t b_base = smem_u32(b_smem + b_stage * B_STAGE_BYTES);
                        if (elect_one()) {
#pragma unroll
                            for (int ks = 0; ks < 4; ks++) {
                                const uint64_t adesc = make_desc(a_tap + ks * 2 * A_KC_BYTES, A_KC_BYTES);
                                const uint64_t bdesc = make_desc(b_base + ks * 2 * B_KC_BYTES, B_KC_BYTES);
                                tc_mma_pair(tmem_d, adesc, bdesc, idesc, (kb | ti | ks) != 0, m0, m1);
                            }
                            tc_commit_pair(b_empty + 8 * b_stage);
                            if (ti == 8) tc_commit_pair(a_empty + 8 * a_stage);
                            if (ti == 8 && kb == kblocks - 1) tc_commit_pair(t_full + 8 * acc);
                            CHAIN(pair == 10 && ti == 8 && kb == kblocks - 1 && (wi.layer == 10 || wi.layer == 11), wi.layer, 3, 1);
                        }
                        __syncwarp();
                        if (++b_stage == B_STAGES) {
                            b_stage = 0;
                            b_phase ^= 1;
                        }
                    }
                    if (++a_stage == A_STAGES) {
                        a_stage = 0;
                        a_phase ^= 1;
                    }
                }
            }
#ifdef TZ_DEBUG_TIMING
            if ((pair == 0 || pair == 40) && lane == 0)
            {
                long long gt_end;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_end));
                printf("pair %d mma: tiles %d total %lld wait_tmem %lld wait_a %lld (encoded tiles %lld) wait_b %lld ns %lld\n",
                       pair, it, clock64() - mma_start, w_t, w_a, w_a0, w_b, gt_end - gt_start);
            }
#endif
        }
    } else if (warp >= W_EPI0 && warp < W_EPI0 + 4) {
        // ---- epilogue (both CTAs): own 128 rows from own TMEM; t_empty lives in the leader
        const int wq = warp & 3;
        const uint32_t t_empty_leader = map_to_rank(t_empty, 0);
        const size_t plane = (size_t)p.rows_set * 8;  // elements per chunk plane of an activation set
        float* bias_w = s_bias + wq * N_OUT;          // this warp's copy of the current layer's bias
        const uint32_t a_full_leader = map_to_rank(a_full, 0);
        int it = 0, bias_layer = -1, a_consumed = 0;
        for (int item = item0; item < items; item += istep, it++) {
            const Item wi = sched.at(item);
            const int layer = wi.layer, pt = wi.pt;
            const Layer& L = p.layers[layer];
            if (layer != bias_layer) {
                __syncwarp();
#pragma unroll
                for (int j = 0; j < N_OUT / 32; j++) bias_w[lane + 32 * j] = __ldg(L.bias + lane + 32 * j);
                __syncwarp();
                bias_layer = layer;
            }
            const size_t set_off = (size_t)(wi.chunk & 1) * p.set_stride;
            const __nv_bfloat16* residual = L.residual ? L.residual + set_off : nullptr;
            // local chain: the output becomes the next item's A tile in shared memory (stage of its k-block kb =
            // (a_next + kb) mod A_STAGES); global memory only keeps what a later layer reads as its residual
            a_consumed += L.cin >> 6;
            const int a_next = a_consumed % A_STAGES;
            const bool to_smem = local && L.out_act != nullptr && layer + 1 < p.n_layers;
            __nv_bfloat16* out_act = L.out_act && (!local || L.out_is_residual) ? L.out_act + set_off : nullptr;
            const float* head_w = L.head_w;
            const int relu = L.relu;
            const int acc = it & 1;
            const int t = pt * 2 + (int)rank;
            const int rel = t * TILE_M + wq * 32 + lane;  // row of the chunk = position * n*n + square (dense rows)
            const size_t grow = (size_t)(p.guard + rel) * 8;  // element offset of the row inside a plane
            const size_t grel = (size_t)wi.chunk * sched.chunk_rows + (size_t)rel;  // row among all rows of the launch
            int q_row = 0, sq_row = 0;  // the position (evaluation-queue slot) and square this thread's row belongs to
            const bool valid = rel < wi.rows && row_position(p.pack, nn, count, (unsigned)grel, &q_row, &sq_row);
            float head_v = 0.0f, head_u = 0.0f;
            // policy layer: the legal moves that start on this thread's square (their logits are all it keeps)
            float* g_row = nullptr;
            const uint16_t* g_act = nullptr;
            const uint16_t* g_perm = nullptr;
            int g_first = 0, g_count = 0;
            // channels of the square's first eight moves (8 bits each; g_valid says which slots hold one), decoded once
            // per tile, and the 32-channel blocks that hold any of the square's moves
            uint32_t g_ch0 = 0, g_ch1 = 0, g_valid = 0, g_blocks = 0;
            int g_q = 0;
            if (L.g_out != nullptr && valid) {
                const int q = q_row, sq = sq_row;
                g_q = q;
                const uint32_t range = L.g_ranges[(size_t)q * 36 + sq];
                g_first = (int)(range & 0xffffu);
                g_count = (int)(range >> 16);
                g_row = L.g_out + (size_t)q * L.g_stride;
                g_act = L.g_actions + (size_t)q * L.g_stride;
                g_perm = L.g_perm ? L.g_perm + (size_t)q * L.g_stride : nullptr;
                for (int m = 0; m < g_count; m++) {
                    const int i = g_perm ? (int)g_perm[g_first + m] : g_first + m;
                    const int ch = enc::move_channel(p.n, g_act[i]);
                    if (ch < 0 || ch >= N_OUT) continue;  // cannot be a move of this board (host lists only)
                    g_blocks |= 1u << (ch >> 5);
                    if (m < 4) g_ch0 |= (uint32_t)ch << (8 * m);
                    else if (m < 8) g_ch1 |= (uint32_t)ch << (8 * (m - 4));
                    if (m < 8) g_valid |= 1u << m;
                }
            }
            float* stg = s_gather + (wq * 32 + lane) * 32;  // this thread's 32 staged channels (XOR-swizzled by lane)
            mbar_wait(t_full + 8 * acc, (it >> 1) & 1);
            tc_fence_after();
#ifndef TZ_NO_DISCARD
            if (p.dead_after[layer] != nullptr && !local && p.n_layers > 1) {
                // All MMAs of the tile are done, so its A rows have been read for the last time -- by this tile.  Rows
                // within HALO of the CTA tile's ends are also read by the neighbours; the 112 in between are dead: tell
                // the L2 not to write them back (a chunk's activation set is dirty from end to end when the chunk has
                // finished, and the next chunk's traffic would evict it to DRAM line by line).  One line = 8 rows of
                // one chunk plane; 14 lines x 32 planes per CTA tile, spread over the 128 epilogue threads.
                const uint8_t* base = reinterpret_cast<const uint8_t*>(p.dead_after[layer] + set_off) +
                                      (size_t)(p.guard + t * TILE_M + HALO) * 16;
                for (int i = wq * 32 + lane; i < 14 * 32; i += 128) {
                    const int plane_i = i / 14, line_i = i - plane_i * 14;
                    if (t * TILE_M + HALO + 8 * line_i < wi.rows)
                        discard_l2_line(base + (size_t)plane_i * (size_t)p.rows_set * 16 + (size_t)line_i * 128);
                }
            }
#endif
            CHAIN(pair == 10 && rank == 0 && wq == 0 && lane == 0 && (layer == 10 || layer == 11), layer, 3, 0);
            // Inside a fused launch the residual rows were written by another SM two layers ago.  One gpu-scope
            // acquire per tile (on the counter that writer released) makes the plain loads below see them: it costs
            // an L1 invalidate per warp and tile, whereas L2-only (ld.cg) loads made the epilogue 1.8x slower
            // than the MMAs of a tile.
            unsigned* prog_tile = p.progress + ((size_t)wi.chunk * sched.chunk_tiles + pt) * 4;
            if (residual != nullptr && layer >= 2 && !local) (void)ld_acquire_gpu(prog_tile + 3);
            const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + acc * N_OUT;
            // the residual rows of the next 32 channels are fetched while the current ones are processed: their L2
            // latency (the rows were written by another SM) would otherwise sit between every two blocks, which is
            // what bounds a layer when there is only one tile per CTA pair (small batches)
            uint4 rnext[4];
            if (residual && valid) {
#pragma unroll
                for (int j = 0; j < 4; j++) rnext[j] = *reinterpret_cast<const uint4*>(residual + (size_t)j * plane + grow);
            }
#pragma unroll 1
            for (int c0 = 0; c0 < N_OUT; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
                uint4 rcur[4];
                if (residual && valid) {
#pragma unroll
                    for (int j = 0; j < 4; j++) rcur[j] = rnext[j];
                    if (c0 + 32 < N_OUT) {
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            rnext[j] = *reinterpret_cast<const uint4*>(residual + (size_t)((c0 + 32) / 8 + j) * plane + grow);
                    }
                }
                tmem_ld_wait();
                if (valid) {
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) f[j] = __uint_as_float(v[j]) + bias_w[c0 + j];
                    if (residual) {
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const uint4 r = rcur[j];
                            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                const float2 x = unpack16(w[e], p.f16);
                                f[j * 8 + e * 2] += x.x;
                                f[j * 8 + e * 2 + 1] += x.y;
                            }
                        }
                    }
                    if (relu) {
#pragma unroll
                        for (int j = 0; j < 32; j++) f[j] = fmaxf(f[j], 0.0f);
                    }
                    if (out_act) {
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            *reinterpret_cast<uint4*>(out_act + (size_t)(c0 / 8 + j) * plane + grow) =
                                make_uint4(pack16(f[j * 8], f[j * 8 + 1], p.f16), pack16(f[j * 8 + 2], f[j * 8 + 3], p.f16),
                                           pack16(f[j * 8 + 4], f[j * 8 + 5], p.f16), pack16(f[j * 8 + 6], f[j * 8 + 7], p.f16));
                    }
                    if (to_smem) {
                        uint8_t* row = a_smem + ((a_next + (c0 >> 6)) % A_STAGES) * A_STAGE_BYTES +
                                       ((c0 & 63) >> 3) * A_KC_BYTES + (HALO + wq * 32 + lane) * 16;
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            *reinterpret_cast<uint4*>(row + j * A_KC_BYTES) =
                                make_uint4(pack16(f[j * 8], f[j * 8 + 1], p.f16), pack16(f[j * 8 + 2], f[j * 8 + 3], p.f16),
                                           pack16(f[j * 8 + 4], f[j * 8 + 5], p.f16), pack16(f[j * 8 + 6], f[j * 8 + 7], p.f16));
                    }
                    if ((g_blocks >> (c0 >> 5)) & 1u) {
                        // own row only: the thread stages its 32 channels and picks the legal ones by dynamic index
#pragma unroll
                        for (int j = 0; j < 32; j++) stg[j ^ lane] = f[j];
#pragma unroll
                        for (int m = 0; m < 8; m++) {
                            const int ch = (int)(((m < 4 ? g_ch0 : g_ch1) >> (8 * (m & 3))) & 0xffu);
                            if (((g_valid >> m) & 1u) && (ch & ~31) == c0) {
                                const int i = g_perm ? (int)g_perm[g_first + m] : g_first + m;
                                g_row[i] = stg[(ch & 31) ^ lane];
                            }
                        }
                    }
                    if (head_w) {  // value / UBE 1x1 convolutions over the tower output (net6_simhash.rs:88-119)
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            head_v = fmaf(f[j], __ldg(head_w + c0 + j), head_v);
                            head_u = fmaf(f[j], __ldg(head_w + N_OUT + c0 + j), head_u);
                        }
                    }
                }
                if (L.g_out != nullptr) {
                    // Squares with more than eight moves (tall stacks late in a game: up to several dozen spreads from
                    // one square) would make their one thread the slowest of the tile: the whole warp takes such a
                    // row's remaining moves, 32 at a time, from the owner's staged channels.
                    uint32_t heavy = __ballot_sync(0xffffffffu, valid && g_count > 8 && ((g_blocks >> (c0 >> 5)) & 1u));
                    if (heavy) {
                        __syncwarp();  // the owners' staged channels are visible to the other lanes
                        while (heavy) {
                            const int src = __ffs(heavy) - 1;
                            heavy &= heavy - 1;
                            const int hq = __shfl_sync(0xffffffffu, g_q, src);
                            const int hfirst = __shfl_sync(0xffffffffu, g_first, src);
                            const int hcount = __shfl_sync(0xffffffffu, g_count, src);
                            float* hrow = L.g_out + (size_t)hq * L.g_stride;
                            const uint16_t* hact = L.g_actions + (size_t)hq * L.g_stride;
                            const uint16_t* hperm = L.g_perm ? L.g_perm + (size_t)hq * L.g_stride : nullptr;
                            const float* hstg = s_gather + (wq * 32 + src) * 32;
                            for (int m = 8 + lane; m < hcount; m += 32) {
                                const int i = hperm ? (int)hperm[hfirst + m] : hfirst + m;
                                const int ch = enc::move_channel(p.n, hact[i]);
                                if (ch >= 0 && (ch & ~31) == c0) hrow[i] = hstg[(ch & 31) ^ src];
                            }
                        }
                        __syncwarp();  // before the owners stage the next block over these channels
                    }
                }
                if (to_smem && (c0 & 32) != 0) {
                    // 64 more channels = one k-block of the next layer are in place: generic-proxy writes -> the tensor
                    // core's reads, then this warp's arrival on the stage (8 epilogue warps of the pair complete it)
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        const uint32_t bar = 8u * (uint32_t)((a_next + (c0 >> 6)) % A_STAGES);
                        if (rank == 0) mbar_arrive(a_full + bar);
                        else mbar_arrive_cluster(a_full_leader + bar);
                    }
                }
                if (!local && p.n_layers > 1 && (fine ? (c0 & 32) != 0 : c0 == N_OUT - 32) && !(p.debug_drop_progress && pair == 0)) {
                    // 64 more output channels (fine) or the whole tile of this warp's rows are visible device-wide
                    __threadfence();
                    __syncwarp();
                    CHAIN(pair == 10 && rank == 0 && wq == 0 && lane == 0 && (layer == 10 || layer == 11), layer, 2, c0 >> 6);
                    if (fine) {
                        if (lane == 0) red_release_gpu_add(prog_tile + (c0 >> 6), 1u);
                    } else if (lane < 4) {
                        red_release_gpu_add(prog_tile + lane, 1u);
                    }
                }
            }
            if (head_w && valid)
                *reinterpret_cast<float2*>(L.head_out + ((size_t)q_row * nn + sq_row) * 2) = make_float2(head_v, head_u);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_cluster(t_empty_leader + 8 * acc);
                if (!local && p.n_layers > 1 && layer == p.n_layers - 1) red_release_gpu_add(p.chunk_done + wi.chunk, 1u);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // nobody leaves (or frees TMEM) while the pair's MMAs may still touch its memory
#ifdef TZ_DEBUG_CHAIN
    if (pair == 10 && rank == 0 && threadIdx.x == 0 && p.n_layers > 12) {
        const long long t0 = g_chain[4];  // layer 10: MMAs of kb 0 start
        for (int l = 0; l < 2; l++) {
            const long long* c = g_chain + l * 32;
            printf("L%d dep_ready %lld %lld %lld %lld | mma_start %lld %lld %lld %lld | issue_end %lld t_full %lld | publish %lld %lld %lld %lld\n",
                   10 + l, c[0] - t0, c[1] - t0, c[2] - t0, c[3] - t0, c[4] - t0, c[5] - t0, c[6] - t0, c[7] - t0, c[13] - t0,
                   c[12] - t0, c[8] - t0, c[9] - t0, c[10] - t0, c[11] - t0);
        }
    }
#endif
    if (warp == W_ALLOC) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

}  // namespace conv
