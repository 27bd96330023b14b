// batch_tc.cu — batched cosine / L2 / dot scan on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
// S = Q · Xᵀ as a tile contraction: the query block is the A operand (M = 128 queries per CTA, K-major,
// resident in shared memory for the CTA's whole life), database row tiles are the B operand
// (N = 256 rows per tile, K-major, streamed by TMA through a ring), and the 128×256 fp32 accumulator
// lives in TMEM, double buffered (2 × 256 of the 512 columns) so the epilogue of tile t overlaps the
// MMAs of tile t+1.
//
// Default: CTA PAIRS (tcgen05 cta_group::2).  The two CTAs of a cluster hold two different query blocks
// and issue ONE M = 256 MMA per K step; each CTA stages only HALF of every row tile (16 KB per K-chunk,
// 6-deep ring), so the bytes pulled from L2 per flop halve — the single-CTA form needs 64 B/clk/SM at
// full tensor rate, more than the ~6300 B/clk the L2 slices deliver to 148 SMs (it ran at 76 % tensor-pipe
// activity, the pair at 90 %; profiles/r02_batch_tc_pair_full.txt).  Batches with an odd number of
// 128-query blocks run single CTAs (optionally multicast clusters, VL_TC_CLUSTER).
//
// TMEM lane == query: every epilogue thread owns one query, keeps that query's running threshold in a
// register, reads its scores with tcgen05.ld (32x32b.x32), rejects a 32-score chunk with its
// NaN-propagating maximum (16 FMNMX3) and pushes the rare survivors (score >= τ) to the per-query
// candidate buffer — the B×N score matrix never reaches memory.  16 epilogue warps (four per TMEM lane
// quarter, 64 columns each) keep the survivor-heavy first filtered stage off the MMA's critical path.
// Inputs are a bf16 mirror of the arena (rows pre-scaled by 1/‖row‖ for cosine); the selected
// candidates are re-scored in f64 from the fp32 arena by the shared rescore/certify kernel, whose
// certificate uses the bf16 error bound (FinalizeParams::eps_scale / tc_abs).
//
// Roofline: tensor pipe.  flops per pass = 2·B·N·K; one 128×256×16 MMA = 128 cycles (4096 MAC/clk/SM).
#include <cuda.h>
#include <cuda_bf16.h>

#include "batch.h"
#include "tc_state.h"

namespace vl {

namespace tc {
constexpr int BM = 128, BN = 256, BK = 64, NSTAGE = 3, KCH_MAX = 6;
#ifndef VL_TC_EPI_WARPS
#define VL_TC_EPI_WARPS 16
#endif
constexpr int EPI_WARPS = VL_TC_EPI_WARPS;   // EPI_WARPS / 4 warps per TMEM lane quarter, an equal share of the tile's columns each
constexpr int CW = BN / (EPI_WARPS / 4);     // columns of a tile owned by one epilogue warp (64 with 16 warps)
static_assert(EPI_WARPS % 4 == 0 && CW % 32 == 0, "whole 32-column chunks per warp");
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = 64 + EPI_THREADS;    // warp 0 = TMA producer, warp 1 = MMA issuer
constexpr int TMEM_BUF_COLS = 256;   // accumulator buffers sit 256 TMEM columns apart (2 × 256 = all 512)
constexpr uint32_t A_CHUNK_BYTES = BM * BK * 2;   // 16 KB
constexpr uint32_t B_STAGE_BYTES = BN * BK * 2;   // 32 KB
constexpr uint32_t SMEM_A = KCH_MAX * A_CHUNK_BYTES;          // 96 KB
constexpr uint32_t SMEM_B = NSTAGE * B_STAGE_BYTES;           // 96 KB
constexpr uint32_t SMEM_XN = 0;                               // (the L2 epilogue keeps its squared norms in registers)
constexpr uint32_t SMEM_BAR = 256;
constexpr int QCAP = 8;                                       // survivor queue entries per epilogue thread
constexpr uint32_t SMEM_PQ = QCAP * EPI_THREADS * 8;          // 16 KB, interleaved [QCAP][EPI_THREADS]
constexpr uint32_t SMEM_TOTAL = SMEM_A + SMEM_B + SMEM_XN + SMEM_BAR + SMEM_PQ + 1024;  // + alignment slack

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
                 ::"r"(dst), "l"(map), "r"(bar), "h"(mask), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
// ---- CTA pair (cta_group::2): one MMA over two SMs, M = 256 (128 queries per CTA), the B tile split across the pair ----
// Same offset in the pair's LEADER (even) CTA: inside a CTA pair the shared::cluster window of the two CTAs differs
// in address bit 24 only, so clearing it names the leader's copy from either CTA (what CUTLASS's 2-SM TMA atoms do).
__device__ __forceinline__ uint32_t leader_addr(uint32_t local_smem_addr) { return local_smem_addr & 0xFEFFFFFFu; }
// TMA load into THIS CTA's shared memory whose bytes are credited to a barrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar_cluster_addr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {   // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(static_cast<uint16_t>(3)) : "memory");
}
// Arrive on a barrier of the pair's other CTA.  Default (cta-scope) semantics on purpose: `.release.cluster` compiles
// to MEMBAR.ALL.GPU + ERRBAR in front of the arrive (30 % of the pair kernel's stall samples, profiles/
// r02_batch_tc_pair_full.txt); what the arrive orders here are tcgen05.ld reads of tensor memory, which
// tcgen05.wait::ld + tcgen05.fence::before_thread_sync already order.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
constexpr int NSTAGE_PAIR = 6;                           // half tiles: twice the ring depth in the same 96 KB
constexpr uint32_t B_STAGE_BYTES_PAIR = (BN / 2) * BK * 2;   // 16 KB
constexpr uint32_t IDESC_PAIR_BITS = (1u << 4) | (1u << 7) | (1u << 10) | ((BN >> 3) << 17) | (((2 * BM) >> 4) << 24);

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
}
// max(a, b, c), NaN if any operand is NaN (one FMNMX3.NAN)
__device__ __forceinline__ float fmax3_nan(float a, float b, float c) {
    float r;
    asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// v[i] for a runtime i without spilling v to local memory (31 selects; survivors are rare)
__device__ __forceinline__ uint32_t select32(const uint32_t (&v)[32], int i) {
    uint32_t r = v[0];
#pragma unroll
    for (int j = 1; j < 32; ++j) r = (j == i) ? v[j] : r;
    return r;
}
__device__ __forceinline__ bool elect_one() {   // one deterministic leader lane of a converged warp
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (rows of 128 B, 8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (static_cast<uint64_t>((saddr >> 4) & 0x3FFF)) | (1ull << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}
// kind::f16: D = f32, A = B = bf16, both K-major, M = 128, N = 256
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((BN >> 3) << 17) | ((BM >> 4) << 24);

struct Params {
    const float* sq_norm;   // [n] (L2)
    const float* qn2;       // [nq_pad] ‖q‖² (L2)
    const float* tau;
    uint64_t* cand;
    uint32_t* count;
    uint32_t* qflags;
    uint32_t capq, nq, row_lo, row_hi, kch, qblocks, tiles;
    uint32_t direct;        // stage 0 of small stores: every score is kept, slot = row − row_lo, no atomics
    uint32_t groupmax;      // threshold-estimation stage: only the maximum of every 32-row group is written
    float* gmax;            // [nq][GMAX_STRIDE]
    uint32_t pdl_first;     // first kernel of the batch's PDL chain: EVERY warp waits for the query conversion
};

// PAIR (CS == 2 only): the two CTAs of a cluster issue ONE tcgen05.mma.cta_group::2 per K step — M = 256 (each CTA's
// 128 queries), N = 256 rows of which each CTA stages HALF (16 KB per K-chunk instead of 32 KB: half the L2 → SM
// traffic, half the shared-memory writes, the B operand read once per pair, a 6-deep ring in the same 96 KB).  Only
// the leader CTA (rank 0) issues MMAs; both issue TMA, whose bytes are credited to the LEADER's barriers; commits
// arrive on the barriers of both CTAs; each CTA's epilogue reads its own TMEM half and releases the accumulator
// buffer on the leader's barrier.
//
// STREAM_A (rows wider than KCH_MAX·64 = 384 elements, e.g. 768 / 1024 / 1536-d embeddings): the query block no longer
// fits in shared memory next to the row ring, so its K-chunks travel through the ring WITH the row chunks — one stage =
// the 16 KB query chunk + the row chunk of the same K range, the whole 192 KB as ring (6 stages of 32 KB for a pair,
// 4 of 48 KB for a single CTA).  The query chunks are re-read from L2 for every row tile: 32 KB per K-chunk per CTA
// of a pair, the L2 → SM rate of the single-CTA kernel at 384-d.
template <int METRIC, int CS, bool PAIR = false, bool STREAM_A = false>
__global__ void __launch_bounds__(THREADS, 1)
batch_scan_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_q, Params p) {
    static_assert(!PAIR || CS == 2, "a CTA pair is a cluster of two");
    static_assert(!STREAM_A || PAIR || CS == 1, "streamed query chunks: single CTAs or CTA pairs only");
    constexpr uint32_t BSTAGE = PAIR ? B_STAGE_BYTES_PAIR : B_STAGE_BYTES;
    constexpr uint32_t STRIDE = STREAM_A ? A_CHUNK_BYTES + BSTAGE : BSTAGE;      // bytes from one ring stage to the next
    constexpr uint32_t B_OFF = STREAM_A ? A_CHUNK_BYTES : 0u;                     // row chunk inside a stage
    constexpr int NST = STREAM_A ? static_cast<int>((SMEM_A + SMEM_B) / STRIDE) : (PAIR ? NSTAGE_PAIR : NSTAGE);
    constexpr uint32_t STAGE_TX = BSTAGE + (STREAM_A ? A_CHUNK_BYTES : 0u);      // bytes one CTA loads per stage
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    unsigned char* s_a = base;                                  // resident query block (!STREAM_A)
    unsigned char* s_ring = STREAM_A ? base : base + SMEM_A;    // stage i: [query chunk (STREAM_A)] [row chunk]
    unsigned char* s_b = s_ring + B_OFF;
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + SMEM_A + SMEM_B + SMEM_XN);
    // bars: [0..NST) full, [NST..2NST) empty, then a_full, tmem_full[2], tmem_empty[2]
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + NST), bar_a = smem_u32(bars + 2 * NST);
    const uint32_t bar_tfull = smem_u32(bars + 2 * NST + 1), bar_tempty = smem_u32(bars + 2 * NST + 3);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * NST + 5);
    unsigned long long* s_pq = reinterpret_cast<unsigned long long*>(base + SMEM_A + SMEM_B + SMEM_XN + SMEM_BAR);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CS > 1 ? cluster_rank() : 0;
    // work: cluster (or CTA) `cid` serves query group `qg` and row tiles t0, t0+tstride, ...
    const uint32_t cid = blockIdx.x / CS, nclusters = gridDim.x / CS;
    const uint32_t qgroups = (p.qblocks + CS - 1) / CS;
    const uint32_t qg = cid % qgroups, t0 = cid / qgroups, tstride = nclusters / qgroups;
    const uint32_t qblock = qg * CS + rank;   // may be >= qblocks (padding CTA of the last group): scores ignored

    if (threadIdx.x == 0) {
        for (int i = 0; i < NST; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, PAIR ? 1 : CS);   // one tcgen05.commit arrival from every MMA-issuing CTA
        }
        mbar_init(bar_a, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, PAIR ? 2 * EPI_WARPS : EPI_WARPS);   // one arrival per epilogue warp (pair: of both CTAs)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (PAIR) {   // the same warp of both CTAs allocates the pair's tensor memory
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    tc_fence_before();
    if (CS > 1) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    const bool active = t0 < p.tiles && tstride > 0;
    // Programmatic dependent launch: the kernels of a batch (convert → scan → select → scan → … → rescore) are
    // chained so that each one is launched while its predecessor is still running; every kernel releases its
    // dependents only AFTER its own wait has returned, so a kernel never overlaps anything older than its direct
    // predecessor.  Here the predecessor is the select that produced tau / count / cand: only the epilogue reads
    // those, so the TMA producer and the MMA issuer start on the (immutable) bf16 rows and queries at once — except
    // in the first scan of a batch, whose predecessor WRITES the bf16 queries.
    if (p.pdl_first) {
        pdl_wait();
        if (threadIdx.x == 0) pdl_launch_dependents();
    }

    if (warp == 0) {
        // ================= TMA producer =================
        if (PAIR) {
            if (lane == 0 && active) {
                // every byte of the pair (both A blocks, both halves of every B stage) is credited to the LEADER's barriers
                if (!STREAM_A) {
                    const uint32_t lead_a = leader_addr(bar_a);
                    if (rank == 0) mbar_expect_tx(bar_a, 2 * p.kch * A_CHUNK_BYTES);
                    for (uint32_t kc = 0; kc < p.kch; ++kc)
                        tma_load_2d_pair(smem_u32(s_a + kc * A_CHUNK_BYTES), &map_q, lead_a, kc * BK, qblock * BM);
                }
                uint32_t stage = 0, phase = 0;
                for (uint32_t t = t0; t < p.tiles; t += tstride) {
                    const int row = static_cast<int>(p.row_lo + t * BN + rank * (BN / 2));   // this CTA's half of the tile
                    for (uint32_t kc = 0; kc < p.kch; ++kc) {
                        mbar_wait(bar_empty + 8 * stage, phase ^ 1);      // own barrier: the commit arrives on both CTAs
                        const uint32_t lead_full = leader_addr(bar_full + 8 * stage);
                        if (rank == 0) mbar_expect_tx(bar_full + 8 * stage, 2 * STAGE_TX);
                        if (STREAM_A) tma_load_2d_pair(smem_u32(s_ring + stage * STRIDE), &map_q, lead_full, kc * BK, qblock * BM);
                        tma_load_2d_pair(smem_u32(s_b + stage * STRIDE), &map_x, lead_full, kc * BK, row);
                        if (++stage == NST) { stage = 0; phase ^= 1; }
                    }
                }
            }
        } else if (lane == 0 && active) {
            if (!STREAM_A) {
                mbar_expect_tx(bar_a, p.kch * A_CHUNK_BYTES);
                for (uint32_t kc = 0; kc < p.kch; ++kc)
                    tma_load_2d(smem_u32(s_a + kc * A_CHUNK_BYTES), &map_q, bar_a, kc * BK, qblock * BM);
            }
            uint32_t stage = 0, phase = 0;
            for (uint32_t t = t0; t < p.tiles; t += tstride) {
                const int row = static_cast<int>(p.row_lo + t * BN);
                for (uint32_t kc = 0; kc < p.kch; ++kc) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    mbar_expect_tx(bar_full + 8 * stage, STAGE_TX);
                    if (STREAM_A) tma_load_2d(smem_u32(s_ring + stage * STRIDE), &map_q, bar_full + 8 * stage, kc * BK, qblock * BM);
                    if (CS == 1) {
                        tma_load_2d(smem_u32(s_b + stage * STRIDE), &map_x, bar_full + 8 * stage, kc * BK, row);
                    } else {
                        constexpr uint32_t SLICE = BN / CS;  // rows loaded by this CTA, multicast to all
                        tma_load_2d_mc(smem_u32(s_b + stage * STRIDE + rank * SLICE * 128), &map_x,
                                       bar_full + 8 * stage, kc * BK, row + rank * SLICE,
                                       static_cast<uint16_t>((1u << CS) - 1));
                    }
                    if (++stage == NST) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // The whole warp runs the loop (warp-uniform control flow and descriptors stay in uniform
        // registers); one elected lane issues the MMAs and their commits.  Descriptors are advanced
        // by adding to the encoded 16-byte address field instead of being rebuilt per instruction.
        if (PAIR) {
            if (active && rank == 0) {   // the leader issues for the pair; its barriers see both CTAs' bytes / epilogues
                if (!STREAM_A) mbar_wait(bar_a, 0);
                tc_fence_after();
                const uint64_t adesc0 = make_desc(smem_u32(STREAM_A ? s_ring : s_a)), bdesc0 = make_desc(smem_u32(s_b));
                uint32_t stage = 0, phase = 0, buf = 0, tphase = 0;
                for (uint32_t t = t0; t < p.tiles; t += tstride) {
                    mbar_wait(bar_tempty + 8 * buf, tphase ^ 1);
                    tc_fence_after();
                    const uint32_t d = tmem_base + buf * TMEM_BUF_COLS;
                    for (uint32_t kc = 0; kc < p.kch; ++kc) {
                        mbar_wait(bar_full + 8 * stage, phase);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t ad = adesc0 + static_cast<uint64_t>(STREAM_A ? stage * (STRIDE >> 4) : kc * (A_CHUNK_BYTES >> 4));
                            const uint64_t bd = bdesc0 + static_cast<uint64_t>(stage * (STRIDE >> 4));
#pragma unroll
                            for (int j = 0; j < BK / 16; ++j)
                                umma_bf16_pair(d, ad + 2 * j, bd + 2 * j, IDESC_PAIR_BITS, (kc | j) != 0 ? 1u : 0u);
                            umma_commit_pair(bar_empty + 8 * stage);    // frees the stage in both CTAs
                        }
                        __syncwarp();
                        if (++stage == NST) { stage = 0; phase ^= 1; }
                    }
                    if (elect_one()) umma_commit_pair(bar_tfull + 8 * buf);   // both CTAs' epilogues
                    __syncwarp();
                    buf ^= 1;
                    if (buf == 0) tphase ^= 1;
                }
            }
        } else if (active) {
            if (!STREAM_A) mbar_wait(bar_a, 0);
            tc_fence_after();
            const uint64_t adesc0 = make_desc(smem_u32(STREAM_A ? s_ring : s_a)), bdesc0 = make_desc(smem_u32(s_b));
            uint32_t stage = 0, phase = 0, buf = 0, tphase = 0;
            for (uint32_t t = t0; t < p.tiles; t += tstride) {
                mbar_wait(bar_tempty + 8 * buf, tphase ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + buf * TMEM_BUF_COLS;
                for (uint32_t kc = 0; kc < p.kch; ++kc) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t ad = adesc0 + static_cast<uint64_t>(STREAM_A ? stage * (STRIDE >> 4) : kc * (A_CHUNK_BYTES >> 4));
                        const uint64_t bd = bdesc0 + static_cast<uint64_t>(stage * (STRIDE >> 4));
#pragma unroll
                        for (int j = 0; j < BK / 16; ++j)
                            umma_bf16(d, ad + 2 * j, bd + 2 * j, IDESC, (kc | j) != 0 ? 1u : 0u);
                        if (CS == 1) umma_commit(bar_empty + 8 * stage);
                        else umma_commit_mc(bar_empty + 8 * stage, static_cast<uint16_t>((1u << CS) - 1));
                    }
                    __syncwarp();
                    if (++stage == NST) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) umma_commit(bar_tfull + 8 * buf);
                __syncwarp();
                buf ^= 1;
                if (buf == 0) tphase ^= 1;
            }
        }
    } else {
        // ================= epilogue: TMEM lane == query =================
        const int quarter = warp & 3;                       // TMEM lanes [32q, 32q+32) belong to this warp
        const int etid = (warp - 2) * 32 + lane;            // 0..255 among the epilogue threads
        const int cbase = ((warp - 2) >> 2) * CW;           // this warp's share of the tile's columns
        const uint32_t q = qblock * BM + quarter * 32 + lane;
        const bool qvalid = qblock < p.qblocks && q < p.nq;
        if (!p.pdl_first) {
            pdl_wait();                                        // the select before this stage is complete and visible
            if (etid == 0) pdl_launch_dependents();
        }
        float tau = qvalid ? __ldcg(p.tau + q) : INFINITY;    // (written by the previous kernel: no read-only path)
        float qn = 0.f;
        if (METRIC == EUCLIDEAN) {
            qn = qvalid ? __ldg(p.qn2 + q) : 0.f;
            tau += qn;                                      // compare 2·acc − ‖x‖² against τ + ‖q‖²
        }
        bool nonfinite = false;
        // Survivors are queued in shared memory ([QCAP][128], conflict-free) and flushed with ONE
        // slot-reserving atomicAdd per QCAP entries: the scan loop never waits on an L2 round trip.
        int nloc = 0;
        unsigned long long* my_q = s_pq + etid;
        auto flush = [&]() {
            if (nloc) {
                const uint32_t idx0 = atomicAdd(p.count + q, static_cast<uint32_t>(nloc));
                for (int i = 0; i < nloc; ++i) {
                    if (idx0 + i < p.capq) p.cand[static_cast<size_t>(q) * p.capq + idx0 + i] = my_q[i * EPI_THREADS];
                    else atomicOr(p.qflags + q, FLAG_OVERFLOW);
                }
                nloc = 0;
            }
        };
        if (active) {
            uint32_t buf = 0, tphase = 0;
            // L2: every warp keeps the ‖x‖² of its CW columns in registers (lane l holds columns cbase + 32j + l),
            // loaded one tile ahead — no shared-memory staging and no barrier between epilogue warps
            constexpr int NXR = CW / 32;
            float xr[NXR] = {}, xp[NXR] = {};
            auto load_xn = [&](uint32_t tile, float (&dst)[NXR]) {
#pragma unroll
                for (int j = 0; j < NXR; ++j) {
                    const uint32_t r = p.row_lo + tile * BN + cbase + 32 * j + lane;
                    dst[j] = (tile < p.tiles && r < p.row_hi) ? __ldg(p.sq_norm + r) : 0.f;
                }
            };
            if (METRIC == EUCLIDEAN) load_xn(t0, xp);
            for (uint32_t t = t0; t < p.tiles; t += tstride) {
                const uint32_t row0 = p.row_lo + t * BN;
                if (METRIC == EUCLIDEAN) {
#pragma unroll
                    for (int j = 0; j < NXR; ++j) xr[j] = xp[j];
                    load_xn(t + tstride, xp);
                }
                mbar_wait(bar_tfull + 8 * buf, tphase);
                tc_fence_after();
                const uint32_t tbase = tmem_base + buf * TMEM_BUF_COLS + (static_cast<uint32_t>(quarter * 32) << 16);
                // one 32-column chunk: branch-free survivor mask (2 instructions per score), then a
                // compact loop over the (rare) set bits — keeps the hot loop inside the instruction cache
                auto process = [&](const uint32_t (&v)[32], int c0, float xn_lane) {   // xn_lane: ‖x‖² of column c0 + lane
                    if (p.groupmax) {
                        // Threshold estimation: the K'-th largest of the maxima of G disjoint row groups is reached by K'
                        // distinct rows, hence a valid lower bound of the K'-th best score — no score is stored, the
                        // rows are scanned again (filtered) by the next stage.
                        float m = -INFINITY;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            float s = __uint_as_float(v[i]);
                            if (METRIC == EUCLIDEAN) s = fmaf(2.f, s, -__shfl_sync(0xFFFFFFFFu, xn_lane, i)) - qn;
                            if (row0 + c0 + i < p.row_hi) m = fmaxf(m, s);   // fmaxf drops NaNs: the bound only gets weaker
                        }
                        if (qvalid) p.gmax[static_cast<size_t>(q) * GMAX_STRIDE + ((row0 - p.row_lo + c0) >> 5)] = m;
                        return;
                    }
                    if (p.direct) {
                        // stage 0 keeps every score: transpose 32 queries × 16 columns through this warp's 2 KB of
                        // the (idle) survivor queue so that each half-warp stores 16 consecutive keys of ONE
                        // query (128 contiguous bytes) instead of 32 keys 8·capq bytes apart
                        float* st = reinterpret_cast<float*>(s_pq) + (warp - 2) * 512;
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                float s = __uint_as_float(v[half * 16 + i]);
                                if (METRIC == EUCLIDEAN) s = fmaf(2.f, s, -__shfl_sync(0xFFFFFFFFu, xn_lane, half * 16 + i)) - qn;
                                if (!isfinite(s)) nonfinite = true;
                                st[lane * 16 + (i ^ (lane & 15))] = s;
                            }
                            __syncwarp();
                            const int c = lane & 15;
                            const uint32_t r = row0 + c0 + half * 16 + c;
#pragma unroll 4
                            for (int it = 0; it < 16; ++it) {
                                const int row = 2 * it + (lane >> 4);
                                const float s = st[row * 16 + (c ^ (row & 15))];
                                const uint32_t qq = qblock * BM + quarter * 32 + row;
                                if (qblock < p.qblocks && qq < p.nq && r < p.row_hi)
                                    p.cand[static_cast<size_t>(qq) * p.capq + (r - p.row_lo)] = make_key(s, r);
                            }
                            __syncwarp();
                        }
                        return;
                    }
                    // L2: s = 2·acc − ‖x‖² >= τ' needs acc >= (τ' + ‖x‖²)/2; with the chunk's SMALLEST ‖x‖² (one
                    // redux per 32 columns; ‖x‖² >= 0 so the uint order is the float order) that is a necessary
                    // condition on acc alone — the hot loop is the same 2 instructions per score as cosine / dot and
                    // the exact predicate is re-applied to the (rare) survivors.  The margin keeps the filter
                    // conservative under rounding.
                    float thr = tau;
                    if (METRIC == EUCLIDEAN) {
                        const float xm = __uint_as_float(__reduce_min_sync(0xFFFFFFFFu, __float_as_uint(xn_lane)));
                        // (a lane without a query carries tau = +inf: inf − inf would make the threshold NaN and send
                        // every chunk of that lane down the survivor path — 2x the batch time at 16 queries)
                        thr = tau < INFINITY ? 0.5f * (tau + xm) - (4e-7f * (fabsf(tau) + xm) + 1e-30f) : INFINITY;
                    }
                    // Fast reject: the chunk's NaN-propagating maximum (16 three-input FMNMX3 for 32 scores, two
                    // independent chains) against the threshold.  Only when some lane of the warp has a candidate (or a
                    // NaN) does the warp build the per-score masks — in the main stage that is about one chunk in four.
                    float mxa = fmax3_nan(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
                    float mxb = fmax3_nan(__uint_as_float(v[3]), __uint_as_float(v[4]), __uint_as_float(v[5]));
#pragma unroll
                    for (int i = 6; i < 30; i += 4) {
                        mxa = fmax3_nan(mxa, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
                        mxb = fmax3_nan(mxb, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
                    }
                    mxa = fmax3_nan(mxa, __uint_as_float(v[30]), __uint_as_float(v[31]));
                    if (!__any_sync(0xFFFFFFFFu, !(fmax3_nan(mxa, mxb, mxb) < thr))) return;
                    uint32_t m4[4] = {0u, 0u, 0u, 0u};   // independent partial masks: no 32-long dependent chain
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        m4[i & 3] |= !(__uint_as_float(v[i]) < thr) ? (1u << i) : 0u;   // acc >= thr, or NaN
                    uint32_t mask = (m4[0] | m4[1]) | (m4[2] | m4[3]);
                    while (mask) {
                        const int i = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const uint32_t r = row0 + c0 + i;
                        if (qvalid && r < p.row_hi) {
                            float s = __uint_as_float(select32(v, i));
                            if (METRIC == EUCLIDEAN) {
                                s = fmaf(2.f, s, -__ldg(p.sq_norm + r));   // rare path: L1/L2 hit
                                if (s < tau) continue;              // passed the chunk-minimum filter only
                                s -= qn;                            // −‖x−q‖²
                            }
                            if (!isfinite(s)) nonfinite = true;
                            my_q[nloc * EPI_THREADS] = make_key(s, r);
                            if (++nloc == QCAP) flush();   // rare: a lane filled its queue inside one chunk
                        }
                    }
                    // Convergent flush: as soon as ANY lane's queue is half full, every lane reserves its slots
                    // in the same instruction, so the warp pays one L2 round trip for 32 queues instead of one
                    // per lane at 32 different (divergent) moments.
                    if (__any_sync(0xFFFFFFFFu, nloc >= QCAP / 2)) flush();
                };
                uint32_t va[32], vb[32];
                tmem_ld32(tbase + cbase, va);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < NXR; j += 2) {   // two chunks per iteration, loads double buffered
                    const int c0 = cbase + 32 * j;
                    tmem_ld32(tbase + c0 + 32, vb);
                    process(va, c0, xr[j]);
                    tmem_ld_wait();
                    if (j + 2 < NXR) tmem_ld32(tbase + c0 + 64, va);
                    process(vb, c0 + 32, xr[j + 1]);
                    if (j + 2 < NXR) tmem_ld_wait();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (PAIR) mbar_arrive_cluster(leader_addr(bar_tempty + 8 * buf));   // the leader's barrier counts both CTAs
                    else mbar_arrive(bar_tempty + 8 * buf);
                }
                buf ^= 1;
                if (buf == 0) tphase ^= 1;
            }
        }
        flush();
        if (nonfinite && qvalid) atomicOr(p.qflags + q, FLAG_NONFINITE);
    }

    tc_fence_before();
    if (CS > 1) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// fp32 arena rows → bf16 mirror [n][KP] (cosine: scaled by 1/‖row‖), zero padded to KP
__global__ void to_bf16_rows_kernel(const float* __restrict__ rows, const float* __restrict__ inv_norm, uint64_t first,
                                    uint64_t n, uint32_t dim, uint32_t pitch, uint32_t KP, int normalise,
                                    __nv_bfloat16* out, float* sq_norm, uint32_t* ex_bits, uint32_t* ex1_bits) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
    for (uint64_t r = warp; r < n; r += nwarps) {
        const float* src = rows + (first + r) * pitch;
        const float sc = normalise ? inv_norm[first + r] : 1.f;
        double ss = 0.0, ref2 = 0.0, err2 = 0.0, err1 = 0.0;   // ‖x‖², ‖x̂‖² (what is rounded), ‖x̃−x̂‖², ‖x̃−x̂‖₁
        for (uint32_t c = lane; c < KP; c += 32) {
            const float x = c < dim ? src[c] : 0.f;
            ss += static_cast<double>(x) * x;
            const float xs = x * sc;
            const __nv_bfloat16 b = __float2bfloat16_rn(xs);
            const double d = static_cast<double>(__bfloat162float(b)) - static_cast<double>(xs);
            ref2 += static_cast<double>(xs) * xs;
            err2 += d * d;
            err1 += fabs(d);
            out[(first + r) * KP + c] = b;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
            ref2 += __shfl_xor_sync(0xFFFFFFFFu, ref2, o);
            err2 += __shfl_xor_sync(0xFFFFFFFFu, err2, o);
            err1 += __shfl_xor_sync(0xFFFFFFFFu, err1, o);
        }
        if (lane == 0) {
            if (sq_norm) sq_norm[first + r] = static_cast<float>(ss);
            // measured relative rounding error of this row, rounded UP to fp32; non-negative floats order like
            // their bit patterns, so an integer atomicMax keeps the maximum over all rows
            if (ex_bits && ref2 > 0.0) atomicMax(ex_bits, __float_as_uint(__double2float_ru(sqrt(err2 / ref2) * 1.000001)));
            if (ex1_bits) atomicMax(ex1_bits, __float_as_uint(__double2float_ru(err1 * 1.000001)));
        }
    }
}

// fp32 queries [nq][pitch] → bf16 [nq_pad][KP] (+ ‖q‖² for L2); rows >= nq are zero
__global__ void to_bf16_queries_kernel(const float* __restrict__ q, uint32_t nq, uint32_t nq_pad, uint32_t dim,
                                       uint32_t pitch, uint32_t KP, __nv_bfloat16* out, float* qn2, float* eq,
                                       uint32_t* count, float* tau, uint32_t* qflags) {
    const uint32_t r = blockIdx.x;
    // Chained batches (launched with the PDL attribute behind the previous batch's rescore kernel, which released us
    // after ITS wait — every scan of that batch is complete, so the bf16 query staging may be overwritten; what the
    // rescore still reads is double buffered).  This kernel only completes after the rescore has (wait at the end),
    // which keeps "kernel j complete ⇒ kernel j−1 complete" true along the whole chain.
    pdl_launch_dependents();            // the first scan may be launched: all of its warps wait for this grid
    if (threadIdx.x == 0 && r < nq) {   // per-query state of the staged scan (was a kernel of its own)
        count[r] = 0u;
        tau[r] = -INFINITY;
        qflags[r] = 0u;
    }
    double ss = 0.0, err2 = 0.0;
    for (uint32_t c = threadIdx.x; c < KP; c += blockDim.x) {
        const float x = (r < nq && c < dim) ? q[static_cast<size_t>(r) * pitch + c] : 0.f;
        ss += static_cast<double>(x) * x;
        const __nv_bfloat16 b = __float2bfloat16_rn(x);
        const double d = static_cast<double>(__bfloat162float(b)) - static_cast<double>(x);
        err2 += d * d;
        out[static_cast<size_t>(r) * KP + c] = b;
    }
    __shared__ double s_red[32], s_err[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
        err2 += __shfl_xor_sync(0xFFFFFFFFu, err2, o);
    }
    if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5] = ss; s_err[threadIdx.x >> 5] = err2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0, e = 0.0;
        for (uint32_t i = 0; i < (blockDim.x + 31) / 32; ++i) { t += s_red[i]; e += s_err[i]; }
        qn2[r] = static_cast<float>(t);
        if (eq) eq[r] = t > 0.0 ? __double2float_ru(sqrt(e / t) * 1.000001) : 0.f;   // E_q = ‖q̃−q‖/‖q‖, rounded up
    }
    pdl_wait();
    (void)nq_pad;
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static bool make_map(CUtensorMap* m, const void* base, uint64_t rows, uint32_t KP, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t gdim[2] = {KP, rows};
    const cuuint64_t gstr[1] = {static_cast<cuuint64_t>(KP) * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(tc::BK), box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// CTA pairs (tcgen05 cta_group::2) are the default whenever the batch has an even number of 128-query blocks
// (an odd count would leave one CTA of a pair scanning for nobody); VL_TC_PAIR=0 turns them off, VL_TC_PAIR=1 forces
// them, VL_TC_CLUSTER=2|4 selects the multicast clusters of independent CTAs instead.
static int tc_pair_mode() {   // -1 = by batch shape, 0 = never, 1 = always
    static const int mode = [] {
        const char* e = std::getenv("VL_TC_PAIR");
        return !e ? -1 : (e[0] == '1' ? 1 : 0);
    }();
    return mode;
}

int tc_cluster_size() {
    static int cs = -1;
    if (cs < 0) {
        cs = 1;
        if (const char* e = std::getenv("VL_TC_CLUSTER")) cs = atoi(e);
        if (cs != 1 && cs != 2 && cs != 4) cs = 1;
    }
    return cs;
}

static bool tc_use_pair(uint32_t qblocks) {
    const int mode = tc_pair_mode();
    if (mode >= 0) return mode == 1;
    return tc_cluster_size() == 1 && qblocks % 2 == 0;
}

void tc_state_free(TcState* t) {
    if (!t) return;
    cudaFree(t->rows_norm); cudaFree(t->rows_raw); cudaFree(t->sq_norm); cudaFree(t->q_bf16); cudaFree(t->qn2);
    cudaFree(t->eq); cudaFree(t->ex_bits);
    *t = TcState();
}

// bring the bf16 mirror(s) needed by `metric` up to date with the arena (rows [0, n))
cudaError_t tc_prepare(TcState* t, const FlatView& v, uint64_t arena_cap, int metric, uint32_t nq, cudaStream_t s) {
    if (v.dim > TC_MAX_DIM || !encode_fn()) { t->usable = false; return cudaSuccess; }
    const uint32_t KP = (v.dim + tc::BK - 1) / tc::BK * tc::BK;
    cudaError_t e;
    if (t->cap < arena_cap || t->KP != KP) {  // (re)allocate lazily per mirror below
        cudaFree(t->rows_norm); cudaFree(t->rows_raw); cudaFree(t->sq_norm);
        t->rows_norm = t->rows_raw = nullptr; t->sq_norm = nullptr;
        if (t->ex_bits) cudaMemsetAsync(t->ex_bits, 0, 3 * sizeof(uint32_t), s);
        t->built_norm = t->built_raw = 0;
        t->cap = arena_cap;
        t->KP = KP;
    }
    const bool cosine = metric == COSINE;
    __nv_bfloat16** mirror = reinterpret_cast<__nv_bfloat16**>(cosine ? &t->rows_norm : &t->rows_raw);
    uint64_t* built = cosine ? &t->built_norm : &t->built_raw;
    if (!*mirror) {
        if ((e = cudaMalloc(mirror, t->cap * KP * 2)) != cudaSuccess) { t->usable = false; cudaGetLastError(); return cudaSuccess; }
        *built = 0;
    }
    if (!t->ex_bits) {
        if ((e = cudaMalloc(&t->ex_bits, 3 * sizeof(uint32_t))) != cudaSuccess) { t->usable = false; cudaGetLastError(); return cudaSuccess; }
        cudaMemsetAsync(t->ex_bits, 0, 3 * sizeof(uint32_t), s);
    }
    if (!cosine && !t->sq_norm) {
        if ((e = cudaMalloc(&t->sq_norm, t->cap * 4)) != cudaSuccess) { t->usable = false; cudaGetLastError(); return cudaSuccess; }
    }
    if (*built < v.n) {
        const uint64_t m = v.n - *built;
        uint64_t blocks = std::min<uint64_t>((m + 7) / 8, 148 * 16);
        tc::to_bf16_rows_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(
            v.rows, v.inv_norm, *built, m, v.dim, v.pitch, KP, cosine ? 1 : 0, *mirror, cosine ? nullptr : t->sq_norm,
            t->ex_bits + (cosine ? 0 : 1), cosine ? nullptr : t->ex_bits + 2);
        *built = v.n;
        t->maps_n = 0;  // force re-encode
    }
    const uint32_t nq_pad = (nq + tc::BM - 1) / tc::BM * tc::BM;
    if (t->q_cap < nq_pad) {
        cudaFree(t->q_bf16); cudaFree(t->qn2); cudaFree(t->eq);
        t->q_bf16 = nullptr; t->qn2 = nullptr; t->eq = nullptr;
        if ((e = cudaMalloc(&t->q_bf16, static_cast<size_t>(nq_pad) * KP * 2)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&t->qn2, static_cast<size_t>(nq_pad) * 4)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&t->eq, 2 * static_cast<size_t>(nq_pad) * 4)) != cudaSuccess) return e;   // [2][q_cap]: see BatchTensor::parity
        t->q_cap = nq_pad;
    }
    t->usable = true;
    return cudaGetLastError();
}

template <int METRIC, int CS, bool PAIR = false, bool STREAM_A = false>
static cudaError_t launch_tc(const CUtensorMap& mx, const CUtensorMap& mq, const tc::Params& p, int grid, cudaStream_t s) {
    auto kern = tc::batch_scan_tc_kernel<METRIC, CS, PAIR, STREAM_A>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_TOTAL);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(tc::THREADS);
    cfg.dynamicSmemBytes = tc::SMEM_TOTAL;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = CS;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CS > 1 ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kern, mx, mq, p);
}

cudaError_t batch_scan_tensor(const FlatView& v, const BatchTensor& bt, const float* d_q, uint32_t nq, int metric,
                              uint32_t lo, uint32_t hi, const BatchWork& w, cudaStream_t s, bool first, uint32_t mode) {
    TcState* t = static_cast<TcState*>(bt.scratch);
    if (!t || !t->usable) return cudaErrorNotSupported;
    const uint32_t KP = t->KP;
    const uint32_t nq_pad = (nq + tc::BM - 1) / tc::BM * tc::BM;
    const bool pair = tc_use_pair(nq_pad / tc::BM);
    const bool stream_a = KP > tc::KCH_MAX * tc::BK;   // wide rows: the query block's K-chunks go through the ring
    const int CS = pair ? 2 : (stream_a ? 1 : tc_cluster_size());
    if (first) {  // first scan of a batch: convert the queries, (re)encode the maps
        {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(nq_pad);
            cfg.blockDim = dim3(128);
            cfg.stream = s;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = bt.chain_batches ? 1 : 0;
            cudaError_t e = cudaLaunchKernelEx(&cfg, tc::to_bf16_queries_kernel, d_q, nq, nq_pad, v.dim, v.pitch, KP,
                                               static_cast<__nv_bfloat16*>(t->q_bf16), t->qn2,
                                               t->eq + static_cast<size_t>(bt.parity & 1u) * t->q_cap, w.count, w.tau, w.qflags);
            if (e != cudaSuccess) return e;
        }
        const void* mirror = metric == COSINE ? t->rows_norm : t->rows_raw;
        if (t->maps_n != v.n || t->maps_base != mirror || t->maps_cs != CS) {
            if (!make_map(&t->map_x, mirror, v.n, KP, tc::BN / CS)) return cudaErrorUnknown;
            t->maps_n = v.n; t->maps_base = mirror; t->maps_cs = CS;
        }
        if (t->mapq_rows != nq_pad || t->mapq_base != t->q_bf16) {
            if (!make_map(&t->map_q, t->q_bf16, nq_pad, KP, tc::BM)) return cudaErrorUnknown;
            t->mapq_rows = nq_pad; t->mapq_base = t->q_bf16;
        }
    }
    tc::Params p;
    p.sq_norm = t->sq_norm; p.qn2 = t->qn2; p.tau = w.tau; p.cand = w.cand; p.count = w.count; p.qflags = w.qflags;
    p.capq = w.capq; p.nq = nq; p.row_lo = lo; p.row_hi = hi; p.kch = KP / tc::BK;
    p.qblocks = nq_pad / tc::BM;
    p.tiles = (hi - lo + tc::BN - 1) / tc::BN;
    p.direct = mode == SCAN_DIRECT ? 1u : 0u;
    p.groupmax = mode == SCAN_GROUPMAX ? 1u : 0u;
    p.gmax = w.gmax;
    p.pdl_first = first ? 1u : 0u;
    if (p.direct && !(lo == 0 && hi <= w.capq)) return cudaErrorInvalidValue;
    if (p.groupmax && (!w.gmax || (hi - lo + 31) / 32 > GMAX_STRIDE)) return cudaErrorInvalidValue;
    int sms = 148;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t qgroups = (p.qblocks + CS - 1) / CS;
    // clusters per query group: as many as fit (cluster placement may strand a few SMs for CS = 4)
    uint32_t max_clusters = static_cast<uint32_t>(sms) / CS;
    if (CS == 4) max_clusters = std::min<uint32_t>(max_clusters, 33);
    uint32_t per_group = std::max<uint32_t>(1, max_clusters / qgroups);
    per_group = std::min<uint32_t>(per_group, p.tiles);
    const int grid = static_cast<int>(per_group * qgroups * CS);
#define VL_TC_LAUNCH(M)                                                                            \
    (stream_a ? (pair ? launch_tc<M, 2, true, true>(t->map_x, t->map_q, p, grid, s)               \
                      : launch_tc<M, 1, false, true>(t->map_x, t->map_q, p, grid, s))             \
     : CS == 1 ? launch_tc<M, 1>(t->map_x, t->map_q, p, grid, s)                                  \
     : CS == 2 ? (pair ? launch_tc<M, 2, true>(t->map_x, t->map_q, p, grid, s)                    \
                       : launch_tc<M, 2>(t->map_x, t->map_q, p, grid, s))                         \
               : launch_tc<M, 4>(t->map_x, t->map_q, p, grid, s))
    switch (metric) {
        case COSINE: return VL_TC_LAUNCH(COSINE);
        case EUCLIDEAN: return VL_TC_LAUNCH(EUCLIDEAN);
        case DOT: return VL_TC_LAUNCH(DOT);
        default: return cudaErrorNotSupported;
    }
#undef VL_TC_LAUNCH
}

}  // namespace vl
