// batch_scan.cu — batched flat scan (B queries at once), CUDA-core tile kernel + the staged
// threshold pipeline shared with the tensor-core kernel (batch_tc.cu).
//
// The reference has no batched search; this is the additive path the north star
// asks for.  Arithmetic per (row, query) pair is that of src/lib.rs:425-572 in fp32; the B×N score
// matrix (4 GB at B=1024, N=1M) is never materialised:
//
//   stage 0: rows [0, S0)      scored, everything kept     → per-query K'-th best  τ0
//   stage 1: rows [S0, S1)     only scores >= τ0 are kept  → τ1 (K'-th best of the first S1 rows)
//   stage 2: rows [S1, N)      only scores >= τ1 are kept  (≈ K'·N/S1 survivors per query)
//   then    : per query sort the survivors, keep K', fp64 re-score in reference order, rank, certify
//
// Survivors go to per-query candidate buffers with one global atomicAdd each (≈0.1 % of the pairs
// after stage 0).  Manhattan has no tensor-core form and always runs here (bound: FP32 issue rate,
// 2 instructions per element); cosine / L2 / dot run here when the tensor path is not applicable.
//
// Tile kernel: CTA = 256 threads, 128 rows × 128 queries, 8×8 micro-tile per thread with rows and
// queries interleaved by 16 (conflict-free shared-memory reads, query reads are warp broadcasts),
// K chunks of 32 columns staged through registers into transposed, odd-pitch shared tiles.
#include <cmath>
#include <cstdlib>

#include "batch.h"
#include "rescore.cuh"
#include "tc_state.h"
#include "topk.cuh"

namespace vl {

constexpr int BT_THREADS = 256;
constexpr int BT_ROWS = 128, BT_QS = 128, BT_K = 32, BT_LD = 129;

__global__ void batch_init_kernel(uint32_t nq, uint32_t* count, float* tau, uint32_t* qflags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) {
        count[i] = 0u;
        tau[i] = -INFINITY;
        qflags[i] = 0u;
    }
}

template <int METRIC>
__global__ void __launch_bounds__(BT_THREADS, 2)
batch_scan_cc_kernel(const float* __restrict__ rows, const float* __restrict__ inv_norm,
                     const float* __restrict__ queries, uint32_t row_lo, uint32_t row_hi, uint32_t pitch,
                     uint32_t nq, const float* __restrict__ tau, uint64_t* cand, uint32_t* count,
                     uint32_t capq, uint32_t* qflags) {
    extern __shared__ float smem[];
    float* s_x = smem;                          // [2][BT_K][BT_LD]
    float* s_q = smem + 2 * BT_K * BT_LD;       // [2][BT_K][BT_LD]
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;     // rows tx + 16 i, queries ty + 16 j
    const uint32_t row0 = row_lo + blockIdx.x * BT_ROWS;
    const uint32_t q0 = blockIdx.y * BT_QS;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    // global → register staging: 4 float4 of the X tile and 4 of the Q tile per thread
    const int lr = tid >> 3, lc4 = tid & 7;     // 32 tile-rows per pass, 8 float4 per tile-row
    float4 gx[4], gq[4];
    auto gload = [&](uint32_t k0) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const uint32_t r = row0 + p * 32 + lr, c = k0 + lc4 * 4;
            gx[p] = (r < row_hi && c < pitch) ? __ldg(reinterpret_cast<const float4*>(rows + static_cast<size_t>(r) * pitch + c))
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
            const uint32_t qq = q0 + p * 32 + lr;
            gq[p] = (qq < nq && c < pitch) ? __ldg(reinterpret_cast<const float4*>(queries + static_cast<size_t>(qq) * pitch + c))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto sstore = [&](int buf) {
        float* x = s_x + buf * BT_K * BT_LD;
        float* q = s_q + buf * BT_K * BT_LD;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int r = p * 32 + lr, c = lc4 * 4;
            x[(c + 0) * BT_LD + r] = gx[p].x; x[(c + 1) * BT_LD + r] = gx[p].y;
            x[(c + 2) * BT_LD + r] = gx[p].z; x[(c + 3) * BT_LD + r] = gx[p].w;
            q[(c + 0) * BT_LD + r] = gq[p].x; q[(c + 1) * BT_LD + r] = gq[p].y;
            q[(c + 2) * BT_LD + r] = gq[p].z; q[(c + 3) * BT_LD + r] = gq[p].w;
        }
    };

    const uint32_t nchunks = (pitch + BT_K - 1) / BT_K;
    gload(0);
    sstore(0);
    __syncthreads();
    for (uint32_t ch = 0; ch < nchunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < nchunks) gload((ch + 1) * BT_K);
        const float* x = s_x + buf * BT_K * BT_LD + tx;
        const float* q = s_q + buf * BT_K * BT_LD + ty;
#pragma unroll 8
        for (int k = 0; k < BT_K; ++k) {
            float xv[8], qv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) xv[i] = x[k * BT_LD + 16 * i];
#pragma unroll
            for (int j = 0; j < 8; ++j) qv[j] = q[k * BT_LD + 16 * j];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (METRIC == COSINE || METRIC == DOT) {
                        acc[i][j] = fmaf(xv[i], qv[j], acc[i][j]);
                    } else if (METRIC == EUCLIDEAN) {
                        const float d = xv[i] - qv[j];
                        acc[i][j] = fmaf(d, d, acc[i][j]);
                    } else {
                        acc[i][j] += fabsf(xv[i] - qv[j]);
                    }
                }
        }
        if (ch + 1 < nchunks) {
            sstore(buf ^ 1);
            __syncthreads();
        }
    }

    // epilogue: threshold filter, survivors to the per-query candidate buffers
    float tq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t qq = q0 + ty + 16 * j;
        tq[j] = qq < nq ? __ldg(tau + qq) : INFINITY;
    }
    bool nonfinite = false;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t r = row0 + tx + 16 * i;
        if (r >= row_hi) continue;
        const float invn = METRIC == COSINE ? __ldg(inv_norm + r) : 1.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float s = acc[i][j];
            if (METRIC == COSINE) s *= invn;
            if (METRIC == EUCLIDEAN || METRIC == MANHATTAN) s = -s;
            const uint32_t qq = q0 + ty + 16 * j;
            if (qq >= nq) continue;
            if (!isfinite(s)) nonfinite = true;
            if (s >= tq[j] || !(s == s)) {
                const uint32_t idx = atomicAdd(count + qq, 1u);
                if (idx < capq) cand[static_cast<size_t>(qq) * capq + idx] = make_key(s, r);
                else atomicOr(qflags + qq, FLAG_OVERFLOW);
            }
        }
    }
    if (nonfinite) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t qq = q0 + ty + 16 * j;
            if (qq < nq) atomicOr(qflags + qq, FLAG_NONFINITE);  // conservative: whole query tile
        }
    }
}

// per query: keep the best Kp candidates and publish the new threshold.  8-pass MSB radix select
// over the 64-bit keys (unique: they embed the row position) finds the exact Kp-th largest key T in
// shared memory; the Kp keys >= T are written back unordered (the rescore kernel ranks them).
__global__ void __launch_bounds__(256) batch_select_kernel(uint64_t* cand, uint32_t* count, float* tau,
                                                           uint32_t capq, int Kp, uint32_t n_override) {
    extern __shared__ __align__(16) unsigned char sm[];
    uint64_t* s_k = reinterpret_cast<uint64_t*>(sm);  // [capq]
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_wsum[8];
    __shared__ uint32_t s_bin, s_need, s_out, s_all;
    __shared__ unsigned long long s_min;
    const uint32_t q = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_wait();                              // the scan that produced the candidates is complete and visible
    if (tid == 0) pdl_launch_dependents();   // the next stage's scan may be launched (its epilogue waits for us)
    const int n = static_cast<int>(n_override ? n_override : min(count[q], capq));
    if (n <= Kp) {
        if (tid == 0) count[q] = n;  // clamp (overflow already flagged); threshold unchanged
        return;
    }
    uint64_t* mine = cand + static_cast<size_t>(q) * capq;
    for (int i = tid; i < n; i += 256) s_k[i] = mine[i];
    if (tid == 0) { s_need = static_cast<uint32_t>(Kp); s_out = 0; }
    uint64_t prefix = 0;
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        s_hist[tid] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += 256) {
            const uint64_t key = s_k[i];
            if (pass == 0 || (key >> (shift + 8)) == (prefix >> (shift + 8)))
                atomicAdd(&s_hist[(key >> shift) & 255], 1u);
        }
        __syncthreads();
        const uint32_t need = s_need;  // read before anybody may update it (barrier below)
        // suffix sums from the top bin down: thread t owns bin 255 − t
        const uint32_t mycount = s_hist[255 - tid];
        uint32_t incl = mycount;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t offset = 0;
        for (int w2 = 0; w2 < warp; ++w2) offset += s_wsum[w2];
        incl += offset;
        if (incl >= need && incl - mycount < need) {  // exactly one thread: the bin holding the need-th key
            s_bin = 255 - tid;
            s_need = need - (incl - mycount);
            s_all = (mycount == need - (incl - mycount)) ? 1u : 0u;   // every key of the bin is needed
            s_min = ~0ull;
        }
        __syncthreads();
        prefix |= static_cast<uint64_t>(s_bin) << shift;
        if (s_all && pass < 7) {
            // early exit (typically after 3 passes): the Kp-th largest key is the smallest key of this bin
            for (int i = tid; i < n; i += 256) {
                const uint64_t key = s_k[i];
                if ((key >> shift) == (prefix >> shift)) atomicMin(&s_min, static_cast<unsigned long long>(key));
            }
            __syncthreads();
            prefix = s_min;
            break;
        }
    }
    // prefix is now the exact Kp-th largest key
    for (int i = tid; i < n; i += 256) {
        const uint64_t key = s_k[i];
        if (key >= prefix) mine[atomicAdd(&s_out, 1u)] = key;
    }
    if (tid == 0) {
        count[q] = Kp;
        tau[q] = key_score(prefix);
    }
}

// Threshold estimation (tensor path, large stores): tau[q] = the Kp-th largest of the G group maxima written by the
// SCAN_GROUPMAX stage — Kp disjoint groups each hold a row scoring >= it, so it is a valid lower bound of the Kp-th
// best score of the store.  One warp per query; the exact Kp-th largest is built bit by bit on the orderable
// pattern (32 count-and-vote steps over NV = G / 32 values per lane; G = 512 → 16 compares per step, not 64).
template <int NV>
__global__ void __launch_bounds__(256) batch_tau_kernel(const float* __restrict__ gmax, uint32_t G, int Kp, float* tau,
                                                        uint32_t nq) {
    const int lane = threadIdx.x & 31;
    const uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    pdl_wait();
    if (threadIdx.x == 0) pdl_launch_dependents();
    if (q >= nq) return;
    static_assert(NV * 32 <= static_cast<int>(GMAX_STRIDE), "group maxima per query");
    uint32_t v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const uint32_t g = static_cast<uint32_t>(i) * 32 + lane;
        v[i] = g < G ? f32_orderable(__ldcg(gmax + static_cast<size_t>(q) * GMAX_STRIDE + g)) : 0u;
    }
    uint32_t T = 0u;
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t c = T | (1u << bit);
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < NV; ++i) cnt += v[i] >= c;
        cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
        if (cnt >= Kp) T = c;
    }
    if (lane == 0) tau[q] = (G >= static_cast<uint32_t>(Kp) && T != 0u) ? orderable_f32(T) : -INFINITY;
}

// per query: load the (sorted, <= Kp) survivors and run the shared rescore / rank / certify phases
__global__ void __launch_bounds__(FIN_THREADS) batch_rescore_kernel(FinalizeParams p, const uint32_t* count,
                                                                       const uint32_t* qflags, uint32_t capq) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(smem_raw);
    double* s_exact = reinterpret_cast<double*>(s_keys + SCAN_CAP);
    uint32_t* s_pos = reinterpret_cast<uint32_t*>(s_exact + KP_MAX);
    float* s_q = reinterpret_cast<float*>(s_pos + KP_MAX);
    float* s_tile = s_q + p.CH;
    const uint32_t qi = blockIdx.x;
    const int tid = threadIdx.x;
    const int n = static_cast<int>(min(count[qi], static_cast<uint32_t>(p.Kp)));
    // count <= Kp here: either select truncated it, or fewer than Kp survivors exist → sort by counting
    const uint64_t* mine = p.cand + static_cast<size_t>(qi) * capq;
    uint64_t* s_tmp = reinterpret_cast<uint64_t*>(s_tile);  // staging tile is free until the rescore
    uint64_t key = 0;
    if (tid < n) {
        key = mine[tid];
        s_tmp[tid] = key;
    }
    __syncthreads();
    if (tid < n) {
        int r = 0;
        for (int j = 0; j < n; ++j) r += s_tmp[j] > key;
        s_keys[r] = key;
    }
    __syncthreads();
    rescore_rank_certify(p, qi, n, s_keys, s_exact, s_pos, s_q, s_tile, qflags[qi]);
}

// ---------------------------------------------------------------------------------------------
// Batched rescore, streaming form (Kp <= 512): one CTA of KpR + 32 threads per query, KpR = Kp rounded
// up to a warp.  Thread t < nc owns candidate t and walks its row in reference order (a serial f64
// chain, lib.rs:425-572); the row data arrives through a double-buffered cp.async pipeline of CH-column
// chunks ([KpR][CH+4] fp32, (CH+4)/4 odd → conflict-free 128-bit reads), the query is widened to f64 once,
// and the query-only chain Σy² runs on the spare warp.  Small CTAs (9 per SM at Kp = 64) put all B
// queries of a batch on the machine in a single wave, where the 512-thread tile kernel needed 3.5.
constexpr int RS_SPARE = 32;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem))),
                 "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// NBUF: depth of the cp.async chunk ring.  2 keeps the CTA small (9 per SM: a 1024-query batch in one wave); a few
// queries (the combiner's cohorts) are latency-bound instead — 12 chunk loads one DRAM round trip apart — and run with 4.
template <int METRIC, int NBUF>
__global__ void batch_rescore_stream_kernel(FinalizeParams p, const uint32_t* count, const uint32_t* qflags,
                                            uint32_t capq, int KpR) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int CH = p.CH, TS = CH + 4;
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(smem_raw);                 // [KpR]
    double* s_exact = reinterpret_cast<double*>(s_keys + KpR);                // [KpR]
    double* s_qd = s_exact + KpR;                                             // [pitch]
    uint32_t* s_pos = reinterpret_cast<uint32_t*>(s_qd + p.pitch);            // [KpR]
    float* s_tile = reinterpret_cast<float*>(s_pos + KpR);                    // [NBUF][KpR][TS]
    __shared__ double s_qnorm, s_qn2;
    __shared__ int s_nan;
    const uint32_t qi = blockIdx.x;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    pdl_wait();                              // the last select is complete and visible
    if (tid == 0) pdl_launch_dependents();   // chained batches: the next batch's query conversion may start
    const int nc = static_cast<int>(min(count[qi], static_cast<uint32_t>(p.Kp)));
    // ---- sort the (<= Kp, unordered) survivors by key, descending: rank by counting ------------------
    const uint64_t* mine = p.cand + static_cast<size_t>(qi) * capq;
    uint64_t* s_tmp = reinterpret_cast<uint64_t*>(s_exact);
    uint64_t key = 0;
    if (tid < nc) {
        key = mine[tid];
        s_tmp[tid] = key;
    }
    if (tid == 0) s_nan = 0;
    const float* q = p.queries + static_cast<size_t>(qi) * p.pitch;
    for (uint32_t i = tid; i < p.pitch; i += nthreads) s_qd[i] = i < p.dim ? static_cast<double>(q[i]) : 0.0;
    __syncthreads();
    if (tid < nc) {
        int r = 0;
        for (int j = 0; j < nc; ++j) r += s_tmp[j] > key;
        s_keys[r] = key;
        s_pos[r] = key_pos(key);
    }
    __syncthreads();

    // ---- (2) exact f64 rescore ----------------------------------------------------------------------
    const int nchunks = static_cast<int>((p.dim + CH - 1) / CH);
    // Fixed thread → (row, 16-byte column) assignment for the chunk loads: W4 threads per row, nthreads / W4 rows per
    // pass, the row pointers of this thread's (<= 8) passes computed ONCE — the per-chunk issue is then one predicated
    // cp.async per pass (it used to redo an integer division and a 64-bit multiply per 16 bytes, ~12 of this kernel's
    // 23 µs at 16 queries).
    constexpr int MAXP = 8;
    const int W4 = CH >> 2;
    const int lrow = tid / W4, lcc = tid - lrow * W4, rpp = nthreads / W4;
    const float* rp[MAXP];
#pragma unroll
    for (int u = 0; u < MAXP; ++u) {
        const int r = lrow + u * rpp;
        rp[u] = (lrow < rpp && r < nc) ? p.rows + static_cast<size_t>(s_pos[r]) * p.pitch + lcc * 4 : nullptr;
    }
    auto issue = [&](int c, int buf) {
        const uint32_t c0 = static_cast<uint32_t>(c) * CH;
        const int w4 = (min(static_cast<uint32_t>(CH), p.dim - c0) + 3) >> 2;
        float* dst = s_tile + static_cast<size_t>(buf) * KpR * TS + lrow * TS + lcc * 4;
        if (lcc < w4) {
#pragma unroll
            for (int u = 0; u < MAXP; ++u)
                if (rp[u]) cp_async16(dst + u * rpp * TS, rp[u] + c0);
        }
        cp_async_commit();
    };
    double a0 = 0.0, a1 = 0.0, qn2 = 0.0;
#pragma unroll
    for (int c = 0; c < NBUF - 1; ++c) {   // NBUF − 1 groups in flight before the first wait (empty ones past the end)
        if (c < nchunks) issue(c, c); else cp_async_commit();
    }
    for (int c = 0; c < nchunks; ++c) {
        if (c + NBUF - 1 < nchunks) issue(c + NBUF - 1, (c + NBUF - 1) % NBUF); else cp_async_commit();
        cp_async_wait<NBUF - 1>();
        __syncthreads();
        const uint32_t c0 = static_cast<uint32_t>(c) * CH;
        const int w = static_cast<int>(min(static_cast<uint32_t>(CH), p.dim - c0));
        if (tid < nc) {
            const float4* t4 = reinterpret_cast<const float4*>(s_tile + (static_cast<size_t>(c % NBUF) * KpR + tid) * TS);
            const double* yq = s_qd + c0;
            // one element of the reference's chains (lib.rs:425-572): same operations in the same order, no FMA
            auto step = [&](float xf, double y) {
                const double x = static_cast<double>(xf);
                if (METRIC == COSINE) {
                    a0 = __dadd_rn(a0, __dmul_rn(x, y));
                    a1 = __dadd_rn(a1, __dmul_rn(x, x));
                } else if (METRIC == EUCLIDEAN) {
                    const double d = __dsub_rn(x, y);
                    a0 = __dadd_rn(a0, __dmul_rn(d, d));
                } else if (METRIC == MANHATTAN) {
                    a0 = __dadd_rn(a0, fabs(__dsub_rn(x, y)));
                } else {
                    a0 = __dadd_rn(a0, __dmul_rn(x, y));
                }
            };
            // Straight-line blocks of 16 elements (no bounds checks, 128-bit shared loads): the kernel is one serial
            // f64 add chain per candidate, so what matters is that nothing but the DADD latency sits between two
            // links — the checked per-element loop spent ~120 cycles per element (17 instructions, a branch each).
            int j = 0;
            for (; j + 16 <= w; j += 16) {
                float4 v[4];
                double2 y2[8];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = t4[(j >> 2) + u];
#pragma unroll
                for (int u = 0; u < 8; ++u) y2[u] = *reinterpret_cast<const double2*>(yq + j + 2 * u);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    step(v[u].x, y2[2 * u].x);
                    step(v[u].y, y2[2 * u].y);
                    step(v[u].z, y2[2 * u + 1].x);
                    step(v[u].w, y2[2 * u + 1].y);
                }
            }
            if (j < w) {   // tail of the last chunk (dim not a multiple of 16)
                const float* t1 = reinterpret_cast<const float*>(t4);
                for (; j < w; ++j) step(t1[j], yq[j]);
            }
        } else if (tid == nthreads - 1) {   // query-only chain Σy² (cosine: lib.rs:433; others: certificate)
            const double* yq = s_qd + c0;
            int j = 0;
            for (; j + 16 <= w; j += 16) {
                double2 y2[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) y2[u] = *reinterpret_cast<const double2*>(yq + j + 2 * u);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    qn2 = __dadd_rn(qn2, __dmul_rn(y2[u].x, y2[u].x));
                    qn2 = __dadd_rn(qn2, __dmul_rn(y2[u].y, y2[u].y));
                }
            }
            for (; j < w; ++j) qn2 = __dadd_rn(qn2, __dmul_rn(yq[j], yq[j]));
        }
        __syncthreads();   // buffer (c % NBUF) may be refilled by the next iteration's issue
    }
    if (tid == nthreads - 1) {
        s_qn2 = qn2;
        s_qnorm = __dsqrt_rn(qn2);
    }
    __syncthreads();
    if (tid < nc) {
        double sc;
        if (METRIC == COSINE) {
            const double na = __dsqrt_rn(a1), nb = s_qnorm;
            sc = (na == 0.0 || nb == 0.0) ? 0.0 : __ddiv_rn(a0, __dmul_rn(na, nb));
        } else if (METRIC == EUCLIDEAN) {
            sc = sim_from_l2(a0);
        } else if (METRIC == MANHATTAN) {
            sc = sim_from_l1(a0);
        } else {
            sc = a0;
        }
        s_exact[tid] = sc;
        if (sc != sc) s_nan = 1;
    }
    __syncthreads();
    rank_and_certify(p, qi, nc, s_keys, s_exact, s_pos, &s_qnorm, &s_nan, qflags[qi]);
}

static size_t rescore_stream_smem(int KpR, int CH, uint32_t pitch, int nbuf) {
    return static_cast<size_t>(KpR) * (8 + 8 + 4) + static_cast<size_t>(pitch) * 8 +
           static_cast<size_t>(nbuf) * KpR * (CH + 4) * sizeof(float);
}

template <int METRIC, int NBUF>
static cudaError_t launch_rescore_stream(const FinalizeParams& p, const BatchWork& w, uint32_t nq, int KpR,
                                         size_t smem, cudaStream_t s) {
    auto kern = batch_rescore_stream_kernel<METRIC, NBUF>;
    if (smem > 40 * 1024) {   // dynamic + the kernel's static shared memory must stay under the 48 KB default
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nq);
    cfg.blockDim = dim3(KpR + RS_SPARE);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;   // joins the batch's PDL chain (the kernel starts with griddepcontrol.wait)
    return cudaLaunchKernelEx(&cfg, kern, p, static_cast<const uint32_t*>(w.count), static_cast<const uint32_t*>(w.qflags),
                              w.capq, KpR);
}

static size_t rescore_smem(int Kp, int CH) {
    return SCAN_CAP * sizeof(uint64_t) + KP_MAX * sizeof(double) + KP_MAX * sizeof(uint32_t) +
           static_cast<size_t>(CH) * sizeof(float) + static_cast<size_t>(Kp) * (CH + 1) * sizeof(float);
}

template <int METRIC>
static cudaError_t launch_cc(const FlatView& v, const float* d_q, uint32_t nq, uint32_t lo, uint32_t hi,
                             const BatchWork& w, cudaStream_t s) {
    const size_t smem = 4 * BT_K * BT_LD * sizeof(float);
    static bool attr[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(batch_scan_cc_kernel<METRIC>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        attr[dev & 63] = true;
    }
    dim3 grid((hi - lo + BT_ROWS - 1) / BT_ROWS, (nq + BT_QS - 1) / BT_QS);
    batch_scan_cc_kernel<METRIC><<<grid, BT_THREADS, smem, s>>>(v.rows, v.inv_norm, d_q, lo, hi, v.pitch, nq, w.tau,
                                                                w.cand, w.count, w.capq, w.qflags);
    return cudaGetLastError();
}

cudaError_t batch_scan_cuda_cores(const FlatView& v, const float* d_q, uint32_t nq, int metric, uint32_t lo,
                                  uint32_t hi, const BatchWork& w, cudaStream_t s) {
    switch (metric) {
        case COSINE: return launch_cc<COSINE>(v, d_q, nq, lo, hi, w, s);
        case EUCLIDEAN: return launch_cc<EUCLIDEAN>(v, d_q, nq, lo, hi, w, s);
        case MANHATTAN: return launch_cc<MANHATTAN>(v, d_q, nq, lo, hi, w, s);
        case DOT: return launch_cc<DOT>(v, d_q, nq, lo, hi, w, s);
        default: return cudaErrorInvalidValue;
    }
}

// probability that one row beats the Kp-th largest of G group maxima (groups of 32 rows)
static double p_row_of(int Kp, uint32_t G) { return 1.0 - pow(1.0 - static_cast<double>(Kp) / G, 1.0 / 32.0); }

cudaError_t launch_batch_flat(const FlatView& v, const float* d_queries, uint32_t nq, uint32_t k, int metric,
                              int Kp, const BatchWork& w, const SearchOut& out, const BatchTensor* tc,
                              uint64_t* launches, cudaStream_t s, int kp_base) {
    cudaError_t e;
    static bool attr[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr[dev & 63]) {
        if ((e = cudaFuncSetAttribute(batch_rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(batch_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)) != cudaSuccess) return e;
        attr[dev & 63] = true;
    }
    const bool use_tc = tc && tc->usable && metric != MANHATTAN;
    // (tensor path: the per-query state is reset by the query-conversion kernel, batch_tc.cu)
    if (!use_tc) batch_init_kernel<<<(nq + 255) / 256, 256, 0, s>>>(nq, w.count, w.tau, w.qflags);
    uint64_t nl = 1;
    // Geometric stage schedule: after a select, tau is the K'-th best of the first S rows, so the next
    // stage over rows [S, r·S) keeps ≈ K'·(r−1) survivors per query; r is chosen so that this stays
    // below half the candidate buffer (capq = 4096 → r = 16 for K' = 64, 9 for K' = 224, 4 for K' = 512).
    uint32_t sel_len = 2;
    while (sel_len < w.capq) sel_len <<= 1;
    uint32_t ratio = w.capq / (2u * static_cast<uint32_t>(Kp));
    ratio = ratio < 2u ? 2u : (ratio > 16u ? 16u : ratio);
    static const int forced_ratio = [] { const char* e = getenv("VL_BATCH_RATIO"); return e ? atoi(e) : 0; }();
    if (forced_ratio >= 2) ratio = static_cast<uint32_t>(forced_ratio);
    uint32_t lo = 0;
    uint64_t hi64 = min(v.n, min(w.capq, 4096u));   // stage 0: everything is kept
    // Tensor path on large stores: instead of keeping (and selecting from) every score of the first 4096 rows, an
    // ESTIMATION stage scans the first S0 = 32·G rows writing only the maximum of every 32-row group; the Kp-th
    // largest group maximum is the first threshold — as tight as the exact Kp-th best of S_eq rows, where a group
    // exceeds it with probability Kp/G, i.e. a row with p = 1 − (1 − Kp/G)^(1/32): S_eq = Kp/p ≈ 15K rows for
    // Kp = 64, G = 512.  Those rows are scanned again by the first filtered stage (1.6 % of a 1M-row store), which can
    // therefore be ~4x longer at the same survivor count: one stage and the 4096-key select less per batch.
    static const bool no_estimate = getenv("VL_BATCH_NO_ESTIMATE") != nullptr;
    const uint32_t G = Kp <= 256 ? 512u : GMAX_STRIDE;
    const uint32_t S0 = G * 32u;
    double balanced = 0.0;   // > 0: the common ratio of the balanced schedule (estimation path only)
    const bool estimate = use_tc && w.gmax && !no_estimate && v.n >= 4ull * S0 && static_cast<uint32_t>(Kp) * 2u <= G;
    if (estimate) {
        if ((e = batch_scan_tensor(v, *tc, d_queries, nq, metric, 0, S0, w, s, true, SCAN_GROUPMAX)) != cudaSuccess) return e;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((nq + 7) / 8);
        cfg.blockDim = dim3(256);
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        e = G <= 512u ? cudaLaunchKernelEx(&cfg, batch_tau_kernel<16>, static_cast<const float*>(w.gmax), G, Kp, w.tau, nq)
                      : cudaLaunchKernelEx(&cfg, batch_tau_kernel<static_cast<int>(GMAX_STRIDE / 32)>,
                                           static_cast<const float*>(w.gmax), G, Kp, w.tau, nq);
        if (e != cudaSuccess) return e;
        nl += 2;
        // VL_BATCH_STAGES=m (experiments): balanced schedule — the filtered stages end at S_eq·r, S_eq·r², … n with the
        // same ratio r = (n / S_eq)^(1/m).  Measured at 1M × 384, B = 1024 (profiles/r02_batch_schedule_sweeps.txt):
        // m = 2 (r = 8) 0.633 ms, m = 3 (r = 4) 0.618 ms, the default geometric ratio 16 below 0.622 ms — back to back
        // the batch time follows the energy of the MMA stage, not the stage structure.
        const double s_eq = static_cast<double>(Kp) / p_row_of(Kp, G);
        static const int forced_stages = [] { const char* e = getenv("VL_BATCH_STAGES"); return e ? atoi(e) : 0; }();
        if (forced_stages > 0 && static_cast<double>(v.n) > s_eq)
            balanced = std::max(2.0, pow(static_cast<double>(v.n) / s_eq, 1.0 / forced_stages) * 1.0001);
        hi64 = static_cast<uint64_t>(s_eq * (balanced > 0.0 ? balanced : static_cast<double>(ratio)));
    }
    bool first = !estimate;
    while (lo < v.n) {
        uint32_t hi = static_cast<uint32_t>(hi64 > v.n ? v.n : hi64);
        const uint32_t mode = (use_tc && first && hi <= w.capq) ? SCAN_DIRECT : SCAN_FILTER;
        if (lo > 0 || estimate) hi = min(v.n, (hi + 255u) & ~255u);   // filtered stages end on 256-row tile boundaries
        if (use_tc) e = batch_scan_tensor(v, *tc, d_queries, nq, metric, lo, hi, w, s, first, mode);
        else e = batch_scan_cuda_cores(v, d_queries, nq, metric, lo, hi, w, s);
        if (e != cudaSuccess) return e;
        first = false;
        const uint32_t n_override = mode == SCAN_DIRECT ? hi - lo : 0u;  // tensor stage 0 of a small store is atomics-free
        {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(nq);
            cfg.blockDim = dim3(256);
            cfg.dynamicSmemBytes = sel_len * sizeof(uint64_t);
            cfg.stream = s;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            if ((e = cudaLaunchKernelEx(&cfg, batch_select_kernel, w.cand, w.count, w.tau, w.capq, Kp, n_override)) != cudaSuccess)
                return e;
        }
        nl += 2;
        lo = hi;
        hi64 = balanced > 0.0 ? static_cast<uint64_t>(static_cast<double>(hi) * balanced) : static_cast<uint64_t>(hi) * ratio;
        if (balanced > 0.0 && hi64 < v.n && v.n - hi64 < 4096u) hi64 = v.n;   // no sliver stage from rounding
    }
    // rescore + certify
    const size_t budget = 180 * 1024;
    int CH = static_cast<int>((v.dim + 3) / 4 * 4);
    while (CH > 4 && rescore_smem(Kp, CH) > budget) CH = (CH / 2 + 3) / 4 * 4;
    if (nq >= 64 && CH > 128) CH = 128;  // many queries: smaller staging tile → 4 CTAs per SM instead of 1
    FinalizeParams p;
    p.rows = v.rows; p.ids = v.ids; p.stats = v.stats; p.queries = d_queries;
    p.id_base = v.id_base; p.pos_base = v.pos_base;
    p.n = v.n; p.dim = v.dim; p.pitch = v.pitch; p.k = k;
    p.metric = metric; p.Kp = Kp; p.grid_x = 0; p.CH = CH;
    p.cand = w.cand; p.cand_count = nullptr; p.cand_max = nullptr; p.ctl = nullptr;
    p.out_ids = out.ids; p.out_scores = out.scores; p.out_pos = out.pos;
    p.out_counts = out.counts; p.out_flags = out.flags;
    p.eps_scale = 1.0;
    p.tc_abs = use_tc ? tc->tc_abs : 0.0;
    if (use_tc) {   // measured rounding-error norms of the mirror that was scanned and of this batch's queries
        const TcState* ts = static_cast<const TcState*>(tc->scratch);
        p.e_x = ts->ex_bits ? ts->ex_bits + (metric == COSINE ? 0 : 1) : nullptr;
        p.e_q = ts->eq + static_cast<size_t>(tc->parity & 1u) * ts->q_cap;
    }
    p.peers = out.peers;
    p.kp_base = kp_base;
    const int KpR = (Kp + 31) & ~31;
    if (nq >= 8 && KpR + RS_SPARE <= 1024) {   // many queries: small streaming CTAs, one wave
        // (16-column chunks would let the CTA fit beside a resident scan CTA of a chained next batch — measured
        // SLOWER, 0.732 vs 0.705 ms per 1024-query batch: the gathers then compete with the MMA stage's TMA stream)
        p.CH = KpR <= 64 ? 32 : 16;
        const bool deep = nq <= 296;   // at most two CTAs per SM: latency-bound, four chunk loads in flight
        const size_t smem = rescore_stream_smem(KpR, p.CH, v.pitch, deep ? 4 : 2);
#define VL_RESCORE(M) (deep ? launch_rescore_stream<M, 4>(p, w, nq, KpR, smem, s) : launch_rescore_stream<M, 2>(p, w, nq, KpR, smem, s))
        switch (metric) {
            case COSINE: e = VL_RESCORE(COSINE); break;
            case EUCLIDEAN: e = VL_RESCORE(EUCLIDEAN); break;
            case MANHATTAN: e = VL_RESCORE(MANHATTAN); break;
            default: e = VL_RESCORE(DOT); break;
        }
#undef VL_RESCORE
        if (e != cudaSuccess) return e;
    } else {
        batch_rescore_kernel<<<nq, FIN_THREADS, rescore_smem(Kp, CH), s>>>(p, w.count, w.qflags, w.capq);
    }
    nl += 1;
    if (launches) *launches += nl;
    return cudaGetLastError();
}

}  // namespace vl
