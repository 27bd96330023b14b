// flat_scan.cu — single-query flat scan: HBM-bound streaming distance + fused top-K'.
//
// Replaces the hot loop of FlatIndex::search (src/index/flat.rs:106-117: n×calculate(), clone,
// stable sort, truncate) and the four metric loops (src/lib.rs:425-572).
//
// Roofline: HBM.  One kernel template, two scanned copies of the rows:
//   * fp32 arena  [n][pitch] fp32 — algorithmic bytes per query = n·pitch·4 (+ n·4 inv-norms for cosine);
//     one warp reads 8 whole rows per iteration as 128-bit ld.global.nc.L1::no_allocate (3 per lane per
//     384-d row, 512 B coalesced per instruction);
//   * bf16 mirror [n][384] bf16 (cosine: rows pre-scaled by 1/‖row‖; dot / L2 / L1: raw rows, L2 + fp32 ‖row‖²) —
//     HALF the bytes per query (s = 2 B/element, stated in bench.py's roofline); one warp reads
//     16 whole 768-B rows per iteration as 64-bit streaming loads (lane l gets elements 4·(32c + l) … +3 of
//     chunk c, the same elements as the fp32 layout, so the query registers are identical); bf16 → fp32 is a
//     shift / mask, accumulation is fp32 against the fp32 query.  Scores are in the tensor-core path's scan
//     units (cos·‖q‖, x·q, −‖x−q‖² via 2x·q − ‖x‖² − ‖q‖²) and are certified with the same bf16 bound; L1 is
//     −Σ|x̃−q| directly, certified with Σ|x̃−x| <= 2^-8·‖x‖₁ <= 2^-8·√dim·max‖row‖.
// The query lives in registers; per-lane partial sums are reduced with a transposed butterfly; each row's
// fp32 score becomes a 64-bit key (score, ~position) and is appended to a CTA candidate buffer only if it
// beats the running threshold.  The exact f64 score and the final order are produced by flat_finalize.cu.
//
// Fixed cost per launch matters as much as the streaming rate (measured: ~7.3 TB/s streaming + ~26 µs fixed
// at 1M rows before this structure):
//   * every CTA owns an EQUAL contiguous share of the rows (row-granular split: no tile-count imbalance);
//   * EARLY GRID-WIDE THRESHOLD: after its first tile each CTA publishes that tile's best key; two tiles later
//     it reads all G published keys, takes the minimum over K' disjoint groups of the group maxima — K'
//     distinct rows reach it, so it is a valid lower bound of the global K'-th best key — and drops
//     everything below it.  From then on ~0.5 % of the rows are appended, the buffer never fills, and neither
//     the warm-up nor the final compaction has to sort a thousand keys while the SM's loads stall;
//   * the running grid-wide threshold (QueryCtl::tau, an atomicMax lower bound) is read one check ahead.
#include "kernels.h"
#include "topk.cuh"

namespace vl {

template <int METRIC>
__device__ __forceinline__ float accum4(float acc, const float4& v, const float4& q) {
    if (METRIC == COSINE || METRIC == DOT) {
        acc = fmaf(v.x, q.x, acc);
        acc = fmaf(v.y, q.y, acc);
        acc = fmaf(v.z, q.z, acc);
        acc = fmaf(v.w, q.w, acc);
    } else if (METRIC == EUCLIDEAN) {
        float d;
        d = v.x - q.x; acc = fmaf(d, d, acc);
        d = v.y - q.y; acc = fmaf(d, d, acc);
        d = v.z - q.z; acc = fmaf(d, d, acc);
        d = v.w - q.w; acc = fmaf(d, d, acc);
    } else {
        acc += fabsf(v.x - q.x);
        acc += fabsf(v.y - q.y);
        acc += fabsf(v.z - q.z);
        acc += fabsf(v.w - q.w);
    }
    return acc;
}

__device__ __forceinline__ uint2 ldg_stream64(const uint2* p) {
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
// cosine / dot / L2: Σ x̃·q (L2 finishes as 2x̃·q − ‖x‖² − ‖q‖²); manhattan: Σ |x̃ − q|
template <int METRIC>
__device__ __forceinline__ float accum4_bf16(float acc, const uint2& v, const float4& q) {
    const float x0 = __uint_as_float(v.x << 16), x1 = __uint_as_float(v.x & 0xFFFF0000u);
    const float x2 = __uint_as_float(v.y << 16), x3 = __uint_as_float(v.y & 0xFFFF0000u);
    if (METRIC == MANHATTAN) {
        acc += fabsf(x0 - q.x);
        acc += fabsf(x1 - q.y);
        acc += fabsf(x2 - q.z);
        acc += fabsf(x3 - q.w);
    } else {
        acc = fmaf(x0, q.x, acc);
        acc = fmaf(x1, q.y, acc);
        acc = fmaf(x2, q.z, acc);
        acc = fmaf(x3, q.w, acc);
    }
    return acc;
}

// 8 per-lane partial sums (one per row) → lane L holds the full sum of row (L >> 2).
// The addition tree is identical for every row, so equal rows give bit-equal scores.
__device__ __forceinline__ float reduce_transposed(const float (&a)[8], int lane) {
    const unsigned FULL = 0xFFFFFFFFu;
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    float c[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = (b4 ? a[i + 4] : a[i]) + __shfl_xor_sync(FULL, b4 ? a[i] : a[i + 4], 16);
    float d[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) d[i] = (b3 ? c[i + 2] : c[i]) + __shfl_xor_sync(FULL, b3 ? c[i] : c[i + 2], 8);
    float e = (b2 ? d[1] : d[0]) + __shfl_xor_sync(FULL, b2 ? d[0] : d[1], 4);
    e += __shfl_xor_sync(FULL, e, 2);
    e += __shfl_xor_sync(FULL, e, 1);
    return e;
}
// 16 per-lane partial sums → lane L holds the full sum of row (L >> 1); 16 shuffles
__device__ __forceinline__ float reduce_transposed(const float (&a)[16], int lane) {
    const unsigned FULL = 0xFFFFFFFFu;
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
    float c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = (b4 ? a[i + 8] : a[i]) + __shfl_xor_sync(FULL, b4 ? a[i] : a[i + 8], 16);
    float d[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) d[i] = (b3 ? c[i + 4] : c[i]) + __shfl_xor_sync(FULL, b3 ? c[i] : c[i + 4], 8);
    float e[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) e[i] = (b2 ? d[i + 2] : d[i]) + __shfl_xor_sync(FULL, b2 ? d[i] : d[i + 2], 4);
    float f = (b1 ? e[1] : e[0]) + __shfl_xor_sync(FULL, b1 ? e[0] : e[1], 2);
    f += __shfl_xor_sync(FULL, f, 1);
    return f;
}

constexpr int BF_CTAS_PER_SM = 2;   // the bf16 variant keeps 48 64-bit loads per lane in flight (124 registers)

// NCH > 0: pitch == NCH*128 elements exactly, query in registers, fully unrolled.
// NCH == 0: any pitch (multiple of 4; fp32 only), query re-read from shared memory per chunk.
// BF16: rows = bf16 mirror ([n][pitch] bf16, NCH == 3), aux = ‖row‖² (L2); else rows = fp32 arena, aux = 1/‖row‖.
template <int METRIC, int NCH, bool BF16>
__global__ void __launch_bounds__(SCAN_THREADS, BF16 ? BF_CTAS_PER_SM : SCAN_CTAS_PER_SM)
flat_scan_kernel(const void* __restrict__ rows_v, const float* __restrict__ aux,
                 const float4* __restrict__ queries, uint32_t n, uint32_t pitch4, uint64_t* cand,
                 uint32_t* cand_count, uint64_t* cand_max, QueryCtl* ctl_all, unsigned long long* early_all,
                 int Kp, int early_trigger) {
    constexpr int R = BF16 ? 16 : SCAN_ROWS_PER_WARP;
    constexpr int TILE = (SCAN_THREADS / 32) * R;
    constexpr int LIMIT = SCAN_CAP - SCAN_TILES_PER_CHECK * TILE;
    constexpr int OWN_SHIFT = BF16 ? 1 : 2;              // owner lanes: every 2nd (16 rows) / 4th (8 rows)
    static_assert(!BF16 || (NCH >= 1 && NCH <= 12), "the bf16 mirror variant serves rows of 128·NCH elements, query in registers");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(smem_raw);
    float4* s_q = reinterpret_cast<float4*>(smem_raw + SCAN_CAP * sizeof(uint64_t));
    __shared__ int s_count;
    __shared__ unsigned long long s_tau, s_red;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t qi = blockIdx.y;
    QueryCtl* ctl = ctl_all + qi;
    const float4* q4 = queries + static_cast<size_t>(qi) * pitch4;

    // Stand-alone (non-pipelined) launches let their finalize be launched right away: it parks in
    // griddepcontrol.wait until this grid has completed, which takes its launch latency off a lone caller's
    // critical path.  In a pipelined stream the same trigger would also release the NEXT query's scan into SMs
    // that are still streaming (measured 5 % slower, 7711 vs 8089 q/s; a trigger after the last tile: 7950), so it is
    // not used there.
    if (early_trigger) pdl_launch_dependents();
    CtaTopK<SCAN_CAP, SCAN_THREADS> topk{s_keys, &s_count};
    for (uint32_t i = tid; i < pitch4; i += SCAN_THREADS) s_q[i] = q4[i];
    if (tid == 0) s_red = 0ull;
    topk.init();  // contains a barrier → s_q visible

    float4 qreg[NCH > 0 ? NCH : 1];
    if (NCH > 0) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) qreg[c] = s_q[c * 32 + lane];
    }
    const uint32_t nch = NCH > 0 ? NCH : (pitch4 + 31) / 32;
    float qn2 = 0.f;
    if (BF16 && METRIC == EUCLIDEAN) {
#pragma unroll
        for (int c = 0; c < (NCH > 0 ? NCH : 1); ++c)
            qn2 += qreg[c].x * qreg[c].x + qreg[c].y * qreg[c].y + qreg[c].z * qreg[c].z + qreg[c].w * qreg[c].w;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) qn2 += __shfl_xor_sync(0xFFFFFFFFu, qn2, o);
    }

    // this CTA's rows: an equal contiguous share [lo, hi)
    const uint32_t G = gridDim.x, b = blockIdx.x;
    const uint32_t lo = static_cast<uint32_t>(static_cast<uint64_t>(n) * b / G);
    const uint32_t hi = static_cast<uint32_t>(static_cast<uint64_t>(n) * (b + 1) / G);
    const uint32_t my_tiles = (hi - lo + TILE - 1) / TILE;
    unsigned long long* early = early_all + static_cast<size_t>(qi) * EARLY_STRIDE;
    const bool use_early = early_all != nullptr && G <= static_cast<uint32_t>(EARLY_STRIDE) &&
                           static_cast<uint32_t>(Kp) <= G;

    unsigned long long tau = 0ull, tau_local = 0ull, g_next = 0ull;
    bool nonfinite = false, got_early = false;
    for (uint32_t it = 0; it < my_tiles; ++it) {
        if (it == 1 && use_early) {
            // ---- publish the best key of the first tile -------------------------------------------------------
            __syncthreads();
            const int n0 = min(s_count, SCAN_CAP);
            unsigned long long m = 0ull;
            for (int i = tid; i < n0; i += SCAN_THREADS) m = s_keys[i] > m ? s_keys[i] : m;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, m, o);
                m = other > m ? other : m;
            }
            if (lane == 0 && m) atomicMax(&s_red, m);
            __syncthreads();
            if (tid == 0 && s_red) *reinterpret_cast<volatile unsigned long long*>(early + b) = s_red;
        } else if ((it == 3 || it == 5 || it == 7) && use_early && !got_early) {
            // ---- early grid-wide threshold: min over Kp disjoint groups of the published maxima ------------------
            __syncthreads();
            if (tid == 0) s_red = ~0ull;
            __syncthreads();
            for (uint32_t t = tid; t < static_cast<uint32_t>(Kp); t += SCAN_THREADS) {
                unsigned long long m = 0ull;
                for (uint32_t j = t; j < G; j += Kp) {
                    const unsigned long long v = *reinterpret_cast<volatile unsigned long long*>(early + j);
                    m = v > m ? v : m;
                }
                atomicMin(&s_red, m);      // a group whose CTAs have not published yet gives 0: no early filter
            }
            __syncthreads();
            const unsigned long long t1 = s_red;
            if (t1 > tau) {                // uniform
                // drop everything below the bound (every entry is read before the buffer is rewritten)
                const int n0 = min(s_count, SCAN_CAP);
                unsigned long long mine[SCAN_CAP / SCAN_THREADS];
#pragma unroll
                for (int u = 0; u < SCAN_CAP / SCAN_THREADS; ++u) {
                    const int i = tid + u * SCAN_THREADS;
                    mine[u] = i < n0 ? s_keys[i] : 0ull;
                }
                __syncthreads();
                if (tid == 0) s_count = 0;
                __syncthreads();
#pragma unroll
                for (int u = 0; u < SCAN_CAP / SCAN_THREADS; ++u)
                    if (mine[u] >= t1) topk.push(mine[u]);
                tau = t1;
                tau_local = t1 > tau_local ? t1 : tau_local;
            }
            got_early = t1 != 0ull;   // 0: some CTAs had not published yet (staggered start) — try again two tiles on
        } else if ((it & (SCAN_TILES_PER_CHECK - 1)) == 0) {
            __syncthreads();
            if (s_count > LIMIT) {  // uniform: read after the barrier, no pushes in flight
                const unsigned long long t = topk.compact(Kp, false);
                tau_local = t > tau_local ? t : tau_local;
            }
            if (tid == 0) {
                // the grid-wide threshold is read one check AHEAD (g_next is consumed at the next check), so its
                // L2 round trip is never waited for between the two barriers
                unsigned long long g = g_next;
                g_next = *reinterpret_cast<volatile unsigned long long*>(&ctl->tau);
                if (tau_local > g) {
                    atomicMax(&ctl->tau, tau_local);
                    g = tau_local;
                }
                s_tau = g;
            }
            __syncthreads();
            tau = s_tau > tau ? s_tau : tau;
        }
        const uint32_t row0 = lo + it * TILE + warp * R;
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
        const bool full = row0 + R <= hi;
        // owner lanes handle row row0 + (lane >> OWN_SHIFT)
        const uint32_t my_row = row0 + (lane >> OWN_SHIFT);
        const bool owner = (lane & ((1 << OWN_SHIFT) - 1)) == 0 && my_row < hi;
        float a1 = 0.f;   // fp32 cosine: 1/‖row‖; bf16 L2: ‖row‖²
        if ((BF16 ? METRIC == EUCLIDEAN : METRIC == COSINE) && owner) a1 = __ldg(aux + my_row);

        if (BF16) {
            const uint2* base = static_cast<const uint2*>(rows_v) + static_cast<size_t>(row0) * pitch4 + lane;
            if (full) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    uint2 v[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) v[r] = ldg_stream64(base + static_cast<size_t>(r) * pitch4 + c * 32);
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = accum4_bf16<METRIC>(acc[r], v[r], qreg[c]);
                }
            } else {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if (row0 + r < hi) {
                            const uint2 v = ldg_stream64(base + static_cast<size_t>(r) * pitch4 + c * 32);
                            acc[r] = accum4_bf16<METRIC>(acc[r], v, qreg[c]);
                        }
                    }
                }
            }
        } else {
            const float4* base = static_cast<const float4*>(rows_v) + static_cast<size_t>(row0) * pitch4 + lane;
            if (NCH > 0) {
                if (full) {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        float4 v[R];
#pragma unroll
                        for (int r = 0; r < R; ++r) v[r] = ldg_stream(base + static_cast<size_t>(r) * pitch4 + c * 32);
#pragma unroll
                        for (int r = 0; r < R; ++r) acc[r] = accum4<METRIC>(acc[r], v[r], qreg[c]);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            if (row0 + r < hi) {
                                const float4 v = ldg_stream(base + static_cast<size_t>(r) * pitch4 + c * 32);
                                acc[r] = accum4<METRIC>(acc[r], v, qreg[c]);
                            }
                        }
                    }
                }
            } else {
                for (uint32_t c = 0; c < nch; ++c) {
                    const uint32_t col = c * 32 + lane;
                    if (col < pitch4) {
                        const float4 q = s_q[col];
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            if (row0 + r < hi) {
                                const float4 v = ldg_stream(base + static_cast<size_t>(r) * pitch4 + c * 32);
                                acc[r] = accum4<METRIC>(acc[r], v, q);
                            }
                        }
                    }
                }
            }
        }

        float s = reduce_transposed(acc, lane);
        if (owner) {
            if (BF16) {
                if (METRIC == EUCLIDEAN) s = fmaf(2.f, s, -a1) - qn2;   // −‖x−q‖²
                if (METRIC == MANHATTAN) s = -s;                        // −Σ|x̃−q|
            } else {
                if (METRIC == COSINE) s *= a1;
                if (METRIC == EUCLIDEAN || METRIC == MANHATTAN) s = -s;
            }
            if (!isfinite(s)) nonfinite = true;
            const unsigned long long key = make_key(s, my_row);
            if (key > tau) topk.push(key);
        }
    }

    // epilogue: keep this CTA's best Kp, publish its threshold and its candidates
    pdl_wait();  // pipelined mode: the previous query's finalize must be done with cand[] first
    const unsigned long long t = topk.compact(Kp, false);
    tau_local = t > tau_local ? t : tau_local;
    const int any_nf = __syncthreads_or(nonfinite ? 1 : 0);
    const int n_out = s_count;
    const size_t slot = static_cast<size_t>(qi) * gridDim.x + blockIdx.x;
    unsigned long long best = 0ull;
    for (int i = tid; i < n_out; i += SCAN_THREADS) {
        const unsigned long long key = s_keys[i];
        cand[slot * Kp + i] = key;
        best = key > best ? key : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, best, o);
        best = other > best ? other : best;
    }
    if (tid == 0) s_tau = 0ull;
    __syncthreads();
    if (lane == 0 && best) atomicMax(&s_tau, best);
    __syncthreads();
    if (tid == 0) {
        cand_max[slot] = s_tau;
        cand_count[slot] = static_cast<uint32_t>(n_out);
        if (tau_local) atomicMax(&ctl->tau, tau_local);
        if (any_nf) atomicOr(&ctl->flags, FLAG_NONFINITE);
    }
}

size_t flat_scan_smem_bytes(uint32_t pitch) {
    return SCAN_CAP * sizeof(uint64_t) + static_cast<size_t>(pitch) * sizeof(float);
}

template <int METRIC, int NCH, bool BF16>
static cudaError_t launch_one(const void* rows, const float* aux, uint32_t n, uint32_t pitch, const float* d_queries,
                              uint32_t nq, const ScanWork& w, bool pipelined, cudaStream_t s) {
    const size_t smem = flat_scan_smem_bytes(pitch);
    auto kern = flat_scan_kernel<METRIC, NCH, BF16>;
    if (smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem));
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(w.grid_x, nq);
    cfg.blockDim = dim3(SCAN_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pipelined ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, rows, aux, reinterpret_cast<const float4*>(d_queries), n, pitch / 4, w.cand,
                              w.cand_count, w.cand_max, w.ctl, reinterpret_cast<unsigned long long*>(w.early), w.Kp,
                              pipelined ? 0 : 1);
}

template <int METRIC>
static cudaError_t launch_metric(const FlatView& v, const float* q, uint32_t nq, const ScanWork& w,
                                 bool pipelined, cudaStream_t s) {
    switch (v.pitch) {
        case 384: return launch_one<METRIC, 3, false>(v.rows, v.inv_norm, v.n, v.pitch, q, nq, w, pipelined, s);
        case 768: return launch_one<METRIC, 6, false>(v.rows, v.inv_norm, v.n, v.pitch, q, nq, w, pipelined, s);
        default: return launch_one<METRIC, 0, false>(v.rows, v.inv_norm, v.n, v.pitch, q, nq, w, pipelined, s);
    }
}

cudaError_t launch_flat_scan(const FlatView& v, const float* d_queries, uint32_t nq, int metric,
                             const ScanWork& w, bool pipelined, cudaStream_t s) {
    switch (metric) {
        case COSINE: return launch_metric<COSINE>(v, d_queries, nq, w, pipelined, s);
        case EUCLIDEAN: return launch_metric<EUCLIDEAN>(v, d_queries, nq, w, pipelined, s);
        case MANHATTAN: return launch_metric<MANHATTAN>(v, d_queries, nq, w, pipelined, s);
        case DOT: return launch_metric<DOT>(v, d_queries, nq, w, pipelined, s);
        default: return cudaErrorInvalidValue;
    }
}

// single-query scan over the bf16 mirror; requires pitch (== the mirror's padded width) of 128 / 256 / 384 / 768 / 1024 /
// 1536 elements (the widths of common embedding models)
template <int METRIC>
static cudaError_t launch_bf16(const FlatView& v, const void* mirror, const float* sq_norm, const float* d_queries,
                               uint32_t nq, const ScanWork& w, bool pipelined, cudaStream_t s) {
    switch (v.pitch) {
        case 128: return launch_one<METRIC, 1, true>(mirror, sq_norm, v.n, v.pitch, d_queries, nq, w, pipelined, s);
        case 256: return launch_one<METRIC, 2, true>(mirror, sq_norm, v.n, v.pitch, d_queries, nq, w, pipelined, s);
        case 384: return launch_one<METRIC, 3, true>(mirror, sq_norm, v.n, v.pitch, d_queries, nq, w, pipelined, s);
        case 768: return launch_one<METRIC, 6, true>(mirror, sq_norm, v.n, v.pitch, d_queries, nq, w, pipelined, s);
        case 1024: return launch_one<METRIC, 8, true>(mirror, sq_norm, v.n, v.pitch, d_queries, nq, w, pipelined, s);
        case 1536: return launch_one<METRIC, 12, true>(mirror, sq_norm, v.n, v.pitch, d_queries, nq, w, pipelined, s);
        default: return cudaErrorNotSupported;
    }
}

bool flat_scan_bf16_supports(uint32_t pitch) {
    return pitch == 128 || pitch == 256 || pitch == 384 || pitch == 768 || pitch == 1024 || pitch == 1536;
}

cudaError_t launch_flat_scan_bf16(const FlatView& v, const void* mirror, const float* sq_norm, const float* d_queries,
                                  uint32_t nq, int metric, const ScanWork& w, bool pipelined, cudaStream_t s) {
    if (!flat_scan_bf16_supports(v.pitch) || !mirror) return cudaErrorNotSupported;
    switch (metric) {
        case COSINE: return launch_bf16<COSINE>(v, mirror, sq_norm, d_queries, nq, w, pipelined, s);
        case EUCLIDEAN: return sq_norm ? launch_bf16<EUCLIDEAN>(v, mirror, sq_norm, d_queries, nq, w, pipelined, s)
                                       : cudaErrorNotSupported;
        case DOT: return launch_bf16<DOT>(v, mirror, sq_norm, d_queries, nq, w, pipelined, s);
        case MANHATTAN: return launch_bf16<MANHATTAN>(v, mirror, sq_norm, d_queries, nq, w, pipelined, s);
        default: return cudaErrorNotSupported;
    }
}

int flat_scan_bf16_max_grid_x(int device) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    return sms * BF_CTAS_PER_SM;  // one full wave of resident CTAs
}

int flat_scan_max_grid_x(int device, uint32_t pitch) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    (void)pitch;
    return sms * SCAN_CTAS_PER_SM;  // one full wave of resident CTAs
}

}  // namespace vl
