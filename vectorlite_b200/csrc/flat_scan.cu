// flat_scan.cu — single-query flat scan: HBM-bound streaming distance + fused top-K'.
//
// Replaces the hot loop of FlatIndex::search (src/index/flat.rs:106-117: n×calculate(), clone,
// stable sort, truncate) and the four metric loops (src/lib.rs:425-572).
//
// Roofline: HBM.  Algorithmic bytes per query = n·pitch·4 (+ n·4 inv-norms for cosine).
// Layout: rows [n][pitch] fp32 row-major (pitch·4 B multiple of 16), one warp reads 8 whole rows
// per iteration as 128-bit ld.global.nc.L1::no_allocate (3 per lane per 384-d row, fully
// coalesced 512 B per instruction); the query lives in registers; the 8 per-lane partial sums
// are reduced with a transposed butterfly (9 shuffles for 8 rows); each row's fp32 score becomes
// a 64-bit key (score, ~position) and is appended to a CTA candidate buffer only if it beats
// the grid-wide running K'-th best key (QueryCtl::tau, an atomicMax lower bound).  The exact
// f64 score and the final order are produced by flat_finalize.cu.
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

// 8 per-lane partial sums (one per row) → lane L holds the full sum of row (L >> 2).
// The addition tree is identical for every row, so equal rows give bit-equal scores.
__device__ __forceinline__ float reduce8_transposed(const float (&a)[8], int lane) {
    const unsigned FULL = 0xFFFFFFFFu;
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    float c[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = b4 ? a[i] : a[i + 4];
        const float keep = b4 ? a[i + 4] : a[i];
        c[i] = keep + __shfl_xor_sync(FULL, send, 16);
    }
    float d[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b3 ? c[i] : c[i + 2];
        const float keep = b3 ? c[i + 2] : c[i];
        d[i] = keep + __shfl_xor_sync(FULL, send, 8);
    }
    const float send = b2 ? d[0] : d[1];
    const float keep = b2 ? d[1] : d[0];
    float e = keep + __shfl_xor_sync(FULL, send, 4);
    e += __shfl_xor_sync(FULL, e, 2);
    e += __shfl_xor_sync(FULL, e, 1);
    return e;
}

// NCH > 0: pitch == NCH*128 floats exactly, query in registers, fully unrolled.
// NCH == 0: any pitch (multiple of 4), query re-read from shared memory per chunk.
template <int METRIC, int NCH>
__global__ void __launch_bounds__(SCAN_THREADS, SCAN_CTAS_PER_SM)
flat_scan_kernel(const float4* __restrict__ rows, const float* __restrict__ inv_norm,
                 const float4* __restrict__ queries, uint32_t n, uint32_t pitch4, uint64_t* cand,
                 uint32_t* cand_count, uint64_t* cand_max, QueryCtl* ctl_all, int Kp) {
    constexpr int R = SCAN_ROWS_PER_WARP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(smem_raw);
    float4* s_q = reinterpret_cast<float4*>(smem_raw + SCAN_CAP * sizeof(uint64_t));
    __shared__ int s_count;
    __shared__ unsigned long long s_tau;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t qi = blockIdx.y;
    QueryCtl* ctl = ctl_all + qi;
    const float4* q4 = queries + static_cast<size_t>(qi) * pitch4;

    CtaTopK<SCAN_CAP, SCAN_THREADS> topk{s_keys, &s_count};
    for (uint32_t i = tid; i < pitch4; i += SCAN_THREADS) s_q[i] = q4[i];
    topk.init();  // contains a barrier → s_q visible

    float4 qreg[NCH > 0 ? NCH : 1];
    if (NCH > 0) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) qreg[c] = s_q[c * 32 + lane];
    }
    const uint32_t nch = NCH > 0 ? NCH : (pitch4 + 31) / 32;

    unsigned long long tau = 0ull, tau_local = 0ull;
    bool nonfinite = false;
    const uint32_t num_tiles = (n + SCAN_TILE_ROWS - 1) / SCAN_TILE_ROWS;
    uint32_t iter = 0;
    for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
        if ((iter & (SCAN_TILES_PER_CHECK - 1)) == 0) {
            __syncthreads();
            if (s_count > SCAN_LIMIT) {  // uniform: read after the barrier, no pushes in flight
                const unsigned long long t = topk.compact(Kp, false);
                tau_local = t > tau_local ? t : tau_local;
            }
            if (tid == 0) {
                unsigned long long g = *reinterpret_cast<volatile unsigned long long*>(&ctl->tau);
                if (tau_local > g) {
                    atomicMax(&ctl->tau, tau_local);
                    g = tau_local;
                }
                s_tau = g;
            }
            __syncthreads();
            tau = s_tau;
        }
        const uint32_t row0 = tile * SCAN_TILE_ROWS + warp * R;
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
        const float4* base = rows + static_cast<size_t>(row0) * pitch4 + lane;
        const bool full = row0 + R <= n;
        // owner lane (lane & 3) == 0 handles row row0 + (lane >> 2)
        const uint32_t my_row = row0 + (lane >> 2);
        float invn = 0.f;
        if (METRIC == COSINE && (lane & 3) == 0 && my_row < n) invn = __ldg(inv_norm + my_row);

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
                        if (row0 + r < n) {
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
                        if (row0 + r < n) {
                            const float4 v = ldg_stream(base + static_cast<size_t>(r) * pitch4 + c * 32);
                            acc[r] = accum4<METRIC>(acc[r], v, q);
                        }
                    }
                }
            }
        }

        float s = reduce8_transposed(acc, lane);
        if ((lane & 3) == 0 && my_row < n) {
            if (METRIC == COSINE) s *= invn;
            if (METRIC == EUCLIDEAN || METRIC == MANHATTAN) s = -s;
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

// ---------------------------------------------------------------------------------------------------------
// bf16-mirror variant (384-d): the same streaming skeleton over the index's bf16 copy of the rows ([n][384] bf16,
// 768 B per row; cosine: rows pre-scaled by 1/‖row‖; dot / L2: raw rows + fp32 ‖row‖²) — HALF the HBM bytes
// per query (s = 2 B/element, stated in bench.py's roofline).  One warp reads 8 whole rows per
// iteration as three 64-bit streaming loads per lane per row (lane l gets elements 4·(32c + l) … +3 of chunk c,
// the same elements as the fp32 kernel, so the query registers are laid out identically); bf16 → fp32 is a
// shift / mask, accumulation is fp32 against the fp32 query.  Scores are in the tensor-core path's scan units
// (cos·‖q‖, x·q, −‖x−q‖² via 2x·q − ‖x‖² − ‖q‖²) and are certified with the same bf16 bound (tc_abs).
__device__ __forceinline__ uint2 ldg_stream64(const uint2* p) {
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float dot4_bf16(float acc, const uint2& v, const float4& q) {
    acc = fmaf(__uint_as_float(v.x << 16), q.x, acc);
    acc = fmaf(__uint_as_float(v.x & 0xFFFF0000u), q.y, acc);
    acc = fmaf(__uint_as_float(v.y << 16), q.z, acc);
    acc = fmaf(__uint_as_float(v.y & 0xFFFF0000u), q.w, acc);
    return acc;
}

template <int METRIC>
__global__ void __launch_bounds__(SCAN_THREADS, SCAN_CTAS_PER_SM)
flat_scan_bf16_kernel(const uint2* __restrict__ rows, const float* __restrict__ sq_norm,
                      const float4* __restrict__ queries, uint32_t n, uint64_t* cand,
                      uint32_t* cand_count, uint64_t* cand_max, QueryCtl* ctl_all, int Kp) {
    constexpr int NCH = 3;
    constexpr uint32_t pitch4 = 96;    // fp32 query: 96 float4; bf16 row: 96 uint2
    constexpr int R = SCAN_ROWS_PER_WARP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(smem_raw);
    float4* s_q = reinterpret_cast<float4*>(smem_raw + SCAN_CAP * sizeof(uint64_t));
    __shared__ int s_count;
    __shared__ unsigned long long s_tau;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t qi = blockIdx.y;
    QueryCtl* ctl = ctl_all + qi;
    const float4* q4 = queries + static_cast<size_t>(qi) * pitch4;

    CtaTopK<SCAN_CAP, SCAN_THREADS> topk{s_keys, &s_count};
    for (uint32_t i = tid; i < pitch4; i += SCAN_THREADS) s_q[i] = q4[i];
    topk.init();  // contains a barrier → s_q visible

    float4 qreg[NCH > 0 ? NCH : 1];
    if (NCH > 0) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) qreg[c] = s_q[c * 32 + lane];
    }
    float qn2 = 0.f;
    if (METRIC == EUCLIDEAN) {
#pragma unroll
        for (int c = 0; c < NCH; ++c)
            qn2 += qreg[c].x * qreg[c].x + qreg[c].y * qreg[c].y + qreg[c].z * qreg[c].z + qreg[c].w * qreg[c].w;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) qn2 += __shfl_xor_sync(0xFFFFFFFFu, qn2, o);
    }

    unsigned long long tau = 0ull, tau_local = 0ull;
    bool nonfinite = false;
    const uint32_t num_tiles = (n + SCAN_TILE_ROWS - 1) / SCAN_TILE_ROWS;
    uint32_t iter = 0;
    for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
        if ((iter & (SCAN_TILES_PER_CHECK - 1)) == 0) {
            __syncthreads();
            if (s_count > SCAN_LIMIT) {  // uniform: read after the barrier, no pushes in flight
                const unsigned long long t = topk.compact(Kp, false);
                tau_local = t > tau_local ? t : tau_local;
            }
            if (tid == 0) {
                unsigned long long g = *reinterpret_cast<volatile unsigned long long*>(&ctl->tau);
                if (tau_local > g) {
                    atomicMax(&ctl->tau, tau_local);
                    g = tau_local;
                }
                s_tau = g;
            }
            __syncthreads();
            tau = s_tau;
        }
        const uint32_t row0 = tile * SCAN_TILE_ROWS + warp * R;
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
        const uint2* base = rows + static_cast<size_t>(row0) * pitch4 + lane;
        const bool full = row0 + R <= n;
        // owner lane (lane & 3) == 0 handles row row0 + (lane >> 2)
        const uint32_t my_row = row0 + (lane >> 2);
        float xn2 = 0.f;
        if (METRIC == EUCLIDEAN && (lane & 3) == 0 && my_row < n) xn2 = __ldg(sq_norm + my_row);

        if (NCH > 0) {
            if (full) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    uint2 v[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) v[r] = ldg_stream64(base + static_cast<size_t>(r) * pitch4 + c * 32);
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = dot4_bf16(acc[r], v[r], qreg[c]);
                }
            } else {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if (row0 + r < n) {
                            const uint2 v = ldg_stream64(base + static_cast<size_t>(r) * pitch4 + c * 32);
                            acc[r] = dot4_bf16(acc[r], v, qreg[c]);
                        }
                    }
                }
            }
        }

        float s = reduce8_transposed(acc, lane);
        if ((lane & 3) == 0 && my_row < n) {
            if (METRIC == EUCLIDEAN) s = fmaf(2.f, s, -xn2) - qn2;   // −‖x−q‖²
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

template <int METRIC, int NCH>
static cudaError_t launch_one(const FlatView& v, const float* d_queries, uint32_t nq, const ScanWork& w,
                              bool pipelined, cudaStream_t s) {
    const size_t smem = flat_scan_smem_bytes(v.pitch);
    auto kern = flat_scan_kernel<METRIC, NCH>;
    if (smem > 48 * 1024) {
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
    return cudaLaunchKernelEx(&cfg, kern, reinterpret_cast<const float4*>(v.rows), v.inv_norm,
                              reinterpret_cast<const float4*>(d_queries), v.n, v.pitch / 4, w.cand,
                              w.cand_count, w.cand_max, w.ctl, w.Kp);
}

template <int METRIC>
static cudaError_t launch_metric(const FlatView& v, const float* q, uint32_t nq, const ScanWork& w,
                                 bool pipelined, cudaStream_t s) {
    switch (v.pitch) {
        case 384: return launch_one<METRIC, 3>(v, q, nq, w, pipelined, s);
        case 768: return launch_one<METRIC, 6>(v, q, nq, w, pipelined, s);
        default: return launch_one<METRIC, 0>(v, q, nq, w, pipelined, s);
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

template <int METRIC>
static cudaError_t launch_bf16(const FlatView& v, const void* mirror, const float* sq_norm, const float* d_queries,
                               uint32_t nq, const ScanWork& w, bool pipelined, cudaStream_t s) {
    const size_t smem = flat_scan_smem_bytes(v.pitch);
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
    return cudaLaunchKernelEx(&cfg, flat_scan_bf16_kernel<METRIC>, reinterpret_cast<const uint2*>(mirror), sq_norm,
                              reinterpret_cast<const float4*>(d_queries), v.n, w.cand, w.cand_count, w.cand_max,
                              w.ctl, w.Kp);
}

// single-query scan over the bf16 mirror (see flat_scan_bf16_kernel); requires pitch == 384 and metric != manhattan
cudaError_t launch_flat_scan_bf16(const FlatView& v, const void* mirror, const float* sq_norm, const float* d_queries,
                                  uint32_t nq, int metric, const ScanWork& w, bool pipelined, cudaStream_t s) {
    if (v.pitch != 384 || !mirror) return cudaErrorNotSupported;
    switch (metric) {
        case COSINE: return launch_bf16<COSINE>(v, mirror, sq_norm, d_queries, nq, w, pipelined, s);
        case EUCLIDEAN: return sq_norm ? launch_bf16<EUCLIDEAN>(v, mirror, sq_norm, d_queries, nq, w, pipelined, s)
                                       : cudaErrorNotSupported;
        case DOT: return launch_bf16<DOT>(v, mirror, sq_norm, d_queries, nq, w, pipelined, s);
        default: return cudaErrorNotSupported;
    }
}

int flat_scan_max_grid_x(int device, uint32_t pitch) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    (void)pitch;
    return sms * SCAN_CTAS_PER_SM;  // one full wave of resident CTAs
}

}  // namespace vl
