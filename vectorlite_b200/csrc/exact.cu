// exact.cu — the EXACT device path: every row scored in f64 with the reference's arithmetic,
// then a stable radix select.  Used (a) when the optimality certificate of the fast path fails
// (heavy ties, e.g. the all-equal-embeddings case of src/client.rs:665-667), (b) for k beyond
// the over-select capacity, (c) on request (VL_MODE_EXACT).  Never a CPU fallback.
//
// exact_scores_kernel restates src/lib.rs:425-572 per row: strictly sequential f64 accumulation
// in index order, no FMA (__dmul_rn/__dadd_rn), cosine = dot/(sqrt(Σx²)·sqrt(Σy²)) with the
// zero-norm → 0.0 branch.  One lane per row; a warp transposes 32 rows × 32 columns through
// shared memory so global loads stay coalesced.  Roofline: FP64 pipe / F2F conversion, not HBM.
//
// exact_select: keys = orderable(score + 0.0) (−0.0 canonicalised: ±0 compare equal in
// flat.rs:116), values = positions in ascending order; cub::DeviceRadixSort::SortPairsDescending
// is stable, so equal scores keep ascending storage position == the reference's stable sort.
#include <cub/device/device_radix_sort.cuh>

#include "kernels.h"

namespace vl {

constexpr int EX_THREADS = 256;  // 8 warps × 32 rows

template <int METRIC>
__global__ void __launch_bounds__(EX_THREADS) exact_scores_kernel(
    const float* __restrict__ rows, const float* __restrict__ query, uint32_t n, uint32_t dim,
    uint32_t pitch, double* __restrict__ scores, uint32_t* flags) {
    extern __shared__ float s_qx[];                         // [pitch] query
    __shared__ float s_t[EX_THREADS / 32][32][33];          // per-warp 32×32 transpose tile
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t i = tid; i < pitch; i += EX_THREADS) s_qx[i] = query[i];
    __syncthreads();
    const uint32_t groups = (n + 31) / 32;
    bool nan_seen = false;
    for (uint32_t g = blockIdx.x * (EX_THREADS / 32) + warp; g < groups; g += gridDim.x * (EX_THREADS / 32)) {
        const uint32_t row0 = g * 32;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        for (uint32_t c0 = 0; c0 < dim; c0 += 32) {
            // coalesced: 8 lanes × float4 cover 32 columns of one row; 4 rows per instruction
#pragma unroll
            for (int rr = 0; rr < 32; rr += 4) {
                const uint32_t r = rr + (lane >> 3);
                const uint32_t col = c0 + (lane & 7) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row0 + r < n && col < pitch)
                    v = *reinterpret_cast<const float4*>(rows + static_cast<size_t>(row0 + r) * pitch + col);
                float* t = &s_t[warp][r][(lane & 7) * 4];
                t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
            }
            __syncwarp();
            const int w = min(32u, dim - c0);
            for (int j = 0; j < w; ++j) {
                const double x = static_cast<double>(s_t[warp][lane][j]);
                const double y = static_cast<double>(s_qx[c0 + j]);
                if (METRIC == COSINE) {
                    a0 = __dadd_rn(a0, __dmul_rn(x, y));
                    a1 = __dadd_rn(a1, __dmul_rn(x, x));
                    a2 = __dadd_rn(a2, __dmul_rn(y, y));
                } else if (METRIC == EUCLIDEAN) {
                    const double d = __dsub_rn(x, y);
                    a0 = __dadd_rn(a0, __dmul_rn(d, d));
                } else if (METRIC == MANHATTAN) {
                    a0 = __dadd_rn(a0, fabs(__dsub_rn(x, y)));
                } else {
                    a0 = __dadd_rn(a0, __dmul_rn(x, y));
                }
            }
            __syncwarp();
        }
        double sc;
        if (METRIC == COSINE) {
            const double na = __dsqrt_rn(a1), nb = __dsqrt_rn(a2);
            sc = (na == 0.0 || nb == 0.0) ? 0.0 : __ddiv_rn(a0, __dmul_rn(na, nb));
        } else if (METRIC == EUCLIDEAN) {
            sc = __ddiv_rn(1.0, __dadd_rn(1.0, __dsqrt_rn(a0)));
        } else if (METRIC == MANHATTAN) {
            sc = __ddiv_rn(1.0, __dadd_rn(1.0, a0));
        } else {
            sc = a0;
        }
        if (row0 + lane < n) {
            scores[row0 + lane] = sc;
            if (sc != sc) nan_seen = true;
        }
    }
    if (__syncthreads_or(nan_seen ? 1 : 0) && tid == 0) atomicOr(flags, FLAG_NAN);
}

cudaError_t launch_exact_scores(const FlatView& v, const float* d_query, int metric, double* d_scores,
                                uint32_t* d_flags, cudaStream_t s) {
    const uint32_t groups = (v.n + 31) / 32;
    int grid = static_cast<int>((groups + 7) / 8);
    if (grid > 148 * 8) grid = 148 * 8;
    if (grid < 1) grid = 1;
    const size_t smem = static_cast<size_t>(v.pitch) * sizeof(float);
    switch (metric) {
        case COSINE:
            exact_scores_kernel<COSINE><<<grid, EX_THREADS, smem, s>>>(v.rows, d_query, v.n, v.dim, v.pitch, d_scores, d_flags);
            break;
        case EUCLIDEAN:
            exact_scores_kernel<EUCLIDEAN><<<grid, EX_THREADS, smem, s>>>(v.rows, d_query, v.n, v.dim, v.pitch, d_scores, d_flags);
            break;
        case MANHATTAN:
            exact_scores_kernel<MANHATTAN><<<grid, EX_THREADS, smem, s>>>(v.rows, d_query, v.n, v.dim, v.pitch, d_scores, d_flags);
            break;
        case DOT:
            exact_scores_kernel<DOT><<<grid, EX_THREADS, smem, s>>>(v.rows, d_query, v.n, v.dim, v.pitch, d_scores, d_flags);
            break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

__global__ void exact_make_keys_kernel(const double* __restrict__ scores, uint32_t n, uint64_t* keys,
                                       uint32_t* vals) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        keys[i] = f64_orderable(scores[i] + 0.0);  // −0.0 + 0.0 == +0.0
        vals[i] = i;
    }
}

__global__ void exact_gather_kernel(const uint32_t* __restrict__ sorted_pos, const double* __restrict__ scores,
                                    const uint64_t* __restrict__ ids, uint64_t id_base, uint64_t pos_base,
                                    uint32_t n, uint32_t k, uint32_t q_index, uint64_t* out_ids,
                                    double* out_scores, uint64_t* out_pos, uint32_t* out_counts) {
    const uint32_t cnt = k < n ? k : n;
    for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) {
        const size_t o = static_cast<size_t>(q_index) * k + i;
        if (i < cnt) {
            const uint32_t pos = sorted_pos[i];
            out_ids[o] = ids ? ids[pos] : id_base + pos;
            out_scores[o] = scores[pos];
            if (out_pos) out_pos[o] = pos_base + pos;
        } else {
            out_ids[o] = ~0ull;
            out_scores[o] = 0.0;
            if (out_pos) out_pos[o] = ~0ull;
        }
    }
    if (threadIdx.x == 0) out_counts[q_index] = cnt;
}

void exact_scratch_free(ExactScratch& sc) {
    cudaFree(sc.temp);
    cudaFree(sc.keys_in);
    cudaFree(sc.keys_out);
    cudaFree(sc.vals_in);
    cudaFree(sc.vals_out);
    sc = ExactScratch();
}

cudaError_t exact_select(const FlatView& v, const double* d_scores, uint32_t k, ExactScratch& sc,
                         const SearchOut& out, uint32_t q_index, cudaStream_t s) {
    cudaError_t e;
    if (sc.cap < v.n) {
        exact_scratch_free(sc);
        const size_t cap = static_cast<size_t>(v.n) + v.n / 4 + 1024;
        if ((e = cudaMalloc(&sc.keys_in, cap * 8)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&sc.keys_out, cap * 8)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&sc.vals_in, cap * 4)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&sc.vals_out, cap * 4)) != cudaSuccess) return e;
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, sc.keys_in, sc.keys_out, sc.vals_in,
                                                  sc.vals_out, static_cast<int64_t>(cap), 0, 64, s);
        if ((e = cudaMalloc(&sc.temp, tb)) != cudaSuccess) return e;
        sc.temp_bytes = tb;
        sc.cap = cap;
    }
    exact_make_keys_kernel<<<148 * 4, 256, 0, s>>>(d_scores, v.n, sc.keys_in, sc.vals_in);
    size_t tb = sc.temp_bytes;
    e = cub::DeviceRadixSort::SortPairsDescending(sc.temp, tb, sc.keys_in, sc.keys_out, sc.vals_in,
                                                  sc.vals_out, static_cast<int64_t>(v.n), 0, 64, s);
    if (e != cudaSuccess) return e;
    exact_gather_kernel<<<1, 256, 0, s>>>(sc.vals_out, d_scores, v.ids, v.id_base, v.pos_base, v.n, k,
                                          q_index, out.ids, out.scores, out.pos, out.counts);
    return cudaGetLastError();
}

}  // namespace vl
