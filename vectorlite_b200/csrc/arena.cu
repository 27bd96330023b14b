// arena.cu — insert-time kernels of the device-resident vector arena, the synthetic generator
// and the row-sharded top-k merge.
//
// The arena replaces the reference's AoS Vec<Vector> with one heap block per row
// (src/index/flat.rs:59-65, src/lib.rs:163-174) by a dense [n][pitch] fp32 matrix in HBM plus
// side arrays (1/‖row‖ for the cosine scan, ids).  Norms are computed once at insert instead of
// once per row per query (src/lib.rs:430-434 recomputes Σx² and Σy² for every pair).
#include "kernels.h"

namespace vl {

// ---- row norms: one warp per row, f64 tree sum (only needs |rel err| << 2^-24) -------------
__global__ void __launch_bounds__(256) row_norms_kernel(const float* __restrict__ rows, uint64_t first,
                                                        uint64_t n, uint32_t pitch, float* inv_norm,
                                                        ArenaStats* stats) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
    for (uint64_t r = warp; r < n; r += nwarps) {
        const float4* row = reinterpret_cast<const float4*>(rows + (first + r) * pitch);
        double ss = 0.0;
        for (uint32_t c = lane; c < pitch / 4; c += 32) {
            const float4 v = row[c];
            ss += static_cast<double>(v.x) * v.x + static_cast<double>(v.y) * v.y +
                  static_cast<double>(v.z) * v.z + static_cast<double>(v.w) * v.w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
        if (lane == 0) {
            const float inv = ss > 0.0 ? static_cast<float>(1.0 / sqrt(ss)) : 0.f;
            inv_norm[first + r] = (ss == ss) ? inv : __int_as_float(0x7FC00000);
            const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(ss));
            if (!(ss == ss) || isinf(ss)) {
                atomicAdd(&stats->nonfinite_rows, 1u);
                atomicMax(&stats->max_norm_sq_bits, 0x7FF8000000000000ull);
            } else {
                atomicMax(&stats->max_norm_sq_bits, bits);
                if (ss > 0.0) atomicMin(&stats->min_nz_norm_sq_bits, bits);
            }
        }
    }
}

cudaError_t launch_row_norms(float* rows, uint64_t first, uint64_t n, uint32_t dim, uint32_t pitch,
                             float* inv_norm, ArenaStats* stats, cudaStream_t s) {
    (void)dim;
    if (n == 0) return cudaSuccess;
    uint64_t blocks = (n + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    row_norms_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(rows, first, n, pitch, inv_norm, stats);
    return cudaGetLastError();
}

// ---- counter-based synthetic rows (bit-identical to oracle vlo_synth_rows_f32; restated here
// from the recipe, integer arithmetic + correctly rounded f64 sqrt/div only) -------------------
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ long long ih4(uint64_t seed, uint64_t row, uint64_t col) {
    const uint64_t h = mix64(mix64(seed ^ ((row + 1) * 0x9E3779B97F4A7C15ull)) ^
                             ((col + 1) * 0xD1B54A32D192ED03ull));
    return static_cast<long long>((h & 0xFFFF) + ((h >> 16) & 0xFFFF) + ((h >> 32) & 0xFFFF) + (h >> 48)) -
           131070ll;
}
__device__ __forceinline__ long long synth_int(uint64_t seed, uint64_t row, uint64_t centre,
                                               uint32_t clusters, uint32_t col) {
    long long x = ih4(seed, row, col);
    if (clusters) x += 4 * ih4(0x5851F42D4C957F2Dull, centre, col);  // centres do not depend on seed
    return x;
}

__global__ void __launch_bounds__(256) synth_fill_kernel(float* rows, uint64_t first_pos, uint64_t n,
                                                         uint32_t dim, uint32_t pitch, uint64_t seed,
                                                         uint64_t first_row, uint32_t clusters) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
    for (uint64_t r = warp; r < n; r += nwarps) {
        const uint64_t row = first_row + r;
        const uint64_t centre =
            clusters ? mix64(seed ^ 0xC2B2AE3D27D4EB4Full ^ (row * 0x9E3779B97F4A7C15ull)) % clusters : 0;
        unsigned long long ss = 0;
        for (uint32_t c = lane; c < dim; c += 32) {
            const long long x = synth_int(seed, row, centre, clusters, c);
            ss += static_cast<unsigned long long>(x * x);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
        const double norm = __dsqrt_rn(static_cast<double>(ss));
        float* out = rows + (first_pos + r) * pitch;
        for (uint32_t c = lane; c < pitch; c += 32) {
            float v = 0.f;
            if (c < dim && ss)
                v = __double2float_rn(
                    __ddiv_rn(static_cast<double>(synth_int(seed, row, centre, clusters, c)), norm));
            out[c] = v;
        }
    }
}

cudaError_t launch_synth_fill(float* rows, uint64_t first_pos, uint64_t n, uint32_t dim, uint32_t pitch,
                              uint64_t seed, uint64_t first_row, uint32_t clusters, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    uint64_t blocks = (n + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    synth_fill_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(rows, first_pos, n, dim, pitch, seed,
                                                                  first_row, clusters);
    return cudaGetLastError();
}

// ---- merge of G per-shard sorted top-k lists (the exchange step of the row-sharded flat
// index).  Global order = (score desc, global storage position asc), the stable-sort order of
// flat.rs:116 over the concatenated shards.  rank(e) = own index + Σ_other lists #better. -------
__device__ __forceinline__ bool better(double sa, uint64_t pa, double sb, uint64_t pb) {
    return sa > sb || (sa == sb && pa < pb);
}

// rank_stride: bytes between consecutive shards' arrays (0 = dense [G][nq][k] / [G][nq])
__global__ void __launch_bounds__(256) merge_topk_kernel(uint32_t G, uint32_t nq, uint32_t k,
                                                         const uint64_t* __restrict__ ids0,
                                                         const double* __restrict__ scores0,
                                                         const uint64_t* __restrict__ pos0,
                                                         const uint32_t* __restrict__ counts0, uint64_t rank_stride,
                                                         uint64_t* out_ids, double* out_scores,
                                                         uint64_t* out_pos, uint32_t* out_counts) {
    const uint32_t q = blockIdx.x;
    const uint64_t st8 = rank_stride ? rank_stride : static_cast<uint64_t>(nq) * k * 8;
    const uint64_t st4 = rank_stride ? rank_stride : static_cast<uint64_t>(nq) * 4;
    auto ids_of = [&](uint32_t g) { return reinterpret_cast<const uint64_t*>(reinterpret_cast<const char*>(ids0) + g * st8); };
    auto sc_of = [&](uint32_t g) { return reinterpret_cast<const double*>(reinterpret_cast<const char*>(scores0) + g * st8); };
    auto pos_of = [&](uint32_t g) { return reinterpret_cast<const uint64_t*>(reinterpret_cast<const char*>(pos0) + g * st8); };
    auto cnt_of = [&](uint32_t g) { return reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(counts0) + g * st4); };
    uint32_t total = 0;
    for (uint32_t g = 0; g < G; ++g) total += cnt_of(g)[q];
    const uint32_t cnt = total < k ? total : k;
    for (uint32_t e = threadIdx.x; e < G * k; e += blockDim.x) {
        const uint32_t g = e / k, i = e - g * k;
        if (i >= cnt_of(g)[q]) continue;
        const size_t me = static_cast<size_t>(q) * k + i;
        const double sc = sc_of(g)[me];
        const uint64_t pp = pos_of(g)[me];
        uint32_t rank = i;
        for (uint32_t g2 = 0; g2 < G; ++g2) {
            if (g2 == g) continue;
            const size_t base = static_cast<size_t>(q) * k;
            const double* s2 = sc_of(g2);
            const uint64_t* p2 = pos_of(g2);
            uint32_t lo = 0, hi = cnt_of(g2)[q];  // first index in list g2 NOT better than me
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (better(s2[base + mid], p2[base + mid], sc, pp)) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < cnt) {
            const size_t o = static_cast<size_t>(q) * k + rank;
            out_ids[o] = ids_of(g)[me];
            out_scores[o] = sc;
            if (out_pos) out_pos[o] = pp;
        }
    }
    for (uint32_t i = cnt + threadIdx.x; i < k; i += blockDim.x) {
        const size_t o = static_cast<size_t>(q) * k + i;
        out_ids[o] = ~0ull;
        out_scores[o] = 0.0;
        if (out_pos) out_pos[o] = ~0ull;
    }
    if (threadIdx.x == 0) out_counts[q] = cnt;
}

cudaError_t launch_merge_topk(uint32_t G, uint32_t nq, uint32_t k, const uint64_t* ids,
                              const double* scores, const uint64_t* pos, const uint32_t* counts,
                              uint64_t rank_stride, uint64_t* out_ids, double* out_scores, uint64_t* out_pos,
                              uint32_t* out_counts, cudaStream_t s) {
    if (nq == 0 || k == 0) return cudaSuccess;
    merge_topk_kernel<<<nq, 256, 0, s>>>(G, nq, k, ids, scores, pos, counts, rank_stride, out_ids, out_scores,
                                         out_pos, out_counts);
    return cudaGetLastError();
}

}  // namespace vl
