// hnsw_device.cuh — device helpers of the HNSW search kernel (hnsw_search.cu): parameters, distance rounds,
// visited tags, result phase.
#pragma once
#include "hnsw_state.h"
#include "kernels.h"

namespace vl {

constexpr int HN_THREADS = 128;
constexpr int HN_WARPS = HN_THREADS / 32;
constexpr int HN_MAX_DEG = 64;
constexpr int HN_MAX_EXPAND = 4;                 // pool entries expanded per step (throughput / construction kernels)
constexpr int HN_MAX_EXPAND_WIDE = 8;            // 16-warp CTAs (a few queries in flight: latency): twice as many, half the steps
constexpr int HN_MAX_CAND = HN_MAX_DEG * HN_MAX_EXPAND;
constexpr int HN_EF_MAX = 2048;   // widest internal beam
constexpr int HN_REFINE_MAX = 1024;  // head of the beam refined in fp32 when the traversal gathers bf16 rows
constexpr int HN_K_MAX = 2048;      // the reference accepts any k (hnsw.rs:437: ef = min(k, len)); k <= widest beam here
constexpr int HN_BEAM_MULT = 8;     // internal beam = 8 x nominal ef (see hnsw_launch_search)

struct HnswParams {
    HnswDeviceGraph g;
    const float* rows;
    const void* rows_bf16 = nullptr;      // bf16 mirror of the rows ([n][pitch], cosine: pre-scaled by 1/‖row‖): the
                                          // traversal gathers 768 B per evaluated node instead of 1.5 KB; the final k
                                          // are re-scored in f64 from the fp32 rows either way
    const float* queries;
    uint32_t pitch, dim, k, ef, vis_mask, beam_cap, cand_cap;
    uint32_t expand = 1;                  // pool entries expanded per step on the beam level (<= the kernel's MAXE)
    uint32_t refine = 0;                  // bf16 gathers: beam entries re-evaluated from the fp32 rows before k are taken
    uint32_t rerank = 0;                  // beam entries re-scored in f64 before the top k are taken (>= k; = k when the
                                          // traversal's distances are fp32; 4k..64+ when they come from the bf16 mirror,
                                          // whose rounding (~1e-3 on a cosine) reorders near-ties among the best entries)
    uint32_t merge = 0;                   // CTAs of >= 4 warps (search): the pool stays sorted, a step's survivors are merged
                                          // in by the whole CTA (rank counting) instead of inserted one by one by warp 0
    uint32_t score_mode = 0;              // 0: exact flat similarity, 1: the reference's quantised score (hnsw.rs:478,51-75)
    uint64_t* out_ids;
    double* out_scores;
    uint32_t* out_counts;
    unsigned long long* visited;
    uint32_t* visited_per_query = nullptr;   // when set (results written straight to pinned host memory): no atomics
    // ---- construction mode (hnsw_build.cu): the queries are rows of the arena -------------------------
    const uint32_t* order = nullptr;      // [nq] node whose row is query qi
    int stop_level = 0;                   // beam search on this level; greedy descent above it
    uint32_t entry_only = 0;              // skip the descent: seed the beam with the entry point
    unsigned long long* out_keys = nullptr;  // [nq][out_stride] sorted beam (orderable distance << 32 | node << 1 | flag)
    uint32_t out_stride = 0;
};

__device__ __forceinline__ uint32_t vis_hash(uint32_t id) { return (id * 2654435761u) >> 7; }
// 16-bit tag of a node in the visited cache (never 0 = empty); a tag collision in the SAME slot with a
// different node (p ≈ 2^-15 per occupied-slot lookup) makes that node look visited — a negligible recall cost
// that halves the cache's shared memory and doubles the resident CTAs for wide beams
__device__ __forceinline__ uint16_t vis_tag(uint32_t id) { return static_cast<uint16_t>(((id * 0x9E3779B1u) >> 16) | 1u); }

template <int METRIC>
__device__ __forceinline__ float acc4(float acc, const float4& v, const float4& q) {
    if (METRIC == COSINE || METRIC == DOT) {
        acc = fmaf(v.x, q.x, acc); acc = fmaf(v.y, q.y, acc); acc = fmaf(v.z, q.z, acc); acc = fmaf(v.w, q.w, acc);
    } else if (METRIC == EUCLIDEAN) {
        float d;
        d = v.x - q.x; acc = fmaf(d, d, acc); d = v.y - q.y; acc = fmaf(d, d, acc);
        d = v.z - q.z; acc = fmaf(d, d, acc); d = v.w - q.w; acc = fmaf(d, d, acc);
    } else {
        acc += fabsf(v.x - q.x); acc += fabsf(v.y - q.y); acc += fabsf(v.z - q.z); acc += fabsf(v.w - q.w);
    }
    return acc;
}

__device__ __forceinline__ float red8(const float (&a)[8], int lane) {
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

// 4 bf16 elements (one 64-bit load) against 4 fp32 query elements; the same element order as the fp32 layout
template <int METRIC>
__device__ __forceinline__ float acc4_bf16(float acc, const uint2& v, const float4& q) {
    const float x0 = __uint_as_float(v.x << 16), x1 = __uint_as_float(v.x & 0xFFFF0000u);
    const float x2 = __uint_as_float(v.y << 16), x3 = __uint_as_float(v.y & 0xFFFF0000u);
    return acc4<METRIC>(acc, make_float4(x0, x1, x2, x3), q);
}

// distance "lower is closer" from the raw accumulation
template <int METRIC>
__device__ __forceinline__ float to_dist(float acc, float invn, float invq) {
    if (METRIC == COSINE) return 1.0f - acc * invn * invq;
    if (METRIC == DOT) return -acc;
    return acc;
}

__device__ __forceinline__ unsigned long long beam_key(float d, uint32_t node) {
    return (static_cast<unsigned long long>(f32_orderable(d)) << 32) | (static_cast<unsigned long long>(node) << 1);
}

// Warp-wide argmax (MAX) / argmin of 63-bit pool values (orderable distance << 31 | node) with their indices, on
// the hardware reduction unit (redux.sync, 32-bit): distance word first, node word among the ties, then one
// ballot + shuffle for the index — ~8 instructions instead of a 5-round 64-bit shuffle butterfly (~40).
// Lanes without a value pass 0 (MAX) / ~0 (min) and index -1.
template <bool MAX>
__device__ __forceinline__ void warp_arg63(unsigned long long& v, int& idx) {
    const unsigned FULL = 0xFFFFFFFFu;
    const uint32_t hi = static_cast<uint32_t>(v >> 31), lo = static_cast<uint32_t>(v) & 0x7FFFFFFFu;
    const uint32_t mh = MAX ? __reduce_max_sync(FULL, hi) : __reduce_min_sync(FULL, hi);
    const bool c1 = hi == mh;
    const uint32_t ml = MAX ? __reduce_max_sync(FULL, c1 ? lo : 0u) : __reduce_min_sync(FULL, c1 ? lo : 0xFFFFFFFFu);
    const unsigned b = __ballot_sync(FULL, c1 && lo == ml);
    const int src = __ffs(b) - 1;
    v = (static_cast<unsigned long long>(mh) << 31) | ml;
    idx = __shfl_sync(FULL, idx, src);
}

// Result phase shared by the search kernels: s_beam[0..size) holds the pool sorted ascending by (distance, node).
// Construction mode hands the whole beam to the neighbour selection; search mode takes the first k non-deleted
// entries (hnsw.rs:472-475), re-scores them in f64 with the reference's Flat formulae and orders them.
template <int METRIC, bool BUILD, int THREADS>
__device__ __forceinline__ void finish_query(const HnswParams& p, uint32_t qi, const unsigned long long* s_beam,
                                             int size_in, const float* q, double* s_ex, uint32_t* s_rid,
                                             double* s_qs, unsigned long long n_eval) {
    __shared__ int s_rcount;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (BUILD) {   // construction: hand the whole sorted beam (ascending distance) to the neighbour selection
        const int size = size_in;
        for (int i = tid; i < size; i += THREADS) p.out_keys[static_cast<size_t>(qi) * p.out_stride + i] = s_beam[i];
        if (tid == 0) p.out_counts[qi] = static_cast<uint32_t>(size);
        return;
    }
    // ---- results: first R >= k non-deleted beam entries (hnsw.rs:472-475), exact f64 re-score, best k of them ------
    const int R = static_cast<int>(p.rerank > p.k ? p.rerank : p.k);
    if (warp == 0) {
        const int size = size_in;
        int cnt = 0;
        for (int i0 = 0; i0 < size && cnt < R; i0 += 32) {
            const int i = i0 + lane;
            uint32_t node = HNSW_NONE;
            bool ok = false;
            if (i < size) {
                node = static_cast<uint32_t>(s_beam[i] >> 1) & 0x7FFFFFFFu;
                ok = !p.g.deleted[node];
            }
            const unsigned m = __ballot_sync(0xFFFFFFFFu, ok);
            const int my = cnt + __popc(m & ((1u << lane) - 1));
            if (ok && my < R && my < HN_K_MAX) s_rid[my] = node;
            cnt += __popc(m);
        }
        if (lane == 0) s_rcount = min(min(cnt, R), HN_K_MAX);
    }
    __syncthreads();
    const int rc = s_rcount;                                  // entries re-scored
    const int out_n = min(rc, static_cast<int>(p.k));         // entries returned
    for (int t = tid; t < rc; t += THREADS) {
        const uint32_t node = s_rid[t];
        const float* row = p.rows + static_cast<size_t>(node) * p.pitch;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        for (uint32_t j = 0; j < p.dim; ++j) {
            const double x = static_cast<double>(row[j]), y = static_cast<double>(q[j]);
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
        s_ex[t] = sc;
        if (p.score_mode == 1) {
            // reference score mode: the u64 milli-unit distance of the functors (hnsw.rs:113-174; `as u64`
            // truncates toward zero, saturates, NaN -> 0), divided by 1000 (hnsw.rs:478) and pushed through
            // convert_distance_to_similarity (hnsw.rs:51-75), which divides cosine / dot by 1000 AGAIN
            unsigned long long d;
            if (METRIC == COSINE) {
                const double na = __dsqrt_rn(a1), nb = __dsqrt_rn(a2);
                d = (na == 0.0 || nb == 0.0) ? 1000ull
                                             : __double2ull_rz(__dmul_rn(__dsub_rn(1.0, __ddiv_rn(a0, __dmul_rn(na, nb))), 1000.0));
            } else if (METRIC == EUCLIDEAN) {
                d = __double2ull_rz(__dmul_rn(__dsqrt_rn(a0), 1000.0));
            } else if (METRIC == MANHATTAN) {
                d = __double2ull_rz(__dmul_rn(a0, 1000.0));
            } else {
                const double c = a0 != a0 ? a0 : fmin(fmax(a0, -1000.0), 1000.0);   // f64::clamp keeps NaN
                d = c != c ? 0ull : __double2ull_rz(__dsub_rn(1000.0, c));
            }
            const double dist = __ddiv_rn(__ull2double_rn(d), 1000.0);
            double q;
            if (METRIC == COSINE) q = __dsub_rn(1.0, __ddiv_rn(dist, 1000.0));
            else if (METRIC == DOT) q = fmin(fmax(__ddiv_rn(__dsub_rn(1000.0, dist), 1000.0), 0.0), 1.0);
            else q = __ddiv_rn(1.0, __dadd_rn(1.0, dist));
            s_qs[t] = q;
        }
    }
    __syncthreads();
    for (int t = tid; t < rc; t += THREADS) {  // final order: score desc, insertion order asc (hnsw.rs:493)
        const double me = s_ex[t];
        const uint32_t mn = s_rid[t];
        int rank = 0;
        if (p.score_mode == 1) {   // quantised score desc; ties (the reference keeps the crate's order) by exact score
            const double mq = s_qs[t];
            for (int j = 0; j < rc; ++j)
                rank += (s_qs[j] > mq) ||
                        (s_qs[j] == mq && ((s_ex[j] > me) || (s_ex[j] == me && s_rid[j] < mn)));
        } else {
            for (int j = 0; j < rc; ++j) rank += (s_ex[j] > me) || (s_ex[j] == me && s_rid[j] < mn);
        }
        if (rank >= out_n) continue;
        const size_t o = static_cast<size_t>(qi) * p.k + rank;
        p.out_ids[o] = p.g.ids[mn];
        p.out_scores[o] = p.score_mode == 1 ? s_qs[t] : me;
    }
    for (int i = out_n + tid; i < static_cast<int>(p.k); i += THREADS) {
        const size_t o = static_cast<size_t>(qi) * p.k + i;
        p.out_ids[o] = ~0ull;
        p.out_scores[o] = 0.0;
    }
    if (tid == 0) {
        p.out_counts[qi] = static_cast<uint32_t>(out_n);
        if (p.visited_per_query) p.visited_per_query[qi] = static_cast<uint32_t>(n_eval);
        else if (p.visited) atomicAdd(p.visited, n_eval);
    }
}

}  // namespace vl
