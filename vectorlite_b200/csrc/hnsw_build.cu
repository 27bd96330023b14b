// hnsw_build.cu — HNSW construction on the device (bulk build of an empty index).
//
// Reference: HNSWIndex::add (src/index/hnsw.rs:363-399) inserts one vector at a time through crate
// hnsw 0.11 `insert` (search with ef_construction, link, back-link), single threaded: minutes to hours
// for 1M × 384 ("time to first query").  Same published algorithm here — level draw,
// greedy descent, ef_construction beam, Algorithm-4 neighbour selection with pruned fill, M links per
// new node, back-links capped at M / M0 — restructured for a GPU:
//
//   * layer by layer, top down: layer l is a graph over S_l = {level >= l}; its insertion order is
//     S_{l+1} first (the hubs, seeded from the entry point), then the nodes whose level is exactly l,
//     which reach layer l by greedy descent through the already complete layers above;
//   * inside a layer, nodes are inserted in batches of at most 1/32 of the nodes already linked
//     (1, 1, …, 2, 3, … up to BATCH_MAX): every node of a batch searches the same frozen graph with the
//     search kernel (hnsw_search.cu, construction mode: the query is a row of the arena, the output is the
//     whole sorted beam), then one warp per node selects its neighbours and writes its list, and one
//     warp per DISTINCT target applies the batch's back-links (pairs sorted by target with a radix
//     sort, so no locks: a target's list is rewritten by exactly one warp);
//   * all distances are true fp32 metrics on the arena rows, warp-cooperative (8 rows per round,
//     whole-row coalesced 128-bit loads, transposed butterfly reduction).
//
// Nodes of one batch do not see each other (they link to the frozen graph only), which costs at most
// ~3 % of the candidate edges; recall at equal (M, M0, ef_construction, ef) is asserted against the
// host builder in tests/test_hnsw_gpu.py.
#include <algorithm>
#include <cub/device/device_radix_sort.cuh>
#include <vector>

#include "hnsw.h"
#include "hnsw_state.h"
#include "kernels.h"

namespace vl {

namespace {

constexpr int HB_WARPS = 4;                  // warps (= nodes / targets) per CTA
constexpr int HB_THREADS = HB_WARPS * 32;
constexpr int HB_MAX_DEG = 64;
constexpr int HB_MAX_CAND = 128;             // candidates of one back-link re-selection (existing + incoming)
constexpr uint32_t HB_BATCH_MAX = 16384;

struct BuildGraph {
    const float* rows;
    const float* inv_norm;
    uint32_t* adj0;
    const uint32_t* upper_off;
    uint32_t* upper;
    uint32_t pitch4, M, M0;
};

__device__ __forceinline__ uint32_t* adj_of(const BuildGraph& g, uint32_t node, int lvl) {
    return lvl == 0 ? g.adj0 + static_cast<size_t>(node) * g.M0
                    : g.upper + (static_cast<size_t>(__ldg(g.upper_off + node)) + lvl - 1) * g.M;
}

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

// 8 partial sums per lane → lane 4r holds the total of row r
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

// distances (lower is closer, the search kernel's convention) from node `a` to ids[0..cnt), cnt <= 8: the
// value for ids[r] is returned in lanes 4r..4r+3 (all lanes of the warp participate)
template <int METRIC>
__device__ __forceinline__ float dist8(const BuildGraph& g, uint32_t a, const uint32_t* ids, int cnt, int lane) {
    const float4* rows4 = reinterpret_cast<const float4*>(g.rows);
    const float4* ar = rows4 + static_cast<size_t>(a) * g.pitch4;
    float acc[8];
    uint32_t nid[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        acc[r] = 0.f;
        nid[r] = r < cnt ? ids[r] : HNSW_NONE;
    }
    for (uint32_t c0 = 0; c0 < g.pitch4; c0 += 32) {
        const uint32_t col = c0 + lane;
        if (col < g.pitch4) {
            const float4 q = __ldg(ar + col);
            float4 v[8];
#pragma unroll
            for (int r = 0; r < 8; ++r)
                v[r] = nid[r] != HNSW_NONE ? __ldg(rows4 + static_cast<size_t>(nid[r]) * g.pitch4 + col)
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < 8; ++r) acc[r] = acc4<METRIC>(acc[r], v[r], q);
        }
    }
    const float s = red8(acc, lane);
    const int r = lane >> 2;
    if (r >= cnt) return INFINITY;
    if (METRIC == COSINE) return 1.0f - s * __ldg(g.inv_norm + a) * __ldg(g.inv_norm + ids[r]);
    if (METRIC == DOT) return -s;
    return s;
}

__device__ __forceinline__ unsigned long long cand_key(float d, uint32_t node) {
    return (static_cast<unsigned long long>(f32_orderable(d)) << 32) | node;
}
__device__ __forceinline__ float cand_dist(unsigned long long k) { return orderable_f32(static_cast<uint32_t>(k >> 32)); }
__device__ __forceinline__ uint32_t cand_node(unsigned long long k) { return static_cast<uint32_t>(k); }

// Algorithm 4 with keepPrunedConnections (hnsw_host.cpp `select`): candidates ascending by distance to the
// base node; a candidate is kept unless an already kept node is closer to it than the base is.  One warp;
// sel / pruned are per-warp shared arrays of `limit` (<= 64) entries.  Returns the number selected.
template <int METRIC>
__device__ int warp_select(const BuildGraph& g, const unsigned long long* cand, int ncand, int limit, uint32_t* sel,
                           uint32_t* pruned, int lane) {
    int nsel = 0, npr = 0;
    for (int ci = 0; ci < ncand && nsel < limit; ++ci) {
        const unsigned long long ck = cand[ci];
        const uint32_t c = cand_node(ck);
        const float dc = cand_dist(ck);
        bool good = true;
        for (int s0 = 0; s0 < nsel && good; s0 += 8) {
            const float d = dist8<METRIC>(g, c, sel + s0, min(8, nsel - s0), lane);
            good = !__any_sync(0xFFFFFFFFu, d < dc);
        }
        if (good) {
            if (lane == 0) sel[nsel] = c;
            ++nsel;
        } else if (npr < limit) {
            if (lane == 0) pruned[npr] = c;
            ++npr;
        }
        __syncwarp();
    }
    for (int i = 0; i < npr && nsel < limit; ++i, ++nsel)
        if (lane == 0) sel[nsel] = pruned[i];
    __syncwarp();
    return nsel;
}

// ---- forward links: one warp per new node ---------------------------------------------------------
// W[w][0..wcount[w]) = sorted beam of node order[w] (search-kernel keys: distance << 32 | node << 1 | flag).
// Writes the node's list on `lvl` and one (target, slot) pair per link: pair_keys[w·M + i] = target << 32 |
// (w·M + i), pair_dist[w·M + i] = distance; unused slots carry target 0xFFFFFFFF (sorted to the end).
template <int METRIC>
__global__ void __launch_bounds__(HB_THREADS) hnsw_build_select_kernel(BuildGraph g, const uint32_t* order, uint32_t nb,
                                                                       int lvl, const unsigned long long* W,
                                                                       const uint32_t* wcount, uint32_t wstride,
                                                                       unsigned long long* pair_keys, float* pair_dist) {
    extern __shared__ __align__(16) unsigned char hb_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t w = blockIdx.x * HB_WARPS + warp;
    if (w >= nb) return;
    // per warp: cand[wstride] u64 | sel[64] | pruned[64]
    const size_t per_warp = static_cast<size_t>(wstride) * 8 + 2 * HB_MAX_DEG * 4;
    unsigned long long* cand = reinterpret_cast<unsigned long long*>(hb_smem + warp * per_warp);
    uint32_t* sel = reinterpret_cast<uint32_t*>(cand + wstride);
    uint32_t* pruned = sel + HB_MAX_DEG;
    const uint32_t q = order[w];
    const int nc = static_cast<int>(min(wcount[w], wstride));
    for (int i = lane; i < nc; i += 32) {
        const unsigned long long k = W[static_cast<size_t>(w) * wstride + i];
        cand[i] = (k & 0xFFFFFFFF00000000ull) | ((static_cast<uint32_t>(k) >> 1) & 0x7FFFFFFFu);
    }
    __syncwarp();
    const int M = static_cast<int>(g.M);
    const int nsel = warp_select<METRIC>(g, cand, nc, M, sel, pruned, lane);
    uint32_t* a = adj_of(g, q, lvl);
    const int cap = static_cast<int>(lvl == 0 ? g.M0 : g.M);
    for (int i = lane; i < cap; i += 32) a[i] = i < nsel ? sel[i] : HNSW_NONE;
    for (int i = lane; i < M; i += 32) {
        const size_t slot = static_cast<size_t>(w) * M + i;
        if (i < nsel) {
            // distance of the link = the candidate's beam distance (find it: the lists are short)
            float d = 0.f;
            for (int j = 0; j < nc; ++j)
                if (cand_node(cand[j]) == sel[i]) { d = cand_dist(cand[j]); break; }
            pair_keys[slot] = (static_cast<unsigned long long>(sel[i]) << 32) | static_cast<uint32_t>(slot);
            pair_dist[slot] = d;
        } else {
            pair_keys[slot] = (0xFFFFFFFFull << 32) | static_cast<uint32_t>(slot);
            pair_dist[slot] = 0.f;
        }
    }
}

// ---- back-links: one warp per distinct target -------------------------------------------------------
// sorted[i] = target << 32 | slot, ascending; the warp whose pair index starts a run of equal targets owns
// that target: it appends the run's sources to the target's list, or, when the list would overflow,
// re-selects the list from (existing ∪ incoming) with Algorithm 4.
template <int METRIC>
__global__ void __launch_bounds__(HB_THREADS) hnsw_build_backlink_kernel(BuildGraph g, const uint32_t* order, int lvl,
                                                                         const unsigned long long* sorted,
                                                                         const float* pair_dist, uint32_t npairs) {
    __shared__ unsigned long long s_cand[HB_WARPS][HB_MAX_CAND];
    __shared__ uint32_t s_sel[HB_WARPS][HB_MAX_DEG], s_pr[HB_WARPS][HB_MAX_DEG], s_old[HB_WARPS][HB_MAX_DEG];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t i0 = blockIdx.x * HB_WARPS + warp;
    if (i0 >= npairs) return;
    const unsigned long long k0 = sorted[i0];
    const uint32_t t = static_cast<uint32_t>(k0 >> 32);
    if (t == 0xFFFFFFFFu) return;
    if (i0 > 0 && static_cast<uint32_t>(sorted[i0 - 1] >> 32) == t) return;   // not the start of the run
    uint32_t inc = 1;
    while (i0 + inc < npairs && inc < HB_MAX_DEG && static_cast<uint32_t>(sorted[i0 + inc] >> 32) == t) ++inc;
    const uint32_t M = g.M;
    uint32_t* a = adj_of(g, t, lvl);
    const uint32_t cap = lvl == 0 ? g.M0 : g.M;
    // existing neighbours (lists are dense prefixes)
    uint32_t m = 0;
    for (uint32_t j0 = 0; j0 < cap; j0 += 32) {
        const uint32_t j = j0 + lane;
        const uint32_t v = j < cap ? a[j] : HNSW_NONE;
        if (j < cap) s_old[warp][j] = v;
        m += __popc(__ballot_sync(0xFFFFFFFFu, v != HNSW_NONE));
    }
    __syncwarp();
    if (m + inc <= cap) {
        for (uint32_t j = lane; j < inc; j += 32) {
            const uint32_t slot = static_cast<uint32_t>(sorted[i0 + j]);
            a[m + j] = order[slot / M];
        }
        return;
    }
    // overflow: candidates = existing (distances recomputed) ∪ incoming (distance known), sorted ascending
    for (uint32_t s0 = 0; s0 < m; s0 += 8) {
        const int cnt = static_cast<int>(min(8u, m - s0));
        const float d = dist8<METRIC>(g, t, &s_old[warp][s0], cnt, lane);
        const int r = lane >> 2;
        if ((lane & 3) == 0 && r < cnt) s_cand[warp][s0 + r] = cand_key(d, s_old[warp][s0 + r]);
    }
    for (uint32_t j = lane; j < inc; j += 32) {
        const uint32_t slot = static_cast<uint32_t>(sorted[i0 + j]);
        s_cand[warp][m + j] = cand_key(pair_dist[slot], order[slot / M]);
    }
    const int total = static_cast<int>(m + inc);
    for (int j = total + lane; j < HB_MAX_CAND; j += 32) s_cand[warp][j] = ~0ull;
    __syncwarp();
    for (int k2 = 2; k2 <= HB_MAX_CAND; k2 <<= 1)       // bitonic sort of 128 keys by one warp
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int x = lane; x < HB_MAX_CAND / 2; x += 32) {
                const int i = ((x & ~(j - 1)) << 1) | (x & (j - 1));
                const int pp = i | j;
                const bool asc = (i & k2) == 0;
                const unsigned long long u = s_cand[warp][i], v = s_cand[warp][pp];
                if ((u > v) == asc) { s_cand[warp][i] = v; s_cand[warp][pp] = u; }
            }
            __syncwarp();
        }
    const int nsel = warp_select<METRIC>(g, s_cand[warp], total, static_cast<int>(cap), s_sel[warp], s_pr[warp], lane);
    for (uint32_t j = lane; j < cap; j += 32) a[j] = static_cast<int>(j) < nsel ? s_sel[warp][j] : HNSW_NONE;
}

template <int METRIC>
int run_batch(const BuildGraph& bg, const HnswDeviceGraph& dg, const float* d_rows, uint32_t pitch, uint32_t dim,
              const uint32_t* d_order, uint32_t nb, int lvl, bool entry_only, uint32_t efc, unsigned long long* d_W,
              uint32_t* d_wcount, unsigned long long* d_pairs, unsigned long long* d_pairs_sorted, float* d_pair_dist,
              void* d_temp, size_t temp_bytes, cudaStream_t s) {
    int st = hnsw_launch_build_search(dg, d_rows, pitch, dim, METRIC, d_order, nb, lvl, entry_only, efc, d_W, efc,
                                      d_wcount, s);
    if (st) return st;
    const size_t sel_smem = HB_WARPS * (static_cast<size_t>(efc) * 8 + 2 * HB_MAX_DEG * 4);
    auto ksel = hnsw_build_select_kernel<METRIC>;
    if (sel_smem > 40 * 1024) cudaFuncSetAttribute(ksel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sel_smem));
    ksel<<<(nb + HB_WARPS - 1) / HB_WARPS, HB_THREADS, sel_smem, s>>>(bg, d_order, nb, lvl, d_W, d_wcount, efc, d_pairs,
                                                                     d_pair_dist);
    const uint32_t npairs = nb * bg.M;
    if (cub::DeviceRadixSort::SortKeys(d_temp, temp_bytes, d_pairs, d_pairs_sorted, static_cast<int>(npairs), 0, 64, s) !=
        cudaSuccess)
        return 6;
    hnsw_build_backlink_kernel<METRIC><<<(npairs + HB_WARPS - 1) / HB_WARPS, HB_THREADS, 0, s>>>(
        bg, d_order, lvl, d_pairs_sorted, d_pair_dist, npairs);
    return cudaGetLastError() == cudaSuccess ? 0 : 6;
}

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { cudaFree(p); }
    bool alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1) == cudaSuccess; }
};

}  // namespace

// Builds the whole graph of `s` (levels, upper_off and inv_norm already assigned on the host, adjacency
// empty) on the device from the arena rows and copies the adjacency back into the host vectors.
int hnsw_build_device(HnswState* s, const float* d_rows, uint32_t pitch, cudaStream_t stream, uint64_t* launches) {
    const uint32_t n = static_cast<uint32_t>(s->level.size());
    if (n == 0) return 0;
    if (s->M > HB_MAX_DEG || s->M0 > HB_MAX_DEG || s->efc > 2048) return 9;
    const uint32_t efc = std::max<uint32_t>(s->efc, s->M);
    // ---- entry point: the first node of the highest level -------------------------------------------
    int max_level = 0;
    uint32_t entry = 0;
    for (uint32_t i = 0; i < n; ++i)
        if (s->level[i] > max_level) { max_level = s->level[i]; entry = i; }
    // ---- device graph arrays (same buffers the search uses) ------------------------------------------
    if (int st = hnsw_reserve_device(s, n)) return st;
    cudaMemsetAsync(s->d_adj0, 0xFF, static_cast<size_t>(n) * s->M0 * 4, stream);
    if (!s->upper.empty()) cudaMemsetAsync(s->d_upper, 0xFF, s->upper.size() * 4, stream);
    cudaMemcpyAsync(s->d_upper_off, s->upper_off.data(), static_cast<size_t>(n) * 4, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(s->d_level, s->level.data(), n, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(s->d_inv_norm, s->inv_norm.data(), static_cast<size_t>(n) * 4, cudaMemcpyHostToDevice, stream);

    DevBuf order, W, wcount, pairs, pairs_sorted, pair_dist, temp;
    const uint32_t bmax = std::min<uint32_t>(HB_BATCH_MAX, n);
    size_t temp_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, temp_bytes, static_cast<unsigned long long*>(nullptr),
                                   static_cast<unsigned long long*>(nullptr), static_cast<int>(bmax * s->M), 0, 64, stream);
    if (!order.alloc(static_cast<size_t>(n) * 4) || !W.alloc(static_cast<size_t>(bmax) * efc * 8) ||
        !wcount.alloc(static_cast<size_t>(bmax) * 4) || !pairs.alloc(static_cast<size_t>(bmax) * s->M * 8) ||
        !pairs_sorted.alloc(static_cast<size_t>(bmax) * s->M * 8) || !pair_dist.alloc(static_cast<size_t>(bmax) * s->M * 4) ||
        !temp.alloc(temp_bytes))
        return 7;

    BuildGraph bg;
    bg.rows = d_rows; bg.inv_norm = s->d_inv_norm; bg.adj0 = s->d_adj0; bg.upper_off = s->d_upper_off;
    bg.upper = s->d_upper; bg.pitch4 = pitch / 4; bg.M = s->M; bg.M0 = s->M0;
    HnswDeviceGraph dg;
    dg.adj0 = s->d_adj0; dg.upper_off = s->d_upper_off; dg.upper = s->d_upper; dg.level = s->d_level;
    dg.deleted = s->d_deleted; dg.ids = s->d_ids; dg.inv_norm = s->d_inv_norm;
    dg.n = n; dg.M = s->M; dg.M0 = s->M0; dg.entry = entry; dg.max_level = max_level;

    std::vector<uint32_t> ord;
    ord.reserve(n);
    uint64_t nl = 0;
    for (int lvl = max_level; lvl >= 0; --lvl) {
        // insertion order of this layer: the entry point, then the rest of S_{l+1}, then level == l
        ord.clear();
        ord.push_back(entry);
        for (uint32_t i = 0; i < n; ++i)
            if (s->level[i] > lvl && i != entry) ord.push_back(i);
        const uint32_t prefix = static_cast<uint32_t>(ord.size());
        for (uint32_t i = 0; i < n; ++i)
            if (s->level[i] == lvl && i != entry) ord.push_back(i);
        const uint32_t len = static_cast<uint32_t>(ord.size());
        if (cudaMemcpyAsync(order.p, ord.data(), static_cast<size_t>(len) * 4, cudaMemcpyHostToDevice, stream) != cudaSuccess)
            return 6;
        uint32_t pos = 1;   // the entry point is "inserted" with an empty list
        while (pos < len) {
            uint32_t nb = std::max<uint32_t>(1, std::min<uint32_t>(pos / 32, bmax));
            const bool in_prefix = pos < prefix;
            nb = std::min<uint32_t>(nb, (in_prefix ? prefix : len) - pos);
            const uint32_t* d_ord = static_cast<const uint32_t*>(order.p) + pos;
            // hubs (and every node of the top layer) start from the entry point; the others descend
            const bool entry_only = in_prefix || lvl == max_level;
            int st;
#define VL_HB_RUN(MET)                                                                                              \
    run_batch<MET>(bg, dg, d_rows, pitch, s->dim, d_ord, nb, lvl, entry_only, efc,                                  \
                   static_cast<unsigned long long*>(W.p), static_cast<uint32_t*>(wcount.p),                          \
                   static_cast<unsigned long long*>(pairs.p), static_cast<unsigned long long*>(pairs_sorted.p),      \
                   static_cast<float*>(pair_dist.p), temp.p, temp_bytes, stream)
            switch (s->metric) {
                case COSINE: st = VL_HB_RUN(COSINE); break;
                case EUCLIDEAN: st = VL_HB_RUN(EUCLIDEAN); break;
                case MANHATTAN: st = VL_HB_RUN(MANHATTAN); break;
                default: st = VL_HB_RUN(DOT); break;
            }
#undef VL_HB_RUN
            if (st) return st;
            nl += 4;
            pos += nb;
        }
        // the host order buffer is reused by the next layer: wait for this layer's upload + kernels
        if (cudaStreamSynchronize(stream) != cudaSuccess) return 6;
    }
    // ---- adjacency back to the host copy (export / persistence / incremental host inserts) -------------
    cudaMemcpyAsync(s->adj0.data(), s->d_adj0, static_cast<size_t>(n) * s->M0 * 4, cudaMemcpyDeviceToHost, stream);
    if (!s->upper.empty())
        cudaMemcpyAsync(s->upper.data(), s->d_upper, s->upper.size() * 4, cudaMemcpyDeviceToHost, stream);
    if (cudaStreamSynchronize(stream) != cudaSuccess) return 6;
    s->entry = entry;
    s->max_level = max_level;
    s->dirty = true;   // ids / deleted flags still have to be uploaded by hnsw_upload
    if (launches) *launches += nl;
    return 0;
}

}  // namespace vl
