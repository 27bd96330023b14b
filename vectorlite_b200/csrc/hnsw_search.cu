// hnsw_search.cu — HNSW traversal on the device: one CTA per query.
//
// Replaces HNSWIndex::search (src/index/hnsw.rs:415-496) and the per-layer greedy / beam search
// it delegates to (crate hnsw 0.11 `nearest` → search_single_layer / search_zero_layer, SURVEY
// Appendix C): upper layers are walked greedily (beam 1), layer 0 with a beam of `ef`.
//
//  - the query lives in registers (12 floats per lane at 384-d), every warp owns a copy;
//  - the beam is a pool of 64-bit keys (orderable fp32 distance | node<<1 | expanded) in shared memory.
//    CTAs of two or more warps (every search launch except very wide beams at large query counts) keep it
//    SORTED: the survivors of a step are compacted while they are scored and merged in by the whole CTA
//    (rank counting against the other list, exact duplicates dropped), and the entries to expand are read
//    off its front with one ballot.  One-warp CTAs keep the older unsorted pool: warp 0 selects the
//    closest unexpanded entry and replaces the worst entry with warp-wide argmin / argmax reductions,
//    sorted once at the end;
//  - visited set: a direct-mapped, lossy tag cache in shared memory (no probing, no overflow).
//    Losing a tag only costs a redundant distance evaluation: a re-evaluated node is either
//    rejected by the beam threshold or found as an exact duplicate key at its insertion point;
//  - neighbour distances are warp-cooperative: each warp scores 8 neighbours at a time, every lane
//    streaming whole rows (384-d: 3×64-bit loads per row from the bf16 mirror, 3×128-bit from the fp32
//    arena; fully coalesced) and the 8 sums reduced with the transposed butterfly of the flat scan;
//  - the final k candidates are re-scored in f64 with the reference's flat formulae
//    (src/lib.rs:425-572) so HNSW scores equal Flat scores for the same ids (the reference's
//    quantised /1000 score quirk, hnsw.rs:478 + 51-75, is deliberately not reproduced);
//  - soft-deleted nodes stay in the graph and are filtered from the results (hnsw.rs:473-475).
//
// Bound: random 1.5 KB row gathers → HBM/L2 latency and bandwidth; no single roofline.  Reported:
// QPS, visited nodes per query (d_visited), recall@10 vs exact flat.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "hnsw_device.cuh"

namespace vl {

template <int METRIC, int NCH, bool BUILD, int WARPS, bool BF16>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 1 ? 32 : (WARPS == 2 ? 16 : (WARPS == 4 ? 8 : 1))) hnsw_search_kernel(HnswParams p) {
    static_assert(!BF16 || (NCH > 0 && !BUILD), "bf16 gathers: register-resident query, search mode only");
    constexpr int THREADS = WARPS * 32;
    // Entries expanded per step.  A step costs two dependent global round trips (adjacency rows, then the
    // neighbours' vectors) whatever the number of entries, so a lone query's latency is steps x ~3 us: the wide
    // CTA expands up to 8 entries at once (256 candidates = two rounds of gathers over its 16 warps).
    constexpr int MAXE = WARPS >= 16 ? HN_MAX_EXPAND_WIDE : HN_MAX_EXPAND;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [q[pitch] f32 — only when the query is not register-resident (NCH == 0)] | beam[beam_cap] u64 |
    //         vis[vis_mask+1] u16 | candidate keys[cand_cap] u64 | candidate ids[cand_cap] u32
    // Everything is sized by the host (size_pool): a narrow beam at M0 = 32 needs ~6.5–10.5 KB, so 21–32 one-warp
    // CTAs are resident per SM.
    float4* s_q = reinterpret_cast<float4*>(smem_raw);
    unsigned long long* s_beam = reinterpret_cast<unsigned long long*>(smem_raw + (NCH == 0 ? static_cast<size_t>(p.pitch) * 4 : 0));
    uint16_t* s_vis = reinterpret_cast<uint16_t*>(s_beam + p.beam_cap);
    unsigned long long* s_ck = reinterpret_cast<unsigned long long*>(s_vis + p.vis_mask + 1);   // candidate keys of this step
    uint32_t* s_cid = reinterpret_cast<uint32_t*>(s_ck + p.cand_cap);
    // merge mode (CTAs of >= 4 warps, search): second pool buffer + the compacted survivors of a step
    unsigned long long* s_beam2 = reinterpret_cast<unsigned long long*>(s_cid + p.cand_cap);
    unsigned long long* s_surv = s_beam2 + p.beam_cap;
    const bool merge_mode = WARPS >= 2 && p.merge != 0u;
    __shared__ int s_nc, s_size, s_done, s_nlive;
    __shared__ float s_invq;
    // result staging reuses the traversal's scratch (dead by then): exact scores over the candidate keys,
    // result nodes over the candidate ids (cand_cap >= k), quantised scores (reference score mode) over the
    // visited cache (>= 4 KB)
    double* s_ex = reinterpret_cast<double*>(s_ck);
    uint32_t* s_rid = s_cid;
    double* s_qs = reinterpret_cast<double*>(s_vis);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t qi = blockIdx.x;
    const uint32_t pitch4 = p.pitch / 4;
    const float4* rows4 = reinterpret_cast<const float4*>(p.rows);
    const float4* q4 = BUILD ? rows4 + static_cast<size_t>(__ldg(p.order + qi)) * pitch4
                             : reinterpret_cast<const float4*>(p.queries) + static_cast<size_t>(qi) * pitch4;
    const int top_level = BUILD ? (p.entry_only ? p.stop_level : p.g.max_level) : p.g.max_level;
    const int bottom_level = BUILD ? p.stop_level : 0;

    if (NCH == 0)
        for (uint32_t i = tid; i < pitch4; i += THREADS) s_q[i] = q4[i];
    for (uint32_t i = tid; i <= p.vis_mask; i += THREADS) s_vis[i] = 0u;
    float4 qreg[NCH > 0 ? NCH : 1];
    if (NCH > 0) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) qreg[c] = __ldg(q4 + c * 32 + lane);
    }
    __syncthreads();
    const uint32_t nch = NCH > 0 ? NCH : (pitch4 + 31) / 32;
    if (METRIC == COSINE && warp == 0) {  // 1/‖q‖ (fp32 is enough for traversal)
        float a = 0.f;
        if (NCH > 0) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const float4 v = qreg[c];
                a += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            }
        } else {
            for (uint32_t i = lane; i < pitch4; i += 32) {
                const float4 v = s_q[i];
                a += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
        if (lane == 0) s_invq = a > 0.f ? rsqrtf(a) : 0.f;
    }
    __syncthreads();
    const float invq = METRIC == COSINE ? s_invq : 1.f;

    // scores up to 8 nodes ids[0..cnt) → keys out[0..cnt)  (one warp)
    auto score8_from = [&](auto bf16_tag, const uint32_t* ids, int cnt, unsigned long long* out) {
        constexpr bool FROM_BF16 = decltype(bf16_tag)::value;
        float acc[8];
        uint32_t nid[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            acc[r] = 0.f;
            nid[r] = r < cnt ? ids[r] : HNSW_NONE;
        }
        // the owner lane's 1/‖row‖ is requested together with the rows (it used to be a second, dependent round
        // trip after the reduction)
        const int own_r = lane >> 2;
        const bool own = (lane & 3) == 0 && own_r < cnt;
        const uint32_t own_id = own ? ids[own_r] : 0u;
        const float invn = (METRIC == COSINE && own && !FROM_BF16) ? __ldg(p.g.inv_norm + own_id) : 1.f;   // mirror: pre-scaled
        if (FROM_BF16) {
            const uint2* rows2 = reinterpret_cast<const uint2*>(p.rows_bf16);
#pragma unroll
            for (int c = 0; c < (NCH > 0 ? NCH : 1); ++c) {
                uint2 v[8];
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    v[r] = nid[r] != HNSW_NONE ? __ldg(rows2 + static_cast<size_t>(nid[r]) * pitch4 + c * 32 + lane)
                                               : make_uint2(0u, 0u);
#pragma unroll
                for (int r = 0; r < 8; ++r) acc[r] = acc4_bf16<METRIC>(acc[r], v[r], qreg[c]);
            }
        } else if (NCH > 0) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                float4 v[8];
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    v[r] = nid[r] != HNSW_NONE ? __ldg(rows4 + static_cast<size_t>(nid[r]) * pitch4 + c * 32 + lane)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int r = 0; r < 8; ++r) acc[r] = acc4<METRIC>(acc[r], v[r], qreg[c]);
            }
        } else {
            for (uint32_t c = 0; c < nch; ++c) {
                const uint32_t col = c * 32 + lane;
                if (col < pitch4) {
                    const float4 q = s_q[col];
#pragma unroll
                    for (int r = 0; r < 8; ++r)
                        if (nid[r] != HNSW_NONE)
                            acc[r] = acc4<METRIC>(acc[r], __ldg(rows4 + static_cast<size_t>(nid[r]) * pitch4 + col), q);
                }
            }
        }
        const float s = red8(acc, lane);
        if (own) out[own_r] = beam_key(to_dist<METRIC>(s, invn, invq), own_id);
    };
    auto score8 = [&](const uint32_t* ids, int cnt, unsigned long long* out) {   // the traversal's gathers
        score8_from(std::integral_constant<bool, BF16>{}, ids, cnt, out);
    };

    // ---- entry point ---------------------------------------------------------------------
    // The beam is an UNSORTED pool of up to `ef` keys.  Warp 0 picks the closest unexpanded entry
    // (argmin by warp reduction) and replaces the worst entry on insertion (argmax by warp reduction):
    // no binary searches, no shifting; the pool is sorted once at the end.
    __shared__ unsigned long long s_worst;   // largest (key >> 1) in the pool once it is full, else ~0
    __shared__ int s_worst_idx;
    auto pool_recompute_worst = [&](int size, uint32_t ef) {   // warp 0, all lanes
        unsigned long long w = 0ull;
        int wi = -1;
        for (int i = lane; i < size; i += 32) {
            const unsigned long long kk = s_beam[i] >> 1;
            if (kk >= w) { w = kk; wi = i; }
        }
        warp_arg63<true>(w, wi);
        if (lane == 0) {
            s_worst = size == static_cast<int>(ef) ? w : ~0ull;
            s_worst_idx = wi;
        }
        __syncwarp();
    };
    if (warp == 0) {
        if (lane == 0) s_cid[0] = p.g.entry;
        __syncwarp();
        score8(s_cid, 1, s_ck);
        __syncwarp();
        if (lane == 0) {
            s_beam[0] = s_ck[0];
            s_size = 1;
            s_vis[vis_hash(p.g.entry) & p.vis_mask] = vis_tag(p.g.entry);
        }
        __syncwarp();
        pool_recompute_worst(1, top_level > bottom_level ? 1u : p.ef);
    }
    __syncthreads();

    unsigned long long n_eval = 1;
    for (int lvl = top_level; lvl >= bottom_level; --lvl) {
        const uint32_t ef = lvl == bottom_level ? p.ef : 1u;
        const uint32_t deg = lvl == 0 ? p.g.M0 : p.g.M;
        for (;;) {
            // ---- warp 0: the `expand` closest unexpanded pool entries, then their unvisited neighbours
            if (warp == 0) {
                const int size = s_size;
                const int expand = ef > 1u ? static_cast<int>(min(p.expand, static_cast<uint32_t>(MAXE))) : 1;
                int nc = 0, picked = 0;
                // pick first (shared memory only), then fetch: the adjacency rows of all picked entries are
                // requested together, so a step waits for ONE global-memory round trip instead of `expand`
                uint32_t node_e[MAXE];
                if (merge_mode && ef > 1u) {
                    // the pool is SORTED (merged by the whole CTA at the end of every step): the entries to expand are
                    // simply the first `expand` ones without the expanded flag — one ballot per 32 entries
#pragma unroll
                    for (int e = 0; e < MAXE; ++e) node_e[e] = HNSW_NONE;
                    if (lane == 0) s_nlive = 0;
                    for (int i0 = 0; i0 < size && picked < expand; i0 += 32) {
                        const int i = i0 + lane;
                        const unsigned long long kk = i < size ? s_beam[i] : 1ull;
                        unsigned m = __ballot_sync(0xFFFFFFFFu, !(kk & 1ull));
                        while (m && picked < expand) {
                            const int src = __ffs(m) - 1;
                            m &= m - 1;
                            const uint32_t node = __shfl_sync(0xFFFFFFFFu, static_cast<uint32_t>(kk >> 1) & 0x7FFFFFFFu, src);
#pragma unroll
                            for (int e = 0; e < MAXE; ++e)
                                if (e == picked) node_e[e] = node;
                            if (lane == src) s_beam[i] = kk | 1ull;
                            ++picked;
                        }
                    }
                    __syncwarp();
                } else {
#pragma unroll
                for (int e = 0; e < MAXE; ++e) {
                    node_e[e] = HNSW_NONE;
                    if (e >= expand || picked < e) continue;      // uniform
                    unsigned long long best = ~0ull;
                    int bi = -1;
                    for (int i = lane; i < size; i += 32) {
                        const unsigned long long kk = s_beam[i];
                        if (!(kk & 1ull) && (kk >> 1) < best) { best = kk >> 1; bi = i; }
                    }
                    warp_arg63<false>(best, bi);
                    if (bi < 0) continue;
                    ++picked;
                    node_e[e] = static_cast<uint32_t>(best) & 0x7FFFFFFFu;
                    if (lane == 0) s_beam[bi] |= 1ull;
                    __syncwarp();
                }
                }
                const uint32_t* adj_e[MAXE];
                if (lvl == 0) {
#pragma unroll
                    for (int e = 0; e < MAXE; ++e)
                        adj_e[e] = p.g.adj0 + static_cast<size_t>(node_e[e] != HNSW_NONE ? node_e[e] : 0u) * p.g.M0;
                } else {
                    uint32_t off_e[MAXE];
#pragma unroll
                    for (int e = 0; e < MAXE; ++e) off_e[e] = node_e[e] != HNSW_NONE ? __ldg(p.g.upper_off + node_e[e]) : 0u;
#pragma unroll
                    for (int e = 0; e < MAXE; ++e) adj_e[e] = p.g.upper + (static_cast<size_t>(off_e[e]) + lvl - 1) * p.g.M;
                }
                for (uint32_t j0 = 0; j0 < deg; j0 += 32) {
                    const uint32_t j = j0 + lane;
                    uint32_t v_e[MAXE];
#pragma unroll
                    for (int e = 0; e < MAXE; ++e)
                        v_e[e] = (node_e[e] != HNSW_NONE && j < deg) ? __ldg(adj_e[e] + j) : HNSW_NONE;
#pragma unroll
                    for (int e = 0; e < MAXE; ++e) {
                        if (node_e[e] == HNSW_NONE) continue;     // uniform
                        const uint32_t v = v_e[e];
                        bool fresh = false;
                        if (v != HNSW_NONE) {
                            const uint32_t slot = vis_hash(v) & p.vis_mask;
                            const uint16_t tag = vis_tag(v);
                            fresh = s_vis[slot] != tag;      // only warp 0 touches the cache
                            s_vis[slot] = tag;
                        }
                        __syncwarp();                         // entry e's tags are visible to entry e+1's checks
                        const unsigned m = __ballot_sync(0xFFFFFFFFu, fresh);
                        if (fresh) s_cid[nc + __popc(m & ((1u << lane) - 1))] = v;
                        nc += __popc(m);
                    }
                }
                if (lane == 0) {
                    s_nc = nc;
                    s_done = picked == 0;
                }
            }
            __syncthreads();
            if (s_done) break;
            const int nc = s_nc;
            // ---- all warps: distances, 8 candidates per warp per round; candidates that cannot enter
            // the pool (not closer than its current worst entry) are dropped right here
            const unsigned long long worst_now = s_worst;
            const bool merging = merge_mode && ef > 1u;
            for (int g0 = warp * 8; g0 < nc; g0 += WARPS * 8) {
                const int cnt = min(8, nc - g0);
                score8(s_cid + g0, cnt, s_ck + g0);
                __syncwarp();
                if (merging) {   // survivors are compacted as they are scored (order irrelevant: keys are unique per node)
                    const unsigned long long key = lane < cnt ? s_ck[g0 + lane] : ~0ull;
                    const bool live = lane < cnt && (key >> 1) < worst_now;
                    const unsigned m = __ballot_sync(0xFFFFFFFFu, live);
                    int base = 0;
                    if (lane == 0 && m) base = atomicAdd(&s_nlive, __popc(m));
                    base = __shfl_sync(0xFFFFFFFFu, base, 0);
                    if (live) s_surv[base + __popc(m & ((1u << lane) - 1u))] = key;
                } else if (lane < cnt && (s_ck[g0 + lane] >> 1) >= worst_now) {
                    s_ck[g0 + lane] = ~0ull;
                }
            }
            n_eval += (tid == 0) ? nc : 0;
            __syncthreads();
            if (merging) {
                // ---- whole CTA: merge the sorted pool with this step's survivors by rank counting -----------------------
                // (warp 0 used to insert them one by one, re-deriving the pool's worst entry after each: ~300 cycles per
                // survivor, ~200 survivors in the steps that fill the pool — tens of µs of a lone query's latency)
                const int size = s_size, L = s_nlive;
                auto lower_bound_pool = [&](unsigned long long k1) {   // first pool index with (key >> 1) >= k1
                    int lo = 0, hi = size;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if ((s_beam[mid] >> 1) < k1) lo = mid + 1; else hi = mid;
                    }
                    return lo;
                };
                // pass A: drop exact duplicates (the visited cache is lossy): of a pool entry, or of an earlier survivor.
                // Writing ~0 while others still compare is benign: whoever equals a dropped key is dropped by the same witness.
                for (int i = tid; i < L; i += THREADS) {
                    const unsigned long long k1 = s_surv[i] >> 1;
                    const int lb = lower_bound_pool(k1);
                    bool dup = lb < size && (s_beam[lb] >> 1) == k1;
                    for (int j = 0; j < i && !dup; ++j) dup = (s_surv[j] >> 1) == k1;
                    if (dup) s_surv[i] = ~0ull;
                }
                __syncthreads();
                // pass B: rank = entries of the other list below + own position; the first ef ranks form the new pool
                int dropped = 0;
                for (int e = tid; e < size + L; e += THREADS) {
                    const unsigned long long key = e < size ? s_beam[e] : s_surv[e - size];
                    if (key == ~0ull) { ++dropped; continue; }
                    const unsigned long long k1 = key >> 1;
                    int r = e < size ? e : lower_bound_pool(k1);
                    for (int j = 0; j < L; ++j) r += (s_surv[j] >> 1) < k1;   // (~0 >> 1 is never below a key)
                    if (r < static_cast<int>(ef)) s_beam2[r] = key;
                }
                if (dropped) atomicSub(&s_nlive, dropped);
                __syncthreads();
                {
                    unsigned long long* t = s_beam; s_beam = s_beam2; s_beam2 = t;
                }
                if (tid == 0) {
                    const int ns = min(static_cast<int>(ef), size + s_nlive);
                    s_size = ns;
                    s_worst = ns == static_cast<int>(ef) ? (s_beam[ns - 1] >> 1) : ~0ull;
                    s_worst_idx = ns - 1;
                }
                __syncthreads();
                continue;
            }
            // ---- warp 0: insert the survivors (append while the pool is filling, else replace the worst)
            if (warp == 0) {
                int size = s_size;
                // 32 candidates per ballot: the scoring phase already dropped what could not beat the pool's worst
                // entry, so only the set bits (a handful per step once the pool is full) are walked, in order
                for (int j0 = 0; j0 < nc; j0 += 32) {
                    const unsigned long long key_l = j0 + lane < nc ? s_ck[j0 + lane] : ~0ull;
                    unsigned live = __ballot_sync(0xFFFFFFFFu, key_l != ~0ull);
                    while (live) {
                        const int src = __ffs(live) - 1;
                        live &= live - 1;
                        const unsigned long long key = __shfl_sync(0xFFFFFFFFu, key_l, src);
                        if ((key >> 1) >= s_worst) continue;
                        bool dup = false;                   // the visited cache is lossy: exact de-duplication here
                        for (int i = lane; i < size; i += 32) dup |= (s_beam[i] >> 1) == (key >> 1);
                        if (__any_sync(0xFFFFFFFFu, dup)) continue;
                        if (size < static_cast<int>(ef)) {
                            if (lane == 0) s_beam[size] = key;
                            ++size;
                            __syncwarp();
                            if (size == static_cast<int>(ef)) pool_recompute_worst(size, ef);
                        } else {
                            if (lane == 0) s_beam[s_worst_idx] = key;
                            __syncwarp();
                            pool_recompute_worst(size, ef);
                        }
                    }
                }
                if (lane == 0) s_size = size;
            }
            __syncthreads();
        }
        // ---- descend: keep the best entry only (greedy levels), un-expand it, forget the cache
        if (lvl > bottom_level) {
            __syncthreads();
            if (tid == 0) {
                s_beam[0] &= ~1ull;          // upper levels run with a pool of one entry
                s_size = 1;
                const uint32_t next_ef = lvl == bottom_level + 1 ? p.ef : 1u;
                s_worst = next_ef == 1u ? (s_beam[0] >> 1) : ~0ull;
                s_worst_idx = 0;
            }
            for (uint32_t i = tid; i <= p.vis_mask; i += THREADS) s_vis[i] = 0u;
            __syncthreads();
            if (tid == 0) {
                const uint32_t node = static_cast<uint32_t>(s_beam[0] >> 1) & 0x7FFFFFFFu;
                s_vis[vis_hash(node) & p.vis_mask] = vis_tag(node);
            }
            __syncthreads();
        }
    }

    // ---- sort the pool once (ascending by distance, node) for the result extraction ------------
    {
        const int size = s_size;
        int len = 2;
        while (len < size) len <<= 1;
        for (int i = size + tid; i < len; i += THREADS) s_beam[i] = ~0ull;
        __syncthreads();
        for (int k2 = 2; k2 <= len; k2 <<= 1)
            for (int j = k2 >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < (len >> 1); t += THREADS) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int pp = i | j;
                    const bool asc = (i & k2) == 0;
                    const unsigned long long x = s_beam[i], y = s_beam[pp];
                    if ((x > y) == asc) { s_beam[i] = y; s_beam[pp] = x; }
                }
                __syncthreads();
            }
    }

    // ---- bf16 gathers: refine the head of the beam in fp32 before the top k are taken ------------------------------
    // The mirror's rounding (~1e-3 on a cosine) reorders near-ties among the best entries (1M clustered rows, beam 80:
    // recall 0.933 when the first k by bf16 distance are returned, 0.980 with fp32 gathers).  The first `refine` beam
    // entries are re-evaluated from the fp32 rows — warp-cooperative, coalesced, 64 rows against the ~1200 the
    // traversal gathered — and re-ordered; the f64 re-score below still only touches k rows.
    if (BF16) {
        const int R = min(s_size, static_cast<int>(p.refine));
        __syncthreads();
        for (int i = tid; i < R; i += THREADS) s_cid[i] = static_cast<uint32_t>(s_beam[i] >> 1) & 0x7FFFFFFFu;
        __syncthreads();
        for (int g0 = warp * 8; g0 < R; g0 += WARPS * 8)
            score8_from(std::false_type{}, s_cid + g0, min(8, R - g0), s_ck + g0);
        __syncthreads();
        // re-order the head: ranks are a permutation of [0, R) (keys are unique: they embed the node) and nothing reads
        // s_beam[0, R) any more (nodes sit in s_cid, keys in s_ck), so every key goes straight to its place
        for (int t = tid; t < R; t += THREADS) {
            const unsigned long long key = s_ck[t];
            int r = 0;
            for (int j = 0; j < R; ++j) r += s_ck[j] < key;
            s_beam[r] = key;
        }
        __syncthreads();
    }
    finish_query<METRIC, BUILD, THREADS>(p, qi, s_beam, s_size, reinterpret_cast<const float*>(q4), s_ex, s_rid, s_qs, n_eval);
}

template <int METRIC, bool BUILD, int WARPS>
static int launch_warps(const HnswParams& p, uint32_t nq, size_t smem, cudaStream_t s) {
    if (p.pitch == 384 && p.rows_bf16 && !BUILD) {
        auto k = hnsw_search_kernel<METRIC, 3, false, WARPS, true>;
        if (smem > 40 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        k<<<nq, WARPS * 32, smem, s>>>(p);
    } else if (p.pitch == 384) {
        auto k = hnsw_search_kernel<METRIC, 3, BUILD, WARPS, false>;
        if (smem > 40 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        k<<<nq, WARPS * 32, smem, s>>>(p);
    } else {
        auto k = hnsw_search_kernel<METRIC, 0, BUILD, WARPS, false>;
        if (smem > 40 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        k<<<nq, WARPS * 32, smem, s>>>(p);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : 6;
}

// CTA width.  Search (the whole-CTA rank merge keeps pool maintenance off warp 0, so extra warps only add gather
// parallelism): measured on 1M x 384, beam 80 (scripts/hnsw_probe.py, registers capped so that 32 warps stay resident
// per SM whatever the width) — 1 024 queries: 0.76 / 1.27 / 1.55 M q/s with 1 / 2 / 4 warps per query; 4 096: 1.65 /
// 2.02 / 1.97; 16 384: 2.35 / 2.64 / 2.37.  A query that finishes sooner beats a query that is merely resident.
// Few queries: latency matters, a whole 16-warp CTA per query scores the up-to-256 candidates of a step in one round
// of gathers (scripts/hnsw_latency.py: 168 / 263 / 310 us per call at nq = 1 / 16 / 128).
// Construction keeps the one-warp kernel it was tuned with (profiles/r01_hnsw_tune.jsonl).
static int hnsw_cta_warps(uint32_t beam, uint32_t nq, bool build = false) {
    if (const char* e = std::getenv("VL_HNSW_WARPS")) {
        const int w = atoi(e);
        if (w == 1 || w == 2 || w == 4 || w == 8 || w == 16) return w;
    }
    if (build) {
        static const bool build2 = std::getenv("VL_HNSW_BUILD_ONE_WARP") == nullptr;
        return nq <= 296 ? 16 : (nq < 1024 ? 4 : (build2 && beam <= 512 ? 2 : 1));
    }
    // (beams above 512 entries: the 64 threads of a 2-warp CTA spend longer on the rank merge than warp 0 did inserting —
    // beam 2048 at 4 096 queries: 49 K vs 59 K q/s — so those keep the one-warp kernel)
    return nq <= 296 ? 16 : (nq < 2400 ? 4 : (beam <= 512 ? 2 : 1));
}

template <int METRIC, bool BUILD>
static int launch_metric(const HnswParams& p, uint32_t nq, size_t smem, cudaStream_t s) {
    // (a register-resident sorted pool — ballot-ranked insertion, shfl_up shifts — was measured and is NOT
    // faster than this shared-memory pool at any beam width: profiles/r01_hnsw_tune.jsonl, "regpool" rows)
    switch (hnsw_cta_warps(p.ef, nq, BUILD)) {
        case 1: return launch_warps<METRIC, BUILD, 1>(p, nq, smem, s);
        case 2: return launch_warps<METRIC, BUILD, 2>(p, nq, smem, s);
        case 8: return launch_warps<METRIC, BUILD, 8>(p, nq, smem, s);
        case 16: return launch_warps<METRIC, BUILD, 16>(p, nq, smem, s);
        default: return launch_warps<METRIC, BUILD, 4>(p, nq, smem, s);
    }
}

// beam / visited-cache sizing shared by search and construction; returns the dynamic smem bytes
static size_t size_pool(HnswParams& p, uint32_t W, uint32_t M0, uint32_t max_deg, uint32_t k, uint32_t pitch,
                        bool wide = false, bool merge = false) {
    const bool query_in_smem = pitch != 384;   // launch_warps: pitch 384 runs the register-resident (NCH = 3) kernels
    p.ef = W;
    uint32_t bcap = 64;
    while (bcap < p.ef) bcap <<= 1;
    p.beam_cap = bcap;
    // visited tag cache (16-bit tags): ~2 slots per expected evaluation (~W·M0/2 fresh nodes), 2K..32K entries
    // The kernel is occupancy-bound by shared memory (one-warp CTAs, 32 per SM at <= 7 KB): a cache of ~ef·M0/2
    // tags (/4 for wide beams) loses a few tags — each costs one re-evaluation — but keeps more queries resident
    // (measured: profiles/r01_hnsw_tune.jsonl, vis_div rows)
    uint32_t vis_div = p.ef >= 256 ? 4 : 2;
    if (const char* e = std::getenv("VL_HNSW_VIS_DIV")) vis_div = static_cast<uint32_t>(std::max(1, atoi(e)));
    const uint32_t want = p.ef * M0 / vis_div;
    uint32_t cap = 2048;   // the quantised-score staging (k doubles) aliases the cache: >= 4·k tags of 2 bytes
    while ((cap < want || cap < 4u * std::max(k, p.rerank)) && cap < 32768) cap <<= 1;
    p.vis_mask = cap - 1;
    // candidates of one step: up to HN_MAX_EXPAND expanded entries x degree; the arrays also stage the results
    // throughput / construction: 1 / 2 / 4 entries per step; wide (latency) CTAs: ef/5 entries, at most 8
    uint32_t expand = p.ef >= 64 ? HN_MAX_EXPAND : (p.ef >= 32 ? 2 : 1);
    if (wide) expand = std::max(1u, std::min<uint32_t>(HN_MAX_EXPAND_WIDE, p.ef / 5));
    if (const char* e = std::getenv("VL_HNSW_EXPAND")) expand = static_cast<uint32_t>(std::max(1, std::min(atoi(e), wide ? HN_MAX_EXPAND_WIDE : HN_MAX_EXPAND)));
    p.expand = expand;
    uint32_t cc = expand * max_deg;
    if (cc < k) cc = k;
    if (cc < p.rerank) cc = p.rerank;    // the result staging (exact scores, nodes) aliases the candidate arrays
    if (cc < p.refine) cc = p.refine;    // ... and so does the fp32 refinement of the beam's head
    if (cc < 8) cc = 8;
    p.cand_cap = (cc + 7u) & ~7u;
    p.merge = merge ? 1u : 0u;   // second pool buffer + compacted survivors (whole-CTA rank merge)
    return (query_in_smem ? static_cast<size_t>(pitch) * 4 : 0) + static_cast<size_t>(bcap) * 8 +
           static_cast<size_t>(cap) * 2 + static_cast<size_t>(p.cand_cap) * 12 +
           (merge ? (static_cast<size_t>(bcap) + p.cand_cap) * 8 : 0);
}

// construction-time search (hnsw_build.cu): node order[i]'s row is query i; greedy descent from the entry
// point through the levels above `level` (or none when entry_only), beam of exactly `ef` on `level`; the
// sorted beam goes to out_keys[i][0..out_counts[i]).
int hnsw_launch_build_search(const HnswDeviceGraph& g, const float* d_rows, uint32_t pitch, uint32_t dim, int metric,
                             const uint32_t* d_order, uint32_t nq, int level, bool entry_only, uint32_t ef,
                             unsigned long long* d_out_keys, uint32_t out_stride, uint32_t* d_out_counts,
                             cudaStream_t stream) {
    if (ef > HN_EF_MAX || ef > out_stride || ef == 0) return 9;
    HnswParams p;
    p.g = g;
    p.rows = d_rows;
    p.queries = nullptr;
    p.pitch = pitch;
    p.dim = dim;
    p.k = 0;
    p.out_ids = nullptr; p.out_scores = nullptr; p.out_counts = d_out_counts; p.visited = nullptr;
    p.order = d_order; p.stop_level = level; p.entry_only = entry_only ? 1u : 0u;
    p.out_keys = d_out_keys; p.out_stride = out_stride;
    static const bool no_merge = std::getenv("VL_HNSW_NO_MERGE") != nullptr;
    const int cta_warps = hnsw_cta_warps(ef, nq, true);
    const size_t smem = size_pool(p, ef, level == 0 ? g.M0 : g.M, std::max(g.M, g.M0), 0, pitch, false,
                                  cta_warps >= 2 && !no_merge);
    switch (metric) {
        case COSINE: return launch_metric<COSINE, true>(p, nq, smem, stream);
        case EUCLIDEAN: return launch_metric<EUCLIDEAN, true>(p, nq, smem, stream);
        case MANHATTAN: return launch_metric<MANHATTAN, true>(p, nq, smem, stream);
        case DOT: return launch_metric<DOT, true>(p, nq, smem, stream);
        default: return 5;
    }
}

int hnsw_launch_search(const HnswDeviceGraph& g, const float* d_rows, uint32_t pitch, uint32_t dim, int metric,
                       const float* d_queries, uint32_t nq, uint32_t k, uint32_t ef, uint64_t* d_out_ids,
                       double* d_out_scores, uint32_t* d_out_counts, unsigned long long* d_visited,
                       cudaStream_t stream, uint32_t score_mode, uint32_t beam_mult, const void* rows_bf16,
                       uint32_t* visited_per_query) {
    if (k > HN_K_MAX) return 9;
    HnswParams p;
    p.g = g;
    p.rows = d_rows;
    p.rows_bf16 = rows_bf16;
    p.queries = d_queries;
    p.pitch = pitch;
    p.dim = dim;
    p.k = k;
    // `ef` is the reference's ef (hnsw.rs:437: min(k, len); > 0 = the additive sweep knob).  The device beam is
    // W = beam_mult x ef.  beam_mult = 1 is EQUAL ef; the reference's layer search (crate hnsw 0.11) pops a LIFO
    // stack with no distance-based early exit and evaluates far more nodes per unit of ef than a sorted beam
    // (measured visit counts sit next to the recall figures in bench.py / tests/golden), so larger factors trade
    // the visit-count gap for recall (vl_hnsw_set_beam_factor).
    uint64_t W = static_cast<uint64_t>(ef < 1 ? 1 : ef) * (beam_mult < 1 ? 1 : beam_mult);
    if (W < k) W = k;
    if (W > HN_EF_MAX) W = HN_EF_MAX;
    p.out_ids = d_out_ids;
    p.out_scores = d_out_scores;
    p.out_counts = d_out_counts;
    p.visited = d_visited;
    p.visited_per_query = visited_per_query;
    p.score_mode = score_mode;
    // bf16 gathers: the best 4k (at least 64) beam entries are refined in fp32 before k are taken (see the kernel)
    p.rerank = k;
    p.refine = 0;
    if (rows_bf16 && pitch == 384)
        p.refine = static_cast<uint32_t>(std::min<uint64_t>(std::max<uint64_t>(4ull * k, 64), std::min<uint64_t>(W, HN_REFINE_MAX)));
    const int cta_warps = hnsw_cta_warps(static_cast<uint32_t>(W), nq);
    static const bool no_merge = std::getenv("VL_HNSW_NO_MERGE") != nullptr;
    const size_t smem = size_pool(p, static_cast<uint32_t>(W), g.M0, std::max(g.M, g.M0), k, pitch, cta_warps >= 16,
                                  cta_warps >= 2 && !no_merge);
    switch (metric) {
        case COSINE: return launch_metric<COSINE, false>(p, nq, smem, stream);
        case EUCLIDEAN: return launch_metric<EUCLIDEAN, false>(p, nq, smem, stream);
        case MANHATTAN: return launch_metric<MANHATTAN, false>(p, nq, smem, stream);
        case DOT: return launch_metric<DOT, false>(p, nq, smem, stream);
        default: return 5;
    }
}

}  // namespace vl
