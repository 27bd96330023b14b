// vl_oracle_hnsw.cpp — CPU ORACLE (test infrastructure, never on the product path).
//
// Restates (a) the reference's HNSW wrapper, src/index/hnsw.rs:51-75,113-174,363-496, and
// (b) the THIRD-PARTY graph it drives: crates.io `hnsw` 0.11.0 + `space` 0.17.0 +
// `rand` 0.8.5 StdRng (Cargo.lock:1111-1124, 2216-2223, 1837-1845).  The crate sources are
// NOT under /root/reference; (b) follows the published algorithm as recorded in SURVEY.md
// Appendix C.  PARITY UNPINNED at graph level: the reference's own tests only pin toy cases
// (hnsw.rs:605-634, 679-749, 776-805) — those are checked in tests/test_oracle_golden.py.
// Used for: the recall@10 baseline at equal (M, M0, ef_construction, ef) and the CPU timing.
#include "vl_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

constexpr size_t NONE = ~static_cast<size_t>(0);  // `!0` empty-slot marker

// ---- src/index/hnsw.rs:113-174: distance functors, `as u64` = truncate toward zero,
// saturating, NaN → 0 (Rust float→int cast semantics).
inline uint64_t as_u64(double d) {
    if (std::isnan(d)) return 0;
    if (d <= 0.0) return 0;
    if (d >= 18446744073709551616.0) return ~0ULL;
    return static_cast<uint64_t>(d);
}

inline uint64_t dist_euclidean(const double* a, const double* b, size_t n) {  // hnsw.rs:116-122
    double sum_sq = 0.0;
    for (size_t i = 0; i < n; ++i) {
        const double d = a[i] - b[i];
        sum_sq += d * d;
    }
    return as_u64(std::sqrt(sum_sq) * 1000.0);
}
inline uint64_t dist_cosine(const double* a, const double* b, size_t n) {  // hnsw.rs:128-147
    double dot = 0.0, a_sq = 0.0, b_sq = 0.0;
    for (size_t i = 0; i < n; ++i) {
        const double x = a[i], y = b[i];
        dot = dot + x * y;
        a_sq = a_sq + x * x;
        b_sq = b_sq + y * y;
    }
    const double norm_a = std::sqrt(a_sq), norm_b = std::sqrt(b_sq);
    if (norm_a == 0.0 || norm_b == 0.0) return 1000;
    const double cosine_sim = dot / (norm_a * norm_b);
    return as_u64((1.0 - cosine_sim) * 1000.0);
}
inline uint64_t dist_manhattan(const double* a, const double* b, size_t n) {  // hnsw.rs:153-159
    double dist = 0.0;
    for (size_t i = 0; i < n; ++i) dist += std::fabs(a[i] - b[i]);
    return as_u64(dist * 1000.0);
}
inline uint64_t dist_dot(const double* a, const double* b, size_t n) {  // hnsw.rs:165-173
    double dot = 0.0;
    for (size_t i = 0; i < n; ++i) dot += a[i] * b[i];
    // f64::clamp: NaN stays NaN
    double c = dot;
    if (c < -1000.0) c = -1000.0;
    if (c > 1000.0) c = 1000.0;
    return as_u64(1000.0 - c);
}
inline uint64_t distance(int metric, const double* a, const double* b, size_t n) {
    switch (metric) {
        case VLO_EUCLIDEAN: return dist_euclidean(a, b, n);
        case VLO_COSINE: return dist_cosine(a, b, n);
        case VLO_MANHATTAN: return dist_manhattan(a, b, n);
        default: return dist_dot(a, b, n);
    }
}

// ---- Accelerated evaluation of the SAME functors (results identical to the functions above) ----
// The restated insert evaluates ~10^4-10^5 distances per row at ef_construction = 400, each a serial
// f64 chain (~0.6 us at 384-d); a 1M-row fixture would take days.  The functors only expose
// floor(v * 1000) of a real value v.  `fast_*` evaluates the sums with 8 independent accumulators
// (vectorisable); re-association moves an f64 sum of <= 4096 products by < 1e-12 relative to
// sum|terms|, i.e. the pre-floor value by < 1e-8 for |v|*1000 < 1e4.  If the fast value lies further
// than GUARD from every integer (and from the clamp / zero edges), the strict left-to-right
// evaluation has the same floor and the fast result is returned; otherwise the strict functor is
// evaluated.  Either way the returned u64 is exactly what `distance()` returns
// (tests/test_oracle_golden.py::test_hnsw_fast_functors_identical compares whole graphs).
constexpr double GUARD = 1e-6;

#define VLO_FAST_ATTR __attribute__((optimize("O3"), target_clones("avx2", "default")))

typedef double v4d __attribute__((vector_size(32)));
typedef float v4f __attribute__((vector_size(16)));
inline __attribute__((always_inline)) v4d load4(const double* p) { v4d v; std::memcpy(&v, p, 32); return v; }
inline __attribute__((always_inline)) v4d load4(const float* p) { v4f v; std::memcpy(&v, p, 16); return __builtin_convertvector(v, v4d); }
inline __attribute__((always_inline)) double hsum(v4d a, v4d b) { const v4d t = a + b; return (t[0] + t[1]) + (t[2] + t[3]); }

template <class TA, class TB>
inline __attribute__((always_inline)) void sums3(const TA* a, const TB* b, size_t n, double& dot, double& aa, double& bb) {
    v4d d0 = {0, 0, 0, 0}, d1 = d0, p0 = d0, p1 = d0, q0 = d0, q1 = d0;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        const v4d x0 = load4(a + i), y0 = load4(b + i), x1 = load4(a + i + 4), y1 = load4(b + i + 4);
        d0 += x0 * y0; p0 += x0 * x0; q0 += y0 * y0;
        d1 += x1 * y1; p1 += x1 * x1; q1 += y1 * y1;
    }
    dot = hsum(d0, d1); aa = hsum(p0, p1); bb = hsum(q0, q1);
    for (; i < n; ++i) {
        const double x = static_cast<double>(a[i]), y = static_cast<double>(b[i]);
        dot += x * y; aa += x * x; bb += y * y;
    }
}
// kind 0: sum (a-b)^2, 1: sum |a-b|, 2: sum a*b
template <int KIND, class TA, class TB>
inline __attribute__((always_inline)) double sum1(const TA* a, const TB* b, size_t n) {
    v4d d0 = {0, 0, 0, 0}, d1 = d0;
    size_t i = 0;
    auto term = [](v4d x, v4d y) -> v4d {
        if (KIND == 0) { const v4d t = x - y; return t * t; }
        if (KIND == 1) { const v4d t = x - y; return t < 0 ? -t : t; }
        return x * y;
    };
    for (; i + 8 <= n; i += 8) {
        d0 += term(load4(a + i), load4(b + i));
        d1 += term(load4(a + i + 4), load4(b + i + 4));
    }
    double r = hsum(d0, d1);
    for (; i < n; ++i) {
        const double x = static_cast<double>(a[i]), y = static_cast<double>(b[i]);
        r += KIND == 0 ? (x - y) * (x - y) : KIND == 1 ? std::fabs(x - y) : x * y;
    }
    return r;
}

// pre-floor value of the functor from fast sums; returns false when the strict functor must decide
template <class TA, class TB>
inline __attribute__((always_inline)) bool fast_value(int metric, const TA* a, const TB* b, size_t n, double& v) {
    switch (metric) {
        case VLO_EUCLIDEAN: v = std::sqrt(sum1<0>(a, b, n)) * 1000.0; break;
        case VLO_MANHATTAN: v = sum1<1>(a, b, n) * 1000.0; break;
        case VLO_COSINE: {
            double dot, aa, bb;
            sums3(a, b, n, dot, aa, bb);
            if (!(aa > 1e-300) || !(bb > 1e-300)) return false;
            v = (1.0 - dot / (std::sqrt(aa) * std::sqrt(bb))) * 1000.0;
            break;
        }
        default: {
            const double dot = sum1<2>(a, b, n);
            if (!(std::fabs(dot) < 999.0)) return false;  // near / beyond the clamp
            v = 1000.0 - dot;
            break;
        }
    }
    if (!(v > GUARD) || !(v < 1e9)) return false;      // NaN, <= 0 edge, huge
    const double f = v - std::floor(v);
    return f > GUARD && f < 1.0 - GUARD;
}
VLO_FAST_ATTR bool fast_value_dd(int metric, const double* a, const double* b, size_t n, double& v) { return fast_value(metric, a, b, n, v); }
VLO_FAST_ATTR bool fast_value_df(int metric, const double* a, const float* b, size_t n, double& v) { return fast_value(metric, a, b, n, v); }
VLO_FAST_ATTR bool fast_value_ff(int metric, const float* a, const float* b, size_t n, double& v) { return fast_value(metric, a, b, n, v); }

// ---- rand 0.8.5 StdRng == ChaCha12Rng (rand_chacha 0.3), from_seed([0u8;32]):
// 64-bit block counter in words 12-13, stream 0; BlockRng yields the u32 words of
// consecutive blocks in order; next_u64 = lo word first, then hi word.
struct ChaCha12 {
    uint32_t key[8];
    uint64_t counter = 0;
    uint32_t buf[16];
    int idx = 16;
    ChaCha12() { std::memset(key, 0, sizeof key); }
    static inline uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
    static inline void qr(uint32_t* s, int a, int b, int c, int d) {
        s[a] += s[b]; s[d] ^= s[a]; s[d] = rotl(s[d], 16);
        s[c] += s[d]; s[b] ^= s[c]; s[b] = rotl(s[b], 12);
        s[a] += s[b]; s[d] ^= s[a]; s[d] = rotl(s[d], 8);
        s[c] += s[d]; s[b] ^= s[c]; s[b] = rotl(s[b], 7);
    }
    void refill() {
        uint32_t in[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574};
        for (int i = 0; i < 8; ++i) in[4 + i] = key[i];
        in[12] = static_cast<uint32_t>(counter);
        in[13] = static_cast<uint32_t>(counter >> 32);
        in[14] = 0;
        in[15] = 0;
        uint32_t s[16];
        std::memcpy(s, in, sizeof s);
        for (int r = 0; r < 6; ++r) {  // 12 rounds = 6 double rounds
            qr(s, 0, 4, 8, 12); qr(s, 1, 5, 9, 13); qr(s, 2, 6, 10, 14); qr(s, 3, 7, 11, 15);
            qr(s, 0, 5, 10, 15); qr(s, 1, 6, 11, 12); qr(s, 2, 7, 8, 13); qr(s, 3, 4, 9, 14);
        }
        for (int i = 0; i < 16; ++i) buf[i] = s[i] + in[i];
        ++counter;
        idx = 0;
    }
    uint32_t next_u32() {
        if (idx >= 16) refill();
        return buf[idx++];
    }
    uint64_t next_u64() {
        const uint64_t lo = next_u32();
        const uint64_t hi = next_u32();
        return (hi << 32) | lo;
    }
};

struct Neighbor {  // space::Neighbor<u64>
    size_t index;
    uint64_t distance;
};

struct Searcher {  // hnsw::Searcher
    std::vector<Neighbor> candidates;  // LIFO stack
    std::vector<Neighbor> nearest;     // ascending by distance
    std::vector<uint32_t> seen_stamp;  // HashSet<usize> semantics via epoch stamps
    uint32_t epoch = 0;
    uint64_t evals = 0;
    void clear(size_t n) {
        candidates.clear();
        nearest.clear();
        if (seen_stamp.size() < n) seen_stamp.resize(n + n / 2 + 1024, 0);
        if (++epoch == 0) {
            std::fill(seen_stamp.begin(), seen_stamp.end(), 0);
            epoch = 1;
        }
    }
    bool seen_insert(size_t i) {
        if (seen_stamp[i] == epoch) return false;
        seen_stamp[i] = epoch;
        return true;
    }
};

struct Graph {  // hnsw::Hnsw<Met, Vec<f64>, StdRng, M, M0>
    size_t dim, M, M0, efc;
    int metric;
    std::vector<double> features;               // [n][dim]
    std::vector<float> features32;              // same values when every inserted row is f32-representable
    bool all32 = true;                          // (then the fast functors read half the bytes)
    bool fast = true;                           // VLO_HNSW_STRICT=1 disables the accelerated evaluation
    mutable uint64_t strict_evals = 0;          // distances the guard band sent to the strict functor
    std::vector<size_t> zero;                   // [n][M0]
    struct Layer {
        std::vector<size_t> zero_node, next_node;
        std::vector<size_t> neighbors;          // [len][M]
        size_t len() const { return zero_node.size(); }
    };
    std::vector<Layer> layers;                  // layers[0] is level 1
    ChaCha12 prng;

    size_t len() const { return zero.size() / M0; }
    const double* feat(size_t i) const { return features.data() + i * dim; }
    void push_feature(const double* q) {
        features.insert(features.end(), q, q + dim);
        if (all32) {
            for (size_t i = 0; i < dim && all32; ++i) all32 = static_cast<double>(static_cast<float>(q[i])) == q[i];
            if (all32) for (size_t i = 0; i < dim; ++i) features32.push_back(static_cast<float>(q[i]));
            else std::vector<float>().swap(features32);
        }
    }
    // functor(query, stored node) and functor(stored, stored): == distance(metric, ., ., dim)
    uint64_t dist_q(const double* q, size_t node) const {
        if (fast) {
            double v;
            const bool ok = all32 ? fast_value_df(metric, q, features32.data() + node * dim, dim, v)
                                  : fast_value_dd(metric, q, feat(node), dim, v);
            if (ok) return static_cast<uint64_t>(v);
            ++strict_evals;
        }
        return distance(metric, q, feat(node), dim);
    }
    uint64_t dist_nodes(size_t a, size_t b) const {
        if (fast) {
            double v;
            const bool ok = all32 ? fast_value_ff(metric, features32.data() + a * dim, features32.data() + b * dim, dim, v)
                                  : fast_value_dd(metric, feat(a), feat(b), dim, v);
            if (ok) return static_cast<uint64_t>(v);
            ++strict_evals;
        }
        return distance(metric, feat(a), feat(b), dim);
    }

    size_t random_level() {
        const double uniform =
            static_cast<double>(prng.next_u64()) / 18446744073709551616.0;  // u64::MAX as f64 == 2^64
        const double v = -std::log(uniform) * (1.0 / std::log(static_cast<double>(M)));
        if (std::isnan(v) || v <= 0.0) return 0;
        if (v >= 1e18) return static_cast<size_t>(1) << 40;
        return static_cast<size_t>(v);
    }

    void initialize_searcher(const double* q, Searcher& s) const {
        s.clear(len());
        const size_t entry = layers.empty() ? 0 : layers.back().zero_node[0];
        const Neighbor c{0, dist_q(q, entry)};
        ++s.evals;
        s.candidates.push_back(c);
        s.nearest.push_back(c);
        s.seen_insert(entry);
    }

    // search_single_layer (layer != nullptr) / search_zero_layer (layer == nullptr)
    void search_layer_impl(const double* q, Searcher& s, const Layer* layer, size_t cap) const {
        while (!s.candidates.empty()) {
            const size_t index = s.candidates.back().index;
            s.candidates.pop_back();
            const size_t deg = layer ? M : M0;
            const size_t* nb = layer ? layer->neighbors.data() + index * M : zero.data() + index * M0;
            for (size_t j = 0; j < deg && nb[j] != NONE; ++j) {  // take_while(|n| n != !0)
                const size_t neighbor = nb[j];
                const size_t node_to_visit = layer ? layer->zero_node[neighbor] : neighbor;
                if (!s.seen_insert(node_to_visit)) continue;  // one seen-set shared by all layers
                const uint64_t d = dist_q(q, node_to_visit);
                ++s.evals;
                // partition_point(|n| n.distance <= d): ties go AFTER equals
                const size_t pos =
                    std::upper_bound(s.nearest.begin(), s.nearest.end(), d,
                                     [](uint64_t dd, const Neighbor& n) { return dd < n.distance; }) -
                    s.nearest.begin();
                if (pos != cap) {
                    if (s.nearest.size() == cap) s.nearest.pop_back();
                    const Neighbor c{neighbor, d};
                    s.nearest.insert(s.nearest.begin() + pos, c);
                    s.candidates.push_back(c);
                }
            }
        }
    }

    void lower_search(const Layer& layer, Searcher& s) const {
        s.candidates.clear();
        const Neighbor best = s.nearest.front();
        s.nearest.clear();
        const Neighbor c{layer.next_node[best.index], best.distance};
        s.nearest.push_back(c);
        s.candidates.push_back(c);
    }

    void add_neighbor(const double* q, size_t node_ix, size_t target_ix, size_t layer) {
        const size_t deg = layer == 0 ? M0 : M;
        size_t* tn = layer == 0 ? zero.data() + target_ix * M0
                                : layers[layer - 1].neighbors.data() + target_ix * M;
        const size_t tz = layer == 0 ? target_ix : layers[layer - 1].zero_node[target_ix];
        size_t empty_point = 0;  // partition_point(|&n| n != !0): slots fill left → right
        while (empty_point < deg && tn[empty_point] != NONE) ++empty_point;
        if (empty_point != deg) {
            tn[empty_point] = node_ix;
            return;
        }
        size_t worst_ix = 0;
        uint64_t worst_d = 0;
        bool have = false;
        for (size_t ix = 0; ix < deg; ++ix) {  // min_by_key(Reverse(d)) → FIRST of the equally-worst
            const size_t n = tn[ix];
            const size_t nz = layer == 0 ? n : layers[layer - 1].zero_node[n];
            const uint64_t d = dist_nodes(tz, nz);
            if (!have || d > worst_d) {
                have = true;
                worst_d = d;
                worst_ix = ix;
            }
        }
        if (dist_q(q, tz) < worst_d) tn[worst_ix] = node_ix;  // strict <
    }

    void create_node(const double* q, const std::vector<Neighbor>& nearest, size_t layer) {
        const size_t deg = layer == 0 ? M0 : M;
        std::vector<size_t> nb(deg, NONE);
        for (size_t i = 0; i < deg && i < nearest.size(); ++i) nb[i] = nearest[i].index;  // verbatim
        if (layer == 0) {
            const size_t new_index = len();
            for (size_t i = 0; i < deg && nb[i] != NONE; ++i) add_neighbor(q, new_index, nb[i], 0);
            zero.insert(zero.end(), nb.begin(), nb.end());
        } else {
            Layer& L = layers[layer - 1];
            const size_t new_index = L.len();
            const size_t zn = len();
            const size_t next = layer == 1 ? len() : layers[layer - 2].len();
            for (size_t i = 0; i < deg && nb[i] != NONE; ++i)
                add_neighbor(q, new_index, nb[i], layer);
            L.zero_node.push_back(zn);
            L.next_node.push_back(next);
            L.neighbors.insert(L.neighbors.end(), nb.begin(), nb.end());
        }
    }

    size_t insert(const double* q, Searcher& s) {
        const size_t level = random_level();
        size_t cap = level >= layers.size() ? efc : 1;
        if (len() == 0) {
            zero.insert(zero.end(), M0, NONE);
            push_feature(q);
            while (layers.size() < level) {
                Layer L;
                L.zero_node.push_back(0);
                L.next_node.push_back(0);
                L.neighbors.assign(M, NONE);
                layers.push_back(std::move(L));
            }
            return 0;
        }
        // the new row must already be addressable by feat() during add_neighbor? No: the crate
        // pushes the feature AFTER linking; `q` is passed explicitly wherever it is needed.
        initialize_searcher(q, s);
        for (size_t ix = layers.size(); ix-- > level;) {
            search_layer_impl(q, s, &layers[ix], cap);
            lower_search(layers[ix], s);
            cap = ix == level ? efc : 1;
        }
        for (size_t ix = std::min(level, layers.size()); ix-- > 0;) {
            search_layer_impl(q, s, &layers[ix], cap);
            create_node(q, s.nearest, ix + 1);
            lower_search(layers[ix], s);
            cap = efc;
        }
        search_layer_impl(q, s, nullptr, cap);
        create_node(q, s.nearest, 0);
        push_feature(q);
        const size_t zero_node = len() - 1;
        while (layers.size() < level) {
            Layer L;
            L.zero_node.push_back(zero_node);
            L.next_node.push_back(layers.empty() ? zero_node : layers.back().len() - 1);
            L.neighbors.assign(M, NONE);
            layers.push_back(std::move(L));
        }
        return zero_node;
    }

    // nearest(q, ef, searcher, dest) == search_layer(q, ef, 0, ..)
    void nearest(const double* q, size_t ef, Searcher& s, std::vector<Neighbor>& dest) const {
        const size_t want = dest.size();
        dest.clear();
        if (len() == 0) return;
        initialize_searcher(q, s);
        for (size_t ix = layers.size(); ix-- > 0;) {
            search_layer_impl(q, s, &layers[ix], 1);
            lower_search(layers[ix], s);
        }
        search_layer_impl(q, s, nullptr, ef);
        const size_t found = std::min(want, s.nearest.size());
        dest.assign(s.nearest.begin(), s.nearest.begin() + found);
    }
};

double convert_distance_to_similarity(double d, int metric) {  // hnsw.rs:51-75
    switch (metric) {
        case VLO_EUCLIDEAN: return 1.0 / (1.0 + d);
        case VLO_COSINE: {
            const double cos_distance = d / 1000.0;
            return 1.0 - cos_distance;
        }
        case VLO_MANHATTAN: return 1.0 / (1.0 + d);
        default: {
            double v = (1000.0 - d) / 1000.0;
            if (v < 0.0) v = 0.0;
            if (v > 1.0) v = 1.0;
            return v;
        }
    }
}

}  // namespace

struct vlo_hnsw {  // HNSWIndex, hnsw.rs:197-213
    Graph g;
    Searcher build_searcher;
    std::unordered_map<uint64_t, size_t> id_to_index;
    std::unordered_map<size_t, uint64_t> index_to_id;
    size_t live = 0;  // metadata.len()
};

static int hnsw_search_impl(const vlo_hnsw* h, Searcher& s, const double* q, size_t qdim, size_t k,
                            int metric, size_t ef, uint64_t* out_ids, double* out_scores,
                            size_t* out_count) {
    if (out_count) *out_count = 0;
    if (qdim != h->g.dim) return VLO_ERR_DIM;                 // hnsw.rs:416-421 (always)
    if (metric != h->g.metric) return VLO_ERR_METRIC_MISMATCH;  // hnsw.rs:425-430
    if (h->live == 0) return VLO_OK;                          // hnsw.rs:432-434
    const size_t max_candidates = std::min(k, h->live);      // hnsw.rs:437
    if (max_candidates == 0) return VLO_OK;
    const size_t ef_search = ef == 0 ? max_candidates : std::max(ef, max_candidates);
    std::vector<Neighbor> neighbors(ef_search, Neighbor{NONE, ~0ULL});
    h->g.nearest(q, ef_search, s, neighbors);                 // hnsw.rs:454-466
    struct R {
        uint64_t id;
        double score;
    };
    std::vector<R> res;
    for (const Neighbor& n : neighbors) {                     // hnsw.rs:472-490
        if (n.index == NONE) continue;
        auto it = h->index_to_id.find(n.index);
        if (it == h->index_to_id.end()) continue;             // soft-deleted
        const double d = static_cast<double>(n.distance) / 1000.0;
        res.push_back(R{it->second, convert_distance_to_similarity(d, metric)});
    }
    if (res.size() >= 2)
        for (const R& r : res)
            if (std::isnan(r.score)) return VLO_ERR_NAN;
    std::stable_sort(res.begin(), res.end(), [](const R& a, const R& b) { return a.score > b.score; });
    const size_t m = std::min(k, res.size());
    for (size_t i = 0; i < m; ++i) {
        out_ids[i] = res[i].id;
        out_scores[i] = res[i].score;
    }
    if (out_count) *out_count = m;
    return VLO_OK;
}

extern "C" {

uint64_t vlo_hnsw_distance(int metric, const double* a, const double* b, size_t n) {
    return distance(metric, a, b, n);
}
double vlo_convert_distance_to_similarity(double d, int metric) {
    return convert_distance_to_similarity(d, metric);
}

vlo_hnsw* vlo_hnsw_create(size_t dim, int metric, size_t M, size_t M0, size_t efc) {
    if (dim == 0 || metric < 0 || metric > 3 || M < 2 || M0 < 1) return nullptr;  // hnsw.rs:217-219
    vlo_hnsw* h = new vlo_hnsw();
    h->g.dim = dim;
    h->g.M = M;
    h->g.M0 = M0;
    h->g.efc = efc ? efc : 400;  // crate default Params::ef_construction
    h->g.metric = metric;
    const char* strict = std::getenv("VLO_HNSW_STRICT");
    h->g.fast = !(strict && strict[0] == '1');
    return h;
}
void vlo_hnsw_destroy(vlo_hnsw* h) { delete h; }

int vlo_hnsw_add(vlo_hnsw* h, uint64_t id, const double* v, size_t len) {
    if (len != h->g.dim) return VLO_ERR_DIM;                       // hnsw.rs:364-366
    if (h->id_to_index.count(id)) return VLO_ERR_DUP_ID;           // hnsw.rs:368-370
    const size_t ix = h->g.insert(v, h->build_searcher);           // hnsw.rs:372-385
    h->id_to_index[id] = ix;
    h->index_to_id[ix] = id;
    ++h->live;
    return VLO_OK;
}

int vlo_hnsw_add_batch_f32(vlo_hnsw* h, const uint64_t* ids, const float* rows, size_t n) {
    std::vector<double> v(h->g.dim);
    for (size_t r = 0; r < n; ++r) {
        for (size_t i = 0; i < h->g.dim; ++i) v[i] = static_cast<double>(rows[r * h->g.dim + i]);
        const int st = vlo_hnsw_add(h, ids ? ids[r] : static_cast<uint64_t>(r), v.data(), h->g.dim);
        if (st != VLO_OK) return st;
    }
    return VLO_OK;
}

int vlo_hnsw_delete(vlo_hnsw* h, uint64_t id) {  // hnsw.rs:400-414 — soft delete
    auto it = h->id_to_index.find(id);
    if (it == h->id_to_index.end()) return VLO_ERR_NOT_FOUND;
    h->index_to_id.erase(it->second);
    h->id_to_index.erase(it);
    --h->live;
    return VLO_OK;
}

size_t vlo_hnsw_len(const vlo_hnsw* h) { return h->live; }

int vlo_hnsw_search(const vlo_hnsw* h, const double* q, size_t qdim, size_t k, int metric,
                    size_t ef, uint64_t* out_ids, double* out_scores, size_t* out_count,
                    uint64_t* out_visited) {
    Searcher s;  // fresh Searcher per call, hnsw.rs:453
    const int st = hnsw_search_impl(h, s, q, qdim, k, metric, ef, out_ids, out_scores, out_count);
    if (out_visited) *out_visited = s.evals;
    return st;
}

int vlo_hnsw_search_batch_f32(const vlo_hnsw* h, const float* queries, size_t nq, size_t k,
                              size_t ef, int nthreads, uint64_t* out_ids, double* out_scores,
                              uint32_t* out_counts, uint64_t* out_visited_total) {
    if (nthreads < 1) nthreads = 1;
    const size_t dim = h->g.dim;
    std::vector<uint64_t> visited(nthreads, 0);
    std::vector<int> status(nthreads, VLO_OK);
    auto work = [&](int t) {
        Searcher s;
        std::vector<double> q(dim);
        std::vector<uint64_t> oi(k ? k : 1);
        std::vector<double> os(k ? k : 1);
        for (size_t qi = t; qi < nq; qi += nthreads) {
            for (size_t i = 0; i < dim; ++i) q[i] = static_cast<double>(queries[qi * dim + i]);
            size_t cnt = 0;
            const int st = hnsw_search_impl(h, s, q.data(), dim, k, h->g.metric, ef, oi.data(),
                                            os.data(), &cnt);
            if (st != VLO_OK) status[t] = st;
            for (size_t i = 0; i < k; ++i) {
                out_ids[qi * k + i] = i < cnt ? oi[i] : ~0ULL;
                out_scores[qi * k + i] = i < cnt ? os[i] : 0.0;
            }
            if (out_counts) out_counts[qi] = static_cast<uint32_t>(cnt);
        }
        visited[t] = s.evals;
    };
    if (nthreads == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    uint64_t tot = 0;
    for (uint64_t v : visited) tot += v;
    if (out_visited_total) *out_visited_total = tot;
    for (int s : status)
        if (s != VLO_OK) return s;
    return VLO_OK;
}

size_t vlo_hnsw_num_layers(const vlo_hnsw* h) { return h->g.layers.size(); }
size_t vlo_hnsw_layer_len(const vlo_hnsw* h, size_t l) {
    if (l == 0) return h->g.len();
    return l - 1 < h->g.layers.size() ? h->g.layers[l - 1].len() : 0;
}

/* layer-0 adjacency [len][M0] (empty slot = UINT64_MAX) and the number of guard-band fallbacks */
void vlo_hnsw_export_zero(const vlo_hnsw* h, uint64_t* out) {
    for (size_t i = 0; i < h->g.zero.size(); ++i) out[i] = static_cast<uint64_t>(h->g.zero[i]);
}
/* upper layer l (1-based): zero_node[len], next_node[len], neighbors[len][M] */
void vlo_hnsw_export_layer(const vlo_hnsw* h, size_t l, uint64_t* zero_node, uint64_t* next_node, uint64_t* nb) {
    const auto& L = h->g.layers[l - 1];
    for (size_t i = 0; i < L.len(); ++i) { zero_node[i] = L.zero_node[i]; next_node[i] = L.next_node[i]; }
    for (size_t i = 0; i < L.neighbors.size(); ++i) nb[i] = static_cast<uint64_t>(L.neighbors[i]);
}
uint64_t vlo_hnsw_strict_evals(const vlo_hnsw* h) { return h->g.strict_evals; }

void vlo_hnsw_levels(size_t M, size_t n, uint32_t* out_levels) {
    Graph g;
    g.M = M;
    for (size_t i = 0; i < n; ++i) out_levels[i] = static_cast<uint32_t>(g.random_level());
}

}  // extern "C"
