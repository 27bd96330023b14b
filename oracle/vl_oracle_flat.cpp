// vl_oracle_flat.cpp — CPU ORACLE (test infrastructure, never on the product path).
// Restates the reference's Flat search and the four similarity metrics; see vl_oracle.h
// for the file:line map.  Build: g++ -O2 -ffp-contract=off (NO -ffast-math, NO -march FMA
// contraction): Rust never fuses a*b+c, and the accumulation order is part of the contract.
#include "vl_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

// src/lib.rs:425-444 — one pass, three accumulators, zero-norm → 0.0,
// dot / (sqrt(na) * sqrt(nb)) (product of two square roots, not sqrt of product).
inline double cosine_similarity(const double* a, const double* b, size_t n) {
    double dot = 0.0, na = 0.0, nb = 0.0;
    for (size_t i = 0; i < n; ++i) {
        const double x = a[i], y = b[i];
        dot += x * y;
        na += x * x;
        nb += y * y;
    }
    const double norm_a = std::sqrt(na);
    const double norm_b = std::sqrt(nb);
    if (norm_a == 0.0 || norm_b == 0.0) return 0.0;
    return dot / (norm_a * norm_b);
}

// src/lib.rs:476-489 — Iterator::sum::<f64>() folds from 0.0 left to right; powi(2) == x*x.
inline double euclidean_similarity(const double* a, const double* b, size_t n) {
    double sum_sq = 0.0;
    for (size_t i = 0; i < n; ++i) {
        const double d = a[i] - b[i];
        sum_sq += d * d;
    }
    return 1.0 / (1.0 + std::sqrt(sum_sq));
}

// src/lib.rs:521-532
inline double manhattan_similarity(const double* a, const double* b, size_t n) {
    double dist = 0.0;
    for (size_t i = 0; i < n; ++i) dist += std::fabs(a[i] - b[i]);
    return 1.0 / (1.0 + dist);
}

// src/lib.rs:565-572
inline double dot_product(const double* a, const double* b, size_t n) {
    double s = 0.0;
    for (size_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

inline double calculate(int metric, const double* a, const double* b, size_t n) {
    switch (metric) {  // src/lib.rs:384-389
        case VLO_COSINE: return cosine_similarity(a, b, n);
        case VLO_EUCLIDEAN: return euclidean_similarity(a, b, n);
        case VLO_MANHATTAN: return manhattan_similarity(a, b, n);
        default: return dot_product(a, b, n);
    }
}

struct Scored {
    uint64_t id;
    double score;
};

// flat.rs:116-117: sort_by(|a,b| b.score.partial_cmp(&a.score).unwrap()) — stable, descending;
// equal scores (incl. +0.0 vs -0.0) keep storage order; any NaN comparison panics.
int sort_truncate(std::vector<Scored>& v, size_t k, uint64_t* out_ids, double* out_scores,
                  size_t* out_count) {
    if (v.size() >= 2) {  // a merge sort compares every element at least once when n >= 2
        for (const Scored& s : v)
            if (std::isnan(s.score)) return VLO_ERR_NAN;
    }
    std::stable_sort(v.begin(), v.end(),
                     [](const Scored& a, const Scored& b) { return a.score > b.score; });
    const size_t m = std::min(k, v.size());
    for (size_t i = 0; i < m; ++i) {
        out_ids[i] = v[i].id;
        out_scores[i] = v[i].score;
    }
    if (out_count) *out_count = m;
    return VLO_OK;
}

template <typename T>
int flat_search_impl(const T* rows, const uint64_t* ids, size_t n, size_t dim, const T* q,
                     size_t qdim, size_t k, int metric, size_t clone_bytes, uint64_t* out_ids,
                     double* out_scores, size_t* out_count) {
    if (metric < 0 || metric > 3) return VLO_ERR_INVALID;
    if (n != 0 && qdim != dim) return VLO_ERR_DIM;  // flat.rs:99-104 (no check when empty)
    std::vector<double> qd(q, q + qdim), rd(dim);
    std::vector<Scored> sims;
    sims.reserve(n);
    std::vector<char*> clones;
    if (clone_bytes) clones.reserve(n);
    for (size_t r = 0; r < n; ++r) {
        const T* row = rows + r * dim;
        const double* a;
        if constexpr (sizeof(T) == sizeof(double)) {
            a = reinterpret_cast<const double*>(row);
        } else {
            for (size_t i = 0; i < dim; ++i) rd[i] = static_cast<double>(row[i]);
            a = rd.data();
        }
        sims.push_back(Scored{ids ? ids[r] : static_cast<uint64_t>(r),
                              calculate(metric, a, qd.data(), dim)});  // flat.rs:110
        if (clone_bytes) {  // flat.rs:111-112: text.clone() + metadata.clone() for EVERY row
            char* c = static_cast<char*>(std::malloc(clone_bytes));
            std::memset(c, static_cast<int>(r), clone_bytes);
            clones.push_back(c);
        }
    }
    const int st = sort_truncate(sims, k, out_ids, out_scores, out_count);
    for (char* c : clones) std::free(c);
    return st;
}

inline uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
inline int64_t ih4(uint64_t seed, uint64_t row, uint64_t col) {
    const uint64_t h = mix64(mix64(seed ^ ((row + 1) * 0x9E3779B97F4A7C15ULL)) ^
                             ((col + 1) * 0xD1B54A32D192ED03ULL));
    return static_cast<int64_t>((h & 0xFFFF) + ((h >> 16) & 0xFFFF) + ((h >> 32) & 0xFFFF) +
                                (h >> 48)) - 131070;
}

}  // namespace

extern "C" {

double vlo_metric(int metric, const double* a, const double* b, size_t n) {
    return calculate(metric, a, b, n);
}

int vlo_flat_search(const double* rows, const uint64_t* ids, size_t n, size_t dim,
                    const double* q, size_t qdim, size_t k, int metric, uint64_t* out_ids,
                    double* out_scores, size_t* out_count) {
    return flat_search_impl<double>(rows, ids, n, dim, q, qdim, k, metric, 0, out_ids, out_scores,
                                    out_count);
}

int vlo_flat_search_f32(const float* rows, const uint64_t* ids, size_t n, size_t dim,
                        const float* q, size_t qdim, size_t k, int metric, uint64_t* out_ids,
                        double* out_scores, size_t* out_count) {
    return flat_search_impl<float>(rows, ids, n, dim, q, qdim, k, metric, 0, out_ids, out_scores,
                                   out_count);
}

int vlo_flat_search_batch_f32(const float* rows, const uint64_t* ids, size_t n, size_t dim,
                              const float* queries, size_t nq, size_t k, int metric, int nthreads,
                              size_t clone_bytes, uint64_t* out_ids, double* out_scores) {
    if (nthreads < 1) nthreads = 1;
    const size_t kk = std::min(k, n);
    std::vector<int> status(nthreads, VLO_OK);
    auto work = [&](int t) {
        std::vector<uint64_t> oi(kk ? kk : 1);
        std::vector<double> os(kk ? kk : 1);
        for (size_t qi = t; qi < nq; qi += nthreads) {
            size_t cnt = 0;
            const int st = flat_search_impl<float>(rows, ids, n, dim, queries + qi * dim, dim, k,
                                                   metric, clone_bytes, oi.data(), os.data(), &cnt);
            if (st != VLO_OK) status[t] = st;
            for (size_t i = 0; i < k; ++i) {
                out_ids[qi * k + i] = i < cnt ? oi[i] : ~0ULL;
                out_scores[qi * k + i] = i < cnt ? os[i] : 0.0;
            }
        }
    };
    if (nthreads == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    for (int s : status)
        if (s != VLO_OK) return s;
    return VLO_OK;
}

void vlo_synth_rows_f32(uint64_t seed, uint64_t row0, size_t n, size_t dim, uint32_t clusters,
                        float* out) {
    std::vector<int64_t> v(dim);
    for (size_t r = 0; r < n; ++r) {
        const uint64_t row = row0 + r;
        uint64_t ss = 0;
        const uint64_t centre =
            clusters ? mix64(seed ^ 0xC2B2AE3D27D4EB4FULL ^ (row * 0x9E3779B97F4A7C15ULL)) % clusters
                     : 0;
        for (size_t c = 0; c < dim; ++c) {
            int64_t x = ih4(seed, row, c);
            if (clusters) x += 4 * ih4(0x5851F42D4C957F2DULL, centre, c);  // centres do not depend on seed
            v[c] = x;
            ss += static_cast<uint64_t>(x * x);
        }
        const double norm = std::sqrt(static_cast<double>(ss));
        for (size_t c = 0; c < dim; ++c)
            out[r * dim + c] = ss ? static_cast<float>(static_cast<double>(v[c]) / norm) : 0.0f;
    }
}

}  // extern "C"
