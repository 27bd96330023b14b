// native_callers.cpp — bench-only helper: N host threads calling vl_index_search with ONE query each on the same
// handle, the way the reference's server calls VectorIndex::search from its tokio workers under a read lock
// (src/client.rs:398, src/server.rs:258-275).  bench.py's Python callers measure the interpreter lock as much as
// the library; this drives the same C-ABI entry point (passed in as a function pointer, host buffers in, host
// results out) from plain threads.  Not part of the product library.
//   g++ -O2 -shared -fPIC -pthread -o scripts/libvl_native_callers.so scripts/native_callers.cpp
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

extern "C" {

typedef int (*vl_search_fn)(void* h, const float* queries, uint32_t nq, uint32_t qdim, uint32_t k, int metric,
                            uint32_t ef, uint64_t* out_ids, double* out_scores, uint32_t* out_counts);

typedef int (*vl_group_search_fn)(void* g, const float* queries, uint32_t nq, uint32_t qdim, uint32_t k, int metric,
                                  uint64_t* out_ids, double* out_scores, uint32_t* out_counts);
}

namespace {
// Runs `total` single-query searches (query i % nq_distinct) from `n_threads` threads sharing one cursor.
// out_* [nq_distinct][k] receive the answer of the first search of every distinct query (for checking).
// Returns the elapsed wall time in seconds, or a negative status if any call failed.
template <typename Call>
double run_callers(const Call& call, const float* queries, uint32_t nq_distinct, uint32_t dim, uint32_t k,
                   uint32_t n_threads, uint64_t total, uint64_t* out_ids, double* out_scores, uint32_t* out_counts) {
    std::atomic<uint64_t> cursor{0};
    std::atomic<int> failed{0};
    auto work = [&]() {
        std::vector<uint64_t> ids(k);
        std::vector<double> sc(k);
        uint32_t cnt = 0;
        for (;;) {
            const uint64_t i = cursor.fetch_add(1, std::memory_order_relaxed);
            if (i >= total || failed.load(std::memory_order_relaxed)) return;
            const uint32_t q = static_cast<uint32_t>(i % nq_distinct);
            const int rc = call(queries + static_cast<size_t>(q) * dim, ids.data(), sc.data(), &cnt);
            if (rc != 0) { failed.store(rc); return; }
            if (i < nq_distinct && out_ids) {
                memcpy(out_ids + static_cast<size_t>(q) * k, ids.data(), k * sizeof(uint64_t));
                memcpy(out_scores + static_cast<size_t>(q) * k, sc.data(), k * sizeof(double));
                out_counts[q] = cnt;
            }
        }
    };
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < n_threads; ++t) th.emplace_back(work);
    for (auto& t : th) t.join();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const int f = failed.load();
    return f ? -static_cast<double>(f < 0 ? -f : f) : dt;
}
}  // namespace

extern "C" {

double vl_native_callers(vl_search_fn search, void* h, const float* queries, uint32_t nq_distinct, uint32_t dim,
                         uint32_t k, int metric, uint32_t ef, uint32_t n_threads, uint64_t total, uint64_t* out_ids,
                         double* out_scores, uint32_t* out_counts) {
    return run_callers([=](const float* q, uint64_t* ids, double* sc, uint32_t* cnt) {
        return search(h, q, 1u, dim, k, metric, ef, ids, sc, cnt); },
        queries, nq_distinct, dim, k, n_threads, total, out_ids, out_scores, out_counts);
}

// the same callers on a shard group (vl_group_search: one process driving several GPUs)
double vl_native_callers_group(vl_group_search_fn search, void* g, const float* queries, uint32_t nq_distinct,
                               uint32_t dim, uint32_t k, int metric, uint32_t n_threads, uint64_t total,
                               uint64_t* out_ids, double* out_scores, uint32_t* out_counts) {
    return run_callers([=](const float* q, uint64_t* ids, double* sc, uint32_t* cnt) {
        return search(g, q, 1u, dim, k, metric, ids, sc, cnt); },
        queries, nq_distinct, dim, k, n_threads, total, out_ids, out_scores, out_counts);
}

}  // extern "C"
