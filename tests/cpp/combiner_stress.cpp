// CPU-only stress test of the caller combiner (vectorlite_b200/csrc/combiner.h) with a stand-in search: many threads
// call search() with ONE query each; every caller must get the answer to ITS query, batch-level failures must stay
// with the offending caller, and nobody may hang — with the polling waiters on (default) and off.
//   g++ -O2 -std=c++17 -pthread -o combiner_stress combiner_stress.cpp && ./combiner_stress
#include <atomic>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../../vectorlite_b200/csrc/combiner.h"

static thread_local std::string g_err;
namespace vl { void set_last_error(const char* msg) { g_err = msg ? msg : ""; } }
extern "C" const char* vl_last_error(void) { return g_err.c_str(); }

int main(int argc, char** argv) {
    const int threads = argc > 1 ? atoi(argv[1]) : 8, per_thread = argc > 2 ? atoi(argv[2]) : 1500;
    const uint32_t qdim = 4, k = 3;
    vl::Combiner comb;
    std::atomic<uint64_t> combined{0}, batches{0}, max_batch{0};
    // the "index": result id j of query q is q[0] * 10 + j, score q[1] + j; a query with q[2] < 0 fails the whole batch
    vl::Combiner::Impl impl = [&](const float* q, uint32_t m, uint32_t kk, int, uint32_t, uint64_t* ids, double* sc, uint32_t* cnt) {
        batches++;
        uint64_t mb = max_batch.load();
        while (m > mb && !max_batch.compare_exchange_weak(mb, m)) {}
        std::this_thread::sleep_for(std::chrono::microseconds(60));
        for (uint32_t i = 0; i < m; ++i)
            if (q[i * qdim + 2] < 0.f) { vl::set_last_error("bad query"); return VL_ERR_INVALID; }
        for (uint32_t i = 0; i < m; ++i) {
            for (uint32_t j = 0; j < kk; ++j) {
                ids[i * kk + j] = static_cast<uint64_t>(q[i * qdim]) * 10 + j;
                sc[i * kk + j] = q[i * qdim + 1] + j;
            }
            cnt[i] = kk;
        }
        return VL_OK;
    };
    std::atomic<int> failures{0};
    std::vector<std::thread> ts;
    for (int t = 0; t < threads; ++t)
        ts.emplace_back([&, t] {
            for (int i = 0; i < per_thread; ++i) {
                const bool bad = (i % 97) == 13 && (t % 3) == 1;
                float q[4] = {static_cast<float>(t * 100000 + i), static_cast<float>(t) + 0.5f, bad ? -1.f : 1.f, 0.f};
                uint64_t ids[3] = {0, 0, 0};
                double sc[3] = {0, 0, 0};
                uint32_t cnt = 0;
                // two (k, metric) classes in flight at once: callers of different classes must never share a batch
                const int metric = t & 1;
                const int rc = comb.search(q, qdim, k, metric, 0u, ids, sc, &cnt, impl, &combined);
                if (bad) {
                    if (rc != VL_ERR_INVALID || std::string(vl_last_error()) != "bad query") failures++;
                } else if (rc != VL_OK || cnt != k || ids[0] != static_cast<uint64_t>(q[0]) * 10 || ids[2] != ids[0] + 2 ||
                           sc[1] != q[1] + 1) {
                    failures++;
                }
            }
        });
    for (auto& t : ts) t.join();
    std::printf("{\"threads\": %d, \"queries\": %d, \"batches\": %llu, \"combined\": %llu, \"max_batch\": %llu, \"failures\": %d}\n",
                threads, threads * per_thread, (unsigned long long)batches.load(), (unsigned long long)combined.load(),
                (unsigned long long)max_batch.load(), failures.load());
    return failures.load() == 0 ? 0 : 1;
}
