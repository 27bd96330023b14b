// CPU-only stress test of the shard group (vectorlite_b200/csrc/group.cpp) over stand-in shards: the fan-out to the
// helper threads (polling or blocking), the completion hand-shake, the caller combiner in front of it and the stable
// merge in shard order — many callers, no GPU.
//   g++ -O2 -std=c++17 -pthread -I. -o group_stress group_stress.cpp ../../vectorlite_b200/csrc/group.cpp
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "../../include/vectorlite_cuda.h"

struct vl_index { uint32_t shard; uint32_t shards; uint64_t n; uint32_t dim; };

static thread_local std::string g_err;
namespace vl { void set_last_error(const char* msg) { g_err = msg ? msg : ""; } }
static std::atomic<uint64_t> g_calls{0}, g_queries{0};

extern "C" {
const char* vl_last_error(void) { return g_err.c_str(); }
uint64_t vl_index_len(const vl_index* h) { return h->n; }
uint32_t vl_index_dim(const vl_index* h) { return h->dim; }
int vl_index_type_of(const vl_index*) { return VL_INDEX_FLAT; }
// stand-in shard: rank r of shard s scores base − r (TIES across shards at equal r: the merge must keep shard order)
// and has id r·S + s, so the merged top-k of every query is 0, 1, 2, … ; query[1] < 0 fails the call.
int vl_index_search(vl_index* h, const float* q, uint32_t nq, uint32_t qdim, uint32_t k, int, uint32_t, uint64_t* ids,
                    double* sc, uint32_t* cnt) {
    g_calls++; g_queries += nq;
    std::this_thread::sleep_for(std::chrono::microseconds(40 + 7 * h->shard));
    for (uint32_t i = 0; i < nq; ++i) {
        if (q[i * qdim + 1] < 0.f) { vl::set_last_error("bad query"); return VL_ERR_INVALID; }
        for (uint32_t r = 0; r < k; ++r) {
            ids[i * k + r] = static_cast<uint64_t>(q[i * qdim]) * 1000 + r * h->shards + h->shard;
            sc[i * k + r] = static_cast<double>(q[i * qdim]) - r;
        }
        cnt[i] = k;
    }
    return VL_OK;
}
}

int main(int argc, char** argv) {
    const uint32_t S = argc > 1 ? atoi(argv[1]) : 4;
    const int threads = argc > 2 ? atoi(argv[2]) : 16, per_thread = argc > 3 ? atoi(argv[3]) : 400;
    const uint32_t qdim = 4, k = 10;
    std::vector<vl_index> shards(S);
    std::vector<vl_index*> ptrs(S);
    for (uint32_t s = 0; s < S; ++s) { shards[s] = vl_index{s, S, 1000, qdim}; ptrs[s] = &shards[s]; }
    vl_group* g = nullptr;
    if (vl_group_create(ptrs.data(), S, &g) != VL_OK) return 2;
    std::atomic<int> failures{0};
    std::vector<std::thread> ts;
    for (int t = 0; t < threads; ++t)
        ts.emplace_back([&, t] {
            for (int i = 0; i < per_thread; ++i) {
                const bool bad = (i % 89) == 7 && (t % 4) == 2;
                const float base = static_cast<float>(t * 1000 + i % 1000);
                float q[4] = {base, bad ? -1.f : 1.f, 0.f, 0.f};
                uint64_t ids[10]; double sc[10]; uint32_t cnt = 0;
                const int rc = vl_group_search(g, q, 1, qdim, k, 0, ids, sc, &cnt);
                if (bad) { if (rc != VL_ERR_INVALID) failures++; continue; }
                if (rc != VL_OK || cnt != k) { failures++; continue; }
                for (uint32_t r = 0; r < k; ++r)   // merged order: score desc, shard order among ties → ids base·1000 + 0, 1, 2, …
                    if (ids[r] != static_cast<uint64_t>(base) * 1000 + r || sc[r] != static_cast<double>(base) - r / S) { failures++; break; }
            }
        });
    for (auto& t : ts) t.join();
    // a multi-query call goes straight to the fan-out (no combiner)
    std::vector<float> qs(8 * qdim, 0.f);
    for (int i = 0; i < 8; ++i) { qs[i * qdim] = 5000.f + i; qs[i * qdim + 1] = 1.f; }
    std::vector<uint64_t> ids(8 * k); std::vector<double> sc(8 * k); std::vector<uint32_t> cnt(8);
    if (vl_group_search(g, qs.data(), 8, qdim, k, 0, ids.data(), sc.data(), cnt.data()) != VL_OK) failures++;
    for (int i = 0; i < 8; ++i)
        for (uint32_t r = 0; r < k; ++r)
            if (ids[i * k + r] != static_cast<uint64_t>(5000 + i) * 1000 + r) { failures++; break; }
    vl_group_destroy(g);
    std::printf("{\"shards\": %u, \"threads\": %d, \"queries\": %d, \"shard_calls\": %llu, \"shard_queries\": %llu, \"failures\": %d}\n",
                S, threads, threads * per_thread, (unsigned long long)g_calls.load(), (unsigned long long)g_queries.load(), failures.load());
    return failures.load() == 0 ? 0 : 1;
}
