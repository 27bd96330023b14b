// group.cpp — a flat store ROW-SHARDED over several handles of one process (one per GPU) searched as ONE index.
//
// The reference is a single-process server: `Collection` owns one index behind an RwLock and calls
// VectorIndex::search from its worker threads (src/client.rs:243-247,398).  On a multi-GPU box the CUDA-backed
// collection owns a shard group instead: shard g holds the contiguous storage-order range [base_g, base_g + n_g)
// (SURVEY §8e), every shard answers the query exactly (certified, f64 scores), and the per-shard top-k lists are
// merged by a STABLE sort in shard order — score descending, global insertion order ascending, the tie-break of
// flat.rs:116.  Routing of inserts / deletes to shards is host bookkeeping above the ABI (multi_gpu.py; the Rust
// shim in INTEGRATION.md).
//
// Fan-out: the calling thread runs shard 0 itself and posts the other shards to a small pool of helper threads, so
// all GPUs scan at the same time; several callers may be inside vl_group_search at once (their per-shard calls meet
// in each handle's combiner exactly like direct callers).  Host-only code: everything on the device happens
// inside vl_index_search.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/vectorlite_cuda.h"
#include "combiner.h"

namespace {

struct Call {
    const float* queries;
    uint32_t nq, qdim, k;
    int metric;
    std::vector<std::vector<uint64_t>> ids;      // [shard][nq*k]
    std::vector<std::vector<double>> scores;
    std::vector<std::vector<uint32_t>> counts;   // [shard][nq]
    std::vector<int> rc;
    std::vector<std::string> err;
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<uint32_t> remaining{0};
};

struct Task {
    Call* call;
    uint32_t shard;
};

}  // namespace

struct vl_group {
    std::vector<vl_index*> shards;
    uint32_t dim = 0;
    std::vector<std::thread> helpers;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Task> queue;
    bool stop = false;
    std::atomic<int> pending{0};     // tasks in `queue` (read by polling helpers without the lock)
    std::atomic<int> spinners{0};    // helpers currently polling
    std::atomic<bool> stopping{false};
    vl::Combiner comb;   // concurrent single-query callers → ONE fan-out of a small batch to all shards

    // A fan-out every few hundred µs: waking the helpers through the condition variable costs a futex round trip and
    // a reschedule per shard and per batch — comparable to the device time of a small batch.  Up to one helper per
    // remote shard therefore keeps polling for the next task for a bounded time (VL_GROUP_SPIN_US, default 400 µs)
    // before it blocks; the caller polls the completion count the same way.  Only when helpers + callers fit the cores.
    static int spin_us() {
        static const int us = [] { const char* e = std::getenv("VL_GROUP_SPIN_US"); return e ? std::max(0, atoi(e)) : 400; }();
        return us;
    }
    static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#elif defined(__aarch64__)
        asm volatile("yield" ::: "memory");
#else
        std::this_thread::yield();
#endif
    }
    int max_spinners() const {   // one polling helper per remote shard in groups of at most 4 shards: measured +9 % / +10 %
        // at 2 / 4 GPUs, −39 % at 8 (7 polling helpers + 16 polling callers, 32-core host: 270 K vs 439 K q/s x shards)
        const int cores = static_cast<int>(std::max(1u, std::thread::hardware_concurrency()));
        const int remote = static_cast<int>(shards.size()) - 1;
        return (remote <= 3 && remote <= cores / 4) ? remote : 0;
    }

    void run_shard(Call* c, uint32_t s) {
        c->ids[s].resize(static_cast<size_t>(c->nq) * c->k);
        c->scores[s].resize(static_cast<size_t>(c->nq) * c->k);
        c->counts[s].assign(c->nq, 0u);
        c->rc[s] = vl_index_search(shards[s], c->queries, c->nq, c->qdim, c->k, c->metric, 0u, c->ids[s].data(),
                                   c->scores[s].data(), c->counts[s].data());
        if (c->rc[s] != VL_OK) c->err[s] = vl_last_error();
    }

    void helper_loop() {
        for (;;) {
            Task t;
            bool got = false;
            {
                std::unique_lock<std::mutex> lk(mu);
                if (!queue.empty()) {
                    t = queue.front();
                    queue.pop_front();
                    pending.fetch_sub(1, std::memory_order_relaxed);
                    got = true;
                } else if (stop) {
                    return;
                }
            }
            if (!got) {
                if (spin_us() > 0 && spinners.fetch_add(1) < max_spinners()) {
                    const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(spin_us());
                    while (pending.load(std::memory_order_acquire) == 0 && !stopping.load(std::memory_order_relaxed) &&
                           std::chrono::steady_clock::now() < deadline)
                        cpu_relax();
                    spinners.fetch_sub(1);
                    if (pending.load(std::memory_order_acquire) > 0) continue;   // try to take it
                } else if (spin_us() > 0) {
                    spinners.fetch_sub(1);
                }
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || !queue.empty(); });
                continue;
            }
            run_shard(t.call, t.shard);
            {   // decrement UNDER the call's mutex: the caller (which polls `remaining`, then takes this mutex before it
                // lets the Call go out of scope) cannot destroy it while this thread is still inside
                std::lock_guard<std::mutex> lk(t.call->mu);
                if (t.call->remaining.fetch_sub(1, std::memory_order_acq_rel) == 1) t.call->cv.notify_one();
            }
        }
    }
};

extern "C" {

int vl_group_create(vl_index* const* shards, uint32_t n, vl_group** out) {
    if (!out) { vl::set_last_error("out is null"); return VL_ERR_INVALID; }
    *out = nullptr;
    if (!shards || n == 0) { vl::set_last_error("a shard group needs at least one index"); return VL_ERR_INVALID; }
    for (uint32_t i = 0; i < n; ++i) {
        if (!shards[i] || vl_index_type_of(shards[i]) != VL_INDEX_FLAT || vl_index_dim(shards[i]) != vl_index_dim(shards[0])) {
            vl::set_last_error("every shard must be a flat index of the same dimension");
            return VL_ERR_INVALID;
        }
    }
    vl_group* g = new (std::nothrow) vl_group();
    if (!g) { vl::set_last_error("host allocation failed"); return VL_ERR_OOM; }
    g->shards.assign(shards, shards + n);
    g->dim = vl_index_dim(shards[0]);
    // helpers: enough for a few concurrent callers to have all their shards in flight
    const uint32_t n_helpers = n > 1 ? std::min<uint32_t>(64u, (n - 1) * 8u) : 0u;
    for (uint32_t i = 0; i < n_helpers; ++i) g->helpers.emplace_back([g] { g->helper_loop(); });
    g->comb.reserve_cores(n - 1);
    *out = g;
    return VL_OK;
}

void vl_group_destroy(vl_group* g) {   // the shards stay alive: they belong to the caller
    if (!g) return;
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->stop = true;
        g->stopping = true;
    }
    g->cv.notify_all();
    for (auto& t : g->helpers) t.join();
    delete g;
}

uint32_t vl_group_size(const vl_group* g) { return g ? static_cast<uint32_t>(g->shards.size()) : 0u; }

static int group_search_impl(vl_group* g, const float* queries, uint32_t nq, uint32_t qdim, uint32_t k, int metric,
                             uint64_t* out_ids, double* out_scores, uint32_t* out_counts);

int vl_group_search(vl_group* g, const float* queries, uint32_t nq, uint32_t qdim, uint32_t k, int metric,
                    uint64_t* out_ids, double* out_scores, uint32_t* out_counts) {
    if (!g || !out_ids || !out_scores || !out_counts || (!queries && nq)) {
        vl::set_last_error("null argument");
        return VL_ERR_INVALID;
    }
    // One query per call from many threads (the reference's serving pattern, client.rs:398): the first caller in
    // runs whatever queued up behind the running search as one batch — ONE fan-out to the shards (where it takes
    // the tensor-core pipeline) and one merge, instead of callers x shards helper hand-offs.
    if (vl::Combiner::enabled() && nq == 1 && k > 0 && qdim == g->dim && g->shards.size() > 1)
        return g->comb.search(queries, qdim, k, metric, 0u, out_ids, out_scores, out_counts,
                              [g, qdim](const float* q, uint32_t m, uint32_t bk, int bm, uint32_t, uint64_t* ids, double* sc,
                                        uint32_t* cnt) { return group_search_impl(g, q, m, qdim, bk, bm, ids, sc, cnt); },
                              nullptr);
    return group_search_impl(g, queries, nq, qdim, k, metric, out_ids, out_scores, out_counts);
}

static int group_search_impl(vl_group* g, const float* queries, uint32_t nq, uint32_t qdim, uint32_t k, int metric,
                             uint64_t* out_ids, double* out_scores, uint32_t* out_counts) {
    for (size_t i = 0; i < static_cast<size_t>(nq) * k; ++i) { out_ids[i] = ~0ull; out_scores[i] = 0.0; }
    for (uint32_t q = 0; q < nq; ++q) out_counts[q] = 0u;
    // shards that hold rows; flat.rs:99-104: the dimension is only checked when the store is non-empty
    std::vector<uint32_t> live;
    for (uint32_t s = 0; s < g->shards.size(); ++s)
        if (vl_index_len(g->shards[s]) > 0) live.push_back(s);
    if (live.empty() || nq == 0 || k == 0) return VL_OK;
    if (live.size() == 1)
        return vl_index_search(g->shards[live[0]], queries, nq, qdim, k, metric, 0u, out_ids, out_scores, out_counts);

    const uint32_t S = static_cast<uint32_t>(g->shards.size());
    Call c;
    c.queries = queries; c.nq = nq; c.qdim = qdim; c.k = k; c.metric = metric;
    c.ids.resize(S); c.scores.resize(S); c.counts.resize(S); c.rc.assign(S, VL_OK); c.err.resize(S);
    c.remaining = static_cast<uint32_t>(live.size()) - 1;
    {
        std::lock_guard<std::mutex> lk(g->mu);
        for (size_t i = 1; i < live.size(); ++i) g->queue.push_back(Task{&c, live[i]});
        g->pending.fetch_add(static_cast<int>(live.size()) - 1, std::memory_order_release);
    }
    g->cv.notify_all();
    g->run_shard(&c, live[0]);
    if (vl_group::spin_us() > 0 && g->max_spinners() > 0) {   // the other shards finish within µs of this one: poll before blocking
        const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(vl_group::spin_us());
        while (c.remaining.load(std::memory_order_acquire) != 0 && std::chrono::steady_clock::now() < deadline)
            vl_group::cpu_relax();
    }
    {
        std::unique_lock<std::mutex> lk(c.mu);
        c.cv.wait(lk, [&] { return c.remaining.load(std::memory_order_acquire) == 0; });
    }
    for (uint32_t s : live)
        if (c.rc[s] != VL_OK) {   // every shard sees the same query: report the first failure (dimension, NaN, …)
            vl::set_last_error(c.err[s].c_str());
            return c.rc[s];
        }
    // stable merge in shard order == (score desc, global storage position asc)
    struct Hit { double score; uint64_t id; };
    std::vector<Hit> hits;
    for (uint32_t q = 0; q < nq; ++q) {
        hits.clear();
        for (uint32_t s : live) {
            const uint32_t cnt = c.counts[s][q];
            for (uint32_t r = 0; r < cnt; ++r)
                hits.push_back(Hit{c.scores[s][static_cast<size_t>(q) * k + r], c.ids[s][static_cast<size_t>(q) * k + r]});
        }
        std::stable_sort(hits.begin(), hits.end(), [](const Hit& a, const Hit& b) { return a.score > b.score; });
        const uint32_t cnt = static_cast<uint32_t>(std::min<size_t>(hits.size(), k));
        for (uint32_t r = 0; r < cnt; ++r) {
            out_ids[static_cast<size_t>(q) * k + r] = hits[r].id;
            out_scores[static_cast<size_t>(q) * k + r] = hits[r].score;
        }
        out_counts[q] = cnt;
    }
    return VL_OK;
}

}  // extern "C"
