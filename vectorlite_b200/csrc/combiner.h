// combiner.h — coalesces concurrent single-query callers into one batched search (host-only).
//
// The reference serves searches from many worker threads under a read lock (src/client.rs:398,
// src/server.rs:258-275), one query per call.  On the device one query is a memory-bound scan and a batch is a
// tensor-core contraction that serves up to 128 queries in the time of 1.6 single-query scans; HNSW runs one
// CTA per query, all in flight together.  So the first caller in becomes the leader, runs whatever has queued up
// behind the running launch as ONE batched search and hands the results back.  A lone caller runs its own query
// at once: no added latency, no timer.  Used by every handle (flat and HNSW, api.cu) and by the shard group
// (group.cpp), where the combined batch is fanned out to all shards at once.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/vectorlite_cuda.h"

namespace vl {

void set_last_error(const char* msg);   // api.cu: the calling thread's vl_last_error()

class Combiner {
public:
    // impl(queries [m][qdim], m, k, metric, ef, ids [m][k], scores [m][k], counts [m]) = the uncombined search
    using Impl = std::function<int(const float*, uint32_t, uint32_t, int, uint32_t, uint64_t*, double*, uint32_t*)>;
    static constexpr size_t MAX_BATCH = 128;

    int search(const float* query, uint32_t qdim, uint32_t k, int metric, uint32_t ef, uint64_t* out_ids,
               double* out_scores, uint32_t* out_counts, const Impl& impl, std::atomic<uint64_t>* combined_stat) {
        Pending me;
        me.q = query; me.k = k; me.metric = metric; me.ef = ef; me.ids = out_ids; me.scores = out_scores; me.count = out_counts;
        std::unique_lock<std::mutex> lk(mu_);
        queue_.push_back(&me);
        while (!me.done) {
            if (leader_) {
                // A batch is in flight (a few hundred µs on the device).  Waking through the condition variable costs a
                // futex round trip plus a reschedule per caller — tens of µs that the NEXT batch waits for, since its
                // leader gathers the re-forming cohort first.  When the callers fit the host's cores, wait for the
                // running batch by polling (bounded: spin_us, default 400 µs) and only then block.
                if (spin_us() > 0 && reserved_ <= 3 && expect_ + reserved_ <= cores() + cores() / 4) {
                    lk.unlock();
                    const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(spin_us());
                    while (!me.done.load(std::memory_order_acquire) && leader_.load(std::memory_order_acquire) &&
                           std::chrono::steady_clock::now() < deadline)
                        cpu_relax();
                    lk.lock();
                    if (me.done || !leader_) continue;
                }
                cv_.wait(lk);
                continue;
            }
            // leader: take the head of the queue and everything behind it with the same (k, metric, ef)
            leader_ = true;
            if (expect_ > 1 && queue_.size() < expect_) {
                // Re-forming cohort: the callers of the batch that just completed are waking up and coming back with
                // their next query.  Without this the first one back (usually the old leader) runs a batch of ONE while
                // the other callers queue behind it — measured at 16 callers: batches alternate 1, 15, 1, 15 …
                // Bounded (wait_us, default 100 µs), and only after a batch that DID combine callers: a lone
                // caller (expect_ == 1) never waits.
                const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(wait_us());
                while (queue_.size() < expect_ && std::chrono::steady_clock::now() < deadline) {
                    lk.unlock();
                    std::this_thread::yield();
                    lk.lock();
                }
            }
            std::vector<Pending*> batch;
            const uint32_t bk = queue_.front()->k, bef = queue_.front()->ef;
            const int bm = queue_.front()->metric;
            for (auto it = queue_.begin(); it != queue_.end() && batch.size() < MAX_BATCH;) {
                if ((*it)->k == bk && (*it)->metric == bm && (*it)->ef == bef) {
                    batch.push_back(*it);
                    it = queue_.erase(it);
                } else {
                    ++it;
                }
            }
            lk.unlock();
            const uint32_t m = static_cast<uint32_t>(batch.size());
            int rc;
            std::string err;
            if (m == 1) {
                Pending* p = batch[0];
                rc = impl(p->q, 1u, bk, bm, bef, p->ids, p->scores, p->count);
                if (rc != VL_OK) err = vl_last_error();
            } else {
                std::vector<float> qs(static_cast<size_t>(m) * qdim);
                std::vector<uint64_t> ids(static_cast<size_t>(m) * bk);
                std::vector<double> sc(static_cast<size_t>(m) * bk);
                std::vector<uint32_t> cnt(m);
                for (uint32_t i = 0; i < m; ++i) memcpy(qs.data() + static_cast<size_t>(i) * qdim, batch[i]->q, qdim * sizeof(float));
                rc = impl(qs.data(), m, bk, bm, bef, ids.data(), sc.data(), cnt.data());
                if (rc != VL_OK) err = vl_last_error();
                for (uint32_t i = 0; i < m; ++i) {
                    memcpy(batch[i]->ids, ids.data() + static_cast<size_t>(i) * bk, bk * sizeof(uint64_t));
                    memcpy(batch[i]->scores, sc.data() + static_cast<size_t>(i) * bk, bk * sizeof(double));
                    *batch[i]->count = cnt[i];
                }
                if (combined_stat) *combined_stat += m;
            }
            std::vector<int> rcs(m, rc);
            std::vector<std::string> errs(m, err);
            if (rc != VL_OK && m > 1) {   // a batch-level failure (e.g. one NaN query) must not leak to the other callers
                for (uint32_t i = 0; i < m; ++i) {
                    Pending* p = batch[i];
                    rcs[i] = impl(p->q, 1u, bk, bm, bef, p->ids, p->scores, p->count);
                    errs[i] = rcs[i] != VL_OK ? std::string(vl_last_error()) : std::string();
                }
            }
            lk.lock();
            for (uint32_t i = 0; i < m; ++i) {
                batch[i]->rc = rcs[i];
                batch[i]->err = errs[i];
                batch[i]->done = true;
            }
            leader_ = false;
            expect_ = m;
            cv_.notify_all();
        }
        lk.unlock();
        if (me.rc != VL_OK) set_last_error(me.err.c_str());   // re-raise in the caller's thread
        return me.rc;
    }

    // host threads that work for the running batch besides its leader (a shard group's helper threads).  Measured with
    // 16 callers (scripts/group_e2e.py): polling waiters gain 9 % / 10 % with 1 / 3 helpers (2 / 4 GPUs, 16-core host)
    // and LOSE 39 % with 7 (8 GPUs, 32-core host: 270 K vs 439 K q/s x shards) although cores are left over — so
    // polling is kept to handles and groups of at most 4 shards.
    void reserve_cores(size_t n) { reserved_ = n; }

    static bool enabled() {
        static const bool on = std::getenv("VL_DISABLE_COMBINER") == nullptr;
        return on;
    }

private:
    struct Pending {
        const float* q; uint32_t k; int metric; uint32_t ef;
        uint64_t* ids; double* scores; uint32_t* count;
        int rc = VL_OK; std::atomic<bool> done{false}; std::string err;
    };
    static int spin_us() {
        static const int us = [] { const char* e = std::getenv("VL_COMBINE_SPIN_US"); return e ? std::max(0, atoi(e)) : 400; }();
        return us;
    }
    static size_t cores() {
        static const size_t n = std::max(1u, std::thread::hardware_concurrency());
        return n;
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
    static int wait_us() {
        static const int us = [] { const char* e = std::getenv("VL_COMBINE_WAIT_US"); return e ? std::max(0, atoi(e)) : 100; }();
        return us;
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::vector<Pending*> queue_;
    std::atomic<bool> leader_{false};
    size_t expect_ = 1;   // size of the batch that just completed: how many callers the next leader may wait for
    size_t reserved_ = 0;
};

}  // namespace vl
