// hnsw_host.cpp — host side of the HNSW index: graph construction, id bookkeeping, upload.
//
// Boundary mirrored: HNSWIndex (src/index/hnsw.rs:197-518) — add / soft delete / search /
// len / get_vector / max_id, id↔internal-index maps (hnsw.rs:204-207), one metric per index
// (hnsw.rs:216-259).  The reference delegates the graph to crate hnsw 0.11 (not in its tree); this
// is NOT a restatement of that crate (the restatement lives in oracle/ and is only the recall
// baseline).  Construction here is the published HNSW algorithm (Malkov & Yashunin): exponential
// level draw with mult 1/ln(M), greedy descent, ef_construction beam per level, heuristic
// neighbour selection with pruned-fill, M links per new node, caps M (upper) / M0 (layer 0) on
// back-links; inserted in parallel over host threads with per-node spin locks.  Distances are true
// fp32 metrics (not the reference's u64 milli-unit quantisation, hnsw.rs:113-174), which is why
// recall at equal (M, M0, ef_construction, ef) is >= the reference's.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <thread>

#include "hnsw.h"
#include "hnsw_state.h"

namespace vl {

namespace {

enum { COS = 0, EUC = 1, MAN = 2, DOTP = 3 };

// 16 independent partial sums: lets the host compiler vectorise without -ffast-math
__attribute__((target_clones("avx2,fma", "default"))) float dot_f32(const float* a, const float* b, uint32_t n) {
    float acc[16] = {0};
    uint32_t i = 0;
    for (; i + 16 <= n; i += 16)
        for (int j = 0; j < 16; ++j) acc[j] += a[i + j] * b[i + j];
    float s = 0.f;
    for (; i < n; ++i) s += a[i] * b[i];
    for (int j = 0; j < 16; ++j) s += acc[j];
    return s;
}
__attribute__((target_clones("avx2,fma", "default"))) float l2sq_f32(const float* a, const float* b, uint32_t n) {
    float acc[16] = {0};
    uint32_t i = 0;
    for (; i + 16 <= n; i += 16)
        for (int j = 0; j < 16; ++j) {
            const float d = a[i + j] - b[i + j];
            acc[j] += d * d;
        }
    float s = 0.f;
    for (; i < n; ++i) {
        const float d = a[i] - b[i];
        s += d * d;
    }
    for (int j = 0; j < 16; ++j) s += acc[j];
    return s;
}
__attribute__((target_clones("avx2,fma", "default"))) float l1_f32(const float* a, const float* b, uint32_t n) {
    float acc[16] = {0};
    uint32_t i = 0;
    for (; i + 16 <= n; i += 16)
        for (int j = 0; j < 16; ++j) acc[j] += std::fabs(a[i + j] - b[i + j]);
    float s = 0.f;
    for (; i < n; ++i) s += std::fabs(a[i] - b[i]);
    for (int j = 0; j < 16; ++j) s += acc[j];
    return s;
}

struct Builder {
    HnswState& g;
    explicit Builder(HnswState& s) : g(s) {}

    const float* vec(uint32_t i) const { return g.vecs.data() + static_cast<size_t>(i) * g.dim; }
    float dist(uint32_t a, uint32_t b) const {  // lower is closer
        switch (g.metric) {
            case COS: return 1.0f - dot_f32(vec(a), vec(b), g.dim) * g.inv_norm[a] * g.inv_norm[b];
            case EUC: return l2sq_f32(vec(a), vec(b), g.dim);
            case MAN: return l1_f32(vec(a), vec(b), g.dim);
            default: return -dot_f32(vec(a), vec(b), g.dim);
        }
    }
    uint32_t* adj(uint32_t node, int lvl) {
        return lvl == 0 ? g.adj0.data() + static_cast<size_t>(node) * g.M0
                        : g.upper.data() + (static_cast<size_t>(g.upper_off[node]) + lvl - 1) * g.M;
    }
    uint32_t cap(int lvl) const { return lvl == 0 ? g.M0 : g.M; }
    void lock(uint32_t i) {
        uint8_t exp = 0;
        while (!g.locks[i].compare_exchange_weak(exp, 1, std::memory_order_acquire)) {
            exp = 0;
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
    }
    void unlock(uint32_t i) { g.locks[i].store(0, std::memory_order_release); }
    uint32_t copy_adj(uint32_t node, int lvl, uint32_t* out) {
        lock(node);
        const uint32_t* a = adj(node, lvl);
        const uint32_t c = cap(lvl);
        uint32_t m = 0;
        while (m < c && a[m] != HNSW_NONE) { out[m] = a[m]; ++m; }
        unlock(node);
        return m;
    }

    struct Scratch {
        std::vector<uint32_t> stamp;
        uint32_t epoch = 0;
        std::vector<std::pair<float, uint32_t>> W, sel, pruned;
        std::vector<uint32_t> nb;
        std::vector<uint64_t> touched;   // (node << 8) | level of adjacency rows written by this thread
    };
    using DI = std::pair<float, uint32_t>;

    // beam search on one level; results (ascending by distance) in sc.W
    void search_layer(uint32_t q, uint32_t ep, float epd, uint32_t ef, int lvl, Scratch& sc) {
        if (++sc.epoch == 0) { std::fill(sc.stamp.begin(), sc.stamp.end(), 0u); sc.epoch = 1; }
        std::priority_queue<DI, std::vector<DI>, std::greater<DI>> cand;  // closest first
        std::priority_queue<DI> res;                                         // farthest first
        cand.emplace(epd, ep);
        res.emplace(epd, ep);
        sc.stamp[ep] = sc.epoch;
        sc.nb.resize(std::max(g.M, g.M0));
        while (!cand.empty()) {
            const DI c = cand.top();
            if (c.first > res.top().first && res.size() >= ef) break;
            cand.pop();
            const uint32_t m = copy_adj(c.second, lvl, sc.nb.data());
            for (uint32_t j = 0; j < m; ++j) {
                const uint32_t v = sc.nb[j];
                if (sc.stamp[v] == sc.epoch) continue;
                sc.stamp[v] = sc.epoch;
                const float d = dist(q, v);
                if (res.size() < ef || d < res.top().first) {
                    cand.emplace(d, v);
                    res.emplace(d, v);
                    if (res.size() > ef) res.pop();
                }
            }
        }
        sc.W.resize(res.size());
        for (size_t i = res.size(); i-- > 0;) { sc.W[i] = res.top(); res.pop(); }
    }

    // Algorithm 4 (heuristic) with keepPrunedConnections; `cands` ascending by distance to base
    void select(const std::vector<DI>& cands, uint32_t m, Scratch& sc) {
        sc.sel.clear();
        sc.pruned.clear();
        for (const DI& c : cands) {
            if (sc.sel.size() >= m) break;
            bool good = true;
            for (const DI& r : sc.sel)
                if (dist(c.second, r.second) < c.first) { good = false; break; }
            (good ? sc.sel : sc.pruned).push_back(c);
        }
        for (size_t i = 0; i < sc.pruned.size() && sc.sel.size() < m; ++i) sc.sel.push_back(sc.pruned[i]);
    }

    void insert(uint32_t q, Scratch& sc) {
        const int lvl = g.level[q];
        std::unique_lock<std::mutex> top(g.entry_mu, std::defer_lock);
        top.lock();
        const int maxl = g.max_level;
        uint32_t cur = g.entry;
        if (lvl <= maxl) top.unlock();  // only a node that raises the top level holds the entry lock
        if (maxl < 0) {                 // first node
            g.entry = q;
            g.max_level = lvl;
            return;
        }
        float curd = dist(q, cur);
        std::vector<uint32_t> nb(std::max(g.M, g.M0));
        for (int l = maxl; l > lvl; --l) {  // greedy descent above the node's level
            bool changed = true;
            while (changed) {
                changed = false;
                const uint32_t m = copy_adj(cur, l, nb.data());
                for (uint32_t j = 0; j < m; ++j) {
                    const float d = dist(q, nb[j]);
                    if (d < curd) { curd = d; cur = nb[j]; changed = true; }
                }
            }
        }
        std::vector<DI> tmp;
        for (int l = std::min(lvl, maxl); l >= 0; --l) {
            search_layer(q, cur, curd, g.efc, l, sc);
            select(sc.W, g.M, sc);  // new node links to M neighbours on every level
            const std::vector<DI> mine = sc.sel;
            lock(q);
            uint32_t* a = adj(q, l);
            for (size_t i = 0; i < mine.size(); ++i) a[i] = mine[i].second;
            unlock(q);
            sc.touched.push_back((static_cast<uint64_t>(q) << 8) | static_cast<uint64_t>(l));
            for (const DI& e : mine) {  // back-links, capped at M (upper) / M0 (layer 0)
                const uint32_t t = e.second;
                sc.touched.push_back((static_cast<uint64_t>(t) << 8) | static_cast<uint64_t>(l));
                lock(t);
                uint32_t* ta = adj(t, l);
                const uint32_t c = cap(l);
                uint32_t m = 0;
                while (m < c && ta[m] != HNSW_NONE) ++m;
                if (m < c) {
                    ta[m] = q;
                    unlock(t);
                } else {
                    // lock(t) is held across the whole re-selection: dropping it between the copy and the write-back
                    // let a concurrent thread's append / re-selection on t be overwritten by this (then stale)
                    // selection — lost edges, run-to-run recall variation.  select() only reads row data and the
                    // lists of OTHER nodes are not touched, so no second lock is taken while t is held.
                    tmp.clear();
                    tmp.emplace_back(e.first, q);
                    for (uint32_t j = 0; j < m; ++j) tmp.emplace_back(dist(t, ta[j]), ta[j]);
                    std::sort(tmp.begin(), tmp.end());
                    select(tmp, c, sc);
                    for (uint32_t j = 0; j < c; ++j) ta[j] = j < sc.sel.size() ? sc.sel[j].second : HNSW_NONE;
                    unlock(t);
                }
            }
            cur = sc.W.front().second;
            curd = sc.W.front().first;
        }
        if (lvl > maxl) {
            g.entry = q;
            g.max_level = lvl;
        }
    }
};

uint64_t next_rand(uint64_t& s) {  // splitmix64
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <typename T>
bool dev_grow(T*& p, size_t& cap_elems, size_t need, size_t elem = sizeof(T)) {
    (void)elem;
    if (need <= cap_elems && p) return true;
    if (p) cudaFree(p);
    p = nullptr;
    const size_t nc = need + need / 4 + 64;
    if (cudaMalloc(&p, nc * sizeof(T)) != cudaSuccess) return false;
    cap_elems = nc;
    return true;
}

}  // namespace

void HnswDeleter::operator()(HnswState* s) const {
    if (!s) return;
    hnsw_state_release_device(s);
    delete s;
}

HnswState* hnsw_state_create(uint32_t dim, int metric, uint32_t M, uint32_t M0, uint32_t efc) {
    if (dim == 0 || M < 2 || M > 64 || M0 < M || M0 > 64 || efc < 1) return nullptr;
    HnswState* s = new HnswState();
    s->dim = dim;
    s->metric = metric;
    s->M = M;
    s->M0 = M0;
    s->efc = efc;
    if (const char* e = std::getenv("VL_HNSW_BEAM_MULT")) s->beam_mult = static_cast<uint32_t>(std::max(1, atoi(e)));
    return s;
}

void hnsw_state_release_device(HnswState* s) {
    cudaFree(s->d_adj0); cudaFree(s->d_upper_off); cudaFree(s->d_upper); cudaFree(s->d_level);
    cudaFree(s->d_deleted); cudaFree(s->d_ids); cudaFree(s->d_inv_norm);
    for (auto& c : s->ctxs) {
        cudaFree(c->d_q); cudaFree(c->d_out); cudaFreeHost(c->h_out); cudaFree(c->d_visited); cudaFreeHost(c->h_visited);
        if (c->stream) cudaStreamDestroy(c->stream);
    }
    s->ctxs.clear();
    s->d_adj0 = s->d_upper_off = s->d_upper = nullptr;
    s->d_level = s->d_deleted = nullptr;
    s->d_ids = nullptr; s->d_inv_norm = nullptr;
    s->d_n_cap = s->d_upper_cap = 0;
    s->dirty = s->deleted_dirty = s->dirty_full = true;
    s->uploaded_n = s->uploaded_upper = 0;
}

bool hnsw_has_id(const HnswState* s, uint64_t id) { return s->index_of.count(id) != 0; }
bool hnsw_index_of(const HnswState* s, uint64_t id, uint64_t* ix) {
    auto it = s->index_of.find(id);
    if (it == s->index_of.end()) return false;
    *ix = it->second;
    return true;
}
bool hnsw_max_id(const HnswState* s, uint64_t* out) {  // metadata.keys().max() (hnsw.rs:267-269)
    if (s->index_of.empty()) return false;
    uint64_t m = 0;
    for (const auto& kv : s->index_of) m = std::max(m, kv.first);
    *out = m;
    return true;
}
uint64_t hnsw_live(const HnswState* s) { return s->live; }

void hnsw_set_builder(HnswState* s, int builder) { s->builder = builder; }
void hnsw_set_score_mode(HnswState* s, uint32_t mode) { s->score_mode = mode; }
void hnsw_set_beam_mult(HnswState* s, uint32_t mult) { s->beam_mult = mult < 1 ? 1 : mult; }
uint32_t hnsw_beam_mult(const HnswState* s) { return s->beam_mult; }
void hnsw_build_info(const HnswState* s, uint64_t out[2]) {
    out[0] = static_cast<uint64_t>(s->last_builder);
    out[1] = s->last_build_us;
}

int hnsw_add_rows(HnswState* s, const uint64_t* ids, const float* rows, uint64_t n, const float* d_rows,
                  uint32_t pitch, cudaStream_t stream, uint64_t* launches) {
    if (n == 0) return 0;
    const auto t_begin = std::chrono::steady_clock::now();
    const uint32_t first = static_cast<uint32_t>(s->level.size());
    const uint64_t total = first + n;
    if (total >= 0x7FFFFFFFull) return 9;
    // ---- grow flat arrays up front (no reallocation during the parallel phase) ----
    s->vecs.insert(s->vecs.end(), rows, rows + n * s->dim);
    s->inv_norm.resize(total);
    s->level.resize(total);
    s->adj0.resize(total * s->M0, HNSW_NONE);
    s->upper_off.resize(total, HNSW_NONE);
    s->deleted.resize(total, 0);
    s->id_of.insert(s->id_of.end(), ids, ids + n);
    const double mult = 1.0 / std::log(static_cast<double>(s->M));
    size_t slots = s->upper.size() / s->M;
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t q = first + static_cast<uint32_t>(i);
        double ss = 0.0;
        const float* v = rows + i * s->dim;
        for (uint32_t c = 0; c < s->dim; ++c) ss += static_cast<double>(v[c]) * v[c];
        s->inv_norm[q] = ss > 0.0 ? static_cast<float>(1.0 / std::sqrt(ss)) : 0.f;
        const double u = (static_cast<double>(next_rand(s->rng_state) >> 11) + 1.0) / 9007199254740993.0;
        int lvl = static_cast<int>(-std::log(u) * mult);
        lvl = std::min(lvl, 30);
        s->level[q] = static_cast<uint8_t>(lvl);
        if (lvl > 0) {
            s->upper_off[q] = static_cast<uint32_t>(slots);
            slots += lvl;
        }
        s->index_of.emplace(ids[i], q);
    }
    s->upper.resize(slots * s->M, HNSW_NONE);
    if (s->locks_cap < total) {
        const size_t nc = total + total / 2 + 1024;
        s->locks.reset(new std::atomic<uint8_t>[nc]);
        for (size_t i = 0; i < nc; ++i) s->locks[i].store(0);
        s->locks_cap = nc;
    }
    s->live += n;
    s->dirty = true;
    if (first == 0) s->dirty_full = true;
    auto finish = [&](int builder) {
        if (n >= 1024) {
            s->last_builder = builder;
            s->last_build_us = static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::microseconds>(
                                                          std::chrono::steady_clock::now() - t_begin).count());
        }
        return 0;
    };

    // ---- bulk add into an empty graph: build on the device (hnsw_build.cu) ----
    int builder = s->builder;
    if (const char* e = std::getenv("VL_HNSW_BUILDER")) {
        if (!strcmp(e, "host")) builder = HNSW_BUILDER_HOST;
        else if (!strcmp(e, "device")) builder = HNSW_BUILDER_DEVICE;
    }
    if (first == 0 && d_rows && builder != HNSW_BUILDER_HOST && (builder == HNSW_BUILDER_DEVICE || n >= 4096)) {
        const int st = hnsw_build_device(s, d_rows, pitch, stream, launches);
        if (st) return st;
        return finish(HNSW_BUILDER_DEVICE);
    }

    // ---- insert: the first few nodes sequentially, the rest over all host threads ----
    Builder b(*s);
    unsigned nthreads = std::thread::hardware_concurrency();
    if (const char* e = std::getenv("VL_HNSW_BUILD_THREADS")) nthreads = std::max(1, atoi(e));
    nthreads = std::max(1u, std::min(nthreads, 64u));
    uint32_t start = first;
    {
        // sequential path: the visited stamps live in the state, so a single add costs O(visited), not O(total)
        Builder::Scratch sc;
        sc.stamp.swap(s->seq_stamp);
        sc.epoch = s->seq_epoch;
        if (sc.stamp.size() < total) sc.stamp.resize(total + total / 4 + 64, 0);
        struct Keep {
            HnswState* s; Builder::Scratch* sc;
            ~Keep() {
                sc->stamp.swap(s->seq_stamp);
                s->seq_epoch = sc->epoch;
                s->touched.insert(s->touched.end(), sc->touched.begin(), sc->touched.end());
            }
        } keep{s, &sc};
        const uint32_t seq_end = static_cast<uint32_t>(std::min<uint64_t>(total, std::max<uint32_t>(first, 256)));
        for (; start < seq_end; ++start) b.insert(start, sc);
        if (nthreads == 1 || total - start < 1024) {
            for (; start < total; ++start) b.insert(start, sc);
            return finish(HNSW_BUILDER_HOST);
        }
    }
    std::atomic<uint32_t> next(start);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nthreads; ++t)
        th.emplace_back([&]() {
            Builder::Scratch sc;
            sc.stamp.assign(total, 0);
            for (;;) {
                const uint32_t q = next.fetch_add(1);
                if (q >= total) break;
                b.insert(q, sc);
            }
        });
    for (auto& x : th) x.join();
    s->dirty_full = true;   // a parallel bulk insert rewrites a large part of the graph: upload it whole
    return finish(HNSW_BUILDER_HOST);
}

// structural audit of the host copy of the graph (all levels): see vl_hnsw_graph_check
void hnsw_graph_check(const HnswState* s, uint64_t out[6]) {
    const size_t n = s->level.size();
    uint64_t edges0 = 0, self = 0, dup = 0, invalid = 0, isolated = 0;
    for (size_t i = 0; i < n; ++i) {
        for (int l = 0; l <= s->level[i]; ++l) {
            const uint32_t cap = l == 0 ? s->M0 : s->M;
            const uint32_t* a = l == 0 ? s->adj0.data() + i * s->M0
                                       : s->upper.data() + (static_cast<size_t>(s->upper_off[i]) + l - 1) * s->M;
            uint32_t deg = 0;
            bool gap = false;
            for (uint32_t j = 0; j < cap; ++j) {
                const uint32_t v = a[j];
                if (v == HNSW_NONE) { gap = true; continue; }
                if (gap) ++invalid;                       // lists must be dense prefixes
                ++deg;
                if (v == i) ++self;
                if (v >= n || s->level[v] < l) { ++invalid; continue; }
                for (uint32_t j2 = 0; j2 < j; ++j2)
                    if (a[j2] == v) { ++dup; break; }
            }
            if (l == 0) {
                edges0 += deg;
                if (deg == 0 && n > 1) ++isolated;
            }
        }
    }
    out[0] = n; out[1] = edges0; out[2] = self; out[3] = dup; out[4] = invalid; out[5] = isolated;
}

bool hnsw_soft_delete(HnswState* s, uint64_t id) {  // hnsw.rs:400-414
    auto it = s->index_of.find(id);
    if (it == s->index_of.end()) return false;
    s->deleted[it->second] = 1;
    s->index_of.erase(it);
    s->live -= 1;
    s->deleted_dirty = true;
    return true;
}

uint64_t hnsw_export(const HnswState* s, uint64_t first, uint64_t cap, uint64_t* out_ids, float* out_rows) {
    uint64_t live_pos = 0, written = 0;
    const size_t n = s->level.size();
    for (size_t i = 0; i < n && written < cap; ++i) {
        if (s->deleted[i]) continue;
        if (live_pos++ < first) continue;
        if (out_ids) out_ids[written] = s->id_of[i];
        if (out_rows) std::memcpy(out_rows + written * s->dim, s->vecs.data() + i * s->dim, s->dim * sizeof(float));
        ++written;
    }
    return written;
}

// ---- graph persistence ------------------------------------------------------------------------------------
namespace {
struct GraphBlobHeader {
    char magic[8];          // "VLHNSWG1"
    uint32_t dim, metric, M, M0, efc;
    int32_t max_level;
    uint32_t entry, reserved;
    uint64_t n, upper_words, rng_state;
};
}  // namespace

size_t hnsw_graph_blob_bytes(const HnswState* s) {
    const size_t n = s->level.size();
    return sizeof(GraphBlobHeader) + n /*level*/ + n * 4 /*upper_off*/ + n * s->M0 * 4 + s->upper.size() * 4;
}

int hnsw_export_graph(const HnswState* s, void* buf, size_t cap, size_t* written) {
    const size_t n = s->level.size();
    if (s->live != n) return 9;   // soft-deleted nodes: their rows are not exported, the graph cannot be restored
    const size_t need = hnsw_graph_blob_bytes(s);
    if (written) *written = need;
    if (!buf || cap < need) return 5;
    GraphBlobHeader hd;
    std::memset(&hd, 0, sizeof hd);
    std::memcpy(hd.magic, "VLHNSWG1", 8);
    hd.dim = s->dim; hd.metric = static_cast<uint32_t>(s->metric); hd.M = s->M; hd.M0 = s->M0; hd.efc = s->efc;
    hd.max_level = s->max_level; hd.entry = s->entry; hd.n = n; hd.upper_words = s->upper.size();
    hd.rng_state = s->rng_state;
    char* p = static_cast<char*>(buf);
    std::memcpy(p, &hd, sizeof hd); p += sizeof hd;
    std::memcpy(p, s->level.data(), n); p += n;
    std::memcpy(p, s->upper_off.data(), n * 4); p += n * 4;
    std::memcpy(p, s->adj0.data(), n * s->M0 * 4); p += n * s->M0 * 4;
    if (!s->upper.empty()) std::memcpy(p, s->upper.data(), s->upper.size() * 4);
    return 0;
}

int hnsw_import_graph(HnswState* s, const uint64_t* ids, const float* rows, uint64_t n, const void* blob, size_t bytes) {
    if (!s->level.empty() || n == 0 || n >= 0x7FFFFFFFull) return 5;
    if (bytes < sizeof(GraphBlobHeader)) return 5;
    GraphBlobHeader hd;
    std::memcpy(&hd, blob, sizeof hd);
    if (std::memcmp(hd.magic, "VLHNSWG1", 8) != 0 || hd.dim != s->dim || hd.metric != static_cast<uint32_t>(s->metric) ||
        hd.M != s->M || hd.M0 != s->M0 || hd.n != n || hd.entry >= n || hd.max_level < 0 || hd.max_level > 30)
        return 5;
    const size_t need = sizeof(GraphBlobHeader) + n + n * 4 + n * s->M0 * 4 + hd.upper_words * 4;
    if (bytes != need || hd.upper_words % s->M != 0) return 5;
    const char* p = static_cast<const char*>(blob) + sizeof hd;
    s->level.assign(reinterpret_cast<const uint8_t*>(p), reinterpret_cast<const uint8_t*>(p) + n); p += n;
    s->upper_off.resize(n); std::memcpy(s->upper_off.data(), p, n * 4); p += n * 4;
    s->adj0.resize(n * s->M0); std::memcpy(s->adj0.data(), p, n * s->M0 * 4); p += n * s->M0 * 4;
    s->upper.resize(hd.upper_words);
    if (hd.upper_words) std::memcpy(s->upper.data(), p, hd.upper_words * 4);
    // structural validation before anything is trusted: slot ranges, then the audit every builder output passes
    bool ok = s->level[hd.entry] == hd.max_level;
    const size_t slots = hd.upper_words / s->M;
    for (size_t i = 0; i < n && ok; ++i) {
        if (s->level[i] > hd.max_level) ok = false;
        if (s->level[i] > 0) ok = ok && s->upper_off[i] != HNSW_NONE && static_cast<size_t>(s->upper_off[i]) + s->level[i] <= slots;
    }
    uint64_t chk[6] = {0, 0, 0, 0, 0, 0};
    if (ok) {
        hnsw_graph_check(s, chk);
        ok = chk[2] == 0 && chk[3] == 0 && chk[4] == 0;
    }
    if (!ok) {
        s->level.clear(); s->upper_off.clear(); s->adj0.clear(); s->upper.clear();
        return 5;
    }
    s->vecs.assign(rows, rows + n * s->dim);
    s->inv_norm.resize(n);
    s->deleted.assign(n, 0);
    s->id_of.assign(ids, ids + n);
    s->index_of.clear();
    s->index_of.reserve(n * 2);
    for (uint64_t i = 0; i < n; ++i) {
        double ss = 0.0;
        const float* v = rows + i * s->dim;
        for (uint32_t c = 0; c < s->dim; ++c) ss += static_cast<double>(v[c]) * v[c];
        s->inv_norm[i] = ss > 0.0 ? static_cast<float>(1.0 / std::sqrt(ss)) : 0.f;
        s->index_of.emplace(ids[i], static_cast<uint32_t>(i));
    }
    if (s->index_of.size() != n) {   // duplicate ids
        s->level.clear(); s->upper_off.clear(); s->adj0.clear(); s->upper.clear(); s->vecs.clear();
        s->inv_norm.clear(); s->deleted.clear(); s->id_of.clear(); s->index_of.clear();
        return 2;
    }
    const size_t nc = n + n / 2 + 1024;
    s->locks.reset(new std::atomic<uint8_t>[nc]);
    for (size_t i = 0; i < nc; ++i) s->locks[i].store(0);
    s->locks_cap = nc;
    s->live = n;
    s->max_level = hd.max_level;
    s->entry = hd.entry;
    s->rng_state = hd.rng_state;
    s->dirty = s->dirty_full = s->deleted_dirty = true;
    s->last_builder = 0;
    return 0;
}

// device graph arrays for n nodes (capacity tracked in nodes / upper slots; grown geometrically)
int hnsw_reserve_device(HnswState* s, size_t n) {
    if (n > s->d_n_cap) {
        const size_t nc = n + n / 4 + 64;
        cudaFree(s->d_adj0); cudaFree(s->d_upper_off); cudaFree(s->d_level); cudaFree(s->d_deleted);
        cudaFree(s->d_ids); cudaFree(s->d_inv_norm);
        s->d_adj0 = s->d_upper_off = nullptr; s->d_level = s->d_deleted = nullptr; s->d_ids = nullptr; s->d_inv_norm = nullptr;
        s->d_n_cap = 0;
        if (cudaMalloc(&s->d_adj0, nc * s->M0 * 4) != cudaSuccess || cudaMalloc(&s->d_upper_off, nc * 4) != cudaSuccess ||
            cudaMalloc(&s->d_level, nc) != cudaSuccess || cudaMalloc(&s->d_deleted, nc) != cudaSuccess ||
            cudaMalloc(&s->d_ids, nc * 8) != cudaSuccess || cudaMalloc(&s->d_inv_norm, nc * 4) != cudaSuccess)
            return 7;
        s->d_n_cap = nc;
    }
    if (!dev_grow(s->d_upper, s->d_upper_cap, std::max<size_t>(s->upper.size(), 1))) return 7;
    return 0;
}

// caller holds graph_mu exclusively (or is the single writer)
static int upload_locked(HnswState* s, cudaStream_t stream) {
    const size_t n = s->level.size();
    if (n == 0) return 0;
    if (s->dirty) {
        const bool fits = n <= s->d_n_cap && s->upper.size() <= s->d_upper_cap && s->d_adj0 && (s->d_upper || s->upper.empty());
        const bool partial = !s->dirty_full && fits && s->uploaded_n <= n && s->touched.size() <= 4096;
        if (partial) {
            // new nodes: contiguous tails of every per-node array and of the upper-level slots
            const size_t n0 = s->uploaded_n, m = n - n0;
            if (m) {
                cudaMemcpyAsync(s->d_adj0 + n0 * s->M0, s->adj0.data() + n0 * s->M0, m * s->M0 * 4, cudaMemcpyHostToDevice, stream);
                cudaMemcpyAsync(s->d_upper_off + n0, s->upper_off.data() + n0, m * 4, cudaMemcpyHostToDevice, stream);
                cudaMemcpyAsync(s->d_level + n0, s->level.data() + n0, m, cudaMemcpyHostToDevice, stream);
                cudaMemcpyAsync(s->d_ids + n0, s->id_of.data() + n0, m * 8, cudaMemcpyHostToDevice, stream);
                cudaMemcpyAsync(s->d_inv_norm + n0, s->inv_norm.data() + n0, m * 4, cudaMemcpyHostToDevice, stream);
                cudaMemcpyAsync(s->d_deleted + n0, s->deleted.data() + n0, m, cudaMemcpyHostToDevice, stream);
            }
            if (s->upper.size() > s->uploaded_upper)
                cudaMemcpyAsync(s->d_upper + s->uploaded_upper, s->upper.data() + s->uploaded_upper,
                                (s->upper.size() - s->uploaded_upper) * 4, cudaMemcpyHostToDevice, stream);
            // re-written adjacency rows of older nodes (back-links), each once
            std::sort(s->touched.begin(), s->touched.end());
            s->touched.erase(std::unique(s->touched.begin(), s->touched.end()), s->touched.end());
            for (const uint64_t t : s->touched) {
                const size_t node = static_cast<size_t>(t >> 8);
                const int lvl = static_cast<int>(t & 0xFF);
                if (node >= n0) continue;   // part of the tail copied above (its upper slots too)
                if (lvl == 0) {
                    cudaMemcpyAsync(s->d_adj0 + node * s->M0, s->adj0.data() + node * s->M0, s->M0 * 4, cudaMemcpyHostToDevice, stream);
                } else {
                    const size_t off = (static_cast<size_t>(s->upper_off[node]) + lvl - 1) * s->M;
                    if (off < s->uploaded_upper)
                        cudaMemcpyAsync(s->d_upper + off, s->upper.data() + off, s->M * 4, cudaMemcpyHostToDevice, stream);
                }
            }
        } else {
            if (int st = hnsw_reserve_device(s, n)) return st;
            cudaMemcpyAsync(s->d_adj0, s->adj0.data(), n * s->M0 * 4, cudaMemcpyHostToDevice, stream);
            cudaMemcpyAsync(s->d_upper_off, s->upper_off.data(), n * 4, cudaMemcpyHostToDevice, stream);
            cudaMemcpyAsync(s->d_level, s->level.data(), n, cudaMemcpyHostToDevice, stream);
            cudaMemcpyAsync(s->d_ids, s->id_of.data(), n * 8, cudaMemcpyHostToDevice, stream);
            cudaMemcpyAsync(s->d_inv_norm, s->inv_norm.data(), n * 4, cudaMemcpyHostToDevice, stream);
            if (!s->upper.empty())
                cudaMemcpyAsync(s->d_upper, s->upper.data(), s->upper.size() * 4, cudaMemcpyHostToDevice, stream);
            s->deleted_dirty = true;
        }
        s->touched.clear();
        s->uploaded_n = n;
        s->uploaded_upper = s->upper.size();
    }
    if (s->deleted_dirty)
        cudaMemcpyAsync(s->d_deleted, s->deleted.data(), n, cudaMemcpyHostToDevice, stream);
    if (cudaStreamSynchronize(stream) != cudaSuccess) return 6;
    s->dirty = s->deleted_dirty = s->dirty_full = false;
    return 0;
}

int hnsw_upload(HnswState* s, cudaStream_t stream) {
    std::unique_lock<std::shared_mutex> lk(s->graph_mu);
    return upload_locked(s, stream);
}

static HnswState::SearchCtx* acquire_ctx(HnswState* s) {
    std::unique_lock<std::mutex> lk(s->ctx_mu);
    for (;;) {
        for (auto& c : s->ctxs)
            if (!c->busy) { c->busy = true; return c.get(); }
        if (static_cast<int>(s->ctxs.size()) < HnswState::MAX_CTX) {
            auto c = std::make_unique<HnswState::SearchCtx>();
            if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
            if (cudaMalloc(&c->d_visited, 8) != cudaSuccess || cudaMallocHost(&c->h_visited, 8) != cudaSuccess) {
                cudaFree(c->d_visited);
                cudaStreamDestroy(c->stream);
                return nullptr;
            }
            c->busy = true;
            s->ctxs.push_back(std::move(c));
            return s->ctxs.back().get();
        }
        s->ctx_cv.wait(lk);
    }
}
static void release_ctx(HnswState* s, HnswState::SearchCtx* c) {
    { std::lock_guard<std::mutex> lk(s->ctx_mu); c->busy = false; }
    s->ctx_cv.notify_one();
}

int hnsw_search_host(HnswState* s, const float* d_rows, uint32_t pitch, const float* queries, uint32_t nq,
                     uint32_t k, uint32_t ef, uint64_t* out_ids, double* out_scores, uint32_t* out_counts,
                     uint64_t* visited, uint64_t* launches, const void* rows_bf16) {
    HnswState::SearchCtx* c = acquire_ctx(s);
    if (!c) return 6;
    struct Rel { HnswState* s; HnswState::SearchCtx* c; ~Rel() { release_ctx(s, c); } } rel{s, c};
    cudaStream_t stream = c->stream;
    // The first search after an add / delete uploads the graph — exclusively: hnsw_reserve_device may free and
    // reallocate the device arrays, which no concurrent reader may be using (this used to run outside
    // any lock).  Mutators are excluded from readers by the caller (RwLock, client.rs:245), so `dirty` can only be
    // set while no search is in flight; the flags are re-checked under the lock.
    if (s->dirty || s->deleted_dirty) {
        std::unique_lock<std::shared_mutex> up(s->graph_mu);
        if (s->dirty || s->deleted_dirty)
            if (int st = upload_locked(s, stream)) return st;
    }
    std::shared_lock<std::shared_mutex> rd(s->graph_mu);
    // hnsw.rs:437: ef = max_candidates = min(k, live); `ef` > 0 is the additive sweep knob
    const uint32_t maxc = static_cast<uint32_t>(std::min<uint64_t>(k, s->live));
    uint32_t ef_search = ef == 0 ? maxc : std::max(ef, maxc);
    const size_t qf = static_cast<size_t>(nq) * pitch;
    if (qf > c->q_cap) {
        cudaFree(c->d_q);
        c->d_q = nullptr; c->q_cap = 0;
        if (cudaMalloc(&c->d_q, qf * sizeof(float)) != cudaSuccess) return 7;
        c->q_cap = qf;
    }
    const size_t on = static_cast<size_t>(nq) * k;
    const size_t need = on * 16 + static_cast<size_t>(nq) * 8;   // ids | scores | counts | visited (per query)
    // a few queries: the kernel writes its (tiny) results straight into the pinned host block (UVA zero-copy), which
    // takes the memset and the two device-to-host copies off the latency path
    const bool zero_copy = nq <= 16;
    if (need > c->out_cap) {
        cudaFree(c->d_out); cudaFreeHost(c->h_out);
        c->d_out = nullptr; c->h_out = nullptr; c->out_cap = 0;
        if (cudaMalloc(&c->d_out, need) != cudaSuccess) return 7;
        if (cudaMallocHost(&c->h_out, need) != cudaSuccess) return 7;
        c->out_cap = need;
    }
    if (!zero_copy) cudaMemsetAsync(c->d_visited, 0, 8, stream);
    HnswDeviceGraph g;
    g.adj0 = s->d_adj0; g.upper_off = s->d_upper_off; g.upper = s->d_upper; g.level = s->d_level;
    g.deleted = s->d_deleted; g.ids = s->d_ids; g.inv_norm = s->d_inv_norm;
    g.n = static_cast<uint32_t>(s->level.size()); g.M = s->M; g.M0 = s->M0; g.entry = s->entry;
    g.max_level = s->max_level;
    unsigned char* ob = zero_copy ? c->h_out : c->d_out;
    uint64_t* d_ids = reinterpret_cast<uint64_t*>(ob);
    double* d_scores = reinterpret_cast<double*>(ob + on * 8);
    uint32_t* d_counts = reinterpret_cast<uint32_t*>(ob + on * 16);
    uint32_t* d_vis_q = zero_copy ? d_counts + nq : nullptr;
    // One launch for the whole batch: splitting it into wave-sized chunks (to overlap the query upload with the
    // search) was measured 25 % SLOWER — every launch pays its own tail of slow queries.
    if (pitch == s->dim) {
        cudaMemcpyAsync(c->d_q, queries, qf * sizeof(float), cudaMemcpyHostToDevice, stream);
    } else {
        cudaMemsetAsync(c->d_q, 0, qf * sizeof(float), stream);
        cudaMemcpy2DAsync(c->d_q, pitch * sizeof(float), queries, s->dim * sizeof(float), s->dim * sizeof(float),
                          nq, cudaMemcpyHostToDevice, stream);
    }
    int st = hnsw_launch_search(g, d_rows, pitch, s->dim, s->metric, c->d_q, nq, k, ef_search, d_ids, d_scores,
                                d_counts, c->d_visited, stream, s->score_mode, s->beam_mult, rows_bf16, d_vis_q);
    if (st) return st;
    if (!zero_copy) {
        cudaMemcpyAsync(c->h_out, c->d_out, need, cudaMemcpyDeviceToHost, stream);
        cudaMemcpyAsync(c->h_visited, c->d_visited, 8, cudaMemcpyDeviceToHost, stream);
    }
    if (cudaStreamSynchronize(stream) != cudaSuccess) return 6;
    if (zero_copy) {
        unsigned long long tot = 0;
        const uint32_t* vq = reinterpret_cast<const uint32_t*>(c->h_out + on * 16) + nq;
        for (uint32_t i = 0; i < nq; ++i) tot += vq[i];
        *c->h_visited = tot;
    }
    memcpy(out_ids, c->h_out, on * 8);
    memcpy(out_scores, c->h_out + on * 8, on * 8);
    memcpy(out_counts, c->h_out + on * 16, static_cast<size_t>(nq) * 4);
    *visited = *c->h_visited;
    *launches = 1;
    return 0;
}

}  // namespace vl
