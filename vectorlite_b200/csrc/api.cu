// api.cu — the C ABI (include/vectorlite_cuda.h): handle, device arena, workspace pool and the
// search orchestration for the flat index.  Host-side semantics follow the reference's
// FlatIndex (src/index/flat.rs:59-136) behind trait VectorIndex (src/lib.rs:224-245); the HNSW
// half of the handle lives in hnsw_host.cpp / hnsw_search.cu.
#include "../../include/vectorlite_cuda.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "batch.h"
#include "combiner.h"
#include "hnsw.h"
#include "kernels.h"
#include "tc_state.h"

using namespace vl;

// ---------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CU(x)                                                                                   \
    do {                                                                                        \
        cudaError_t _e = (x);                                                                   \
        if (_e != cudaSuccess)                                                                  \
            return fail(_e == cudaErrorMemoryAllocation ? VL_ERR_OOM : VL_ERR_CUDA, "%s: %s", #x, \
                        cudaGetErrorString(_e));                                                \
    } while (0)

enum { ST_LAUNCHES = 0, ST_FAST = 1, ST_EXACT = 2, ST_H2D = 3, ST_D2H = 4, ST_HNSW_VISITED = 5, ST_BF16_SCANS = 6, ST_COMBINED = 7,
       ST_BF16_RETRY = 8, ST_FP32_RETRY = 9, ST_BOOSTED = 10, ST_TENSOR = 11, ST_N = 12 };

namespace {

constexpr uint32_t NQ_CHUNK = 32;  // queries per launch of the per-query scan
constexpr uint32_t DEV_SETS = 4;   // control-block / early-threshold sets rotated by pipelined device searches
// nq >= batch_min → batched tile pipeline.  Measured at 1M x 384 (profiles/r01_small_batch.md): the tensor-core
// pipeline serves 2..128 queries in 0.19 ms, less than two single-query scans, so it takes over from nq = 2; the
// CUDA-core pipeline (manhattan, FP32 mode, wider rows) costs 3.7 ms per 128-query tile and only pays from 16.
// Where single-query scans read the bf16 mirror (0.12 ms each, manhattan included) the CUDA-core pipeline pays from 32.
static uint32_t batch_min_for(bool tensor_path, bool mirror_scans = false) {
    static const int forced = [] { const char* e = getenv("VL_BATCH_MIN"); return e ? atoi(e) : 0; }();
    if (forced >= 2) return static_cast<uint32_t>(forced);
    return tensor_path ? 2u : (mirror_scans ? 32u : 16u);
}
constexpr uint32_t BATCH_CHUNK = 1024;   // queries per batched pass
constexpr uint32_t BATCH_CAPQ = 4096;    // candidate slots per query

struct Boost {   // adaptive over-selection state of one search path (single-query scans / batches), see search_levels
    std::atomic<int> on{0};
    std::atomic<uint32_t> base_ok_run{0};
};

struct Slot {  // one in-flight search: stream + scratch, all sized on demand
    cudaStream_t stream = nullptr;
    // queries
    float* d_q = nullptr; float* h_q = nullptr; size_t q_cap = 0;  // floats
    // scan scratch
    uint64_t* cand = nullptr; size_t cand_cap = 0;
    uint32_t* cand_count = nullptr; size_t cc_cap = 0;
    uint64_t* cand_max = nullptr; size_t cm_cap = 0;
    QueryCtl* ctl = nullptr; size_t ctl_cap = 0;
    uint64_t* early = nullptr; size_t early_cap = 0;   // [2][NQ_CHUNK][EARLY_STRIDE], zero between searches
    // outputs: ONE device block + ONE pinned mirror per call, carved as
    // [ids m*k u64][scores m*k f64][counts m u32][flags m u32] so a single D2H brings everything
    unsigned char* d_out = nullptr; unsigned char* h_out = nullptr; size_t out_bytes = 0;
    uint64_t* d_ids = nullptr; double* d_scores = nullptr; uint32_t* d_counts = nullptr; uint32_t* d_flags = nullptr;
    uint64_t* h_ids = nullptr; double* h_scores = nullptr; uint32_t* h_counts = nullptr; uint32_t* h_flags = nullptr;
    size_t out_used = 0;
    // batched pipeline scratch; two sets so that pipelined device searches can chain batches (batch.h)
    struct BatchSet {
        uint64_t* cand = nullptr; size_t cand_cap = 0;
        uint32_t* count = nullptr; size_t count_cap = 0;
        float* tau = nullptr; size_t tau_cap = 0;
        uint32_t* qflags = nullptr; size_t qflags_cap = 0;
    } bset[2];
    float* b_gmax = nullptr; size_t b_gmax_cap = 0;
    uint32_t batch_parity = 0;
    // exact path
    double* d_exact = nullptr; size_t exact_cap = 0;
    uint32_t* d_exflags = nullptr;
    ExactScratch exs;
    bool busy = false;
};

}  // namespace

struct vl_index {
    int type = VL_INDEX_FLAT;
    int device = 0;
    uint32_t dim = 0, pitch = 0;
    int mode = VL_MODE_AUTO;
    uint64_t pos_base = 0;
    // ---- arena ----
    float* d_rows = nullptr;
    float* d_inv_norm = nullptr;
    uint64_t* d_ids = nullptr;   // only when !identity
    ArenaStats* d_stats = nullptr;
    uint64_t n = 0, cap = 0;
    // ---- host id bookkeeping ----
    bool identity = true;        // id == id_base + pos for every row (no map needed)
    uint64_t id_base = 0;
    bool have_max = false;
    uint64_t max_id = 0;
    std::vector<uint64_t> ids_host;
    std::unordered_map<uint64_t, uint32_t> id_to_pos;
    // ---- workspaces ----
    std::mutex mu;
    std::condition_variable cv;
    std::vector<std::unique_ptr<Slot>> slots;
    Slot dev_slot;               // used by vl_index_search_device (caller-ordered)
    cudaStream_t mut_stream = nullptr;
    int max_grid_x = 148 * 4;
    int max_grid_x_bf16 = 148 * 2;
    std::atomic<uint64_t> stats[ST_N];
    // ---- profiling (roofline reports) ----
    bool pipelined = false;      // vl_index_set_pipelined: PDL overlap between consecutive device searches
    uint32_t dev_parity = 0;     // rotation of the dev_slot control-block sets (DEV_SETS)
    bool profiling = false;
    std::vector<cudaEvent_t> prof_ev;   // pairs
    size_t prof_n = 0;                  // pairs recorded since last read
    // ---- tensor-core batched path ----
    TcState tc;
    std::mutex tc_mu;
    // ---- combiner: concurrent single-query callers are coalesced into one batched launch ----
    Combiner comb;
    // ---- adaptive over-selection (search_levels) ----
    Boost boost_single, boost_batch;
    // ---- hnsw ----
    HnswPtr hnsw;
    int hnsw_metric = -1;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <typename T>
int grow_dev(T*& p, size_t& cap, size_t need) {
    if (need <= cap) return VL_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    CU(cudaMalloc(&p, need * sizeof(T)));
    cap = need;
    return VL_OK;
}

static int reserve_early(Slot& s, cudaStream_t stream) {
    const size_t need = static_cast<size_t>(DEV_SETS) * NQ_CHUNK * EARLY_STRIDE;
    if (s.early_cap >= need) return VL_OK;
    int st = grow_dev(s.early, s.early_cap, need);
    if (st) return st;
    CU(cudaMemsetAsync(s.early, 0, s.early_cap * sizeof(uint64_t), stream));
    return VL_OK;
}

int slot_reserve(vl_index* h, Slot& s, uint32_t nq, uint32_t k, int Kp, int grid_x) {
    int st;
    const size_t qf = static_cast<size_t>(nq) * h->pitch;
    if (qf > s.q_cap) {
        if (s.d_q) cudaFree(s.d_q);
        if (s.h_q) cudaFreeHost(s.h_q);
        s.d_q = nullptr; s.h_q = nullptr; s.q_cap = 0;
        CU(cudaMalloc(&s.d_q, qf * sizeof(float)));
        CU(cudaMallocHost(&s.h_q, qf * sizeof(float)));
        s.q_cap = qf;
    }
    if ((st = grow_dev(s.cand, s.cand_cap, static_cast<size_t>(nq) * grid_x * Kp))) return st;
    if ((st = grow_dev(s.cand_count, s.cc_cap, static_cast<size_t>(nq) * grid_x))) return st;
    if ((st = grow_dev(s.cand_max, s.cm_cap, static_cast<size_t>(nq) * grid_x))) return st;
    if (nq > s.ctl_cap) {
        if ((st = grow_dev(s.ctl, s.ctl_cap, nq))) return st;
        CU(cudaMemsetAsync(s.ctl, 0, nq * sizeof(QueryCtl), s.stream));
    }
    if ((st = reserve_early(s, s.stream))) return st;
    const size_t on = static_cast<size_t>(nq) * std::max<uint32_t>(k, 1);
    const size_t need = on * 16 + static_cast<size_t>(nq) * 8;
    if (need > s.out_bytes) {
        cudaFree(s.d_out); cudaFreeHost(s.h_out);
        s.d_out = nullptr; s.h_out = nullptr; s.out_bytes = 0;
        CU(cudaMalloc(&s.d_out, need));
        CU(cudaMallocHost(&s.h_out, need));
        s.out_bytes = need;
    }
    s.d_ids = reinterpret_cast<uint64_t*>(s.d_out);
    s.d_scores = reinterpret_cast<double*>(s.d_out + on * 8);
    s.d_counts = reinterpret_cast<uint32_t*>(s.d_out + on * 16);
    s.d_flags = s.d_counts + nq;
    s.h_ids = reinterpret_cast<uint64_t*>(s.h_out);
    s.h_scores = reinterpret_cast<double*>(s.h_out + on * 8);
    s.h_counts = reinterpret_cast<uint32_t*>(s.h_out + on * 16);
    s.h_flags = s.h_counts + nq;
    s.out_used = need;
    if (!s.d_exflags) CU(cudaMalloc(&s.d_exflags, 4));
    return VL_OK;
}

int slot_reserve_batch(Slot& s, uint32_t nq, BatchWork* w, uint32_t set = 0) {
    int st;
    Slot::BatchSet& b = s.bset[set & 1u];
    if ((st = grow_dev(b.cand, b.cand_cap, static_cast<size_t>(nq) * BATCH_CAPQ))) return st;
    if ((st = grow_dev(b.count, b.count_cap, nq))) return st;
    if ((st = grow_dev(b.tau, b.tau_cap, nq))) return st;
    if ((st = grow_dev(b.qflags, b.qflags_cap, nq))) return st;
    if ((st = grow_dev(s.b_gmax, s.b_gmax_cap, static_cast<size_t>(nq) * GMAX_STRIDE))) return st;
    w->cand = b.cand; w->count = b.count; w->tau = b.tau; w->qflags = b.qflags; w->capq = BATCH_CAPQ;
    w->gmax = s.b_gmax;
    return VL_OK;
}

void slot_free(Slot& s) {
    cudaFree(s.d_q); cudaFreeHost(s.h_q); cudaFree(s.cand); cudaFree(s.cand_count); cudaFree(s.cand_max); cudaFree(s.ctl); cudaFree(s.early);
    cudaFree(s.d_out); cudaFreeHost(s.h_out);
    for (auto& b : s.bset) { cudaFree(b.cand); cudaFree(b.count); cudaFree(b.tau); cudaFree(b.qflags); }
    cudaFree(s.b_gmax);
    cudaFree(s.d_exact); cudaFree(s.d_exflags);
    exact_scratch_free(s.exs);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = Slot();
}

Slot* acquire_slot(vl_index* h) {
    std::unique_lock<std::mutex> lk(h->mu);
    for (;;) {
        for (auto& s : h->slots)
            if (!s->busy) {
                s->busy = true;
                return s.get();
            }
        if (h->slots.size() < 8) {
            auto s = std::make_unique<Slot>();
            if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
            s->busy = true;
            h->slots.push_back(std::move(s));
            return h->slots.back().get();
        }
        h->cv.wait(lk);
    }
}
void release_slot(vl_index* h, Slot* s) {
    {
        std::lock_guard<std::mutex> lk(h->mu);
        s->busy = false;
    }
    h->cv.notify_one();
}

int pick_kp(uint32_t k) {  // over-select size K'
    uint32_t kp = k + std::max<uint32_t>(k, 32);
    kp = (kp + 31) / 32 * 32;
    return static_cast<int>(std::max<uint32_t>(kp, 64));
}

FlatView view_of(const vl_index* h) {
    FlatView v;
    v.rows = h->d_rows;
    v.inv_norm = h->d_inv_norm;
    v.ids = h->identity ? nullptr : h->d_ids;
    v.stats = h->d_stats;
    v.id_base = h->id_base;
    v.pos_base = h->pos_base;
    v.n = static_cast<uint32_t>(h->n);
    v.dim = h->dim;
    v.pitch = h->pitch;
    return v;
}

int arena_reserve(vl_index* h, uint64_t need) {
    if (need <= h->cap) return VL_OK;
    if (need >= 0xFFFFFFFEull) return fail(VL_ERR_UNSUPPORTED, "a shard holds at most 2^32-2 rows");
    uint64_t nc = std::max<uint64_t>(need, h->cap + h->cap / 2);
    nc = std::max<uint64_t>(nc, 1024);
    float* rows = nullptr; float* inv = nullptr; uint64_t* ids = nullptr;
    CU(cudaMalloc(&rows, nc * h->pitch * sizeof(float)));
    cudaError_t e = cudaMalloc(&inv, nc * sizeof(float));
    if (e == cudaSuccess && !h->identity) e = cudaMalloc(&ids, nc * sizeof(uint64_t));
    if (e != cudaSuccess) {
        cudaFree(rows); cudaFree(inv); cudaFree(ids);
        return fail(VL_ERR_OOM, "arena grow: %s", cudaGetErrorString(e));
    }
    if (h->n) {
        CU(cudaMemcpyAsync(rows, h->d_rows, h->n * h->pitch * sizeof(float), cudaMemcpyDeviceToDevice, h->mut_stream));
        CU(cudaMemcpyAsync(inv, h->d_inv_norm, h->n * sizeof(float), cudaMemcpyDeviceToDevice, h->mut_stream));
        if (ids) CU(cudaMemcpyAsync(ids, h->d_ids, h->n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, h->mut_stream));
        CU(cudaStreamSynchronize(h->mut_stream));
    }
    cudaFree(h->d_rows); cudaFree(h->d_inv_norm); cudaFree(h->d_ids);
    h->d_rows = rows; h->d_inv_norm = inv; h->d_ids = ids;
    h->cap = nc;
    return VL_OK;
}

// leave identity mode: ids become an explicit array + hash map
int materialize_ids(vl_index* h) {
    if (!h->identity) return VL_OK;
    h->ids_host.resize(h->n);
    h->id_to_pos.reserve(h->n * 2 + 16);
    for (uint64_t i = 0; i < h->n; ++i) {
        h->ids_host[i] = h->id_base + i;
        h->id_to_pos.emplace(h->id_base + i, static_cast<uint32_t>(i));
    }
    uint64_t* ids = nullptr;
    if (h->cap) {
        CU(cudaMalloc(&ids, h->cap * sizeof(uint64_t)));
        if (h->n) CU(cudaMemcpy(ids, h->ids_host.data(), h->n * 8, cudaMemcpyHostToDevice));
    }
    h->d_ids = ids;
    h->identity = false;
    return VL_OK;
}

bool find_pos(const vl_index* h, uint64_t id, uint32_t* pos) {
    if (h->identity) {
        if (h->n == 0 || id < h->id_base || id - h->id_base >= h->n) return false;
        *pos = static_cast<uint32_t>(id - h->id_base);
        return true;
    }
    auto it = h->id_to_pos.find(id);
    if (it == h->id_to_pos.end()) return false;
    *pos = it->second;
    return true;
}

int flat_add_batch(vl_index* h, const uint64_t* ids, const float* rows, uint64_t n) {
    if (n == 0) return VL_OK;
    // ---- id checks first: all-or-nothing (flat.rs:86-88) ----
    bool seq = true;
    if (h->identity) {
        const uint64_t base = h->n ? h->id_base + h->n : ids[0];
        for (uint64_t i = 0; i < n && seq; ++i) seq = ids[i] == base + i;
    }
    if (h->identity && !seq) {
        int st = materialize_ids(h);
        if (st) return st;
    }
    if (!h->identity) {
        uint64_t i = 0;
        for (; i < n; ++i) {
            auto r = h->id_to_pos.emplace(ids[i], static_cast<uint32_t>(h->n + i));
            if (!r.second) break;
        }
        if (i < n) {
            const uint64_t dup = ids[i];
            for (uint64_t j = 0; j < i; ++j) h->id_to_pos.erase(ids[j]);
            return fail(VL_ERR_DUP_ID, "Vector ID %llu already exists", static_cast<unsigned long long>(dup));
        }
    }
    int st = arena_reserve(h, h->n + n);
    if (st) {
        if (!h->identity) for (uint64_t j = 0; j < n; ++j) h->id_to_pos.erase(ids[j]);
        return st;
    }
    if (h->identity && h->n == 0) h->id_base = ids[0];
    // ---- upload rows (pad to pitch) ----
    float* dst = h->d_rows + h->n * h->pitch;
    if (h->pitch == h->dim) {
        CU(cudaMemcpyAsync(dst, rows, n * h->dim * sizeof(float), cudaMemcpyHostToDevice, h->mut_stream));
    } else {
        CU(cudaMemsetAsync(dst, 0, n * h->pitch * sizeof(float), h->mut_stream));
        CU(cudaMemcpy2DAsync(dst, h->pitch * sizeof(float), rows, h->dim * sizeof(float),
                             h->dim * sizeof(float), n, cudaMemcpyHostToDevice, h->mut_stream));
    }
    if (!h->identity) {
        CU(cudaMemcpyAsync(h->d_ids + h->n, ids, n * 8, cudaMemcpyHostToDevice, h->mut_stream));
        h->ids_host.insert(h->ids_host.end(), ids, ids + n);
    }
    CU(launch_row_norms(h->d_rows, h->n, n, h->dim, h->pitch, h->d_inv_norm, h->d_stats, h->mut_stream));
    CU(cudaStreamSynchronize(h->mut_stream));
    h->stats[ST_LAUNCHES] += 1;
    h->stats[ST_H2D] += n * h->dim * sizeof(float);
    for (uint64_t i = 0; i < n; ++i)
        if (!h->have_max || ids[i] > h->max_id) { h->max_id = ids[i]; h->have_max = true; }
    h->n += n;
    return VL_OK;
}

// order-preserving removal of one storage position (flat.rs:94 `retain` keeps order)
int flat_remove_pos(vl_index* h, uint32_t pos) {
    const uint64_t tail = h->n - pos - 1;
    // the bf16 mirrors are positional too: shift them with the rows (rebuilding a mirror costs a pass over the
    // whole arena — 4.5 ms on a 12.5M-row shard); a mirror that is only partly built is simply dropped
    const bool keep_norm = h->tc.rows_norm && h->tc.built_norm == h->n;
    const bool keep_raw = h->tc.rows_raw && h->tc.sq_norm && h->tc.built_raw == h->n;
    if (tail) {
        const size_t chunk_bytes = 64ull << 20;
        char* tmp = nullptr;
        const size_t row_bytes = h->pitch * sizeof(float);
        CU(cudaMalloc(&tmp, std::min<uint64_t>(chunk_bytes, tail * row_bytes)));
        // overlapping device-to-device moves are bounced through tmp, chunk by chunk, front to back
        auto shift = [&](void* base, size_t elem_bytes) {
            if (!base) return;
            char* b = static_cast<char*>(base);
            const uint64_t per = std::max<uint64_t>(1, std::min<uint64_t>(chunk_bytes, tail * row_bytes) / elem_bytes);
            for (uint64_t off = 0; off < tail; off += per) {
                const uint64_t m = std::min(per, tail - off);
                cudaMemcpyAsync(tmp, b + (pos + 1 + off) * elem_bytes, m * elem_bytes, cudaMemcpyDeviceToDevice, h->mut_stream);
                cudaMemcpyAsync(b + (pos + off) * elem_bytes, tmp, m * elem_bytes, cudaMemcpyDeviceToDevice, h->mut_stream);
            }
        };
        shift(h->d_rows, row_bytes);
        shift(h->d_inv_norm, sizeof(float));
        shift(h->d_ids, sizeof(uint64_t));
        if (keep_norm) shift(h->tc.rows_norm, static_cast<size_t>(h->tc.KP) * 2);
        if (keep_raw) {
            shift(h->tc.rows_raw, static_cast<size_t>(h->tc.KP) * 2);
            shift(h->tc.sq_norm, sizeof(float));
        }
        cudaError_t e = cudaStreamSynchronize(h->mut_stream);
        cudaFree(tmp);
        CU(e);
    }
    h->tc.built_norm = keep_norm ? h->n - 1 : 0;
    h->tc.built_raw = keep_raw ? h->n - 1 : 0;
    h->tc.maps_n = 0;   // the TMA map encodes the row count
    const uint64_t id = h->ids_host[pos];
    h->id_to_pos.erase(id);
    h->ids_host.erase(h->ids_host.begin() + pos);
    for (uint64_t i = pos; i < h->ids_host.size(); ++i) h->id_to_pos[h->ids_host[i]] = static_cast<uint32_t>(i);
    h->n -= 1;
    if (h->have_max && id == h->max_id) {  // max_id() recomputes (flat.rs:76-78)
        h->have_max = !h->ids_host.empty();
        if (h->have_max) h->max_id = *std::max_element(h->ids_host.begin(), h->ids_host.end());
    }
    return VL_OK;
}

// exact path for one query already resident at d_query; writes slot outputs at q_index
int run_exact_one(vl_index* h, Slot& s, const FlatView& v, const float* d_query, int metric, uint32_t k,
                  uint32_t q_index, cudaStream_t stream) {
    int st = grow_dev(s.d_exact, s.exact_cap, static_cast<size_t>(v.n));
    if (st) return st;
    SearchOut out{s.d_ids, s.d_scores, nullptr, s.d_counts, s.d_flags};
    CU(cudaMemsetAsync(s.d_exflags, 0, 4, stream));
    CU(launch_exact_scores(v, d_query, metric, s.d_exact, s.d_exflags, stream));
    CU(exact_select(v, s.d_exact, k, s.exs, out, q_index, stream));
    CU(cudaMemcpyAsync(s.d_flags + q_index, s.d_exflags, 4, cudaMemcpyDeviceToDevice, stream));
    h->stats[ST_LAUNCHES] += 4;
    h->stats[ST_EXACT] += 1;
    return VL_OK;
}

// Single-query scans read the bf16 mirror of the rows when one applies (AUTO mode, 128/256/384/768/1024/1536-d, all four metrics):
// half the HBM bytes per query; the candidates are re-scored in f64 and certified with the bf16 bound exactly
// as in the tensor-core batched path.  Brings the mirror up to date (lazily, like tc_prepare for batches) and
// returns its pointers; `*mirror` stays null when the fp32 scan has to be used.
static bool mirror_scans_apply(const vl_index* h) {
    static const bool disabled = getenv("VL_DISABLE_BF16_SCAN") != nullptr;
    return !disabled && h->mode == VL_MODE_AUTO && flat_scan_bf16_supports(h->pitch) && h->pitch == h->dim;
}

static int single_query_mirror(vl_index* h, const FlatView& v, int metric, cudaStream_t stream, const void** mirror,
                               const float** sq_norm) {
    *mirror = nullptr;
    *sq_norm = nullptr;
    if (!mirror_scans_apply(h)) return VL_OK;
    std::lock_guard<std::mutex> lk(h->tc_mu);
    const bool cosine = metric == VL_METRIC_COSINE;
    const uint64_t before = cosine ? h->tc.built_norm : h->tc.built_raw;
    CU(tc_prepare(&h->tc, v, h->cap, metric, 1, stream));
    if (!h->tc.usable || h->tc.KP != v.pitch) return VL_OK;
    const uint64_t after = cosine ? h->tc.built_norm : h->tc.built_raw;
    if (after != before) CU(cudaStreamSynchronize(stream));   // other streams may scan the mirror next
    *mirror = cosine ? h->tc.rows_norm : h->tc.rows_raw;
    *sq_norm = h->tc.sq_norm;
    return VL_OK;
}

// scan + finalize of up to NQ_CHUNK single queries (bf16 mirror when allowed and available, else the fp32 arena)
static int launch_single_queries(vl_index* h, const FlatView& v, const float* dq, uint32_t m, uint32_t k, int metric,
                                 const ScanWork& w, const SearchOut& out, bool pipelined, bool allow_bf16,
                                 cudaStream_t stream, bool* used_bf16, int kp_base = 0) {
    const void* mirror = nullptr;
    const float* sqn = nullptr;
    if (allow_bf16) {
        int st = single_query_mirror(h, v, metric, stream, &mirror, &sqn);
        if (st) return st;
    }
    ScanWork ws = w;
    if (mirror) {   // persistent grid of the bf16 kernel: one resident wave (scan and finalize must agree on it)
        const uint32_t tiles = (v.n + 127) / 128;
        ws.grid_x = static_cast<int>(std::min<uint32_t>(std::min<uint32_t>(tiles, static_cast<uint32_t>(h->max_grid_x_bf16)),
                                                         static_cast<uint32_t>(w.grid_x)));
        if (ws.grid_x < 1) ws.grid_x = 1;
    }
    const bool prof = h->profiling && h->prof_n < h->prof_ev.size() / 2;
    if (prof) CU(cudaEventRecord(h->prof_ev[2 * h->prof_n], stream));
    if (mirror) CU(launch_flat_scan_bf16(v, mirror, sqn, dq, m, metric, ws, pipelined, stream));
    else CU(launch_flat_scan(v, dq, m, metric, ws, pipelined, stream));
    if (prof) {
        CU(cudaEventRecord(h->prof_ev[2 * h->prof_n + 1], stream));
        h->prof_n += 1;
    }
    CertAux aux;
    aux.kp_base = kp_base;
    if (mirror) {
        // mirror scans round ONE operand (the rows; the query stays fp32): worst case 2^-8 + pitch·2^-24 < 0.0040,
        // or the rounding error measured while the mirror was built, whichever is smaller (rescore.cuh)
        aux.tc_abs = 0.0040;
        if (h->tc.ex_bits) {
            aux.e_x = h->tc.ex_bits + (metric == VL_METRIC_COSINE ? 0 : 1);
            aux.e_x1 = h->tc.ex_bits + 2;
        }
    }
    CU(launch_flat_finalize(v, dq, m, k, metric, ws, out, 1.0f, stream, aux));
    h->stats[ST_LAUNCHES] += 2;
    if (mirror) h->stats[ST_BF16_SCANS] += m;
    if (used_bf16) *used_bf16 = mirror != nullptr;
    return VL_OK;
}

// ---- over-selection levels of a host search --------------------------------------------------------------
// A search over-selects K' candidates by approximate score, re-scores them in f64 and certifies that no excluded
// row can reach the top-k (rescore.cuh).  When the certificate of a query does not hold, ONLY that query moves on:
//   level 0  approximate scan at the handle's current over-selection (bf16 mirror / tensor cores where they
//            apply, else the fp32 arena): K' = pick_kp(k), or kp_big(K') while the handle is boosted;
//   level 1  the same scan at kp_big(K') — more candidates widen the gap between the k-th exact score and the
//            worst kept approximate score (clustered data: top-10 / top-64 cosine gap 0.003, top-10 / top-256 0.006);
//   level 2  fp32 arena at kp_big(K') (error bound ~2^-24·pitch instead of the bf16 one);
//   level 3  exact path: every row in f64, stable radix sort (heavy ties: all-equal mock embeddings).
// Failing queries are compacted and re-run TOGETHER (as a batch when there are enough of them).  The handle
// remembers tight data: if a quarter of a level-0 pass fails, later searches start at kp_big directly (one pass
// instead of two); the finalize kernel reports whether the base K' alone would have certified (FLAG_BASE_OK), and
// after 256 consecutive such queries the boost is dropped again.
static int kp_big(int kp) { return std::min<int>(KP_MAX, std::max(4 * kp, 256)); }

struct PassCfg {
    bool use_bf16 = true;   // bf16 mirror / tensor cores allowed
    bool exact = false;
    int Kp = 64;
    int kp_base = 0;
};

static bool pass_is_batched(const vl_index* h, uint32_t m, int metric, bool use_bf16, bool exact) {
    const bool tc_ok = use_bf16 && h->mode == VL_MODE_AUTO && metric != VL_METRIC_MANHATTAN && !getenv("VL_DISABLE_TC");
    return !exact && m >= batch_min_for(tc_ok && h->dim <= TC_MAX_DIM, use_bf16 && mirror_scans_apply(h));
}

// one pass over m staged host queries; leaves ids / scores / counts / flags of the m queries in the slot's pinned
// mirrors (s.h_*).  `batched_out` reports which pipeline ran.
static int run_pass(vl_index* h, Slot& s, const FlatView& v, const float* queries, uint32_t m, uint32_t qdim, uint32_t k,
                    int metric, const PassCfg& cfg, bool* bf16_used) {
    const uint32_t tiles = (v.n + SCAN_TILE_ROWS - 1) / SCAN_TILE_ROWS;
    const int grid_x = static_cast<int>(std::min<uint32_t>(tiles, h->max_grid_x));
    const bool tc_ok = cfg.use_bf16 && h->mode == VL_MODE_AUTO && metric != VL_METRIC_MANHATTAN && !getenv("VL_DISABLE_TC");
    const bool batched = pass_is_batched(h, m, metric, cfg.use_bf16, cfg.exact);
    *bf16_used = false;
    int st = slot_reserve(h, s, m, k, cfg.Kp, batched ? 1 : grid_x);
    if (st) return st;
    for (uint32_t q = 0; q < m; ++q) {
        float* d = s.h_q + static_cast<size_t>(q) * h->pitch;
        memcpy(d, queries + static_cast<size_t>(q) * qdim, h->dim * sizeof(float));
        for (uint32_t c = h->dim; c < h->pitch; ++c) d[c] = 0.f;
    }
    const size_t qbytes = static_cast<size_t>(m) * h->pitch * sizeof(float);
    CU(cudaMemcpyAsync(s.d_q, s.h_q, qbytes, cudaMemcpyHostToDevice, s.stream));
    h->stats[ST_H2D] += qbytes;
    SearchOut out{s.d_ids, s.d_scores, nullptr, s.d_counts, s.d_flags};
    // small batches (the combiner's cohorts of concurrent single-query callers): the rescore kernel writes its
    // results straight into the pinned host mirror (UVA zero-copy) — no D2H copy operation, one synchronize
    static const bool no_zero_copy = getenv("VL_DISABLE_ZERO_COPY") != nullptr;
    const bool zero_copy = batched && !cfg.exact && m <= 128 && !no_zero_copy;
    if (zero_copy) out = SearchOut{s.h_ids, s.h_scores, nullptr, s.h_counts, s.h_flags};
    if (cfg.exact) {
        for (uint32_t q = 0; q < m; ++q) {
            st = run_exact_one(h, s, v, s.d_q + static_cast<size_t>(q) * h->pitch, metric, k, q, s.stream);
            if (st) return st;
        }
    } else if (batched) {
        BatchWork bw;
        if ((st = slot_reserve_batch(s, m, &bw))) return st;
        uint64_t nl = 0;
        if (tc_ok) {
            std::lock_guard<std::mutex> lk(h->tc_mu);  // shared bf16 query staging + maps
            CU(tc_prepare(&h->tc, v, h->cap, metric, m, s.stream));
            BatchTensor bt;
            bt.usable = h->tc.usable;
            bt.scratch = &h->tc;
            *bf16_used = bt.usable;
            if (bt.usable) h->stats[ST_TENSOR] += m;
            CU(launch_batch_flat(v, s.d_q, m, k, metric, cfg.Kp, bw, out, &bt, &nl, s.stream, cfg.kp_base));
            CU(cudaStreamSynchronize(s.stream));
        } else {
            CU(launch_batch_flat(v, s.d_q, m, k, metric, cfg.Kp, bw, out, nullptr, &nl, s.stream, cfg.kp_base));
            if (zero_copy) CU(cudaStreamSynchronize(s.stream));
        }
        h->stats[ST_LAUNCHES] += nl;
        if (zero_copy) {
            h->stats[ST_D2H] += static_cast<size_t>(m) * (k * 16 + 8);
            return VL_OK;
        }
    } else {
        // few queries: the finalize kernel writes its (tiny) results straight into the pinned host mirror (UVA
        // zero-copy), which saves the D2H copy operation on the latency path
        SearchOut hout{s.h_ids, s.h_scores, nullptr, s.h_counts, s.h_flags};
        ScanWork w{s.cand, s.cand_count, s.cand_max, s.ctl, grid_x, cfg.Kp};
        w.early = m <= NQ_CHUNK ? s.early : nullptr;
        if ((st = launch_single_queries(h, v, s.d_q, m, k, metric, w, hout, false, cfg.use_bf16, s.stream, bf16_used,
                                        cfg.kp_base)))
            return st;
        CU(cudaStreamSynchronize(s.stream));
        h->stats[ST_D2H] += static_cast<size_t>(m) * (k * 16 + 8);
        return VL_OK;
    }
    CU(cudaMemcpyAsync(s.h_out, s.d_out, s.out_used, cudaMemcpyDeviceToHost, s.stream));
    CU(cudaStreamSynchronize(s.stream));
    h->stats[ST_D2H] += static_cast<size_t>(m) * (k * 16 + 8);
    return VL_OK;
}

// results of m host queries (contiguous [m][qdim]) at `level` and, for the queries whose certificate fails, the
// levels after it.  *nan is set when a NaN similarity was seen (the reference panics, flat.rs:116).
static int search_levels(vl_index* h, Slot& s, const FlatView& v, const float* queries, uint32_t m, uint32_t qdim,
                         uint32_t k, int metric, int level, uint64_t* out_ids, double* out_scores, uint32_t* out_counts,
                         bool* nan) {
    const int Kp0 = pick_kp(k), Kp1 = kp_big(Kp0);
    const bool fast = h->mode != VL_MODE_EXACT && k <= 256;
    if (!fast) level = 3;
    const bool bf16_possible = h->mode == VL_MODE_AUTO;
    PassCfg cfg;
    Boost* boost = nullptr;
    if (level == 0) {
        boost = pass_is_batched(h, m, metric, bf16_possible, false) ? &h->boost_batch : &h->boost_single;
        const bool boosted = boost->on.load(std::memory_order_relaxed) != 0 && Kp1 > Kp0;
        cfg.use_bf16 = bf16_possible;
        cfg.Kp = boosted ? Kp1 : Kp0;
        cfg.kp_base = boosted ? Kp0 : 0;
        if (boosted) h->stats[ST_BOOSTED] += m;
    } else if (level == 1) {
        cfg.use_bf16 = bf16_possible;
        cfg.Kp = Kp1;
    } else if (level == 2) {
        cfg.use_bf16 = false;
        cfg.Kp = Kp1;
    } else {
        cfg.exact = true;
        cfg.Kp = Kp0;
    }
    bool bf16_used = false;
    int st = run_pass(h, s, v, queries, m, qdim, k, metric, cfg, &bf16_used);
    if (st) return st;
    if (level > 0 && !cfg.exact && !bf16_used) h->stats[ST_FP32_RETRY] += m;
    std::vector<uint32_t> failed;
    uint32_t base_ok = 0;
    for (uint32_t q = 0; q < m; ++q) {
        const uint32_t f = s.h_flags[q];
        if (f & FLAG_NAN) *nan = true;
        if (!cfg.exact && (f & FLAG_CERT_FAIL)) {
            failed.push_back(q);
            continue;
        }
        if (!cfg.exact) h->stats[ST_FAST] += 1;
        base_ok += (f & FLAG_BASE_OK) != 0;
        memcpy(out_ids + static_cast<size_t>(q) * k, s.h_ids + static_cast<size_t>(q) * k, k * 8);
        memcpy(out_scores + static_cast<size_t>(q) * k, s.h_scores + static_cast<size_t>(q) * k, k * 8);
        out_counts[q] = s.h_counts[q];
    }
    if (boost && Kp1 > Kp0) {
        if (cfg.kp_base) {   // boosted pass: drop the boost after 256 consecutive queries the base K' would have served
            if (base_ok == m) {
                if (boost->base_ok_run.fetch_add(m) + m >= 256) { boost->on = 0; boost->base_ok_run = 0; }
            } else {
                boost->base_ok_run = 0;
            }
        } else if (failed.size() * 4 >= m) {   // tight data: start the next searches at the larger K'
            boost->on = 1;
            boost->base_ok_run = 0;
        }
    }
    if (failed.empty()) return VL_OK;
    // next level for the failing queries only
    const int next = bf16_used ? (cfg.Kp < Kp1 ? 1 : 2) : (cfg.Kp < Kp1 ? 2 : 3);
    if (bf16_used) h->stats[ST_BF16_RETRY] += failed.size();
    const uint32_t mf = static_cast<uint32_t>(failed.size());
    std::vector<float> fq(static_cast<size_t>(mf) * qdim);
    std::vector<uint64_t> fi(static_cast<size_t>(mf) * k);
    std::vector<double> fs(static_cast<size_t>(mf) * k);
    std::vector<uint32_t> fc(mf);
    for (uint32_t i = 0; i < mf; ++i)
        memcpy(fq.data() + static_cast<size_t>(i) * qdim, queries + static_cast<size_t>(failed[i]) * qdim, qdim * sizeof(float));
    st = search_levels(h, s, v, fq.data(), mf, qdim, k, metric, next, fi.data(), fs.data(), fc.data(), nan);
    if (st) return st;
    for (uint32_t i = 0; i < mf; ++i) {
        memcpy(out_ids + static_cast<size_t>(failed[i]) * k, fi.data() + static_cast<size_t>(i) * k, k * 8);
        memcpy(out_scores + static_cast<size_t>(failed[i]) * k, fs.data() + static_cast<size_t>(i) * k, k * 8);
        out_counts[failed[i]] = fc[i];
    }
    return VL_OK;
}

static int flat_search_impl(vl_index* h, const float* queries, uint32_t nq, uint32_t qdim, uint32_t k, int metric,
                            uint64_t* out_ids, double* out_scores, uint32_t* out_counts) {
    // flat.rs:99-104: the dimension is only checked when the index is non-empty
    if (h->n != 0 && qdim != h->dim)
        return fail(VL_ERR_DIM, "Dimension mismatch: expected %u, got %u", h->dim, qdim);
    for (size_t i = 0; i < static_cast<size_t>(nq) * k; ++i) { out_ids[i] = ~0ull; out_scores[i] = 0.0; }
    for (uint32_t q = 0; q < nq; ++q) out_counts[q] = 0;
    if (h->n == 0 || k == 0 || nq == 0) return VL_OK;

    DeviceGuard dg(h->device);
    Slot* sp = acquire_slot(h);
    if (!sp) return fail(VL_ERR_CUDA, "cannot create a CUDA stream");
    Slot& s = *sp;
    struct Rel { vl_index* h; Slot* s; ~Rel() { release_slot(h, s); } } rel{h, sp};

    const FlatView v = view_of(h);
    const bool tensor_path = h->mode == VL_MODE_AUTO && metric != VL_METRIC_MANHATTAN && h->dim <= TC_MAX_DIM && !getenv("VL_DISABLE_TC");
    const bool batched = h->mode != VL_MODE_EXACT && k <= 256 && nq >= batch_min_for(tensor_path, mirror_scans_apply(h));
    const uint32_t chunk = batched ? BATCH_CHUNK : NQ_CHUNK;
    bool nan = false;
    for (uint32_t q0 = 0; q0 < nq; q0 += chunk) {
        const uint32_t m = std::min(chunk, nq - q0);
        int st = search_levels(h, s, v, queries + static_cast<size_t>(q0) * qdim, m, qdim, k, metric, 0,
                               out_ids + static_cast<size_t>(q0) * k, out_scores + static_cast<size_t>(q0) * k,
                               out_counts + q0, &nan);
        if (st) return st;
    }
    // the reference panics (partial_cmp().unwrap(), flat.rs:116) as soon as n >= 2
    if (nan && h->n >= 2) return fail(VL_ERR_NAN, "similarity is NaN (the reference panics at flat.rs:116)");
    return VL_OK;
}

}  // namespace

namespace vl {
// for the host-only translation units of the library (group.cpp): sets the calling thread's vl_last_error()
void set_last_error(const char* msg) { snprintf(g_err, sizeof g_err, "%s", msg ? msg : ""); }
}  // namespace vl

// =========================================================================================
extern "C" {

const char* vl_last_error(void) { return g_err; }
const char* vl_version(void) { return "vectorlite-b200 0.1.0 (sm_100a)"; }

// Combiner (combiner.h) shared by the flat and HNSW host searches and by the shard group: concurrent SINGLE-query
// callers on one handle are combined into one batched search by the first caller in.
static bool combiner_enabled() { return Combiner::enabled(); }

int flat_search(vl_index* h, const float* queries, uint32_t nq, uint32_t qdim, uint32_t k, int metric,
                uint64_t* out_ids, double* out_scores, uint32_t* out_counts) {
    // Small stores (a scan of a few tens of microseconds, BASELINE config 1: 10K rows) are served per caller on
    // the handle's stream pool: a combined batch would wait for its cohort longer than the scan takes.
    const bool tiny = static_cast<uint64_t>(h->n) * h->pitch < (1ull << 24);
    if (!combiner_enabled() || nq != 1 || h->n == 0 || k == 0 || qdim != h->dim || tiny)
        return flat_search_impl(h, queries, nq, qdim, k, metric, out_ids, out_scores, out_counts);
    return h->comb.search(queries, qdim, k, metric, 0u, out_ids, out_scores, out_counts,
                          [&](const float* q, uint32_t m, uint32_t bk, int bm, uint32_t, uint64_t* ids, double* sc,
                              uint32_t* cnt) { return flat_search_impl(h, q, m, qdim, bk, bm, ids, sc, cnt); },
                          &h->stats[ST_COMBINED]);
}

static int create_common(uint32_t dim, int device, vl_index** out, int type) {
    if (!out) return fail(VL_ERR_INVALID, "out is null");
    *out = nullptr;
    if (dim == 0) return fail(VL_ERR_INVALID, "dimension cannot be 0");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(VL_ERR_CUDA, "no usable CUDA device (%s); there is no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(VL_ERR_INVALID, "device %d out of range [0,%d)", device, ndev);
    vl_index* h = new (std::nothrow) vl_index();
    if (!h) return fail(VL_ERR_OOM, "host allocation failed");
    h->type = type;
    h->device = device;
    h->dim = dim;
    h->pitch = (dim + 3) / 4 * 4;
    for (auto& c : h->stats) c = 0;
    DeviceGuard dg(device);
    ArenaStats init{0ull, 0x7FF0000000000000ull, 0u, 0u};
    if (cudaStreamCreateWithFlags(&h->mut_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&h->d_stats, sizeof(ArenaStats)) != cudaSuccess ||
        cudaMemcpy(h->d_stats, &init, sizeof init, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->dev_slot.stream, cudaStreamNonBlocking) != cudaSuccess) {
        const char* msg = cudaGetErrorString(cudaGetLastError());
        vl_index_destroy(h);
        return fail(VL_ERR_CUDA, "device init failed: %s", msg);
    }
    h->max_grid_x = std::min(flat_scan_max_grid_x(device, h->pitch), SCAN_CAP);
    h->max_grid_x_bf16 = std::min(flat_scan_bf16_max_grid_x(device), SCAN_CAP);
    {
        // The persistent scan grid leaves a few CTA slots FREE: a full wave (2 x 124 registers x 256 threads per SM)
        // fills the register file, so the finalize CTA of query i (and a sharded search's merge CTA, which waits
        // for its peers) would otherwise hold back one scan CTA of query i+1 — and with equal shares per CTA the
        // whole scan.  Measured in a pipelined stream at 1M x 384 (profiles/r01_grid_trim.txt): trim 0 / 1 / 2 / 4 / 8
        // = 7937 / 8302 / 8583 / 8721 / 8655 q/s (bf16 mirror), 4224 / 4447 / 4495 / 4556 / 4509 (fp32 arena).
        const char* e = getenv("VL_SCAN_GRID_TRIM");
        const int t = e ? std::max(0, atoi(e)) : 4;
        h->max_grid_x = std::max(1, h->max_grid_x - t);
        h->max_grid_x_bf16 = std::max(1, h->max_grid_x_bf16 - t);
    }
    *out = h;
    return VL_OK;
}

int vl_flat_create(uint32_t dim, int device, vl_index** out) {
    return create_common(dim, device, out, VL_INDEX_FLAT);
}

int vl_hnsw_create(uint32_t dim, int metric, uint32_t M, uint32_t M0, uint32_t efc, int device,
                   vl_index** out) {
    if (metric < 0 || metric > 3) return fail(VL_ERR_INVALID, "unknown metric %d", metric);
    int st = create_common(dim, device, out, VL_INDEX_HNSW);
    if (st) return st;
    (*out)->hnsw_metric = metric;
    (*out)->hnsw.reset(hnsw_state_create(dim, metric, M ? M : 16, M0 ? M0 : 32, efc ? efc : 400));
    if (!(*out)->hnsw) {
        vl_index_destroy(*out);
        *out = nullptr;
        return fail(VL_ERR_INVALID, "invalid HNSW parameters (M=%u, M0=%u)", M, M0);
    }
    return VL_OK;
}

void vl_index_destroy(vl_index* h) {
    if (!h) return;
    DeviceGuard dg(h->device);
    cudaDeviceSynchronize();
    if (h->hnsw) hnsw_state_release_device(h->hnsw.get());
    for (auto& s : h->slots) slot_free(*s);
    slot_free(h->dev_slot);
    cudaFree(h->d_rows); cudaFree(h->d_inv_norm); cudaFree(h->d_ids); cudaFree(h->d_stats);
    for (auto& e : h->prof_ev) cudaEventDestroy(e);
    tc_state_free(&h->tc);
    if (h->mut_stream) cudaStreamDestroy(h->mut_stream);
    delete h;
}

int vl_index_add_batch(vl_index* h, const uint64_t* ids, const float* rows, uint64_t n) {
    if (!h || (n && (!ids || !rows))) return fail(VL_ERR_INVALID, "null argument");
    DeviceGuard dg(h->device);
    if (h->type == VL_INDEX_HNSW) {
        // hnsw.rs:363-399: dup check against live ids AND inside the batch (the reference adds one vector at a
        // time and rejects the second occurrence, hnsw.rs:369), all-or-nothing, before anything is touched
        {
            std::unordered_set<uint64_t> seen;
            if (n > 1) seen.reserve(n * 2);
            for (uint64_t i = 0; i < n; ++i)
                if (hnsw_has_id(h->hnsw.get(), ids[i]) || (n > 1 && !seen.insert(ids[i]).second))
                    return fail(VL_ERR_DUP_ID, "Vector ID %llu already exists", static_cast<unsigned long long>(ids[i]));
        }
        const uint64_t first = h->n;
        // arena rows are addressed by internal index; ids are tracked by the HNSW state
        std::vector<uint64_t> internal(n);
        for (uint64_t i = 0; i < n; ++i) internal[i] = first + i;
        const bool was_identity = h->identity;
        int st = flat_add_batch(h, internal.data(), rows, n);
        if (st) return st;
        (void)was_identity;
        uint64_t nl = 0;
        st = hnsw_add_rows(h->hnsw.get(), ids, rows, n, h->d_rows, h->pitch, h->mut_stream, &nl);
        if (st) {
            h->n = first;   // roll the arena back: rows past n are never read, the graph was not extended
            return fail(st, "hnsw insert failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
        h->stats[ST_LAUNCHES] += nl;
        return VL_OK;
    }
    return flat_add_batch(h, ids, rows, n);
}

int vl_index_add(vl_index* h, uint64_t id, const float* values, uint32_t len) {
    if (!h || !values) return fail(VL_ERR_INVALID, "null argument");
    if (len != h->dim) {
        if (h->type == VL_INDEX_HNSW)  // hnsw.rs:365
            return fail(VL_ERR_DIM, "Vector dimension mismatch: expected %u, got %u", h->dim, len);
        return fail(VL_ERR_DIM, "Vector dimension mismatch");  // flat.rs:84
    }
    if (h->type == VL_INDEX_FLAT) {
        uint32_t pos;
        if (find_pos(h, id, &pos))
            return fail(VL_ERR_DUP_ID, "Vector ID %llu already exists", static_cast<unsigned long long>(id));
    }
    return vl_index_add_batch(h, &id, values, 1);
}

int vl_index_add_f64(vl_index* h, uint64_t id, const double* values, uint32_t len) {
    if (!h || !values) return fail(VL_ERR_INVALID, "null argument");
    std::vector<float> f(len);
    for (uint32_t i = 0; i < len; ++i) f[i] = static_cast<float>(values[i]);
    return vl_index_add(h, id, f.data(), len);
}

int vl_index_delete(vl_index* h, uint64_t id) {
    if (!h) return fail(VL_ERR_INVALID, "null handle");
    DeviceGuard dg(h->device);
    if (h->type == VL_INDEX_HNSW) {
        if (!hnsw_soft_delete(h->hnsw.get(), id))  // hnsw.rs:401-403
            return fail(VL_ERR_NOT_FOUND, "Vector ID %llu does not exist", static_cast<unsigned long long>(id));
        return VL_OK;
    }
    uint32_t pos;
    if (!find_pos(h, id, &pos)) return VL_OK;  // flat.rs:93-96: deleting a missing id is Ok
    int st = materialize_ids(h);
    if (st) return st;
    return flat_remove_pos(h, pos);
}

int vl_index_fill_synthetic(vl_index* h, uint64_t seed, uint64_t first_row, uint64_t n, uint32_t clusters,
                            uint64_t first_id) {
    if (!h) return fail(VL_ERR_INVALID, "null handle");
    if (h->type != VL_INDEX_FLAT) return fail(VL_ERR_UNSUPPORTED, "fill_synthetic is a flat-index utility");
    if (n == 0) return VL_OK;
    DeviceGuard dg(h->device);
    if (h->identity) {
        if (h->n && first_id != h->id_base + h->n) {
            int st = materialize_ids(h);
            if (st) return st;
        }
    }
    if (!h->identity) {
        for (uint64_t i = 0; i < n; ++i)
            if (h->id_to_pos.count(first_id + i))
                return fail(VL_ERR_DUP_ID, "Vector ID %llu already exists", static_cast<unsigned long long>(first_id + i));
    }
    int st = arena_reserve(h, h->n + n);
    if (st) return st;
    if (h->identity && h->n == 0) h->id_base = first_id;
    CU(launch_synth_fill(h->d_rows, h->n, n, h->dim, h->pitch, seed, first_row, clusters, h->mut_stream));
    CU(launch_row_norms(h->d_rows, h->n, n, h->dim, h->pitch, h->d_inv_norm, h->d_stats, h->mut_stream));
    if (!h->identity) {
        std::vector<uint64_t> ids(n);
        for (uint64_t i = 0; i < n; ++i) {
            ids[i] = first_id + i;
            h->id_to_pos.emplace(ids[i], static_cast<uint32_t>(h->n + i));
        }
        h->ids_host.insert(h->ids_host.end(), ids.begin(), ids.end());
        CU(cudaMemcpyAsync(h->d_ids + h->n, ids.data(), n * 8, cudaMemcpyHostToDevice, h->mut_stream));
    }
    CU(cudaStreamSynchronize(h->mut_stream));
    h->stats[ST_LAUNCHES] += 2;
    const uint64_t last = first_id + n - 1;
    if (!h->have_max || last > h->max_id) { h->max_id = last; h->have_max = true; }
    h->n += n;
    return VL_OK;
}

int vl_hnsw_set_builder(vl_index* h, int builder) {
    if (!h || h->type != VL_INDEX_HNSW) return fail(VL_ERR_INVALID, "not an HNSW index");
    if (builder < 0 || builder > 2) return fail(VL_ERR_INVALID, "builder must be 0 (auto), 1 (host) or 2 (device)");
    hnsw_set_builder(h->hnsw.get(), builder);
    return VL_OK;
}

int vl_hnsw_set_beam_factor(vl_index* h, uint32_t factor) {
    if (!h || h->type != VL_INDEX_HNSW) return fail(VL_ERR_INVALID, "not an HNSW index");
    if (factor < 1 || factor > 64) return fail(VL_ERR_INVALID, "beam factor must be in [1, 64]");
    hnsw_set_beam_mult(h->hnsw.get(), factor);
    return VL_OK;
}

int vl_hnsw_set_score_mode(vl_index* h, int mode) {
    if (!h || h->type != VL_INDEX_HNSW) return fail(VL_ERR_INVALID, "not an HNSW index");
    if (mode != 0 && mode != 1) return fail(VL_ERR_INVALID, "score mode must be 0 (exact) or 1 (reference-quantised)");
    hnsw_set_score_mode(h->hnsw.get(), static_cast<uint32_t>(mode));
    return VL_OK;
}

int vl_hnsw_graph_bytes(const vl_index* h, uint64_t* out_bytes) {
    if (!h || h->type != VL_INDEX_HNSW || !out_bytes) return fail(VL_ERR_INVALID, "not an HNSW index");
    *out_bytes = hnsw_graph_blob_bytes(h->hnsw.get());
    return VL_OK;
}

int vl_hnsw_export_graph(const vl_index* h, void* buf, uint64_t cap, uint64_t* out_written) {
    if (!h || h->type != VL_INDEX_HNSW || !buf) return fail(VL_ERR_INVALID, "not an HNSW index / null buffer");
    size_t w = 0;
    const int st = hnsw_export_graph(h->hnsw.get(), buf, cap, &w);
    if (out_written) *out_written = w;
    if (st == 9) return fail(VL_ERR_UNSUPPORTED, "the graph holds soft-deleted nodes whose rows are not exported; rebuild on load");
    if (st) return fail(VL_ERR_INVALID, "graph buffer too small: need %llu bytes", static_cast<unsigned long long>(w));
    return VL_OK;
}

int vl_hnsw_import_graph(vl_index* h, const uint64_t* ids, const float* rows, uint64_t n, const void* blob, uint64_t bytes) {
    if (!h || h->type != VL_INDEX_HNSW || !ids || !rows || !blob) return fail(VL_ERR_INVALID, "null argument / not an HNSW index");
    if (h->n != 0) return fail(VL_ERR_INVALID, "import_graph needs an empty index");
    DeviceGuard dg(h->device);
    std::vector<uint64_t> internal(n);
    for (uint64_t i = 0; i < n; ++i) internal[i] = i;
    int st = flat_add_batch(h, internal.data(), rows, n);
    if (st) return st;
    st = hnsw_import_graph(h->hnsw.get(), ids, rows, n, blob, bytes);
    if (st) {
        h->n = 0;
        if (st == 2) return fail(VL_ERR_DUP_ID, "duplicate vector ids");
        return fail(VL_ERR_INVALID, "graph blob does not match this index (parameters, row count or structure)");
    }
    return VL_OK;
}

int vl_hnsw_build_info(const vl_index* h, uint64_t* out_builder, uint64_t* out_micros) {
    if (!h || h->type != VL_INDEX_HNSW) return fail(VL_ERR_INVALID, "not an HNSW index");
    uint64_t info[2];
    hnsw_build_info(h->hnsw.get(), info);
    if (out_builder) *out_builder = info[0];
    if (out_micros) *out_micros = info[1];
    return VL_OK;
}

int vl_hnsw_graph_check(const vl_index* h, uint64_t* out6) {
    if (!h || h->type != VL_INDEX_HNSW || !out6) return fail(VL_ERR_INVALID, "not an HNSW index");
    hnsw_graph_check(h->hnsw.get(), out6);
    return VL_OK;
}

int vl_index_build(vl_index* h) {
    if (!h) return fail(VL_ERR_INVALID, "null handle");
    if (h->type != VL_INDEX_HNSW) return VL_OK;
    DeviceGuard dg(h->device);
    int st = hnsw_upload(h->hnsw.get(), h->mut_stream);
    if (st) return fail(st, "hnsw graph upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    return VL_OK;
}

int vl_index_search(vl_index* h, const float* queries, uint32_t nq, uint32_t qdim, uint32_t k, int metric,
                    uint32_t ef, uint64_t* out_ids, double* out_scores, uint32_t* out_counts) {
    if (!h || (nq && (!queries || !out_counts)) || (nq && k && (!out_ids || !out_scores)))
        return fail(VL_ERR_INVALID, "null argument");
    if (metric < 0 || metric > 3) return fail(VL_ERR_INVALID, "unknown metric %d", metric);
    if (h->type == VL_INDEX_HNSW) {
        // hnsw.rs:416-434: dimension (always), metric, empty
        if (qdim != h->dim) return fail(VL_ERR_DIM, "Dimension mismatch: expected %u, got %u", h->dim, qdim);
        if (metric != h->hnsw_metric)
            return fail(VL_ERR_METRIC_MISMATCH, "Metric mismatch: requested %d, index built for %d", metric, h->hnsw_metric);
        for (size_t i = 0; i < static_cast<size_t>(nq) * k; ++i) { out_ids[i] = ~0ull; out_scores[i] = 0.0; }
        for (uint32_t q = 0; q < nq; ++q) out_counts[q] = 0;
        if (hnsw_live(h->hnsw.get()) == 0 || k == 0 || nq == 0) return VL_OK;
        auto impl = [&](const float* q, uint32_t m, uint32_t bk, int, uint32_t bef, uint64_t* ids, double* sc,
                        uint32_t* cnt) -> int {
            DeviceGuard dg(h->device);
            uint64_t visited = 0, launches = 0;
            // The traversal gathers rows from the bf16 mirror where one applies (AUTO mode, 384-d; cosine: the
            // pre-normalised mirror, so no 1/‖row‖ is fetched either): half the bytes per evaluated node.  The mirror
            // is brought up to date lazily, like the flat scans do; the final k are re-scored in f64 from the fp32 rows.
            const void* mirror = nullptr;
            const float* sqn = nullptr;
            static const bool fp32_gather = getenv("VL_HNSW_FP32_GATHER") != nullptr;
            if (!fp32_gather && h->pitch == 384) {
                const FlatView v = view_of(h);
                int stm = single_query_mirror(h, v, h->hnsw_metric, h->mut_stream, &mirror, &sqn);
                if (stm) return stm;
            }
            // uploads a changed graph under its exclusive lock, then searches on a stream of its own
            int st = hnsw_search_host(h->hnsw.get(), h->d_rows, h->pitch, q, m, bk, bef, ids, sc, cnt, &visited, &launches, mirror);
            h->stats[ST_HNSW_VISITED] = visited;
            h->stats[ST_LAUNCHES] += launches;
            if (st) return fail(st, "hnsw search failed: %s", cudaGetErrorString(cudaGetLastError()));
            return VL_OK;
        };
        if (combiner_enabled() && nq == 1)   // concurrent single-query callers share one launch
            return h->comb.search(queries, qdim, k, metric, ef, out_ids, out_scores, out_counts, impl, &h->stats[ST_COMBINED]);
        return impl(queries, nq, k, metric, ef, out_ids, out_scores, out_counts);
    }
    return flat_search(h, queries, nq, qdim, k, metric, out_ids, out_scores, out_counts);
}

int vl_index_search_f64(vl_index* h, const double* queries, uint32_t nq, uint32_t qdim, uint32_t k, int metric,
                        uint32_t ef, uint64_t* out_ids, double* out_scores, uint32_t* out_counts) {
    if (!queries && nq) return fail(VL_ERR_INVALID, "null argument");
    std::vector<float> f(static_cast<size_t>(nq) * qdim);
    for (size_t i = 0; i < f.size(); ++i) f[i] = static_cast<float>(queries[i]);
    return vl_index_search(h, f.data(), nq, qdim, k, metric, ef, out_ids, out_scores, out_counts);
}

// device-resident search shared by vl_index_search_device and vl_index_search_exchange: `o` carries the
// output pointers for query 0 (and, for a row-sharded exchange, the peer mirrors of those outputs)
static int search_device_impl(vl_index* h, const float* d_queries, uint32_t nq, uint32_t k, int metric,
                              const SearchOut& o, cudaStream_t stream) {
    Slot& s = h->dev_slot;
    const FlatView v = view_of(h);
    const int Kp = pick_kp(k);
    const uint32_t tiles = (v.n + SCAN_TILE_ROWS - 1) / SCAN_TILE_ROWS;
    const int grid_x = static_cast<int>(std::min<uint32_t>(tiles, h->max_grid_x));
    if (h->pitch != h->dim) {
        // The kernels read rows and queries at the arena pitch (dim rounded up to 4 floats, zero padded, 16-byte
        // aligned float4 loads).  The ABI takes queries as [nq][dim]: repack them into a padded staging buffer.
        const size_t need = static_cast<size_t>(nq) * h->pitch;
        if (need > s.q_cap) {
            CU(cudaStreamSynchronize(stream));   // an earlier search may still read the old staging buffer
            if (s.d_q) cudaFree(s.d_q);
            s.d_q = nullptr; s.q_cap = 0;
            CU(cudaMalloc(&s.d_q, need * sizeof(float)));
            s.q_cap = need;
        }
        CU(cudaMemsetAsync(s.d_q, 0, need * sizeof(float), stream));
        CU(cudaMemcpy2DAsync(s.d_q, h->pitch * sizeof(float), d_queries, h->dim * sizeof(float), h->dim * sizeof(float), nq,
                             cudaMemcpyDeviceToDevice, stream));
        d_queries = s.d_q;
    }
    auto out_at = [&](uint32_t q0) {
        SearchOut out = o;
        out.ids = o.ids + static_cast<size_t>(q0) * k;
        out.scores = o.scores + static_cast<size_t>(q0) * k;
        out.pos = o.pos ? o.pos + static_cast<size_t>(q0) * k : nullptr;
        out.counts = o.counts + q0;
        out.flags = o.flags + q0;
        out.peers.q_off = o.peers.q_off + q0;
        return out;
    };
    const bool tensor_path = h->mode == VL_MODE_AUTO && metric != VL_METRIC_MANHATTAN && h->dim <= TC_MAX_DIM && !getenv("VL_DISABLE_TC");
    if (nq >= batch_min_for(tensor_path, mirror_scans_apply(h))) {
        const bool want_tc = h->mode == VL_MODE_AUTO && metric != VL_METRIC_MANHATTAN && !getenv("VL_DISABLE_TC");
        for (uint32_t q0 = 0; q0 < nq; q0 += BATCH_CHUNK) {
            const uint32_t m = std::min(BATCH_CHUNK, nq - q0);
            BatchWork bw;
            // pipelined handles chain consecutive batches (PDL): alternate the scratch set the rescore kernel reads
            const uint32_t parity = h->pipelined ? (s.batch_parity++ & 1u) : 0u;
            int st = VL_OK;
            if (h->pipelined) st = slot_reserve_batch(s, m, &bw, parity ^ 1u);   // both sets up front: no allocation in the 2nd batch
            if (!st) st = slot_reserve_batch(s, m, &bw, parity);
            if (st) return st;
            const SearchOut out = out_at(q0);
            uint64_t nl = 0;
            BatchTensor bt;
            if (want_tc) {  // device API: calls on one handle are caller-ordered, no lock needed
                CU(tc_prepare(&h->tc, v, h->cap, metric, m, stream));
                bt.usable = h->tc.usable;
                bt.scratch = &h->tc;
                bt.chain_batches = h->pipelined;
                bt.parity = parity;
                if (bt.usable) h->stats[ST_TENSOR] += m;
            }
            CU(launch_batch_flat(v, d_queries + static_cast<size_t>(q0) * h->pitch, m, k, metric, Kp, bw, out,
                                 want_tc ? &bt : nullptr, &nl, stream));
            h->stats[ST_LAUNCHES] += nl;
        }
        return VL_OK;
    }
    for (uint32_t q0 = 0; q0 < nq; q0 += NQ_CHUNK) {
        const uint32_t m = std::min(NQ_CHUNK, nq - q0);
        // scratch only (no query / output staging): reserve with k = 0 sized outputs
        int st;
        if ((st = grow_dev(s.cand, s.cand_cap, static_cast<size_t>(m) * grid_x * Kp))) return st;
        if ((st = grow_dev(s.cand_count, s.cc_cap, static_cast<size_t>(m) * grid_x))) return st;
        if ((st = grow_dev(s.cand_max, s.cm_cap, static_cast<size_t>(m) * grid_x))) return st;
        if (s.ctl_cap < DEV_SETS * NQ_CHUNK) {
            if ((st = grow_dev(s.ctl, s.ctl_cap, DEV_SETS * NQ_CHUNK))) return st;
            CU(cudaMemsetAsync(s.ctl, 0, DEV_SETS * NQ_CHUNK * sizeof(QueryCtl), stream));
        }
        // DEV_SETS control-block sets, rotated per launch: in pipelined mode the next scan starts
        // while the previous finalize (which re-arms its own set at the end) may still be running
        const uint32_t parity = h->dev_parity % DEV_SETS;
        QueryCtl* ctl = s.ctl + parity * NQ_CHUNK;
        h->dev_parity += 1;
        if ((st = reserve_early(s, stream))) return st;
        ScanWork w{s.cand, s.cand_count, s.cand_max, ctl, grid_x, Kp};
        w.early = s.early + static_cast<size_t>(parity) * NQ_CHUNK * EARLY_STRIDE;
        const SearchOut out = out_at(q0);
        const float* dq = d_queries + static_cast<size_t>(q0) * h->pitch;
        if ((st = launch_single_queries(h, v, dq, m, k, metric, w, out, h->pipelined, true, stream, nullptr))) return st;
    }
    return VL_OK;
}

int vl_index_search_device(vl_index* h, const float* d_queries, uint32_t nq, uint32_t k, int metric, uint32_t ef,
                           uint64_t* d_out_ids, double* d_out_scores, uint64_t* d_out_pos,
                           uint32_t* d_out_counts, uint32_t* d_out_flags, void* cuda_stream) {
    (void)ef;
    if (!h || !d_queries || !d_out_ids || !d_out_scores || !d_out_counts || !d_out_flags)
        return fail(VL_ERR_INVALID, "null argument");
    if (metric < 0 || metric > 3) return fail(VL_ERR_INVALID, "unknown metric %d", metric);
    if (h->type != VL_INDEX_FLAT) return fail(VL_ERR_UNSUPPORTED, "search_device: flat indexes only");
    if (h->n == 0 || k == 0 || nq == 0) return fail(VL_ERR_INVALID, "empty index, k == 0 or nq == 0");
    if (k > 256) return fail(VL_ERR_UNSUPPORTED, "search_device supports k <= 256");
    DeviceGuard dg(h->device);
    cudaStream_t stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->dev_slot.stream;
    SearchOut o{d_out_ids, d_out_scores, d_out_pos, d_out_counts, d_out_flags};
    return search_device_impl(h, d_queries, nq, k, metric, o, stream);
}

// ---- peer-memory exchange of a row-sharded flat index ---------------------------------------------------
struct vl_exchange {
    static constexpr uint32_t DEPTH = 4;
    int device = 0;
    uint32_t G = 0, rank = 0, nq_cap = 0, k_cap = 0;
    uint64_t blk_cap = 0;
    char* base = nullptr;
    size_t bytes = 0, ready_off = 0, ack_off = 0;
    char* peer[EXCH_MAX_PEERS] = {};
    bool ipc_open[EXCH_MAX_PEERS] = {};
    bool connected = false;
    uint64_t seq = 0;
    uint32_t merged[DEPTH] = {};   // merge CTAs launched on each slot so far (same on every rank)
};

int vl_exchange_create(int device, uint32_t world, uint32_t rank, uint32_t max_nq, uint32_t max_k, vl_exchange** out) {
    if (!out) return fail(VL_ERR_INVALID, "out is null");
    *out = nullptr;
    if (world < 1 || world > static_cast<uint32_t>(EXCH_MAX_PEERS) || rank >= world || !max_nq || !max_k || max_k > 256)
        return fail(VL_ERR_INVALID, "exchange: need 1 <= world <= %d, rank < world, 1 <= max_k <= 256", EXCH_MAX_PEERS);
    DeviceGuard dg(device);
    vl_exchange* x = new (std::nothrow) vl_exchange();
    if (!x) return fail(VL_ERR_OOM, "host allocation failed");
    x->device = device; x->G = world; x->rank = rank; x->nq_cap = max_nq; x->k_cap = max_k;
    x->blk_cap = (vl_packed_result_bytes(max_nq, max_k) + 255) / 256 * 256;
    const size_t data = static_cast<size_t>(vl_exchange::DEPTH) * world * x->blk_cap;
    const size_t sig = static_cast<size_t>(vl_exchange::DEPTH) * world * max_nq * sizeof(uint32_t);
    x->ready_off = data;
    x->ack_off = data + (sig + 255) / 256 * 256;
    x->bytes = x->ack_off + static_cast<size_t>(vl_exchange::DEPTH) * world * sizeof(uint32_t);
    cudaError_t e = cudaMalloc(&x->base, x->bytes);   // plain cudaMalloc: exportable with cudaIpcGetMemHandle
    if (e == cudaSuccess) e = cudaMemset(x->base, 0, x->bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        const int code = e == cudaErrorMemoryAllocation ? VL_ERR_OOM : VL_ERR_CUDA;
        vl_exchange_destroy(x);
        return fail(code, "exchange create: %s", cudaGetErrorString(e));
    }
    x->peer[rank] = x->base;
    x->connected = world == 1;
    *out = x;
    return VL_OK;
}

void vl_exchange_destroy(vl_exchange* x) {
    if (!x) return;
    DeviceGuard dg(x->device);
    cudaDeviceSynchronize();
    for (uint32_t g = 0; g < x->G; ++g)
        if (x->ipc_open[g]) cudaIpcCloseMemHandle(x->peer[g]);
    cudaFree(x->base);
    delete x;
}

int vl_exchange_local_handle(const vl_exchange* x, void* out_handle) {
    if (!x || !out_handle) return fail(VL_ERR_INVALID, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == VL_EXCHANGE_HANDLE_BYTES, "handle size");
    DeviceGuard dg(x->device);
    cudaIpcMemHandle_t hnd;
    CU(cudaIpcGetMemHandle(&hnd, x->base));
    memcpy(out_handle, &hnd, sizeof hnd);
    return VL_OK;
}

int vl_exchange_connect(vl_exchange* x, const void* handles) {
    if (!x || !handles) return fail(VL_ERR_INVALID, "null argument");
    DeviceGuard dg(x->device);
    for (uint32_t g = 0; g < x->G; ++g) {
        if (g == x->rank || x->ipc_open[g]) continue;
        cudaIpcMemHandle_t hnd;
        memcpy(&hnd, static_cast<const char*>(handles) + static_cast<size_t>(g) * VL_EXCHANGE_HANDLE_BYTES, sizeof hnd);
        void* p = nullptr;
        CU(cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess));
        x->peer[g] = static_cast<char*>(p);
        x->ipc_open[g] = true;
    }
    x->connected = true;
    return VL_OK;
}

int vl_exchange_connect_local(vl_exchange** all, uint32_t n) {
    if (!all || !n || n > static_cast<uint32_t>(EXCH_MAX_PEERS)) return fail(VL_ERR_INVALID, "bad argument");
    for (uint32_t i = 0; i < n; ++i)
        if (!all[i] || all[i]->G != n || all[i]->rank != i || all[i]->bytes != all[0]->bytes)
            return fail(VL_ERR_INVALID, "connect_local: exchange %u does not match (world, rank, capacities)", i);
    for (uint32_t i = 0; i < n; ++i) {
        DeviceGuard dg(all[i]->device);
        for (uint32_t g = 0; g < n; ++g) {
            if (all[g]->device != all[i]->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(all[g]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return fail(VL_ERR_CUDA, "peer access %d -> %d: %s", all[i]->device, all[g]->device, cudaGetErrorString(e));
                cudaGetLastError();
            }
            all[i]->peer[g] = all[g]->base;
        }
        all[i]->connected = true;
    }
    return VL_OK;
}

int vl_index_search_exchange(vl_index* h, vl_exchange* x, const float* d_queries, uint32_t nq, uint32_t k, int metric,
                             uint64_t* d_out_ids, double* d_out_scores, uint64_t* d_out_pos, uint32_t* d_out_counts,
                             uint32_t* d_out_flags, void* cuda_stream) {
    if (!h || !x || !d_queries || !d_out_ids || !d_out_scores || !d_out_counts || !d_out_flags)
        return fail(VL_ERR_INVALID, "null argument");
    if (metric < 0 || metric > 3) return fail(VL_ERR_INVALID, "unknown metric %d", metric);
    if (h->type != VL_INDEX_FLAT) return fail(VL_ERR_UNSUPPORTED, "search_exchange: flat indexes only");
    if (h->n == 0 || k == 0 || nq == 0) return fail(VL_ERR_INVALID, "empty shard, k == 0 or nq == 0");
    if (!x->connected) return fail(VL_ERR_INVALID, "exchange is not connected to its peers");
    if (x->device != h->device) return fail(VL_ERR_INVALID, "exchange and index live on different devices");
    if (nq > x->nq_cap || k > x->k_cap)
        return fail(VL_ERR_INVALID, "exchange capacity exceeded: nq %u > %u or k %u > %u", nq, x->nq_cap, k, x->k_cap);
    DeviceGuard dg(h->device);
    cudaStream_t stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->dev_slot.stream;
    const uint32_t G = x->G, r = x->rank;
    const uint32_t slot = static_cast<uint32_t>(x->seq % vl_exchange::DEPTH);
    const uint32_t stamp = static_cast<uint32_t>(x->seq / vl_exchange::DEPTH) + 1u;
    x->seq += 1;
    const size_t my_block = (static_cast<size_t>(slot) * G + r) * x->blk_cap;
    const size_t nk8 = static_cast<size_t>(nq) * k * 8;
    char* lb = x->base + my_block;
    SearchOut o{reinterpret_cast<uint64_t*>(lb), reinterpret_cast<double*>(lb + nk8),
                reinterpret_cast<uint64_t*>(lb + 2 * nk8), reinterpret_cast<uint32_t*>(lb + 3 * nk8),
                reinterpret_cast<uint32_t*>(lb + 3 * nk8 + static_cast<size_t>(nq) * 4)};
    o.peers.G = G; o.peers.self = r; o.peers.stamp = stamp; o.peers.q_off = 0;
    o.peers.ack_want = x->merged[slot];
    x->merged[slot] += nq;
    for (uint32_t g = 0; g < G; ++g) {
        o.peers.delta[g] = static_cast<long long>(x->peer[g] - x->base);
        o.peers.ready[g] = reinterpret_cast<uint32_t*>(x->peer[g] + x->ready_off) + (static_cast<size_t>(slot) * G + r) * x->nq_cap;
        o.peers.ack[g] = reinterpret_cast<const uint32_t*>(x->base + x->ack_off) + (static_cast<size_t>(slot) * G + g);
    }
    int st = search_device_impl(h, d_queries, nq, k, metric, o, stream);
    if (st) return st;
    ExchangeMerge m;
    m.G = G; m.self = r; m.nq = nq; m.k = k; m.stamp = stamp;
    m.slot = x->base + static_cast<size_t>(slot) * G * x->blk_cap;
    m.blk = x->blk_cap;
    m.ready = reinterpret_cast<const uint32_t*>(x->base + x->ready_off) + static_cast<size_t>(slot) * G * x->nq_cap;
    for (uint32_t g = 0; g < static_cast<uint32_t>(EXCH_MAX_PEERS); ++g)
        m.ack[g] = g < G ? reinterpret_cast<uint32_t*>(x->peer[g] + x->ack_off) + (static_cast<size_t>(slot) * G + r)
                         : nullptr;
    m.nq_cap = x->nq_cap;
    m.out_ids = d_out_ids; m.out_scores = d_out_scores; m.out_pos = d_out_pos; m.out_counts = d_out_counts;
    m.out_flags = d_out_flags;
    // Pipelined handles: the merge joins the programmatic-dependent-launch chain scan → finalize → merge →
    // next scan, so the next search's scan streams rows while this (tiny) kernel waits for the peers.
    CU(launch_exchange_merge(m, h->pipelined && nq == 1, stream));
    h->stats[ST_LAUNCHES] += 1;
    return VL_OK;
}

int vl_merge_topk_device(int device, uint32_t G, uint32_t nq, uint32_t k, const uint64_t* d_ids,
                         const double* d_scores, const uint64_t* d_pos, const uint32_t* d_counts,
                         uint64_t* d_out_ids, double* d_out_scores, uint64_t* d_out_pos,
                         uint32_t* d_out_counts, void* cuda_stream) {
    if (!d_ids || !d_scores || !d_pos || !d_counts || !d_out_ids || !d_out_scores || !d_out_counts)
        return fail(VL_ERR_INVALID, "null argument");
    DeviceGuard dg(device);
    CU(launch_merge_topk(G, nq, k, d_ids, d_scores, d_pos, d_counts, 0, d_out_ids, d_out_scores, d_out_pos,
                         d_out_counts, static_cast<cudaStream_t>(cuda_stream)));
    return VL_OK;
}

int vl_merge_topk_packed_device(int device, uint32_t G, uint32_t nq, uint32_t k, const void* d_packed,
                                uint64_t* d_out_ids, double* d_out_scores, uint64_t* d_out_pos,
                                uint32_t* d_out_counts, void* cuda_stream) {
    if (!d_packed || !d_out_ids || !d_out_scores || !d_out_counts) return fail(VL_ERR_INVALID, "null argument");
    DeviceGuard dg(device);
    const uint64_t nk8 = static_cast<uint64_t>(nq) * k * 8;
    const uint64_t stride = vl_packed_result_bytes(nq, k);
    const char* b = static_cast<const char*>(d_packed);
    CU(launch_merge_topk(G, nq, k, reinterpret_cast<const uint64_t*>(b), reinterpret_cast<const double*>(b + nk8),
                         reinterpret_cast<const uint64_t*>(b + 2 * nk8), reinterpret_cast<const uint32_t*>(b + 3 * nk8),
                         stride, d_out_ids, d_out_scores, d_out_pos, d_out_counts,
                         static_cast<cudaStream_t>(cuda_stream)));
    return VL_OK;
}

uint64_t vl_packed_result_bytes(uint32_t nq, uint32_t k) {
    return static_cast<uint64_t>(nq) * k * 24 + static_cast<uint64_t>(nq) * 8;
}

uint64_t vl_index_len(const vl_index* h) {
    if (!h) return 0;
    return h->type == VL_INDEX_HNSW ? hnsw_live(h->hnsw.get()) : h->n;
}
uint32_t vl_index_dim(const vl_index* h) { return h ? h->dim : 0; }
int vl_index_type_of(const vl_index* h) { return h ? h->type : -1; }
int vl_index_metric(const vl_index* h) { return h && h->type == VL_INDEX_HNSW ? h->hnsw_metric : -1; }
int vl_index_device(const vl_index* h) { return h ? h->device : -1; }

int vl_index_max_id(const vl_index* h, uint64_t* out_id) {
    if (!h || !out_id) return fail(VL_ERR_INVALID, "null argument");
    if (h->type == VL_INDEX_HNSW) {
        if (!hnsw_max_id(h->hnsw.get(), out_id)) return fail(VL_ERR_NOT_FOUND, "index is empty");
        return VL_OK;
    }
    if (!h->have_max || h->n == 0) return fail(VL_ERR_NOT_FOUND, "index is empty");
    *out_id = h->max_id;
    return VL_OK;
}

int vl_index_get_vector(const vl_index* h, uint64_t id, float* out_values) {
    if (!h || !out_values) return fail(VL_ERR_INVALID, "null argument");
    uint32_t pos;
    if (h->type == VL_INDEX_HNSW) {
        uint64_t ix;
        if (!hnsw_index_of(h->hnsw.get(), id, &ix)) return fail(VL_ERR_NOT_FOUND, "Vector ID %llu does not exist", static_cast<unsigned long long>(id));
        pos = static_cast<uint32_t>(ix);
    } else if (!find_pos(h, id, &pos)) {
        return fail(VL_ERR_NOT_FOUND, "Vector ID %llu does not exist", static_cast<unsigned long long>(id));
    }
    DeviceGuard dg(h->device);
    CU(cudaMemcpy(out_values, h->d_rows + static_cast<size_t>(pos) * h->pitch, h->dim * sizeof(float), cudaMemcpyDeviceToHost));
    return VL_OK;
}

int vl_index_export(const vl_index* h, uint64_t first, uint64_t cap, uint64_t* out_ids, float* out_rows,
                    uint64_t* out_n) {
    if (!h || !out_n) return fail(VL_ERR_INVALID, "null argument");
    if (h->type == VL_INDEX_HNSW) {  // live rows only, insertion order (what HNSWIndex serialises, hnsw.rs:197-213)
        *out_n = hnsw_export(h->hnsw.get(), first, cap, out_ids, out_rows);
        return VL_OK;
    }
    const uint64_t m = first >= h->n ? 0 : std::min(cap, h->n - first);
    *out_n = m;
    if (m == 0) return VL_OK;
    DeviceGuard dg(h->device);
    if (out_rows)
        CU(cudaMemcpy2D(out_rows, h->dim * sizeof(float), h->d_rows + first * h->pitch, h->pitch * sizeof(float),
                        h->dim * sizeof(float), m, cudaMemcpyDeviceToHost));
    if (out_ids)
        for (uint64_t i = 0; i < m; ++i) out_ids[i] = h->identity ? h->id_base + first + i : h->ids_host[first + i];
    return VL_OK;
}

int vl_index_set_mode(vl_index* h, int mode) {
    if (!h || mode < 0 || mode > 2) return fail(VL_ERR_INVALID, "bad mode");
    h->mode = mode;
    return VL_OK;
}
int vl_index_set_pos_base(vl_index* h, uint64_t base) {
    if (!h) return fail(VL_ERR_INVALID, "null handle");
    h->pos_base = base;
    return VL_OK;
}
int vl_index_stats(const vl_index* h, uint64_t* out, uint32_t n) {
    if (!h || !out) return fail(VL_ERR_INVALID, "null argument");
    for (uint32_t i = 0; i < n; ++i) out[i] = i < ST_N ? h->stats[i].load() : 0;
    return VL_OK;
}
int vl_index_set_pipelined(vl_index* h, int enabled) {
    if (!h) return fail(VL_ERR_INVALID, "null handle");
    h->pipelined = enabled != 0;
    return VL_OK;
}
int vl_index_set_profiling(vl_index* h, int enabled) {
    if (!h) return fail(VL_ERR_INVALID, "null handle");
    DeviceGuard dg(h->device);
    if (enabled && h->prof_ev.empty()) {
        h->prof_ev.resize(2048);
        for (auto& e : h->prof_ev) CU(cudaEventCreate(&e));
    }
    h->profiling = enabled != 0;
    h->prof_n = 0;
    return VL_OK;
}
int vl_index_profile_read(vl_index* h, double* out_ms_total, uint64_t* out_launches) {
    if (!h || !out_ms_total || !out_launches) return fail(VL_ERR_INVALID, "null argument");
    DeviceGuard dg(h->device);
    double tot = 0.0;
    for (size_t i = 0; i < h->prof_n; ++i) {
        CU(cudaEventSynchronize(h->prof_ev[2 * i + 1]));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, h->prof_ev[2 * i], h->prof_ev[2 * i + 1]));
        tot += ms;
    }
    *out_ms_total = tot;
    *out_launches = h->prof_n;
    h->prof_n = 0;
    return VL_OK;
}
int vl_index_device_rows(const vl_index* h, const float** d_rows, uint32_t* pitch) {
    if (!h || !d_rows || !pitch) return fail(VL_ERR_INVALID, "null argument");
    *d_rows = h->d_rows;
    *pitch = h->pitch;
    return VL_OK;
}

}  // extern "C"
