// hnsw_state.h — shared definition of the HNSW half of a handle (host builder ↔ device search).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <unordered_map>
#include <vector>

namespace vl {

constexpr uint32_t HNSW_NONE = 0xFFFFFFFFu;

// Flattened graph as the device kernel sees it.
struct HnswDeviceGraph {
    const uint32_t* adj0 = nullptr;       // [n][M0] layer-0 adjacency, HNSW_NONE padded
    const uint32_t* upper_off = nullptr;  // [n] first slot in `upper` (HNSW_NONE when level == 0)
    const uint32_t* upper = nullptr;      // [slots][M]: levels 1..L of a node are consecutive slots
    const uint8_t* level = nullptr;       // [n]
    const uint8_t* deleted = nullptr;     // [n] soft-delete flags (hnsw.rs:407-411)
    const uint64_t* ids = nullptr;        // [n] internal index → caller id
    const float* inv_norm = nullptr;      // [n] (cosine)
    uint32_t n = 0, M = 0, M0 = 0, entry = 0;
    int max_level = -1;
};

struct HnswState {
    uint32_t dim = 0, M = 16, M0 = 32, efc = 400;
    int metric = 0;
    // ---- host graph (internal index == arena position == insertion order) ----
    std::vector<float> vecs;           // [n][dim]
    std::vector<float> inv_norm;       // [n]
    std::vector<uint8_t> level;        // [n]
    std::vector<uint32_t> adj0;        // [n][M0]
    std::vector<uint32_t> upper_off;   // [n]
    std::vector<uint32_t> upper;       // [slots][M]
    std::vector<uint8_t> deleted;      // [n]
    std::vector<uint64_t> id_of;       // [n]
    std::unordered_map<uint64_t, uint32_t> index_of;  // live ids only (id_to_index, hnsw.rs:204)
    uint64_t live = 0;
    int max_level = -1;
    uint32_t entry = 0;
    uint64_t rng_state = 0x9E3779B97F4A7C15ull;
    std::unique_ptr<std::atomic<uint8_t>[]> locks;
    size_t locks_cap = 0;
    std::mutex entry_mu;
    int builder = 0;                   // HNSW_BUILDER_*
    uint32_t score_mode = 0;           // 0 exact flat similarity, 1 reference-quantised (hnsw.rs:478 + 51-75)
    uint32_t beam_mult = 8;            // device beam width = beam_mult x ef (vl_hnsw_set_beam_factor)
    int last_builder = 0;              // builder used by the last bulk add
    uint64_t last_build_us = 0;
    // ---- device copy ----
    uint32_t* d_adj0 = nullptr; uint32_t* d_upper_off = nullptr; uint32_t* d_upper = nullptr;
    uint8_t* d_level = nullptr; uint8_t* d_deleted = nullptr; uint64_t* d_ids = nullptr; float* d_inv_norm = nullptr;
    size_t d_n_cap = 0, d_upper_cap = 0;
    bool dirty = true;          // graph changed since last upload
    bool deleted_dirty = true;
    // Incremental adds (host builder) only touch the new nodes' rows and the adjacency rows of their back-link
    // targets: those are listed here so the next search uploads a few hundred bytes instead of the whole graph
    // (~140 MB at 1M nodes).  `dirty_full` forces the full upload (device build, reallocation, too many rows).
    bool dirty_full = true;
    size_t uploaded_n = 0, uploaded_upper = 0;          // nodes / upper-array words already on the device
    std::vector<uint64_t> touched;                      // (node << 8) | level of re-written adjacency rows
    // ---- reader side (src/client.rs:398: many searches under one read lock) ----
    // graph_mu: a search holds it shared while its kernel reads the device graph; the first search after a
    // mutation takes it exclusively to upload (hnsw_upload may reallocate the device arrays).  Searches run on
    // their own streams / scratch (a small pool), never on the handle's mutation stream.
    std::shared_mutex graph_mu;
    struct SearchCtx {
        cudaStream_t stream = nullptr;
        float* d_q = nullptr; size_t q_cap = 0;
        unsigned char* d_out = nullptr; unsigned char* h_out = nullptr; size_t out_cap = 0;
        unsigned long long* d_visited = nullptr;
        unsigned long long* h_visited = nullptr;
        bool busy = false;
    };
    static constexpr int MAX_CTX = 4;
    std::mutex ctx_mu;
    std::condition_variable ctx_cv;
    std::vector<std::unique_ptr<SearchCtx>> ctxs;
    // ---- host-builder scratch kept across single adds (hnsw.rs:363-399 inserts one vector at a time) ----
    std::vector<uint32_t> seq_stamp;
    uint32_t seq_epoch = 0;
};

int hnsw_launch_search(const HnswDeviceGraph& g, const float* d_rows, uint32_t pitch, uint32_t dim, int metric,
                       const float* d_queries, uint32_t nq, uint32_t k, uint32_t ef, uint64_t* d_out_ids,
                       double* d_out_scores, uint32_t* d_out_counts, unsigned long long* d_visited,
                       cudaStream_t stream, uint32_t score_mode = 0, uint32_t beam_mult = 1,
                       const void* rows_bf16 = nullptr, uint32_t* visited_per_query = nullptr);

// hnsw_search.cu, construction mode: node d_order[i]'s row is query i; beam of `ef` on `level`
int hnsw_launch_build_search(const HnswDeviceGraph& g, const float* d_rows, uint32_t pitch, uint32_t dim, int metric,
                             const uint32_t* d_order, uint32_t nq, int level, bool entry_only, uint32_t ef,
                             unsigned long long* d_out_keys, uint32_t out_stride, uint32_t* d_out_counts,
                             cudaStream_t stream);
// device graph arrays for n nodes (hnsw_host.cpp)
int hnsw_reserve_device(HnswState* s, size_t n);
// hnsw_build.cu: build the whole (empty) graph of `s` on the device; adjacency is copied back to the host
int hnsw_build_device(HnswState* s, const float* d_rows, uint32_t pitch, cudaStream_t stream, uint64_t* launches);

}  // namespace vl
