// hnsw.h — internal interface of the HNSW half of a vl_index handle (host graph builder +
// device search).  Mirrors HNSWIndex (src/index/hnsw.rs:197-518) on top of the shared arena.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <memory>

namespace vl {

struct HnswState;
struct HnswDeleter {
    void operator()(HnswState* s) const;
};
using HnswPtr = std::unique_ptr<HnswState, HnswDeleter>;

HnswState* hnsw_state_create(uint32_t dim, int metric, uint32_t M, uint32_t M0, uint32_t ef_construction);
void hnsw_state_release_device(HnswState* s);
bool hnsw_has_id(const HnswState* s, uint64_t id);
bool hnsw_index_of(const HnswState* s, uint64_t id, uint64_t* internal_index);
bool hnsw_max_id(const HnswState* s, uint64_t* out);
uint64_t hnsw_live(const HnswState* s);
// append n rows (host f32 [n][dim]) to the graph; internal index == arena position.  d_rows/pitch: the same
// rows in the device arena (already uploaded on `stream`).  A bulk add into an EMPTY graph is built on the
// device (hnsw_build.cu) unless the builder is pinned to the host; everything else is inserted by the
// parallel host builder.
enum { HNSW_BUILDER_AUTO = 0, HNSW_BUILDER_HOST = 1, HNSW_BUILDER_DEVICE = 2 };
int hnsw_add_rows(HnswState* s, const uint64_t* ids, const float* rows, uint64_t n, const float* d_rows,
                  uint32_t pitch, cudaStream_t stream, uint64_t* launches);
void hnsw_set_builder(HnswState* s, int builder);
void hnsw_set_score_mode(HnswState* s, uint32_t mode);
void hnsw_set_beam_mult(HnswState* s, uint32_t mult);
uint32_t hnsw_beam_mult(const HnswState* s);
// [0] builder used by the last bulk add (1 host, 2 device), [1] its wall time in microseconds
void hnsw_build_info(const HnswState* s, uint64_t out[2]);
bool hnsw_soft_delete(HnswState* s, uint64_t id);
void hnsw_graph_check(const HnswState* s, uint64_t out[6]);
// live rows in insertion order: up to `cap` starting at live position `first` (rows come from the host copy)
uint64_t hnsw_export(const HnswState* s, uint64_t first, uint64_t cap, uint64_t* out_ids, float* out_rows);
// Graph persistence.  The reference does not serialise its graph (#[serde(skip)], hnsw.rs:199-200) and
// re-inserts every vector on load in HashMap order (hnsw.rs:322-348): a different graph after every load.  The blob
// holds levels + adjacency of all nodes (versioned header, little endian); import installs it over the same rows
// in the same order without building.  Export refuses graphs with soft-deleted nodes (their rows are not part of
// vl_index_export): 9 = unsupported.  Import validates sizes, parameters and the structure (graph_check).
size_t hnsw_graph_blob_bytes(const HnswState* s);
int hnsw_export_graph(const HnswState* s, void* buf, size_t cap, size_t* written);
int hnsw_import_graph(HnswState* s, const uint64_t* ids, const float* rows, uint64_t n, const void* blob, size_t bytes);
// flatten + upload the graph if it changed since the last upload (takes the graph lock exclusively)
int hnsw_upload(HnswState* s, cudaStream_t stream);
// Re-entrant: uploads a changed graph under the exclusive graph lock, then searches under the shared lock on a
// stream / scratch set of its own (pool of HnswState::MAX_CTX).
// rows_bf16 (may be null): bf16 mirror of the rows for the traversal's gathers (cosine: pre-normalised rows)
int hnsw_search_host(HnswState* s, const float* d_rows, uint32_t pitch, const float* queries, uint32_t nq,
                     uint32_t k, uint32_t ef, uint64_t* out_ids, double* out_scores, uint32_t* out_counts,
                     uint64_t* visited, uint64_t* launches, const void* rows_bf16 = nullptr);

}  // namespace vl
