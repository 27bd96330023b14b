/*
 * vectorlite_cuda.h — C ABI of the B200-native VectorLite search hot path.
 *
 * Drop-in boundary: the reference's `trait VectorIndex` (src/lib.rs:224-245) and the two
 * implementations behind `enum VectorIndexWrapper` (src/lib.rs:270-327):
 *   FlatIndex  (src/index/flat.rs:59-136)  → vl_flat_create  + vl_index_*
 *   HNSWIndex  (src/index/hnsw.rs:197-518) → vl_hnsw_create  + vl_index_*
 * A Rust `extern "C"` block binds exactly these symbols (INTEGRATION.md shows the shim that
 * re-creates the reference's error strings / enum variants from the status codes).
 *
 * Conventions
 *  - Plain pointers and sizes only; every pointer argument is caller-owned and borrowed for
 *    the duration of the call; outputs are caller-allocated.  No callbacks, no exceptions
 *    cross the boundary, nothing aborts: every entry point returns a vl_status.
 *  - Vectors cross the boundary as f32 (device storage format) or f64 (the reference's
 *    interface type, src/lib.rs:232; narrowed to f32 on entry).  Scores are returned as f64
 *    and are bit-identical to the reference formulae evaluated on the stored f32 values
 *    widened to f64 (sequential accumulation, no FMA).
 *  - Threading (src/client.rs:243-247: Arc<RwLock<VectorIndexWrapper>>): vl_index_search,
 *    vl_index_get_vector, vl_index_len, ... are re-entrant on one handle (reader side);
 *    vl_index_add*, vl_index_delete, vl_index_fill_synthetic may assume external exclusion
 *    (writer side).  The last-error string is thread-local.
 *  - There is NO CPU fallback: every search runs on the CUDA device the handle was created
 *    on; if no device is usable vl_*_create fails with VL_ERR_CUDA.
 */
#ifndef VECTORLITE_CUDA_H
#define VECTORLITE_CUDA_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vl_index vl_index;

/* SimilarityMetric, declaration order of src/lib.rs:364-378 */
typedef enum vl_metric {
    VL_METRIC_COSINE = 0,
    VL_METRIC_EUCLIDEAN = 1,
    VL_METRIC_MANHATTAN = 2,
    VL_METRIC_DOT = 3
} vl_metric;

typedef enum vl_status {
    VL_OK = 0,
    VL_ERR_DIM = 1,             /* "Vector dimension mismatch" (flat.rs:84, hnsw.rs:365) /
                                   VectorLiteError::DimensionMismatch (flat.rs:100, hnsw.rs:417) */
    VL_ERR_DUP_ID = 2,          /* "Vector ID {} already exists" (flat.rs:87, hnsw.rs:369) */
    VL_ERR_NOT_FOUND = 3,       /* "Vector ID {} does not exist" (hnsw.rs:402); get_vector → None */
    VL_ERR_METRIC_MISMATCH = 4, /* VectorLiteError::MetricMismatch (hnsw.rs:426-429) */
    VL_ERR_INVALID = 5,         /* bad argument (null pointer, dim == 0, unknown metric, ...) */
    VL_ERR_CUDA = 6,            /* CUDA runtime failure, or no usable device */
    VL_ERR_OOM = 7,             /* host or device allocation failed */
    VL_ERR_NAN = 8,             /* a similarity is NaN: the reference panics (flat.rs:116) */
    VL_ERR_UNSUPPORTED = 9
} vl_status;

typedef enum vl_index_type { VL_INDEX_FLAT = 0, VL_INDEX_HNSW = 1 } vl_index_type; /* lib.rs:248-268 */

/* Search-path selector for vl_index_set_mode (flat indexes).
 * AUTO: fp32 (B=1) / bf16 tensor-core (batched) scan → over-select → fp64 rescore in reference
 *       summation order → optimality certificate; a query whose certificate fails is re-run on
 *       the EXACT path.  Both paths return identical, oracle-exact results.
 * EXACT: every row scored in f64 in reference order on the device, stable radix select. */
/* AUTO: approximate scans may read the bf16 mirror of the rows (tensor-core batches for rows up to 2048 elements;
 * single-query scans at 128 / 256 / 384 / 768 / 1024 / 1536-d) —
 * results stay bit-exact through the f64 rescore + certificate.  FP32: scans read the fp32 arena only.  EXACT: every
 * row scored in f64. */
typedef enum vl_mode { VL_MODE_AUTO = 0, VL_MODE_EXACT = 1, VL_MODE_FP32 = 2 } vl_mode;

/* ---- lifecycle ---------------------------------------------------------------------- */
/* FlatIndex::new(dim, vec![]) (flat.rs:68).  device = CUDA ordinal. */
int vl_flat_create(uint32_t dim, int device, vl_index** out);
/* HNSWIndex::new(dim, metric) (hnsw.rs:216-259).  The reference's compile-time profiles
 * (hnsw.rs:95-109: M/M0 = 16/32 default, 8/16 memory-optimized, 32/64 high-accuracy) and the
 * crate's ef_construction (400) are runtime parameters here; 0 selects the default. */
int vl_hnsw_create(uint32_t dim, int metric, uint32_t M, uint32_t M0, uint32_t ef_construction,
                   int device, vl_index** out);
void vl_index_destroy(vl_index* h);

/* ---- mutation (writer side) ----------------------------------------------------------- */
/* VectorIndex::add (flat.rs:82-91, hnsw.rs:363-399): VL_ERR_DIM / VL_ERR_DUP_ID. */
int vl_index_add(vl_index* h, uint64_t id, const float* values, uint32_t len);
int vl_index_add_f64(vl_index* h, uint64_t id, const double* values, uint32_t len);
/* Bulk load, == FlatIndex::new(dim, data) / the persistence load path (persistence.rs:149-176).
 * rows is [n][dim] row-major.  All-or-nothing on VL_ERR_DUP_ID. */
int vl_index_add_batch(vl_index* h, const uint64_t* ids, const float* rows, uint64_t n);
/* VectorIndex::delete.  Flat: deleting a missing id is VL_OK (flat.rs:93-96), order of the
 * remaining rows is preserved (it is the tie-break).  HNSW: VL_ERR_NOT_FOUND (hnsw.rs:401-403),
 * soft delete. */
int vl_index_delete(vl_index* h, uint64_t id);
/* Bench / test utility: append n rows generated on the device by the counter-based generator
 * (bit-identical to oracle vlo_synth_rows_f32), ids = first_id + i, generator row = first_row + i. */
int vl_index_fill_synthetic(vl_index* h, uint64_t seed, uint64_t first_row, uint64_t n,
                            uint32_t clusters, uint64_t first_id);
/* HNSW only: (re)build / finish the device graph after bulk adds.  Called implicitly by search. */
int vl_index_build(vl_index* h);
/* HNSW construction (src/index/hnsw.rs:363-399 inserts one vector at a time on one CPU thread).  A bulk
 * vl_index_add_batch into an EMPTY index (>= 4096 rows) is built on the device, layer by layer in batches
 * (csrc/hnsw_build.cu); single adds and adds to a non-empty index use the parallel host builder.
 * builder: 0 = auto (default), 1 = always host, 2 = device for any bulk add into an empty index. */
int vl_hnsw_set_builder(vl_index* h, int builder);
/* Builder used by the last bulk add (1 = host, 2 = device) and its wall time in microseconds. */
int vl_hnsw_build_info(const vl_index* h, uint64_t* out_builder, uint64_t* out_micros);
/* Scores returned by HNSW searches.  0 (default): the exact Flat similarity of each returned id (f64, lib.rs:425-572),
 * so Flat and HNSW agree.  1: the reference's own HNSW score — the functor's u64 milli-unit distance
 * (hnsw.rs:113-174) / 1000 (hnsw.rs:478) through convert_distance_to_similarity (hnsw.rs:51-75), bit for bit,
 * including its second division by 1000 for cosine and dot product; results ordered by that score. */
int vl_hnsw_set_score_mode(vl_index* h, int mode);
/* Graph persistence.  The reference does not serialise its graph (#[serde(skip)], hnsw.rs:199-200): on
 * load it re-inserts every vector in HashMap order (hnsw.rs:322-348) and gets a different graph every time.  Here a
 * saved index can restore the very graph it was searched with: vl_hnsw_export_graph writes levels + adjacency of
 * all nodes into a caller buffer of vl_hnsw_graph_bytes bytes (VL_ERR_UNSUPPORTED if the graph holds soft-deleted
 * nodes: their rows are not part of vl_index_export — rebuild on load instead, as the reference does);
 * vl_hnsw_import_graph installs such a blob into an EMPTY index together with the rows in the exported order
 * (vl_index_export), without building.  The blob is validated (parameters, sizes, structure audit). */
int vl_hnsw_graph_bytes(const vl_index* h, uint64_t* out_bytes);
int vl_hnsw_export_graph(const vl_index* h, void* buf, uint64_t cap, uint64_t* out_written);
int vl_hnsw_import_graph(vl_index* h, const uint64_t* ids, const float* rows, uint64_t n, const void* blob,
                         uint64_t bytes);
/* Device beam width of a search = factor x ef, where ef is the reference's ef (hnsw.rs:437: min(k, len), or the
 * `ef` argument of vl_index_search when > 0).  1 = equal ef.  Range [1, 64]. */
int vl_hnsw_set_beam_factor(vl_index* h, uint32_t factor);
/* Structural audit of the graph (all levels): out6 = nodes, layer-0 edges, self loops, duplicate edges,
 * invalid targets (out of range / absent from the level / after a gap), isolated layer-0 nodes. */
int vl_hnsw_graph_check(const vl_index* h, uint64_t* out6);

/* ---- queries (reader side) ------------------------------------------------------------ */
/* VectorIndex::search (flat.rs:98-119, hnsw.rs:415-496) for nq queries at once (nq = 1 is the
 * reference call).  queries is [nq][qdim] row-major HOST memory.  out_ids / out_scores are
 * [nq][k]; out_counts[q] = number of valid entries for query q (min(k, len); HNSW may return
 * fewer after soft deletes); unused slots hold id = UINT64_MAX, score = 0.  Results are ordered
 * by score descending, ties by insertion order (the stable sort of flat.rs:116).
 * Flat: VL_ERR_DIM only when the index is non-empty (flat.rs:99).  HNSW: VL_ERR_DIM always,
 * VL_ERR_METRIC_MISMATCH when metric != index metric.  ef: HNSW beam width; 0 = the reference's
 * ef = min(k, len) (hnsw.rs:437); ignored by flat indexes. */
int vl_index_search(vl_index* h, const float* queries, uint32_t nq, uint32_t qdim, uint32_t k,
                    int metric, uint32_t ef, uint64_t* out_ids, double* out_scores,
                    uint32_t* out_counts);
int vl_index_search_f64(vl_index* h, const double* queries, uint32_t nq, uint32_t qdim, uint32_t k,
                        int metric, uint32_t ef, uint64_t* out_ids, double* out_scores,
                        uint32_t* out_counts);
/* Same computation with DEVICE-resident queries [nq][dim] (dense, as the host API takes them; when dim is not
 * a multiple of 4 they are repacked on the stream into a zero-padded staging buffer at the arena pitch) and
 * DEVICE outputs, enqueued on `cuda_stream` (a cudaStream_t; NULL = the handle's own stream) without host
 * synchronisation.
 * d_out_pos (may be NULL) receives storage positions (+ the handle's position base, see
 * vl_index_set_pos_base) — what a row-sharded merge tie-breaks on.  d_out_flags[q]: bit0 = the
 * optimality certificate failed (caller must re-run that query through vl_index_search or in
 * VL_MODE_EXACT), bit1 = non-finite fp32 score seen, bit2 = NaN similarity, bit3 = candidate buffer
 * overflow (implies bit0), bit4 = a peer shard's results did not arrive (vl_index_search_exchange; treat like
 * bit0: the merged list is incomplete).  No retry happens here: the host API (vl_index_search) moves a failing
 * query through larger over-selections, the fp32 arena and finally the exact path by itself. */
int vl_index_search_device(vl_index* h, const float* d_queries, uint32_t nq, uint32_t k, int metric,
                           uint32_t ef, uint64_t* d_out_ids, double* d_out_scores,
                           uint64_t* d_out_pos, uint32_t* d_out_counts, uint32_t* d_out_flags,
                           void* cuda_stream);
/* Merge G per-shard result lists (each [nq][k], device memory, laid out [G][nq][k] as an
 * all-gather leaves them) into the global top-k ordered by (score desc, position asc).
 * counts is [G][nq]. */
int vl_merge_topk_device(int device, uint32_t G, uint32_t nq, uint32_t k, const uint64_t* d_ids,
                         const double* d_scores, const uint64_t* d_pos, const uint32_t* d_counts,
                         uint64_t* d_out_ids, double* d_out_scores, uint64_t* d_out_pos,
                         uint32_t* d_out_counts, void* cuda_stream);

/* Packed variant for a single-collective exchange.  Each shard's result block is
 *   [ids nq*k u64][scores nq*k f64][positions nq*k u64][counts nq u32][flags nq u32]
 * (vl_packed_result_bytes(nq,k) bytes; point vl_index_search_device's outputs at the five sections of
 * this rank's block), the G blocks are laid out back to back as ONE all-gather leaves them. */
uint64_t vl_packed_result_bytes(uint32_t nq, uint32_t k);
int vl_merge_topk_packed_device(int device, uint32_t G, uint32_t nq, uint32_t k, const void* d_packed,
                                uint64_t* d_out_ids, double* d_out_scores, uint64_t* d_out_pos,
                                uint32_t* d_out_counts, void* cuda_stream);

/* ---- row-sharded exchange over NVLink peer memory (one process per GPU) ----------------------
 * The reference has a single in-process index (client.rs:243-247); shard-aware routing is the
 * addition the north star names.  A vl_exchange owns, on this rank's device, a ring of slots with one
 * block per shard plus per-(shard, query) stamps, exported to the peer processes with CUDA IPC.
 * vl_index_search_exchange runs the local shard search with the kernel that produces the final top-k
 * storing it directly into every peer's slot (peer-mapped HBM, NVLink stores + system-scope release),
 * then a merge kernel waits for the G stamps of each query and merges from local memory: no
 * collective call and no host synchronisation on the data path.  Every rank must issue the same
 * sequence of calls (same nq, k).  Bounded waits: a peer that never arrives sets bit4 of the flags.
 *
 * Setup: every rank creates its exchange, publishes vl_exchange_local_handle (64 bytes) to all ranks
 * through any host channel, calls vl_exchange_connect with the [world][64] handle table, and passes a
 * host barrier before the first search.  vl_exchange_connect_local wires exchanges that live in ONE
 * process (tests; several GPUs driven by one process). */
#define VL_EXCHANGE_HANDLE_BYTES 64
typedef struct vl_exchange vl_exchange;
int vl_exchange_create(int device, uint32_t world, uint32_t rank, uint32_t max_nq, uint32_t max_k,
                       vl_exchange** out);
void vl_exchange_destroy(vl_exchange* x);
int vl_exchange_local_handle(const vl_exchange* x, void* out_handle);
int vl_exchange_connect(vl_exchange* x, const void* handles);
int vl_exchange_connect_local(vl_exchange** all, uint32_t n);
/* Sharded search: d_queries [nq][dim] replicated on every rank; outputs (device, [nq][k] / [nq]) hold
 * the GLOBAL top-k on every rank; d_out_flags[q] = OR of all shards' flags (bit0: some shard's
 * certificate failed → re-run exact).  Everything is enqueued on `cuda_stream`; on a pipelined handle
 * (vl_index_set_pipelined) the merge joins the programmatic-dependent-launch chain, so the next
 * search's scan streams rows while the merge is still waiting for the slowest peer. */
int vl_index_search_exchange(vl_index* h, vl_exchange* x, const float* d_queries, uint32_t nq, uint32_t k,
                             int metric, uint64_t* d_out_ids, double* d_out_scores, uint64_t* d_out_pos,
                             uint32_t* d_out_counts, uint32_t* d_out_flags, void* cuda_stream);

/* ---- shard group: a flat store row-sharded over several handles of ONE process ----------------------
 * The reference's server is a single process whose Collection owns one index (src/client.rs:243-247); on a
 * multi-GPU box it owns a group of flat handles, one per device.  Shard g holds the contiguous storage-order
 * range [base_g, base_g + n_g) (the caller routes inserts: bulk loads split evenly, appends to the shard that
 * owns the tail), so merging the per-shard top-k lists by a STABLE sort in shard order reproduces the stable
 * sort of flat.rs:116 over the whole store (score desc, insertion order asc).  vl_group_search runs
 * vl_index_search on every non-empty shard concurrently (host buffers in, host results out, same outputs and
 * errors as vl_index_search) and is re-entrant like it.  The group borrows the handles: destroy it first. */
typedef struct vl_group vl_group;
int vl_group_create(vl_index* const* shards, uint32_t n, vl_group** out);
void vl_group_destroy(vl_group* g);
uint32_t vl_group_size(const vl_group* g);
int vl_group_search(vl_group* g, const float* queries, uint32_t nq, uint32_t qdim, uint32_t k, int metric,
                    uint64_t* out_ids, double* out_scores, uint32_t* out_counts);

/* ---- accessors ------------------------------------------------------------------------ */
uint64_t vl_index_len(const vl_index* h);          /* VectorIndex::len */
uint32_t vl_index_dim(const vl_index* h);          /* VectorIndex::dimension */
int vl_index_type_of(const vl_index* h);           /* VectorIndexWrapper::index_type */
int vl_index_metric(const vl_index* h);            /* VectorIndexWrapper::metric: -1 for flat (None) */
int vl_index_device(const vl_index* h);
/* max_id() (flat.rs:76-78, hnsw.rs:267-269): VL_ERR_NOT_FOUND when empty. */
int vl_index_max_id(const vl_index* h, uint64_t* out_id);
/* VectorIndex::get_vector: the stored f32 values. */
int vl_index_get_vector(const vl_index* h, uint64_t id, float* out_values);
/* Bulk export in storage order (persistence / Clone): up to `cap` rows starting at storage
 * position `first`; returns rows written through out_n. */
int vl_index_export(const vl_index* h, uint64_t first, uint64_t cap, uint64_t* out_ids,
                    float* out_rows, uint64_t* out_n);

/* ---- tuning / introspection ------------------------------------------------------------ */
int vl_index_set_mode(vl_index* h, int mode);
/* Row-sharded deployments: global storage position = base + local position. */
int vl_index_set_pos_base(vl_index* h, uint64_t base);
/* Counters since creation: [0] kernels launched, [1] searches served by the certified fast path,
 * [2] queries re-run on the exact path, [3] bytes H2D, [4] bytes D2H, [5] last HNSW visited,
 * [6] single-query scans served from the bf16 mirror of the rows (AUTO mode, 128/256/384/768/1024/1536-d, all four metrics),
 * [7] single-query host searches that were combined with concurrent callers into a batched launch,
 * [8] queries whose certificate did not hold after a bf16 scan (mirror / tensor cores) and that were re-run at the
 *     next level (larger over-selection K', then the fp32 arena, then the exact path),
 * [9] queries re-run on the fp32 arena, [10] queries served at the boosted over-selection from the start (the
 *     handle remembers data whose top-k gaps are below the bf16 bound), [11] queries whose batched scan ran on the
 *     tensor cores (tcgen05 kernel over the bf16 mirror; rows up to 2048 elements wide). */
int vl_index_stats(const vl_index* h, uint64_t* out, uint32_t n);
/* Pipelined device searches (flat, vl_index_search_device only).  When enabled, consecutive searches
 * enqueued on one stream overlap through programmatic dependent launch: the scan of search i+1
 * starts while the small rescore/certify kernel of search i is still running (results of each
 * search still become visible in stream order).  Contract: the queries of a search must not be
 * produced by the kernel enqueued immediately before it on that stream (resident query batches,
 * H2D copies and earlier kernels are fine).  Off by default. */
int vl_index_set_pipelined(vl_index* h, int enabled);
/* Kernel timing for roofline reports: while enabled, vl_index_search_device brackets every launch
 * of the dominant scan kernel with CUDA events on the launching stream (ring of 1024 pairs).
 * vl_index_profile_read synchronises, returns the summed duration (ms) and number of the bracketed
 * launches since the last read, and clears the ring. */
int vl_index_set_profiling(vl_index* h, int enabled);
int vl_index_profile_read(vl_index* h, double* out_ms_total, uint64_t* out_launches);
/* Raw device pointers for profiling harnesses (rows, pitch in floats). */
int vl_index_device_rows(const vl_index* h, const float** d_rows, uint32_t* pitch);

const char* vl_last_error(void);
const char* vl_version(void);

#ifdef __cplusplus
}
#endif
#endif
