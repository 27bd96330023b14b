/*
 * vl_oracle.h — CPU ORACLE for the VectorLite search hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (vectorlite_b200/) never links, imports or calls anything in this directory.
 *
 * It is a C++ restatement (the reference is Rust; no Rust toolchain exists in this
 * image, so oracle/_ref cannot be built — see DESIGN.md) of:
 *   - src/lib.rs:380-391   SimilarityMetric::calculate  (dispatch, arg order (stored, query))
 *   - src/lib.rs:425-444   cosine_similarity
 *   - src/lib.rs:476-489   euclidean_similarity
 *   - src/lib.rs:521-532   manhattan_similarity
 *   - src/lib.rs:565-572   dot_product
 *   - src/index/flat.rs:98-119  FlatIndex::search (dim rule, score all rows, stable sort desc, truncate)
 *   - src/index/hnsw.rs:113-174 u64 milli-unit distance functors
 *   - src/index/hnsw.rs:51-75   convert_distance_to_similarity
 *   - src/index/hnsw.rs:363-496 HNSWIndex add / delete / search
 *   - crate hnsw 0.11.0 (+ space 0.17.0, rand 0.8.5 StdRng) insert / nearest —
 *     THIRD-PARTY, source NOT in /root/reference; restated from the published algorithm
 *     (SURVEY.md Appendix C).  Graph-level parity is therefore UNPINNED; only the
 *     toy KATs of hnsw.rs:605-634 etc. pin it (see tests/test_oracle_golden.py).
 *
 * Parity pinning: Flat + metrics are pinned against every known-answer test the
 * reference's own unit tests hold for this path (SURVEY.md §8c ①-⑩).
 *
 * Arithmetic rules (must hold for bit-equality with rustc output): IEEE f64, strict
 * left-to-right accumulation from 0.0, no FMA contraction (compile with
 * -ffp-contract=off, no -ffast-math), powi(2) == x*x.
 */
#ifndef VL_ORACLE_H
#define VL_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* declaration order of src/lib.rs:364-378 */
enum { VLO_COSINE = 0, VLO_EUCLIDEAN = 1, VLO_MANHATTAN = 2, VLO_DOT = 3 };

/* status codes shared with the product's C ABI numbering */
enum { VLO_OK = 0, VLO_ERR_DIM = 1, VLO_ERR_DUP_ID = 2, VLO_ERR_NOT_FOUND = 3,
       VLO_ERR_METRIC_MISMATCH = 4, VLO_ERR_INVALID = 5, VLO_ERR_NAN = 8 };

/* src/lib.rs:380-391 — calculate(a = stored, b = query) */
double vlo_metric(int metric, const double* a, const double* b, size_t n);

/* src/index/flat.rs:98-119 on f64 rows [n][dim].  Returns VLO_ERR_DIM when n>0 and
 * qdim != dim (flat.rs:99-104); VLO_ERR_NAN where the reference would panic
 * (partial_cmp().unwrap() on NaN, flat.rs:116).  out_* hold min(k,n) entries. */
int vlo_flat_search(const double* rows, const uint64_t* ids, size_t n, size_t dim,
                    const double* q, size_t qdim, size_t k, int metric,
                    uint64_t* out_ids, double* out_scores, size_t* out_count);

/* Same, rows/query given as f32 and widened to f64 (what the device stores). */
int vlo_flat_search_f32(const float* rows, const uint64_t* ids, size_t n, size_t dim,
                        const float* q, size_t qdim, size_t k, int metric,
                        uint64_t* out_ids, double* out_scores, size_t* out_count);

/* nq queries, one query per thread on nthreads threads (the HTTP server's best case
 * under the collection read lock, src/client.rs:398).  ids may be NULL (id = position).
 * clone_bytes > 0 additionally performs one heap allocation + copy of that many bytes
 * per row per query, bracketing the text/metadata clones of flat.rs:111-112. */
int vlo_flat_search_batch_f32(const float* rows, const uint64_t* ids, size_t n, size_t dim,
                              const float* queries, size_t nq, size_t k, int metric,
                              int nthreads, size_t clone_bytes,
                              uint64_t* out_ids, double* out_scores);

/* src/index/hnsw.rs:113-174 — quantised u64 distance functors */
uint64_t vlo_hnsw_distance(int metric, const double* a, const double* b, size_t n);
/* src/index/hnsw.rs:51-75 */
double vlo_convert_distance_to_similarity(double distance, int metric);

/* HNSWIndex (src/index/hnsw.rs:197-518) over a restated crate-hnsw-0.11 graph. */
typedef struct vlo_hnsw vlo_hnsw;
vlo_hnsw* vlo_hnsw_create(size_t dim, int metric, size_t M, size_t M0, size_t ef_construction);
void vlo_hnsw_destroy(vlo_hnsw* h);
int vlo_hnsw_add(vlo_hnsw* h, uint64_t id, const double* v, size_t len);      /* hnsw.rs:363-399 */
int vlo_hnsw_add_batch_f32(vlo_hnsw* h, const uint64_t* ids, const float* rows, size_t n);
int vlo_hnsw_delete(vlo_hnsw* h, uint64_t id);                                 /* hnsw.rs:400-414 */
size_t vlo_hnsw_len(const vlo_hnsw* h);
/* hnsw.rs:415-496.  ef == 0 reproduces the reference (ef = min(k, len)); ef > 0 is
 * the additive sweep knob (ef_search = max(ef, min(k,len)), results truncated to k).
 * out_visited (may be NULL) receives the number of distance evaluations. */
int vlo_hnsw_search(const vlo_hnsw* h, const double* q, size_t qdim, size_t k, int metric,
                    size_t ef, uint64_t* out_ids, double* out_scores, size_t* out_count,
                    uint64_t* out_visited);
int vlo_hnsw_search_batch_f32(const vlo_hnsw* h, const float* queries, size_t nq, size_t k,
                              size_t ef, int nthreads, uint64_t* out_ids, double* out_scores,
                              uint32_t* out_counts, uint64_t* out_visited_total);
/* graph introspection for tests */
size_t vlo_hnsw_num_layers(const vlo_hnsw* h);           /* upper layers */
size_t vlo_hnsw_layer_len(const vlo_hnsw* h, size_t l);  /* l=0 → zero layer */
/* layer-0 adjacency [len][M0], empty slot = UINT64_MAX; upper layer l >= 1 */
void vlo_hnsw_export_zero(const vlo_hnsw* h, uint64_t* out);
void vlo_hnsw_export_layer(const vlo_hnsw* h, size_t l, uint64_t* zero_node, uint64_t* next_node, uint64_t* nb);
/* distances that fell inside the guard band of the accelerated functors and were decided by the strict ones */
uint64_t vlo_hnsw_strict_evals(const vlo_hnsw* h);
/* first n levels drawn by the crate's random_level() for a given M (ChaCha12, zero seed) */
void vlo_hnsw_levels(size_t M, size_t n, uint32_t* out_levels);

/* Counter-based synthetic vectors shared bit-exactly with the device generator
 * (vectorlite_b200/csrc/synth.cuh restates the same integer recipe independently):
 * splitmix64(seed,row,col/4) → 4×u16 Irwin–Hall(4) ints → exact integer ‖·‖² →
 * f64 sqrt / div → f32.  Unit L2 norm up to f32 rounding (mimics embeddings.rs:173-179).
 * clusters == 0: i.i.d. directions.  clusters > 0: Gaussian-mixture-like data, row r
 * belongs to centre hash(r) % clusters and v = 4*centre + noise (noise σ = centre σ / 4). */
void vlo_synth_rows_f32(uint64_t seed, uint64_t row0, size_t n, size_t dim, uint32_t clusters,
                        float* out);

#ifdef __cplusplus
}
#endif
#endif
