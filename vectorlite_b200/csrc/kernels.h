// kernels.h — internal launch interface between the C ABI (api.cu) and the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace vl {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_CTAS_PER_SM = 3;   // 85 regs/thread: no spills with 8 rows × float4 in flight
constexpr int SCAN_ROWS_PER_WARP = 8;
constexpr int SCAN_TILE_ROWS = (SCAN_THREADS / 32) * SCAN_ROWS_PER_WARP;  // 64
constexpr int SCAN_CAP = 1024;            // candidate buffer entries per CTA
constexpr int SCAN_TILES_PER_CHECK = 4;   // CTA-wide barrier every 4 tiles
constexpr int SCAN_LIMIT = SCAN_CAP - SCAN_TILES_PER_CHECK * SCAN_TILE_ROWS;  // 768
constexpr int KP_MAX = 512;               // max over-selected candidates per query (K')
constexpr int FIN_THREADS = 512;
constexpr int EARLY_STRIDE = 1024;        // per-query slots of the scan's early grid-wide threshold exchange (>= grid_x)

struct FlatView {          // device-resident flat store (one shard)
    const float* rows;     // [n][pitch] fp32, zero padded to pitch
    const float* inv_norm; // [n] fp32 1/‖row‖ (0 for zero rows)
    const uint64_t* ids;   // [n] or nullptr when id == id_base + pos
    const ArenaStats* stats;
    uint64_t id_base;
    uint64_t pos_base;
    uint32_t n;
    uint32_t dim;
    uint32_t pitch;        // floats, multiple of 4
};

struct ScanWork {          // per workspace slot
    uint64_t* cand;        // [nq][grid_x][Kp]
    uint32_t* cand_count;  // [nq][grid_x]
    uint64_t* cand_max;    // [nq][grid_x] best key of each CTA's list (0 when empty)
    QueryCtl* ctl;         // [nq]
    int grid_x;
    int Kp;
    uint64_t* early = nullptr;  // [nq][EARLY_STRIDE] first-tile maxima, all zero between searches (the finalize
                           // kernel clears what its scan published); nullptr disables the early threshold
};

// Peer-memory exchange of a row-sharded search (exchange.cu).  When G > 0 the kernel that writes a
// query's final results stores them not only at the local SearchOut pointers (this shard's block of
// the local exchange slot) but also, through NVLink peer mappings, at the same offsets of every
// peer's slot (ptr + delta[g]), then publishes `stamp` in ready[g][q] with system-scope release.
constexpr int EXCH_MAX_PEERS = 8;
struct PeerPush {
    uint32_t G = 0;        // shards (0 = no exchange)
    uint32_t self = 0;
    uint32_t stamp = 0;    // use count of the slot (1, 2, ...)
    uint32_t q_off = 0;    // index of the launch's first query inside the signal arrays
    uint32_t ack_want = 0; // merge CTAs every peer must have completed on this slot before it is rewritten
    long long delta[EXCH_MAX_PEERS] = {};    // peer g's copy of my block − my local block (bytes)
    uint32_t* ready[EXCH_MAX_PEERS] = {};    // peer g's ready[slot][self][·]  (ready[self] is local)
    const uint32_t* ack[EXCH_MAX_PEERS] = {};// local ack[slot][g]: number of merge CTAs peer g has completed on this slot
};

struct SearchOut {         // device outputs, [nq][k]
    uint64_t* ids;
    double* scores;
    uint64_t* pos;         // may be nullptr
    uint32_t* counts;      // [nq]
    uint32_t* flags;       // [nq]
    PeerPush peers{};      // row-sharded exchange (G == 0: local outputs only)
};

// single-query-per-CTA-column fp32 streaming scan (grid = grid_x × nq)
// pipelined: launch with the PDL attribute so the scan may start while the PREVIOUS kernel in the
// stream (the previous query's finalize) is still running; the scan only waits for it right before
// publishing its candidates.  The caller guarantees the queries were not produced by that kernel.
cudaError_t launch_flat_scan(const FlatView& v, const float* d_queries, uint32_t nq, int metric,
                             const ScanWork& w, bool pipelined, cudaStream_t s);
// the same scan over the index's bf16 mirror of the rows (384-d, cosine / dot / L2): half the HBM bytes per
// query; scores in the tensor-core path's scan units, certified with its bf16 bound (finalize: tc_abs > 0)
cudaError_t launch_flat_scan_bf16(const FlatView& v, const void* mirror, const float* sq_norm, const float* d_queries,
                                  uint32_t nq, int metric, const ScanWork& w, bool pipelined, cudaStream_t s);
// Certificate inputs of a scan over the bf16 mirror (rescore.cuh).  tc_abs > 0: the candidates come from a bf16 scan,
// |approx − exact| <= tc_abs·‖x‖·‖q‖ in the worst case; e_x / e_q / e_x1: the rounding-error norms measured when the
// mirror was built and the queries were converted (the smaller of the two bounds is used).  kp_base > 0: the kernel
// also reports FLAG_BASE_OK when the first kp_base candidates alone would have certified the result.
struct CertAux {
    double tc_abs = 0.0;
    const uint32_t* e_x = nullptr;
    const float* e_q = nullptr;
    const uint32_t* e_x1 = nullptr;
    int kp_base = 0;
};
// merge per-CTA candidates, fp64 rescore in reference order, rank, certify (grid = nq)
cudaError_t launch_flat_finalize(const FlatView& v, const float* d_queries, uint32_t nq, uint32_t k,
                                 int metric, const ScanWork& w, const SearchOut& out, float eps_scale,
                                 cudaStream_t s, const CertAux& aux = CertAux());
size_t flat_scan_smem_bytes(uint32_t pitch);
int flat_scan_max_grid_x(int device, uint32_t pitch);
bool flat_scan_bf16_supports(uint32_t pitch);   // rows of 128 / 256 / 384 / 768 / 1024 / 1536 elements
int flat_scan_bf16_max_grid_x(int device);   // the bf16 kernel keeps 2 CTAs per SM resident (124 registers)

// exact path: every row scored in f64 in reference order
cudaError_t launch_exact_scores(const FlatView& v, const float* d_query, int metric, double* d_scores,
                                uint32_t* d_flags, cudaStream_t s);
// stable select of the top-k of d_scores[n] → out (single query q_index of the out arrays)
struct ExactScratch {
    void* temp = nullptr;
    size_t temp_bytes = 0;
    uint64_t* keys_in = nullptr;
    uint64_t* keys_out = nullptr;
    uint32_t* vals_in = nullptr;
    uint32_t* vals_out = nullptr;
    size_t cap = 0;
};
cudaError_t exact_select(const FlatView& v, const double* d_scores, uint32_t k, ExactScratch& sc,
                         const SearchOut& out, uint32_t q_index, cudaStream_t s);
void exact_scratch_free(ExactScratch& sc);

// insert-time kernels
cudaError_t launch_row_norms(float* rows, uint64_t first, uint64_t n, uint32_t dim, uint32_t pitch,
                             float* inv_norm, ArenaStats* stats, cudaStream_t s);
cudaError_t launch_synth_fill(float* rows, uint64_t first_pos, uint64_t n, uint32_t dim, uint32_t pitch,
                              uint64_t seed, uint64_t first_row, uint32_t clusters, cudaStream_t s);
// exchange.cu: wait for every shard's block of the slot (ready stamps), merge, acknowledge.
struct ExchangeMerge {
    uint32_t G, self, nq, k, stamp;
    const char* slot;              // local slot: G blocks of `blk` bytes, packed layout (vl_packed_result_bytes)
    uint64_t blk;
    const uint32_t* ready;         // local ready[slot][g][q], stride nq_cap
    uint32_t* ack[EXCH_MAX_PEERS]; // peer g's ack[slot][self] counter
    uint32_t nq_cap;
    uint64_t* out_ids; double* out_scores; uint64_t* out_pos; uint32_t* out_counts; uint32_t* out_flags;
};
// pipelined: launch with the PDL attribute (the kernel releases its dependents at once and ends with
// griddepcontrol.wait, so "merge complete" still implies "the finalize before it is complete")
cudaError_t launch_exchange_merge(const ExchangeMerge& m, bool pipelined, cudaStream_t s);

cudaError_t launch_merge_topk(uint32_t G, uint32_t nq, uint32_t k, const uint64_t* ids,
                              const double* scores, const uint64_t* pos, const uint32_t* counts,
                              uint64_t rank_stride, uint64_t* out_ids, double* out_scores, uint64_t* out_pos,
                              uint32_t* out_counts, cudaStream_t s);

}  // namespace vl
