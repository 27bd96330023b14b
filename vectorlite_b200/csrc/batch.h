// batch.h — internal interface of the batched flat scan (batch_scan.cu, batch_tc.cu).
#pragma once
#include "kernels.h"

namespace vl {

struct BatchWork {        // per workspace slot, sized for nq queries
    uint64_t* cand;       // [nq][capq] candidate keys
    uint32_t* count;      // [nq]
    float* tau;           // [nq] running threshold in scan units (score of the K'-th best so far)
    uint32_t* qflags;     // [nq] FLAG_* bits raised by the scan kernels
    uint32_t capq;
    float* gmax = nullptr;   // [nq][GMAX_STRIDE] group maxima of the threshold-estimation stage (tensor path)
};
constexpr uint32_t GMAX_STRIDE = 1024;   // groups of 32 rows → estimation stages of up to 32768 rows
enum BatchScanMode : uint32_t { SCAN_FILTER = 0, SCAN_DIRECT = 1, SCAN_GROUPMAX = 2 };

struct BatchTensor {      // bf16 mirror of the arena for the tcgen05 path (batch_tc.cu)
    bool usable = false;
    const void* rows_bf16 = nullptr;   // [n][pitch] bf16 (cosine: rows pre-scaled by 1/‖row‖)
    const void* rows_bf16_raw = nullptr;  // [n][pitch] bf16, unscaled (dot, L2)
    const float* sq_norm = nullptr;    // [n] ‖row‖² fp32 (L2 via ‖x‖²+‖q‖²−2x·q)
    // certificate: |approx − exact| <= tc_abs·‖x‖·‖q‖.  BOTH operands are rounded to bf16 here (worst case 2^-8
    // relative per element, each side): (1 + 2^-8)² − 1 = 0.0078278, + K·2^-23 for the fp32 accumulation → 0.0079.
    double tc_abs = 0.0079;
    void* scratch = nullptr;           // TcState*
    // Pipelined device searches: batch i+1's query conversion is launched (PDL) while batch i's rescore kernel is
    // still running, so everything the rescore reads is double buffered: `parity` selects the E_q array here, the
    // caller alternates the BatchWork set.
    bool chain_batches = false;
    uint32_t parity = 0;
};

cudaError_t batch_scan_cuda_cores(const FlatView& v, const float* d_q, uint32_t nq, int metric, uint32_t lo,
                                  uint32_t hi, const BatchWork& w, cudaStream_t s);
// first == true: the first scan of a batch (converts the queries, resets the per-query state, heads the PDL chain)
cudaError_t batch_scan_tensor(const FlatView& v, const BatchTensor& tc, const float* d_q, uint32_t nq, int metric,
                              uint32_t lo, uint32_t hi, const BatchWork& w, cudaStream_t s, bool first, uint32_t mode);
// full pipeline: init → 3 staged scans with per-query selects → rescore/certify.  tc may be null.
cudaError_t launch_batch_flat(const FlatView& v, const float* d_queries, uint32_t nq, uint32_t k, int metric,
                              int Kp, const BatchWork& w, const SearchOut& out, const BatchTensor* tc,
                              uint64_t* launches, cudaStream_t s, int kp_base = 0);

}  // namespace vl
