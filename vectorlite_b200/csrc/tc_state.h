// tc_state.h — per-index state of the tensor-core batched path (bf16 mirrors, TMA maps).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace vl {

// widest rows the tensor-core batched path takes (bf16 mirror, K padded to 64): up to 384 elements the query block is
// resident in shared memory, wider rows stream its K-chunks through the ring (batch_tc.cu, STREAM_A)
constexpr uint32_t TC_MAX_DIM = 2048;

struct TcState {
    bool usable = false;
    uint32_t KP = 0;               // K padded to a multiple of 64 bf16 (one 128-byte swizzle row)
    uint64_t cap = 0;              // rows allocated in the mirrors
    void* rows_norm = nullptr;     // bf16 [cap][KP], rows scaled by 1/‖row‖ (cosine)
    void* rows_raw = nullptr;      // bf16 [cap][KP], unscaled (dot, L2)
    float* sq_norm = nullptr;      // fp32 [cap] ‖row‖² (L2)
    uint64_t built_norm = 0, built_raw = 0;
    void* q_bf16 = nullptr;        // bf16 [q_cap][KP]
    float* qn2 = nullptr;          // fp32 [q_cap]
    float* eq = nullptr;           // fp32 [q_cap] E_q = ‖q̃−q‖/‖q‖ of the last converted batch (certificate)
    uint32_t* ex_bits = nullptr;   // [3] float bits, maxima over rows (integer atomicMax): [0] ‖x̃−x̂‖/‖x̂‖ of the normalised
                                   // mirror, [1] the same for the raw mirror, [2] ‖x̃−x‖₁ of the raw mirror (manhattan)
    uint32_t q_cap = 0;
    alignas(64) CUtensorMap map_x;
    alignas(64) CUtensorMap map_q;
    uint64_t maps_n = 0; const void* maps_base = nullptr; int maps_cs = 0;
    uint32_t mapq_rows = 0; const void* mapq_base = nullptr;
};

cudaError_t tc_prepare(TcState* t, const FlatView& v, uint64_t arena_cap, int metric, uint32_t nq, cudaStream_t s);
void tc_state_free(TcState* t);
int tc_cluster_size();

}  // namespace vl
