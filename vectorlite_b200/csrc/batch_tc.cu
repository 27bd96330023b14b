// batch_tc.cu — tensor-core (tcgen05 / TMEM / TMA) batched scan.  Placeholder until the kernel lands:
// reports "not usable" so launch_batch_flat stays on the CUDA-core tile kernel.
#include "batch.h"

namespace vl {

cudaError_t batch_scan_tensor(const FlatView&, const BatchTensor&, const float*, uint32_t, int, uint32_t,
                              uint32_t, const BatchWork&, cudaStream_t) {
    return cudaErrorNotSupported;
}

}  // namespace vl
