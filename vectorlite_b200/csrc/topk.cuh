// topk.cuh — CTA-level streaming top-K' selection over 64-bit keys in shared memory.
//
// Replaces the reference's "score everything, stable-sort n results, truncate"
// (src/index/flat.rs:106-117) with a threshold-filtered candidate buffer: rows whose key beats
// the running K'-th best key are appended (rare after warm-up: ~K' ln(n/K') appends for n rows),
// and the buffer is compacted by an in-smem bitonic sort when it fills.  n scores are never
// materialised.
#pragma once
#include "common.cuh"

namespace vl {

template <int CAP, int THREADS>
struct CtaTopK {
    static_assert((CAP & (CAP - 1)) == 0, "CAP must be a power of two");
    uint64_t* keys;  // shared memory [CAP]
    int* count;      // shared memory

    __device__ __forceinline__ void init() {
        for (int i = threadIdx.x; i < CAP; i += THREADS) keys[i] = 0ull;
        if (threadIdx.x == 0) *count = 0;
        __syncthreads();
    }

    // Any thread, any time between barriers.  The caller guarantees room (see room_for()).
    __device__ __forceinline__ void push(uint64_t key) {
        const int i = atomicAdd(count, 1);
        if (i < CAP) keys[i] = key;
    }

    // Descending bitonic sort of keys[0..len), len a power of two <= CAP.  All threads.
    __device__ __forceinline__ void sort_desc(int len) {
        for (int k = 2; k <= len; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = threadIdx.x; t < (len >> 1); t += THREADS) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int p = i | j;
                    const bool desc = (i & k) == 0;
                    const uint64_t a = keys[i], b = keys[p];
                    if ((a < b) == desc) {
                        keys[i] = b;
                        keys[p] = a;
                    }
                }
                __syncthreads();
            }
        }
    }

    // Keep the `keep` largest keys (sorted descending when a sort was needed or `force_sort`).
    // Returns the keep-th largest key (0 while fewer than `keep` keys are held): a valid lower
    // bound for membership in the global top-`keep`.  All threads must call; contains barriers.
    __device__ __forceinline__ uint64_t compact(int keep, bool force_sort) {
        __syncthreads();
        const int n = min(*count, CAP);
        uint64_t tau = 0ull;
        if (n > keep || force_sort) {
            int len = 2;
            while (len < n) len <<= 1;
            for (int i = n + threadIdx.x; i < len; i += THREADS) keys[i] = 0ull;
            __syncthreads();
            sort_desc(len);
            const int n2 = min(n, keep);
            if (n2 == keep) tau = keys[keep - 1];
            __syncthreads();
            if (threadIdx.x == 0) *count = n2;
            __syncthreads();
        }
        return tau;
    }
};

}  // namespace vl
