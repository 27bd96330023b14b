// exchange.cu — the exchange step of the row-sharded flat index over NVLink peer memory.
//
// The reference has one process and one index (src/client.rs:243-247); the row-sharded deployment
// keeps the result of FlatIndex::search (src/index/flat.rs:98-119) exact by merging per-shard top-k
// lists under the global order (score desc, global storage position asc).  Instead of a collective
// call between "search" and "merge", the kernel that produces a shard's final top-k
// (rescore_rank_certify, rescore.cuh) stores it straight into every peer's exchange slot through
// peer-mapped HBM and publishes a per-(shard, query) stamp with system-scope release; the kernel
// below waits for the G stamps of its query, merges the G lists that are by then in LOCAL memory,
// and acknowledges to the peers (one counter per slot and peer) that the slot may be reused.  No host synchronisation, no NCCL call
// on the data path; every wait is bounded (FLAG_EXCHANGE on time-out) so a lost peer cannot hang
// the device.
#include "kernels.h"

namespace vl {

static __device__ __forceinline__ bool better(double sa, uint64_t pa, double sb, uint64_t pb) {
    return sa > sb || (sa == sb && pa < pb);
}

// one CTA per query; dynamic smem: scores[G*k] f64 | pos[G*k] u64 | cnt[G] u32
__global__ void __launch_bounds__(256) exchange_merge_kernel(ExchangeMerge m) {
    extern __shared__ __align__(16) unsigned char xsm[];
    double* s_sc = reinterpret_cast<double*>(xsm);
    uint64_t* s_pos = reinterpret_cast<uint64_t*>(s_sc + m.G * m.k);
    uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_pos + m.G * m.k);
    __shared__ uint32_t s_flags;
    __shared__ int s_fail;
    const uint32_t q = blockIdx.x, tid = threadIdx.x, G = m.G, k = m.k;
    const size_t nk8 = static_cast<size_t>(m.nq) * k * 8;
    pdl_launch_dependents();   // pipelined streams: the next search's scan may start now
    if (tid == 0) { s_flags = 0u; s_fail = 0; }
    __syncthreads();
    // ---- wait: shard g's block for this query has landed in local memory ------------------------------
    if (tid < G) {
        if (!wait_stamp(m.ready + static_cast<size_t>(tid) * m.nq_cap + q, m.stamp, EXCH_TIMEOUT_NS)) s_fail = 1;
    }
    __syncthreads();
    // ---- stage the G lists (L2 loads: the data was written by peers, never cached in this SM's L1) -----
    if (tid < G) {
        const char* b = m.slot + static_cast<size_t>(tid) * m.blk;
        const uint32_t c = __ldcg(reinterpret_cast<const uint32_t*>(b + 3 * nk8) + q);
        s_cnt[tid] = c < k ? c : k;
        atomicOr(&s_flags, __ldcg(reinterpret_cast<const uint32_t*>(b + 3 * nk8 + static_cast<size_t>(m.nq) * 4) + q));
    }
    for (uint32_t e = tid; e < G * k; e += blockDim.x) {
        const uint32_t g = e / k, i = e - g * k;
        const char* b = m.slot + static_cast<size_t>(g) * m.blk;
        const size_t at = static_cast<size_t>(q) * k + i;
        s_sc[e] = __ldcg(reinterpret_cast<const double*>(b + nk8) + at);
        s_pos[e] = __ldcg(reinterpret_cast<const unsigned long long*>(b + 2 * nk8) + at);
    }
    __syncthreads();
    uint32_t total = 0;
    for (uint32_t g = 0; g < G; ++g) total += s_cnt[g];
    const uint32_t cnt = total < k ? total : k;
    // ---- merge: rank(e) = own index + Σ_{other lists} #entries better than e (binary search) ----------
    for (uint32_t e = tid; e < G * k; e += blockDim.x) {
        const uint32_t g = e / k, i = e - g * k;
        if (i >= s_cnt[g]) continue;
        const double sc = s_sc[e];
        const uint64_t pp = s_pos[e];
        uint32_t rank = i;
        for (uint32_t g2 = 0; g2 < G; ++g2) {
            if (g2 == g) continue;
            uint32_t lo = 0, hi = s_cnt[g2];
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (better(s_sc[g2 * k + mid], s_pos[g2 * k + mid], sc, pp)) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < cnt) {
            const size_t o = static_cast<size_t>(q) * k + rank;
            const char* b = m.slot + static_cast<size_t>(g) * m.blk;
            m.out_ids[o] = __ldcg(reinterpret_cast<const unsigned long long*>(b) + static_cast<size_t>(q) * k + i);
            m.out_scores[o] = sc;
            if (m.out_pos) m.out_pos[o] = pp;
        }
    }
    for (uint32_t i = cnt + tid; i < k; i += blockDim.x) {
        const size_t o = static_cast<size_t>(q) * k + i;
        m.out_ids[o] = ~0ull;
        m.out_scores[o] = 0.0;
        if (m.out_pos) m.out_pos[o] = ~0ull;
    }
    __syncthreads();   // every read of the slot is done
    if (tid == 0) {
        m.out_counts[q] = cnt;
        // a block that never arrived leaves the merged list incomplete: raise the "do not use" bit with it
        m.out_flags[q] = s_flags | (s_fail ? (FLAG_EXCHANGE | FLAG_CERT_FAIL) : 0u);
    }
    // ---- acknowledge: peer g may overwrite its block of this slot (next use) --------------------------
    if (tid < G && tid != m.self) red_add_release_sys(m.ack[tid], 1u);
    pdl_wait();   // no-op unless launched with the PDL attribute: do not complete before the finalize has
}

cudaError_t launch_exchange_merge(const ExchangeMerge& m, bool pipelined, cudaStream_t s) {
    if (m.nq == 0 || m.k == 0) return cudaSuccess;
    const size_t smem = static_cast<size_t>(m.G) * m.k * 16 + static_cast<size_t>(m.G) * 4;
    // small lists: a 64-thread CTA (< 4K registers) fits beside three resident scan CTAs, so a merge that
    // is still waiting for its peers does not take an SM slot away from the next search's scan
    const unsigned threads = m.G * m.k <= 128 ? 64u : 256u;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(m.nq);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pipelined ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, exchange_merge_kernel, m);
}

}  // namespace vl
