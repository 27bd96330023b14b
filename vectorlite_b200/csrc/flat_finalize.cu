// flat_finalize.cu — merge, exact fp64 rescore, final order and optimality certificate.
//
// One CTA per query.  (1) Merges the per-CTA candidate lists of flat_scan.cu into the global
// top-K' by approximate key.  (2) Re-scores those K' rows in f64 with the reference's exact
// arithmetic — sequential accumulation in index order, no FMA — i.e. src/lib.rs:425-444
// (cosine), 476-489 (euclidean), 521-532 (manhattan), 565-572 (dot) evaluated on the stored f32
// values widened to f64.  (3) Orders them as the stable descending sort of flat.rs:116 does
// (score desc, storage position asc; ±0 equal).  (4) Certifies that no row outside the K'
// candidates can reach the top-k: every excluded row has approximate score <= the worst kept
// one, and the fp32 evaluation error is bounded (see bound_* below), so when
// exact_kth > bound(worst kept) strictly, the returned ids and scores are exactly the
// reference's.  Otherwise FLAG_CERT_FAIL is raised and the host re-runs the query on the exact
// path (exact.cu) — still on the GPU.
#include "kernels.h"
#include "rescore.cuh"
#include "topk.cuh"

namespace vl {

__global__ void __launch_bounds__(FIN_THREADS, 1) flat_finalize_kernel(FinalizeParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: keys[SCAN_CAP] u64 | exact[KP_MAX] f64 | cpos[KP_MAX] u32 | qs[CH] f32 | tile[Kp][CH+1] f32
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(smem_raw);
    double* s_exact = reinterpret_cast<double*>(s_keys + SCAN_CAP);
    uint32_t* s_pos = reinterpret_cast<uint32_t*>(s_exact + KP_MAX);
    float* s_q = reinterpret_cast<float*>(s_pos + KP_MAX);
    float* s_tile = s_q + p.CH;
    __shared__ int s_count;

    const int tid = threadIdx.x;
    const uint32_t qi = blockIdx.x;
    const int Kp = p.Kp;
    QueryCtl* ctl = p.ctl + qi;
    pdl_launch_dependents();  // the next query's scan may start now (it never reads our outputs)
    pdl_wait();               // the scan that produced our candidates is complete and visible

    // ---- (1) merge -------------------------------------------------------------------
    // The Kp-th largest of the per-CTA maxima (tau1) is a lower bound of the global Kp-th best key
    // (Kp distinct rows reach it); exactly Kp lists have a maximum >= tau1, so only those Kp·Kp
    // slots can hold a member of the global top-Kp, and only ~Kp entries actually pass tau1.
    CtaTopK<SCAN_CAP, FIN_THREADS> topk{s_keys, &s_count};
    __shared__ unsigned long long s_max[SCAN_CAP];
    __shared__ uint32_t s_ccount[SCAN_CAP];
    __shared__ uint16_t s_qual[KP_MAX];
    __shared__ unsigned long long s_tau1;
    __shared__ int s_overflow, s_nz;
    const int G = p.grid_x;  // <= SCAN_CAP (host guarantees)
    const uint32_t slots = static_cast<uint32_t>(G) * Kp;
    const uint64_t* cand = p.cand + static_cast<size_t>(qi) * slots;
    const uint32_t* ccount = p.cand_count + static_cast<size_t>(qi) * G;
    const uint64_t* cmax = p.cand_max + static_cast<size_t>(qi) * G;
    if (tid == 0) { s_overflow = 0; s_count = 0; s_tau1 = 0ull; s_nz = 0; }
    if (p.early)   // the scan is complete: clear what it published so the slot is all-zero for its next user
        for (int i = tid; i < G; i += FIN_THREADS) p.early[static_cast<size_t>(qi) * EARLY_STRIDE + i] = 0ull;
    __syncthreads();
    {
        int nz = 0;
        for (int i = tid; i < G; i += FIN_THREADS) {
            const unsigned long long m = cmax[i];
            s_max[i] = m;
            s_ccount[i] = ccount[i];
            nz += m != 0ull;
        }
        if (nz) atomicAdd(&s_nz, nz);
    }
    __syncthreads();
    const int nqual = min(Kp, s_nz);
    for (int i = tid; i < G; i += FIN_THREADS) {  // rank by counting: no barriers, broadcast reads
        const unsigned long long me = s_max[i];
        if (me == 0ull) continue;
        int r = 0;
        for (int j = 0; j < G; ++j) r += s_max[j] > me;
        if (r < Kp) s_qual[r] = static_cast<uint16_t>(i);
        if (r == Kp - 1) s_tau1 = me;
    }
    __syncthreads();
    unsigned long long tau = ctl->tau;  // every key of the global top-K' is >= tau
    tau = s_tau1 > tau ? s_tau1 : tau;
    {   // one pass over the qualifying lists: coalesced, barrier-free, 8 independent loads in flight
        constexpr int U = 8;
        const uint32_t qslots = static_cast<uint32_t>(nqual) * Kp;
        for (uint32_t s0 = 0; s0 < qslots; s0 += FIN_THREADS * U) {
            unsigned long long key[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t s = s0 + u * FIN_THREADS + tid;
                key[u] = 0ull;
                if (s < qslots) {
                    const uint32_t l = s / Kp, e = s - l * Kp;
                    const uint32_t c = s_qual[l];
                    if (e < s_ccount[c]) key[u] = cand[static_cast<size_t>(c) * Kp + e];
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (key[u] != 0ull && key[u] >= tau) {
                    const int i = atomicAdd(&s_count, 1);
                    if (i < SCAN_CAP) s_keys[i] = key[u]; else s_overflow = 1;
                }
            }
        }
    }
    __syncthreads();
    if (s_overflow) {  // rare: more than SCAN_CAP entries above tau1 → bounded rounds with compaction
        __syncthreads();
        topk.init();
        for (uint32_t s0 = 0; s0 < slots; s0 += FIN_THREADS) {
            __syncthreads();
            if (s_count > SCAN_CAP - FIN_THREADS) {
                const unsigned long long t = topk.compact(Kp, false);
                tau = t > tau ? t : tau;
            }
            const uint32_t s = s0 + tid;
            if (s < slots) {
                const uint32_t c = s / Kp, e = s - c * Kp;
                if (e < s_ccount[c]) {
                    const unsigned long long key = cand[s];
                    if (key >= tau) topk.push(key);
                }
            }
        }
        topk.compact(Kp, true);  // sorted descending, count <= Kp
    } else {
        // rank-sort the survivors (typically ~Kp..2Kp of them) by counting: no barrier-laden network
        const int n = s_count;
        unsigned long long mine[2];
        int rk[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = tid + u * FIN_THREADS;
            rk[u] = -1;
            if (i < n) {
                mine[u] = s_keys[i];
                int r = 0;
                for (int j = 0; j < n; ++j) r += s_keys[j] > mine[u];
                rk[u] = r;
            }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 2; ++u)
            if (rk[u] >= 0 && rk[u] < Kp) s_keys[rk[u]] = mine[u];
        if (tid == 0) s_count = min(n, Kp);
        __syncthreads();
    }
    rescore_rank_certify(p, qi, s_count, s_keys, s_exact, s_pos, s_q, s_tile, 0u);
}

static size_t finalize_smem(int Kp, int CH) {
    return SCAN_CAP * sizeof(uint64_t) + KP_MAX * sizeof(double) + KP_MAX * sizeof(uint32_t) +
           static_cast<size_t>(CH) * sizeof(float) + static_cast<size_t>(Kp) * (CH + 1) * sizeof(float);
}

cudaError_t launch_flat_finalize(const FlatView& v, const float* d_queries, uint32_t nq, uint32_t k,
                                 int metric, const ScanWork& w, const SearchOut& out, float eps_scale,
                                 cudaStream_t s, const CertAux& aux) {
    // choose the staging chunk so that the tile fits in ~180 KB of shared memory
    const size_t budget = 180 * 1024;
    int CH = static_cast<int>((v.dim + 3) / 4 * 4);
    while (CH > 4 && finalize_smem(w.Kp, CH) > budget) CH = (CH / 2 + 3) / 4 * 4;
    const size_t smem = finalize_smem(w.Kp, CH);
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(flat_finalize_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr_set[dev & 63] = true;
    }
    FinalizeParams p;
    p.rows = v.rows; p.ids = v.ids; p.stats = v.stats; p.queries = d_queries;
    p.id_base = v.id_base; p.pos_base = v.pos_base;
    p.n = v.n; p.dim = v.dim; p.pitch = v.pitch; p.k = k;
    p.metric = metric; p.Kp = w.Kp; p.grid_x = w.grid_x; p.CH = CH;
    p.cand = w.cand; p.cand_count = w.cand_count; p.cand_max = w.cand_max; p.ctl = w.ctl;
    p.early = w.early;
    p.out_ids = out.ids; p.out_scores = out.scores; p.out_pos = out.pos;
    p.out_counts = out.counts; p.out_flags = out.flags;
    p.eps_scale = eps_scale;
    p.tc_abs = aux.tc_abs; p.e_x = aux.e_x; p.e_q = aux.e_q; p.e_x1 = aux.e_x1; p.kp_base = aux.kp_base;
    p.peers = out.peers;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nq);
    cfg.blockDim = dim3(FIN_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;  // always: launch latency overlaps the scan's tail; pdl_wait() orders the data
    return cudaLaunchKernelEx(&cfg, flat_finalize_kernel, p);
}

}  // namespace vl
