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
#include "topk.cuh"

namespace vl {

struct FinalizeParams {
    const float* rows;
    const uint64_t* ids;
    const ArenaStats* stats;
    const float* queries;
    uint64_t id_base, pos_base;
    uint32_t n, dim, pitch, k;
    int metric, Kp, grid_x, CH;  // CH = columns staged per chunk (multiple of 4)
    const uint64_t* cand;
    const uint32_t* cand_count;
    const uint64_t* cand_max;
    QueryCtl* ctl;
    uint64_t* out_ids;
    double* out_scores;
    uint64_t* out_pos;
    uint32_t* out_counts;
    uint32_t* out_flags;
    double eps_scale;  // multiplies the per-term rounding unit (1 = fp32 scan, larger for bf16)
};

__device__ __forceinline__ double sim_from_l2(double ss) {  // lib.rs:485-488
    return __ddiv_rn(1.0, __dadd_rn(1.0, __dsqrt_rn(ss)));
}
__device__ __forceinline__ double sim_from_l1(double s) {   // lib.rs:528-531
    return __ddiv_rn(1.0, __dadd_rn(1.0, s));
}

__global__ void __launch_bounds__(FIN_THREADS, 1) flat_finalize_kernel(FinalizeParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: keys[SCAN_CAP] u64 | exact[KP_MAX] f64 | cpos[KP_MAX] u32 | qs[CH] f32 | tile[Kp][CH+1] f32
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(smem_raw);
    double* s_exact = reinterpret_cast<double*>(s_keys + SCAN_CAP);
    uint32_t* s_pos = reinterpret_cast<uint32_t*>(s_exact + KP_MAX);
    float* s_q = reinterpret_cast<float*>(s_pos + KP_MAX);
    float* s_tile = s_q + p.CH;
    __shared__ int s_count;
    __shared__ double s_kth, s_qnorm;
    __shared__ int s_nan;

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
    if (tid == 0) { s_nan = 0; s_overflow = 0; s_count = 0; s_tau1 = 0ull; s_nz = 0; }
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
    const int nc = s_count;
    if (tid < nc) s_pos[tid] = key_pos(s_keys[tid]);
    __syncthreads();

    // ---- (2) exact f64 rescore, reference summation order --------------------------------
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;  // cosine: dot, Σx², Σy²; others: a0 only
    double qn2 = 0.0;                     // ‖q‖² for the dot-product bound (last thread)
    const float* q = p.queries + static_cast<size_t>(qi) * p.pitch;
    const int CH = p.CH, TS = CH + 1;
    for (uint32_t c0 = 0; c0 < p.dim; c0 += CH) {
        const int w = min(static_cast<uint32_t>(CH), p.dim - c0);   // live columns in this chunk
        const int w4 = (w + 3) >> 2;                                // float4s (pitch is padded)
        {
            constexpr int U = 4;
            const int total = nc * w4;
            for (int i0 = 0; i0 < total; i0 += FIN_THREADS * U) {
                float4 v[U];
                int rr[U], cc[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = i0 + u * FIN_THREADS + tid;
                    rr[u] = -1;
                    if (i < total) {
                        rr[u] = i / w4;
                        cc[u] = i - rr[u] * w4;
                        v[u] = *reinterpret_cast<const float4*>(
                            p.rows + static_cast<size_t>(s_pos[rr[u]]) * p.pitch + c0 + cc[u] * 4);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (rr[u] >= 0) {
                        float* t = s_tile + rr[u] * TS + cc[u] * 4;
                        t[0] = v[u].x; t[1] = v[u].y; t[2] = v[u].z; t[3] = v[u].w;
                    }
                }
            }
        }
        for (int i = tid; i < w; i += FIN_THREADS) s_q[i] = q[c0 + i];
        __syncthreads();
        if (tid < nc) {
            const float* t = s_tile + tid * TS;
            if (p.metric == COSINE) {
                for (int j = 0; j < w; ++j) {
                    const double x = static_cast<double>(t[j]), y = static_cast<double>(s_q[j]);
                    a0 = __dadd_rn(a0, __dmul_rn(x, y));
                    a1 = __dadd_rn(a1, __dmul_rn(x, x));
                    a2 = __dadd_rn(a2, __dmul_rn(y, y));
                }
            } else if (p.metric == EUCLIDEAN) {
                for (int j = 0; j < w; ++j) {
                    const double d = __dsub_rn(static_cast<double>(t[j]), static_cast<double>(s_q[j]));
                    a0 = __dadd_rn(a0, __dmul_rn(d, d));
                }
            } else if (p.metric == MANHATTAN) {
                for (int j = 0; j < w; ++j) {
                    const double d = __dsub_rn(static_cast<double>(t[j]), static_cast<double>(s_q[j]));
                    a0 = __dadd_rn(a0, fabs(d));
                }
            } else {
                for (int j = 0; j < w; ++j)
                    a0 = __dadd_rn(a0, __dmul_rn(static_cast<double>(t[j]), static_cast<double>(s_q[j])));
            }
        }
        if (p.metric == DOT && tid == FIN_THREADS - 1) {
            for (int j = 0; j < w; ++j) {
                const double y = static_cast<double>(s_q[j]);
                qn2 = __dadd_rn(qn2, __dmul_rn(y, y));
            }
        }
        __syncthreads();
    }
    if (tid < nc) {
        double sc;
        if (p.metric == COSINE) {
            const double na = __dsqrt_rn(a1), nb = __dsqrt_rn(a2);
            sc = (na == 0.0 || nb == 0.0) ? 0.0 : __ddiv_rn(a0, __dmul_rn(na, nb));
            if (tid == 0) s_qnorm = nb;
        } else if (p.metric == EUCLIDEAN) {
            sc = sim_from_l2(a0);
        } else if (p.metric == MANHATTAN) {
            sc = sim_from_l1(a0);
        } else {
            sc = a0;
        }
        s_exact[tid] = sc;
        if (sc != sc) s_nan = 1;
    }
    if (p.metric == DOT && tid == FIN_THREADS - 1) s_qnorm = __dsqrt_rn(qn2);
    __syncthreads();

    // ---- (3) final order: score desc, position asc (stable sort of flat.rs:116) -------------
    const int cnt = min(static_cast<int>(p.k), nc);
    if (tid < nc) {
        const double me = s_exact[tid];
        const uint32_t mp = s_pos[tid];
        int rank = 0;
        for (int j = 0; j < nc; ++j) {
            const double o = s_exact[j];
            rank += (o > me) || (o == me && s_pos[j] < mp);
        }
        if (rank < cnt) {
            const size_t o = static_cast<size_t>(qi) * p.k + rank;
            p.out_ids[o] = p.ids ? p.ids[mp] : p.id_base + mp;
            p.out_scores[o] = me;
            if (p.out_pos) p.out_pos[o] = p.pos_base + mp;
            if (rank == cnt - 1) s_kth = me;
        }
    }
    for (int i = cnt + tid; i < static_cast<int>(p.k); i += FIN_THREADS) {
        const size_t o = static_cast<size_t>(qi) * p.k + i;
        p.out_ids[o] = ~0ull;
        p.out_scores[o] = 0.0;
        if (p.out_pos) p.out_pos[o] = ~0ull;
    }
    __syncthreads();

    // ---- (4) certificate ---------------------------------------------------------------
    if (tid == 0) {
        uint32_t flags = ctl->flags;
        if (s_nan) flags |= FLAG_NAN;
        const bool excluded_exist = p.n > static_cast<uint32_t>(nc);
        if (excluded_exist && cnt > 0) {
            // every excluded row has approximate score <= worst (in scan units)
            const double worst = static_cast<double>(key_score(s_keys[nc - 1]));
            const double u = 5.9604644775390625e-08 * p.eps_scale;  // 2^-24 × scale
            const double nn = static_cast<double>(p.pitch);
            const double kth = s_kth;
            bool ok;
            if (p.metric == COSINE) {
                // |fl32(dot)·fl32(1/‖a‖) − dot/‖a‖| <= ((nn+8)·u)·‖q‖ ; cosine = that / ‖q‖
                const double qn = s_qnorm;
                const double min_nz = __longlong_as_double(p.stats->min_nz_norm_sq_bits);
                const bool scale_ok = qn >= 1e-15 && !(min_nz < 1e-30);
                const double bound = worst / qn + (nn + 8.0) * u * 1.01 + 1e-30;
                ok = scale_ok && kth > bound;
            } else if (p.metric == DOT) {
                const double maxn = sqrt(__longlong_as_double(p.stats->max_norm_sq_bits));
                const double bound = worst + (nn + 2.0) * u * 1.01 * maxn * s_qnorm + 1e-30;
                ok = kth > bound;
            } else if (p.metric == EUCLIDEAN) {
                // Σ(a−q)² has only non-negative terms → RELATIVE error <= (nn+4)·u
                double L = (-worst) * (1.0 - (nn + 4.0) * u * 1.01) - 1e-36;
                L = L > 0.0 ? L * (1.0 - 1e-12) : 0.0;
                ok = kth > sim_from_l2(L);
            } else {
                double L = (-worst) * (1.0 - (nn + 2.0) * u * 1.01) - 1e-36;
                L = L > 0.0 ? L * (1.0 - 1e-12) : 0.0;
                ok = kth > sim_from_l1(L);
            }
            if (!ok) flags |= FLAG_CERT_FAIL;
        }
        if (flags & FLAG_NONFINITE) flags |= FLAG_CERT_FAIL;
        p.out_counts[qi] = static_cast<uint32_t>(cnt);
        p.out_flags[qi] = flags;
        ctl->tau = 0ull;  // re-arm the control block for the next search on this slot
        ctl->flags = 0u;
        ctl->done = 0u;
    }
}

static size_t finalize_smem(int Kp, int CH) {
    return SCAN_CAP * sizeof(uint64_t) + KP_MAX * sizeof(double) + KP_MAX * sizeof(uint32_t) +
           static_cast<size_t>(CH) * sizeof(float) + static_cast<size_t>(Kp) * (CH + 1) * sizeof(float);
}

cudaError_t launch_flat_finalize(const FlatView& v, const float* d_queries, uint32_t nq, uint32_t k,
                                 int metric, const ScanWork& w, const SearchOut& out, float eps_scale,
                                 cudaStream_t s) {
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
    p.out_ids = out.ids; p.out_scores = out.scores; p.out_pos = out.pos;
    p.out_counts = out.counts; p.out_flags = out.flags;
    p.eps_scale = eps_scale;
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
