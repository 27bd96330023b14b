// rescore.cuh — phases (2)-(4) shared by the single-query finalize and the batched pipeline:
// exact f64 re-score in reference summation order, final (score desc, position asc) order,
// optimality certificate.  See flat_finalize.cu for the argument.
#pragma once
#include "kernels.h"

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
    uint64_t* early = nullptr;   // [nq][EARLY_STRIDE] published by the scan; cleared by the finalize kernel
    uint64_t* out_ids;
    double* out_scores;
    uint64_t* out_pos;
    uint32_t* out_counts;
    uint32_t* out_flags;
    double eps_scale;  // multiplies the per-term rounding unit (1 = fp32 scan)
    double tc_abs;     // > 0: bf16 tensor-core scan, |approx − exact| <= tc_abs·‖x‖·‖q‖ (absolute)
    const uint32_t* e_x = nullptr;   // tensor-core batches: float bits of max over rows of ‖x̃−x̂‖/‖x̂‖ (measured, batch_tc.cu)
    const float* e_q = nullptr;      // tensor-core batches: [nq] ‖q̃−q‖/‖q‖ (measured); null: the query was not rounded
    const uint32_t* e_x1 = nullptr;  // mirror scans, manhattan: float bits of max over rows of ‖x̃−x‖₁ (measured)
    int kp_base = 0;                 // > 0: also report (FLAG_BASE_OK) whether the first kp_base candidates alone certify
    PeerPush peers;    // G > 0: mirror the results into every peer shard's exchange slot (NVLink stores)
};

// result store: local, plus the same offset inside every peer's copy of this shard's block
template <typename T>
__device__ __forceinline__ void out_store(const FinalizeParams& p, T* ptr, T val, uint32_t skip_mask = 0u) {
    *ptr = val;
    for (uint32_t g = 0; g < p.peers.G; ++g)
        if (g != p.peers.self && !((skip_mask >> g) & 1u))
            *reinterpret_cast<T*>(reinterpret_cast<char*>(ptr) + p.peers.delta[g]) = val;
}

__device__ __forceinline__ double sim_from_l2(double ss) {  // lib.rs:485-488
    return __ddiv_rn(1.0, __dadd_rn(1.0, __dsqrt_rn(ss)));
}
__device__ __forceinline__ double sim_from_l1(double s) {   // lib.rs:528-531
    return __ddiv_rn(1.0, __dadd_rn(1.0, s));
}


// Phases (3)-(4) + the exchange hand-shake.  Expects s_exact[0..nc) (exact scores), s_pos[0..nc), s_keys[0..nc)
// (sorted by key, descending), *s_qnorm = ‖q‖ and *s_nan already written and made visible by a barrier.
// All threads of the block call; blockDim.x >= nc.
__device__ __forceinline__ void rank_and_certify(const FinalizeParams& p, uint32_t qi, int nc, const uint64_t* s_keys,
                                                 const double* s_exact, const uint32_t* s_pos, const double* s_qnorm,
                                                 const int* s_nan, uint32_t extra_flags) {
    __shared__ double s_kth;
    const int tid = threadIdx.x;
    QueryCtl* ctl = p.ctl ? p.ctl + qi : nullptr;
    // ---- exchange: the peers must have finished reading the previous use of this slot ----------
    // A peer that has not acknowledged in time may still be reading the slot: its copy is NOT overwritten (it will
    // see no stamp for this search and flag the query itself), and the query is flagged here.
    __shared__ unsigned int s_xfail;   // bit g: peer g never released the slot
    if (p.peers.G) {
        if (tid == 0) s_xfail = 0u;
        __syncthreads();
        if (tid < static_cast<int>(p.peers.G) && tid != static_cast<int>(p.peers.self) && p.peers.stamp > 1u)
            if (!wait_stamp(p.peers.ack[tid], p.peers.ack_want, EXCH_TIMEOUT_NS)) atomicOr(&s_xfail, 1u << tid);
        __syncthreads();
    }
    const uint32_t skip = p.peers.G ? s_xfail : 0u;

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
            out_store<uint64_t>(p, p.out_ids + o, p.ids ? p.ids[mp] : p.id_base + mp, skip);
            out_store<double>(p, p.out_scores + o, me, skip);
            if (p.out_pos) out_store<uint64_t>(p, p.out_pos + o, p.pos_base + mp, skip);
            if (rank == cnt - 1) s_kth = me;
        }
    }
    for (int i = cnt + tid; i < static_cast<int>(p.k); i += static_cast<int>(blockDim.x)) {
        const size_t o = static_cast<size_t>(qi) * p.k + i;
        out_store<uint64_t>(p, p.out_ids + o, ~0ull, skip);
        out_store<double>(p, p.out_scores + o, 0.0, skip);
        if (p.out_pos) out_store<uint64_t>(p, p.out_pos + o, ~0ull, skip);
    }
    __syncthreads();

    // ---- (4) certificate ---------------------------------------------------------------
    if (tid == 0) {
        uint32_t flags = (ctl ? ctl->flags : 0u) | extra_flags;
        if (*s_nan) flags |= FLAG_NAN;
        const bool excluded_exist = p.n > static_cast<uint32_t>(nc);
        if (excluded_exist && cnt > 0) {
            const double u = 5.9604644775390625e-08 * p.eps_scale;  // 2^-24 × scale
            const double nn = static_cast<double>(p.pitch);
            const double kth = s_kth;
            const double qn = *s_qnorm;
            const double maxn = sqrt(__longlong_as_double(p.stats->max_norm_sq_bits));
            // bf16 scans.  Round-to-nearest moves an element by at most 2^-8 relative (a value just above a power of
            // two), ~2^-9.5 RMS.  Worst-case constants (p.tc_abs): single-query mirror scans round ONE operand (fp32
            // query), |Σx̃q − Σxq| <= 2^-8·‖x‖‖q‖ = 0.00391 (+ fp32 accumulation) < 0.0040; tensor-core batches round
            // BOTH, ((1+2^-8)² − 1)·‖x‖‖q‖ = 0.00783 (+ K·2^-23) < 0.0079 (experiments/adversarial_bf16_rounding.py
            // builds the input that needs them).  MEASURED bound: Cauchy–Schwarz with the rounding-error norms that
            // actually occurred, E_x = max over rows of ‖x̃−x̂‖/‖x̂‖ (kept while the mirror is built, batch_tc.cu),
            // E_q = ‖q̃−q‖/‖q‖ of this query (0 when the query stays fp32): |x̃·q̃ − x̂·q| <= (E_x + E_q + E_x·E_q)·‖x̂‖‖q‖,
            // plus the fp32 accumulation and, for the pre-normalised cosine mirror, the fp32 rounding of x·(1/‖x‖)
            // (pitch·2^-23 covers both).  Both bounds are valid: the smaller one is used (≈ 0.0017 per rounded
            // operand on real-valued data, the constant on adversarial data).
            double tc_abs = p.tc_abs;
            if (p.tc_abs > 0.0 && p.e_x) {
                const double ex = static_cast<double>(__uint_as_float(*p.e_x));
                const double eq = p.e_q ? static_cast<double>(p.e_q[qi]) : 0.0;
                const double measured = (ex + eq + ex * eq) * 1.0001 + nn * 1.1920928955078125e-07 + 2e-7;
                tc_abs = measured < tc_abs ? measured : tc_abs;
            }
            // every excluded row has approximate score <= worst (in scan units): can one of them reach kth?
            auto holds = [&](double worst) -> bool {
                if (p.tc_abs > 0.0) {
                    if (p.metric == COSINE) {       // rows pre-normalised: scan units are cos·‖q‖
                        return qn >= 1e-15 && kth > worst / qn + tc_abs;
                    } else if (p.metric == DOT) {
                        return kth > worst + tc_abs * maxn * qn + 1e-30;
                    } else if (p.metric == MANHATTAN) {
                        // −Σ|x̃−q| with fp32 query: rounding the rows moves the sum by at most Σ|x̃−x| = ‖x̃−x‖₁ — measured
                        // (max over rows, e_x1) or <= 2^-8·‖x‖₁ <= 2^-8·√dim·‖x‖ in the worst case; the fp32 sum of
                        // non-negative terms adds a RELATIVE (nn+2)·u.  Every excluded row has exact Σ|x−q| >= L.
                        double e1 = 0.00390625 * 1.01 * sqrt(static_cast<double>(p.dim)) * maxn;
                        if (p.e_x1) {
                            const double m1 = static_cast<double>(__uint_as_float(*p.e_x1)) * 1.0001;
                            e1 = m1 < e1 ? m1 : e1;
                        }
                        double L = (-worst) * (1.0 - (nn + 2.0) * u * 1.01) - e1 - 1e-30;
                        L = L > 0.0 ? L * (1.0 - 1e-12) : 0.0;
                        return kth > sim_from_l1(L);
                    } else {                        // −‖x−q‖² from ‖x‖² + ‖q‖² − 2x·q
                        double L = (-worst) - 2.0 * tc_abs * maxn * qn - 2e-6 * (maxn * maxn + qn * qn) - 1e-36;
                        L = L > 0.0 ? L * (1.0 - 1e-12) : 0.0;
                        return kth > sim_from_l2(L);
                    }
                } else if (p.metric == COSINE) {
                    // |fl32(dot)·fl32(1/‖a‖) − dot/‖a‖| <= ((nn+8)·u)·‖q‖ ; cosine = that / ‖q‖
                    const double min_nz = __longlong_as_double(p.stats->min_nz_norm_sq_bits);
                    const bool scale_ok = qn >= 1e-15 && !(min_nz < 1e-30);
                    const double bound = worst / qn + (nn + 8.0) * u * 1.01 + 1e-30;
                    return scale_ok && kth > bound;
                } else if (p.metric == DOT) {
                    const double bound = worst + (nn + 2.0) * u * 1.01 * maxn * qn + 1e-30;
                    return kth > bound;
                } else if (p.metric == EUCLIDEAN) {
                    // Σ(a−q)² has only non-negative terms → RELATIVE error <= (nn+4)·u
                    double L = (-worst) * (1.0 - (nn + 4.0) * u * 1.01) - 1e-36;
                    L = L > 0.0 ? L * (1.0 - 1e-12) : 0.0;
                    return kth > sim_from_l2(L);
                } else {
                    double L = (-worst) * (1.0 - (nn + 2.0) * u * 1.01) - 1e-36;
                    L = L > 0.0 ? L * (1.0 - 1e-12) : 0.0;
                    return kth > sim_from_l1(L);
                }
            };
            const bool ok = holds(static_cast<double>(key_score(s_keys[nc - 1])));
            if (!ok) flags |= FLAG_CERT_FAIL;
            // adaptive over-selection (api.cu): would the first kp_base candidates alone have certified this top-k?
            if (p.kp_base > 0 && ok &&
                (nc <= p.kp_base || holds(static_cast<double>(key_score(s_keys[p.kp_base - 1])))))
                flags |= FLAG_BASE_OK;
        } else if (p.kp_base > 0) {
            flags |= FLAG_BASE_OK;
        }
        if (flags & (FLAG_NONFINITE | FLAG_OVERFLOW)) flags |= FLAG_CERT_FAIL;
        if (skip) flags |= FLAG_EXCHANGE | FLAG_CERT_FAIL;   // a peer never released the slot
        out_store<uint32_t>(p, p.out_counts + qi, static_cast<uint32_t>(cnt), skip);
        out_store<uint32_t>(p, p.out_flags + qi, flags, skip);
        if (p.peers.G) {
            // every thread's remote stores precede the barrier above; this fence makes them (and the two
            // stores just issued) visible system-wide before the stamps that announce them
            __threadfence_system();
            for (uint32_t g = 0; g < p.peers.G; ++g)
                if (!((skip >> g) & 1u)) st_release_sys(p.peers.ready[g] + p.peers.q_off + qi, p.peers.stamp);
        }
        if (ctl) {  // re-arm the control block for the next search on this slot
            ctl->tau = 0ull;
            ctl->flags = 0u;
            ctl->done = 0u;
        }
    }
}


// Shared-memory carve-up expected by rescore_rank_certify (dynamic smem of the calling kernel):
//   keys[SCAN_CAP] u64 | exact[KP_MAX] f64 | pos[KP_MAX] u32 | q[CH] f32 | tile[Kp][CH+1] f32
// s_keys[0..nc) must hold the candidates sorted by key, descending.  All FIN_THREADS threads call.
__device__ __forceinline__ void rescore_rank_certify(const FinalizeParams& p, uint32_t qi, int nc,
                                                     uint64_t* s_keys, double* s_exact, uint32_t* s_pos,
                                                     float* s_q, float* s_tile, uint32_t extra_flags) {
    __shared__ double s_qnorm;
    __shared__ int s_nan;
    const int tid = threadIdx.x;
    QueryCtl* ctl = p.ctl ? p.ctl + qi : nullptr;
    if (tid == 0) s_nan = 0;
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
        if (p.metric != COSINE && p.metric != MANHATTAN && tid == FIN_THREADS - 1) {
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
    if (p.metric != COSINE && p.metric != MANHATTAN && tid == FIN_THREADS - 1) s_qnorm = __dsqrt_rn(qn2);
    __syncthreads();

    rank_and_certify(p, qi, nc, s_keys, s_exact, s_pos, &s_qnorm, &s_nan, extra_flags);
}

}  // namespace vl
