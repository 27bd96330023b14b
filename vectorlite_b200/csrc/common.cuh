// common.cuh — shared device/host helpers for the VectorLite B200 hot path.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace vl {

enum Metric : int { COSINE = 0, EUCLIDEAN = 1, MANHATTAN = 2, DOT = 3 };

// flag bits written per query by the finalize kernels
constexpr uint32_t FLAG_CERT_FAIL = 1u;   // optimality certificate failed → re-run exact
constexpr uint32_t FLAG_NONFINITE = 2u;   // a non-finite approximate score was seen
constexpr uint32_t FLAG_NAN = 4u;         // an exact similarity is NaN (reference panics)
constexpr uint32_t FLAG_OVERFLOW = 8u;    // a candidate buffer overflowed → re-run exact
constexpr uint32_t FLAG_EXCHANGE = 16u;   // a peer's results did not arrive in time (row-sharded exchange)
constexpr uint32_t FLAG_BASE_OK = 32u;    // host searches only: the base over-selection K' alone would have certified

constexpr uint32_t INVALID_POS = 0xFFFFFFFFu;

// ---- order-preserving float ↔ unsigned maps (larger unsigned == larger float) -----------
__host__ __device__ __forceinline__ uint32_t f32_orderable(float f) {
#ifdef __CUDA_ARCH__
    const uint32_t u = __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float orderable_f32(uint32_t o) {
    const uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
__host__ __device__ __forceinline__ uint64_t f64_orderable(double d) {
#ifdef __CUDA_ARCH__
    const uint64_t u = static_cast<uint64_t>(__double_as_longlong(d));
#else
    uint64_t u;
    memcpy(&u, &d, 8);
#endif
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

// 64-bit selection key of the approximate scan: high word = orderable fp32 score, low word =
// ~position, so that "larger key" == "higher score, then EARLIER storage position" — the
// stable-sort tie-break of src/index/flat.rs:116.  Key 0 is never a real key (pos <= 2^32-2).
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t pos) {
    return (static_cast<uint64_t>(f32_orderable(score)) << 32) | (0xFFFFFFFFu - pos);
}
__host__ __device__ __forceinline__ uint32_t key_pos(uint64_t key) {
    return 0xFFFFFFFFu - static_cast<uint32_t>(key);
}
__host__ __device__ __forceinline__ float key_score(uint64_t key) {
    return orderable_f32(static_cast<uint32_t>(key >> 32));
}

// per-(workspace slot, query) control block shared by the scan CTAs
struct QueryCtl {
    unsigned long long tau;  // running lower bound of the K'-th best key over the whole grid
    unsigned int flags;
    unsigned int done;
};

// per-index statistics maintained on the device at insert time (certificate inputs)
struct ArenaStats {
    unsigned long long max_norm_sq_bits;     // max  ‖row‖² (f64 bit pattern; positive ⇒ int-ordered)
    unsigned long long min_nz_norm_sq_bits;  // min  ‖row‖² over rows with non-zero norm
    unsigned int nonfinite_rows;
    unsigned int pad;
};

// 128-bit streaming load: read-only path, do not allocate in L1 (each byte is used once)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

// Programmatic dependent launch (sm_90+): `pdl_wait` blocks until the previous kernel in the stream
// has completed and its writes are visible (no-op unless THIS kernel was launched with the
// programmatic-stream-serialization attribute); `pdl_launch_dependents` lets the next kernel in
// the stream start early (it only matters if THAT kernel carries the attribute).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- system-scope signalling for the peer-memory exchange (flags live in peer-mapped HBM) --------
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// spin until *p reaches `want` (stamps only grow); false after `limit_ns` — never hangs the GPU
__device__ __forceinline__ bool wait_stamp(const uint32_t* p, uint32_t want, unsigned long long limit_ns) {
    if (static_cast<int32_t>(ld_acquire_sys(p) - want) >= 0) return true;
    const unsigned long long t0 = global_timer_ns();
    for (;;) {
        if (static_cast<int32_t>(ld_acquire_sys(p) - want) >= 0) return true;
        if (global_timer_ns() - t0 > limit_ns) return false;
        __nanosleep(64);
    }
}
constexpr unsigned long long EXCH_TIMEOUT_NS = 2000000000ull;  // 2 s


}  // namespace vl
