// common.cuh -- shared device/host helpers for librsb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/rsb200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "librsb200 is written for sm_100a (B200) only"
#endif

namespace rsb {

// ----------------------------------------------------------------------------- host error plumbing
void set_error(const char* fmt, ...);   // api.cu (thread-local buffer)

#define RSB_REQUIRE(cond, code, ...)                 \
    do {                                              \
        if (!(cond)) {                                \
            rsb::set_error(__VA_ARGS__);              \
            return (code);                            \
        }                                             \
    } while (0)

#define RSB_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (call);                                                         \
        if (_e != cudaSuccess) {                                                         \
            rsb::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (int32_t)_e;                                                          \
        }                                                                                \
    } while (0)

#define RSB_LAUNCH_CHECK()                                                                \
    do {                                                                                  \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess) {                                                          \
            rsb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (int32_t)_e;                                                           \
        }                                                                                 \
        rsb::note_launch();                                                               \
    } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

void note_launch();   // api.cu: statistics counter behind rsb200_launch_count()
int sm_count();   // api.cu: multiProcessorCount of the current device (cached per device)

constexpr uint32_t kNoSlot = 0xFFFFFFFFu;      // touch of the padding row / invalid id: no gradient entry
constexpr uint32_t kDirect = 0x80000000u;      // entry flag: value is the coefficient itself

// ----------------------------------------------------------------------------- device helpers
#ifdef __CUDACC__

constexpr unsigned kFull = 0xFFFFFFFFu;

__device__ __forceinline__ float4 ldg128(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}
// streaming 128-bit load for rows that are read once (no L1 allocation)
__device__ __forceinline__ float4 ldg128_stream(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg128_stream(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- L2 eviction-priority hints (createpolicy + .L2::cache_hint).  The gathered table rows are read once
// (evict_first) while the grouping metadata (CSR offsets, entry list, query matrix) is small enough to
// live in the 126 MB L2 between the kernels of one step (evict_last).
__device__ __forceinline__ uint64_t l2_policy(int kind) {   // 0 normal, 1 evict_first, 2 evict_last
    uint64_t pn, pf, pl;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pn));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pf));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pl));
    return kind == 1 ? pf : (kind == 2 ? pl : pn);
}
__device__ __forceinline__ float4 ldg128_stream_hint(const float* p, uint64_t pol) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ float4 ldg128_hint(const float* p, uint64_t pol) {
    float4 r;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ uint32_t ldg32_hint(const uint32_t* p, uint64_t pol) {
    uint32_t r;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ uint64_t ldg64_stream_hint(const uint64_t* p, uint64_t pol) {
    uint64_t r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void stg64_hint(uint64_t* p, uint64_t v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.u64 [%0], %1, %2;" :: "l"(p), "l"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void stg128_stream_hint(float* p, float4 v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

__device__ __forceinline__ float dot4(float4 a, float4 b) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ float sqdist4(float4 a, float4 b) {
    float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z, dw = a.w - b.w;
    return fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw)));
}
__device__ __forceinline__ void fma4(float4& acc, float c, float4 v) {
    acc.x = fmaf(c, v.x, acc.x); acc.y = fmaf(c, v.y, acc.y);
    acc.z = fmaf(c, v.z, acc.z); acc.w = fmaf(c, v.w, acc.w);
}

// softplus(x) = log(1 + e^x), the numerically stable form torch's logsigmoid uses:
// -logsigmoid(-x) = max(x, 0) + log1p(exp(-|x|))
__device__ __forceinline__ float softplusf(float x) {
    return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float sigmoidf(float x) {
    // stable for both signs
    float e = expf(-fabsf(x));
    float s = 1.0f / (1.0f + e);
    return x >= 0.0f ? s : e * s;
}

// ----------------------------------------------------------------------------- Philox4x32-10
// cuRAND-compatible: curand_init(seed, subsequence, offset) + curand4().
struct Philox {
    static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    __device__ __forceinline__ static uint4 round(uint4 c, uint2 k) {
        uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        return make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    }
    // counter = (ctr_lo as u64 in x,y ; subsequence as u64 in z,w), key = seed
    __device__ __forceinline__ static uint4 gen(uint64_t seed, uint64_t subsequence, uint64_t ctr_lo) {
        uint4 c = make_uint4((uint32_t)ctr_lo, (uint32_t)(ctr_lo >> 32), (uint32_t)subsequence, (uint32_t)(subsequence >> 32));
        uint2 k = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
        for (int i = 0; i < 9; ++i) { c = round(c, k); k.x += W0; k.y += W1; }
        return round(c, k);
    }
};

// cuRAND _curand_uniform: (0, 1]
__device__ __forceinline__ float curand_uniform_from_u32(uint32_t x) {
    return x * 2.3283064365386963e-10f + (2.3283064365386963e-10f / 2.0f);
}

#endif  // __CUDACC__
}  // namespace rsb
