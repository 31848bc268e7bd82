// Shared host/device helpers for libosb (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/osb.h"
#include "../../include/osb_debug.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libosb is written for sm_100a (B200) only"
#endif

namespace osb {

// ---- host-side error plumbing --------------------------------------------------------------
void set_error(const char* fmt, ...);
int  check_device();                 // OSB_OK or OSB_ERR_ARCH (cached per device)
int  sm_count();                     // cached
int  cta_limit(int max_ctas);         // sm_count(), or the caller's smaller bound (max_ctas > 0)

#define OSB_REQUIRE(cond, code, ...)                 \
    do {                                             \
        if (!(cond)) {                               \
            ::osb::set_error(__VA_ARGS__);           \
            return (code);                           \
        }                                            \
    } while (0)

#define OSB_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t e__ = (expr);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            ::osb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorName(e__),      \
                             __FILE__, __LINE__);                                        \
            return OSB_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

#define OSB_LAUNCH_CHECK() OSB_CUDA(cudaGetLastError())

#define OSB_ENTRY_GUARD()                         \
    do {                                          \
        int rc__ = ::osb::check_device();         \
        if (rc__ != OSB_OK) return rc__;          \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- device helpers --------------------------------------------------------------------------
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2   = 0.6931471805599453f;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_log2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <typename T> __device__ __forceinline__ float to_float(T v);
template <> __device__ __forceinline__ float to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
// (value, index) arg-max with numpy semantics: larger value wins, ties go to the smaller index.
__device__ __forceinline__ void argmax_combine(float& v, int& i, float ov, int oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}
template <int WIDTH>
__device__ __forceinline__ void group_argmax(float& v, int& i) {
#pragma unroll
    for (int o = WIDTH / 2; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(kFull, v, o);
        int   oi = __shfl_xor_sync(kFull, i, o);
        argmax_combine(v, i, ov, oi);
    }
}

// Order-preserving float32 -> uint32 key (np.unique / '>' compare floats, so -0.0 == +0.0).
__device__ __forceinline__ uint32_t float_to_key(float f) {
    uint32_t b = __float_as_uint(f);
    if (b == 0x80000000u) b = 0u;
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}
constexpr uint32_t kAbsentKey = 0xffffffffu;

// Deterministic "last block finishes" reduction of per-block double partials.
// partials[nslots * gridDim.x] laid out slot-major; the block that draws the last ticket sums every
// slot in block-index order (fixed order => bit-reproducible) and returns true on thread 0 with
// the totals in out[].  The ticket counter is reset so the workspace can be reused without a memset
// once it has been zero-initialised.
template <int NSLOTS>
__device__ __forceinline__ bool last_block_reduce(const double (&mine)[NSLOTS], double* partials,
                                                  unsigned int* ticket, double (&out)[NSLOTS]) {
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NSLOTS; ++s) partials[(size_t)s * gridDim.x + blockIdx.x] = mine[s];
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    // one warp sums in a fixed order: lane l takes blocks l, l+32, ... then a fixed xor tree
    bool done = false;
    if (threadIdx.x < 32) {
#pragma unroll
        for (int s = 0; s < NSLOTS; ++s) {
            double acc = 0.0;
            for (unsigned b = threadIdx.x; b < gridDim.x; b += 32)
                acc += __ldcg(&partials[(size_t)s * gridDim.x + b]);
            out[s] = warp_sum(acc);
        }
        if (threadIdx.x == 0) { *ticket = 0u; done = true; }
    }
    return done;
}

}  // namespace osb
