// Streaming row reductions over a row-major float32 score matrix: LPR lanes cooperate on one row
// (32/LPR rows per warp per step), 128-bit non-allocating loads when rows are 16-byte aligned.
// Used by OSCR stage A (util.py:109-111) and confidence (metrics.py:39).
#pragma once
#include "osb_common.cuh"

namespace osb {

__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream_f1(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

__device__ __forceinline__ void upd1(float& m, int& am, float v, int c) {
    if (v > m) { m = v; am = c; }
}
__device__ __forceinline__ void upd4(float& m, int& am, const float4& v, int c) {
    upd1(m, am, v.x, c); upd1(m, am, v.y, c + 1); upd1(m, am, v.z, c + 2); upd1(m, am, v.w, c + 3);
}

// Max and FIRST arg-max (numpy argmax tie rule) of row[0..ncols).  Every lane of the warp must call
// this (the combine uses full-mask shuffles); groups whose `active` is false do no loads and get
// (-inf, INT_MAX).  Result is valid in every lane of an active group.
template <int LPR>
__device__ __forceinline__ void group_row_argmax(const float* __restrict__ row, int ncols, bool active,
                                                 bool vec4, int sub, float& m, int& am) {
    m = -INFINITY;
    am = 0x7fffffff;
    if (active) {
        if (vec4) {
            const int nvec = ncols >> 2;
            const float4* r4 = reinterpret_cast<const float4*>(row);
            int j = sub;
            for (; j + 3 * LPR < nvec; j += 4 * LPR) {       // 4 independent 16-byte loads in flight
                float4 a = ld_stream_f4(r4 + j), b = ld_stream_f4(r4 + j + LPR);
                float4 c = ld_stream_f4(r4 + j + 2 * LPR), d = ld_stream_f4(r4 + j + 3 * LPR);
                upd4(m, am, a, 4 * j); upd4(m, am, b, 4 * (j + LPR));
                upd4(m, am, c, 4 * (j + 2 * LPR)); upd4(m, am, d, 4 * (j + 3 * LPR));
            }
            for (; j < nvec; j += LPR) { float4 a = ld_stream_f4(r4 + j); upd4(m, am, a, 4 * j); }
            for (int c = (nvec << 2) + sub; c < ncols; c += LPR) upd1(m, am, ld_stream_f1(row + c), c);
        } else {
            for (int c = sub; c < ncols; c += LPR) upd1(m, am, ld_stream_f1(row + c), c);
        }
    }
    group_argmax<LPR>(m, am);
}

static inline bool rows_are_vec4(const void* base, int64_t ld) {
    return (reinterpret_cast<uintptr_t>(base) % 16 == 0) && (ld % 4 == 0);
}

}  // namespace osb
