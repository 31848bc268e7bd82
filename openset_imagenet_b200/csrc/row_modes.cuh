// How a label selects a row's target distribution, shared by the row-wise loss kernel and the fused
// head epilogue so both classify labels identically.
#pragma once
#include "osb_common.cuh"

namespace osb {

enum : int { kRowIgnored = 0, kRowOneHot = 1, kRowUniform = 2, kRowBad = 3 };

struct RowMode {
    int   mode;
    float omega;   // one-hot: row weight (1 or class_w[y]);  uniform: S = sum_c T[i,c] = C * float32(w/C)
};

// entropic (losses.py:18-28): y < 0 -> uniform target, y in [0,C) -> one-hot, y >= C -> IndexError.
// softmax  (train.py:343)   : y == ignore_index -> ignored, else must be a column.
// garbage  (train.py:347)   : same with row weight class_w[y] (default ignore_index -100).
__device__ __forceinline__ RowMode classify_row(int variant, int64_t y, int C, int64_t ignore_index,
                                                const float* __restrict__ class_w, float s_uniform) {
    RowMode r;
    r.mode = kRowBad;
    r.omega = 0.f;
    if (variant == OSB_ENTROPIC) {
        if (y < 0) { r.mode = kRowUniform; r.omega = s_uniform; }
        else if (y < C) { r.mode = kRowOneHot; r.omega = 1.f; }
    } else {
        if (y == ignore_index) r.mode = kRowIgnored;
        else if (y >= 0 && y < C) {
            r.mode = kRowOneHot;
            r.omega = (variant == OSB_GARBAGE) ? class_w[y] : 1.f;
        }
    }
    return r;
}

}  // namespace osb
