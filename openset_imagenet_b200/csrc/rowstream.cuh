// Streaming whole rows of a contiguous float32 matrix through shared memory with 1-D bulk copies
// (cp.async.bulk + mbarrier), and the max-only row reduction used on the staged rows.
// Shared by OSCR stage A (oscr.cu) and metrics.confidence (confidence.cu).
#pragma once
#include "osb_common.cuh"
#include "tc05.cuh"

namespace osb {

using tc::smem_u32;

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// One row in shared memory, LPR lanes per row:
//   m_all = max over the row (util.py:110),
//   m_lo  = max over the aligned 4-element groups (and unaligned head / tail elements) lying entirely BEFORE column lab.
// "first arg-max == lab" (util.py:109,114) is then
//   ts > m_lo  &&  ts == m_all  &&  the (<= 3) elements of lab's own 4-element group before lab are < ts,
// so the hot loop is max instructions only (no index bookkeeping).  lab = -1 leaves m_lo at -inf.
// ALIGNED (C % 4 == 0 and a 16-byte aligned chunk): every row is a whole number of float4.  Otherwise the row may
// start at any 4-byte offset and up to 3 head and 3 tail elements are read as scalars (needs LPR >= 6).
template <int LPR, bool ALIGNED>
__device__ __forceinline__ void smem_row_split_max(const float* row, int C, int lab, bool active, int sub,
                                                   float& m_lo, float& m_all, int& head_out, int& jb_out) {
    m_lo = -INFINITY;
    m_all = -INFINITY;
    int head = 0;
    if (!ALIGNED) {
        head = (int)(((16u - (smem_u32(row) & 15u)) & 15u) >> 2);
        head = head > C ? C : head;
    }
    const int nvec = (C - head) >> 2;
    const int jb = ALIGNED ? ((lab < 0 ? 0 : lab) >> 2)        // aligned groups j < jb lie entirely before lab
                           : (lab < head ? 0 : ((lab - head) >> 2));
    head_out = head; jb_out = jb;
    if (active) {
        if (!ALIGNED) {                                         // unaligned ends, exact per-element split
            const int tail0 = head + 4 * nvec;
            const int c = sub < head ? sub : tail0 + (sub - head);
            if (sub < head + (C - tail0)) {
                const float v = row[c];
                m_all = v;
                if (c < lab) m_lo = v;
            }
        }
        const float4* r4 = reinterpret_cast<const float4*>(row + head);
        // whole sweeps of LPR float4 need no bounds check (same trip count for every row when ALIGNED); one
        // predicated sweep takes the remainder
        const int full = nvec / LPR;
        int j = sub;
#pragma unroll 8
        for (int it = 0; it < full; ++it, j += LPR) {
            const float4 v = r4[j];
            const float t = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
            m_all = fmaxf(m_all, t);
            m_lo = fmaxf(m_lo, j < jb ? t : -INFINITY);
        }
        if (j < nvec) {
            const float4 v = r4[j];
            const float t = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
            m_all = fmaxf(m_all, t);
            m_lo = fmaxf(m_lo, j < jb ? t : -INFINITY);
        }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
        m_lo = fmaxf(m_lo, __shfl_xor_sync(kFull, m_lo, o));
        m_all = fmaxf(m_all, __shfl_xor_sync(kFull, m_all, o));
    }
}


// Rows per bulk-copy chunk and pipeline depth for rows of C floats plus lab_bytes of label per row:
// one sweep of `warps` warps at LPR lanes per row when that fits `stage_cap` bytes, else fewer rows (>= 8).
struct StreamPlan { bool ok; int lpr; int R; int S; size_t stage_bytes; };
static inline StreamPlan plan_row_stream(int C, size_t lab_bytes, int warps, bool aligned4, size_t budget, size_t stage_cap, int max_stages) {
    StreamPlan p = {};
    const size_t row_bytes = (size_t)C * 4;
    p.lpr = (C <= 128 && aligned4) ? 4 : (C <= 256 ? 8 : 32);
    int R = warps * (32 / p.lpr);
    while (R > 8 && (size_t)R * row_bytes > stage_cap) R >>= 1;
    p.R = R;
    p.stage_bytes = (size_t)R * (row_bytes + lab_bytes);
    p.ok = (size_t)R * row_bytes <= budget / 2;
    int S = p.ok ? (int)(budget / p.stage_bytes) : 0;
    p.S = S > max_stages ? max_stages : S;
    if (p.S < 2) p.ok = false;
    return p;
}

}  // namespace osb
