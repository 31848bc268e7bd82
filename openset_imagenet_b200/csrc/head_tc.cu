// Fused open-set head on tcgen05 tensor cores (sm_100a).  Every GEMM runs on CTA pairs (cta_group::2, M = 256 MMAs).
//
//   head_fwd_kernel        : Z = X W^T (+b) accumulated in TMEM (2 x 256 columns, double-buffered), and an epilogue that
//                            never writes Z: online log-softmax statistics, entropic / softmax / garbage loss, confidence
//                            statistics, arg-max, and the un-scaled grad-logits dZ (bf16, TMA store).  Persistent; heads
//                            wider than 256 classes run as a group of up to 4 units (CTA pairs) that share the rows, split
//                            the classes and exchange the per-row softmax statistics through (value, epoch) words in L2.
//                            <..., FUSE>: the same kernel goes on with the tiles of the dX GEMM on the same pipeline state.
//   head_fwd_splitk_kernel : the same for small batches: a cluster of 8 CTAs per 128-row tile, split over K, partial
//                            logits reduce-scattered through distributed shared memory.
//   head_dx2_kernel        : dX = g * dZ W          (persistent, double-buffered TMEM accumulators; its three roles are
//                            the device functions dx_producer / dx_mma / dx_epilogue, shared with the fused kernel)
//   head_dw2_kernel        : dW = g * dZ^T X        (split over the batch; at most two splits: written by the GEMM itself,
//                            else fp32 partials + dw_reduce_kernel, deterministic either way); head_dw_kernel: C <= 128
//
// dZ is stored DENOMINATOR-FREE, omega_i (p_i - t_i), as one bf16 plane or as two (plane 1 = the bf16 rounding residual of
// plane 0; the backward GEMMs then run over both planes: K-concatenation, fp32 accumulation); g = gscale / den is applied
// in the backward epilogues, so the forward kernel never waits for the label statistics.
//
// All operands are 128B-swizzled TMA tiles; one tile layout serves as K-major or MN-major operand
// (tc05.cuh).  Warp roles per CTA: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocator, then the epilogue warps
// (TMEM lane quadrant = warp % 4): 8 of them (two threads per accumulator row) in the forward and dX kernels, 4 in dW.
//
// Reference arithmetic: model.py:23-26,37-38 (linear), losses.py:16-29 / train.py:343,347 (losses),
// train.py:177 + metrics.py:33-40 + util.py:109-111 (the per-row statistics), autograd of those.
#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include "osb_common.cuh"
#include "row_modes.cuh"
#include "tc05.cuh"

namespace osb {
namespace head {

using namespace osb::tc;

constexpr int kThreads = 192;
constexpr int kEpiWarp0 = 2;
constexpr int BM = 128;                 // accumulator rows per tile (UMMA M)
constexpr int BK = 64;                  // elements per 128-byte swizzle row
constexpr int kABytes = BM * BK * 2;    // 16 KB: 128 rows x 128 B
constexpr int kBoxBytes = 64 * BK * 2;  // 8 KB: 64 rows x 128 B
constexpr int kStageWarpBytes = 32 * 128;   // epilogue staging: 32 rows x 128 B per warp
constexpr int kMaxPartials = 2048;
constexpr size_t kHeadTicketBytes = 128;
constexpr int kMaxXid = 80;             // row blocks that exchange statistics at the same time: <= 148 / 2 CTAs
constexpr size_t kXchStatBytes = (size_t)kMaxXid * 2 * 4 * 128 * 6 * sizeof(uint2);
// workspace (zeroed once by the caller): [ticket, exchange epoch][loss partials][exchange statistics][dW partials]
constexpr size_t kWsOffEpoch = 64;
constexpr size_t kWsOffFlagEpoch = 72;
constexpr size_t kWsOffPartials = kHeadTicketBytes;
constexpr size_t kWsOffStats = kWsOffPartials + sizeof(double) * kMaxPartials;
// [dW tile flags]: one word per (output tile, CTA of the pair, epilogue warp): "split 0 has stored these rows" (direct dW)
constexpr int kMaxDwFlagTiles = 1024;
constexpr size_t kWsOffDwFlags = (kWsOffStats + kXchStatBytes + 255) / 256 * 256;
// [row-tile flags]: 8 slots (unit of the group x CTA of the pair) x 8 epilogue warps per 256-row tile of the batch: "this
// warp's grad-logits of the tile have landed" -- lets the dW GEMM behind a forward start on the SMs the forward leaves idle
constexpr int kMaxRowFlagTiles = 1024;
constexpr int kRowFlagWords = 64;
constexpr int kRowFlagTileRows = 256;
constexpr size_t kWsOffRowFlags = kWsOffDwFlags + sizeof(unsigned int) * kMaxDwFlagTiles * 8;
constexpr size_t kWsOffDw = kWsOffRowFlags + sizeof(unsigned int) * kMaxRowFlagTiles * kRowFlagWords;

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
    return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
// Programmatic dependent launch.  A kernel launched with programmatic stream serialization starts once every CTA of the
// kernel in front of it has called pdl_trigger (or exited) and runs its prologue (barrier init, TMEM allocation, descriptor
// prefetch) under that kernel's tail; pdl_wait returns when the kernel in front has completed and its writes are visible.
// Both are no-ops in a plain launch.  Nothing global is read or written before pdl_wait.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- cluster helpers (2-CTA class split) -------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_peer(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_peer_f32(uint32_t raddr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" :: "r"(raddr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_peer_v2(uint32_t raddr, float a, float b) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" :: "r"(raddr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void mbar_arrive_peer(uint32_t raddr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(raddr) : "memory");
}
// 8-byte (value, epoch) words: a reader that sees the epoch it waits for has the value that was stored with it
// (aligned 8-byte accesses are single-copy atomic), so data and flag travel in ONE L2 round trip and need no fence.
__device__ __forceinline__ void st_ll(uint2* p, float v, uint32_t epoch) {
    asm volatile("st.relaxed.gpu.global.v2.b32 [%0], {%1, %2};" :: "l"(p), "r"(__float_as_uint(v)), "r"(epoch) : "memory");
}
// the same 8-byte words in shared memory, between the two threads that own one row inside a CTA (no CTA barrier)
__device__ __forceinline__ void sts_ll(uint2* p, float v, uint32_t epoch) {
    asm volatile("st.volatile.shared.v2.b32 [%0], {%1, %2};" :: "r"(osb::tc::smem_u32(p)), "r"(__float_as_uint(v)), "r"(epoch) : "memory");
}
__device__ __forceinline__ uint2 lds_ll(const uint2* p) {
    uint2 v;
    asm volatile("ld.volatile.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(osb::tc::smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ uint2 ld_ll(const uint2* p) {
    uint2 v;
    asm volatile("ld.relaxed.gpu.global.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}

// pack 8 fp32 -> 8 bf16 and store one 16-byte chunk of a 128B-swizzled staging row
__device__ __forceinline__ void stage_chunk(uint8_t* sbuf, int row, int chunk, const float* f) {
    __nv_bfloat162 p0 = __floats2bfloat162_rn(f[0], f[1]);
    __nv_bfloat162 p1 = __floats2bfloat162_rn(f[2], f[3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(f[4], f[5]);
    __nv_bfloat162 p3 = __floats2bfloat162_rn(f[6], f[7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
    u.z = *reinterpret_cast<uint32_t*>(&p2); u.w = *reinterpret_cast<uint32_t*>(&p3);
    *reinterpret_cast<uint4*>(sbuf + row * 128 + ((chunk ^ (row & 7)) << 4)) = u;
}

constexpr int kDxBN = 256;
// =================================================================================================
// dX = g * dZ W with CTA pairs (cta_group::2): A = dZ tile (K-major over classes), B = W tile viewed MN-major (N = feature
// index); one cluster of 2 CTAs per 256 rows x 256 features, M = 256 MMAs.  The K loop runs over the classes of every dZ
// plane (plane 1 = rounding residual of plane 0, the W k-block is the same).
// CTA r of the pair loads rows [128 r, +128) of the dZ tile and features [128 r, +128) of the W tile (32 KB per stage
// instead of 48 KB), the leader issues the MMAs, each CTA drains its own 128 x 256 accumulator from its TMEM.
// =================================================================================================
// kDx2EW warps per TMEM lane quadrant drain the accumulator.
constexpr int kDx2EW = 2;
constexpr int kDx2Threads = (2 + 4 * kDx2EW) * 32;
constexpr int kDx2Stages = 5;
constexpr int kDx2StageBytes = kABytes + 2 * kBoxBytes;         // 32 KB per CTA
constexpr int kDx2OffStage = kDx2Stages * kDx2StageBytes;
constexpr int kDx2OffBars = kDx2OffStage + 4 * (kDxBN / 64) * kStageWarpBytes;      // one staging buffer per (quadrant, 64-column box)
constexpr int kDx2SmemBytes = kDx2OffBars + (2 * kDx2Stages + 4) * 8 + 16 + 1024;
static_assert(kDx2SmemBytes <= 232448, "dx2 kernel shared memory");

// ---- the three roles of the dX GEMM as device functions: head_dx2_kernel runs them on their own, the fused forward + dX
// kernel runs them behind its forward loops on the SAME pipeline state (stage ring, TMEM buffers, barriers), so that a
// CTA pair's first dX tiles run under the forward's last epilogue ----
struct DxShape {
    int B, K, C, planes;
    int first, stride;                  // tiles first, first + stride, ... (row-tile major: tile = mt * n_tiles + nt)
};

// Wait until the forward has published the 256-row tile `rt` of dZ (row-tile words of the workspace, see
// OSB_FWD_PUBLISH_ROW_TILES): lanes 8j .. 8j+7 look at the (up to) 8 slots x 8 warp words of row tile rt + j, four tiles per
// poll.  `target`: the value a published word holds.  Returns false when nothing arrived for 0.5 s.
__device__ __forceinline__ bool wait_row_tiles(const unsigned int* row_flags, unsigned int target, int rt, int n_row_tiles, int slots,
                                               int lane, int& ready_lo, int& ready_hi) {
    if (rt >= ready_lo && rt < ready_hi) return true;
    const int j = rt + (lane >> 3), slot = lane & 7;
    const bool mine = j < n_row_tiles && slot < slots;
    const unsigned int* fp = row_flags + ((size_t)j * 8 + slot) * 8;
    const long long t0 = gtime();
    unsigned got;
    bool alive = true;
    for (;;) {
        bool ok = true;
        if (mine) {
            uint32_t f[8];
            asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(f[0]), "=r"(f[1]), "=r"(f[2]), "=r"(f[3]) : "l"(fp) : "memory");
            asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(f[4]), "=r"(f[5]), "=r"(f[6]), "=r"(f[7]) : "l"(fp + 4) : "memory");
            ok = f[0] == target && f[1] == target && f[2] == target && f[3] == target &&
                 f[4] == target && f[5] == target && f[6] == target && f[7] == target;
        }
        got = __ballot_sync(0xffffffffu, ok);
        if ((got & 0xffu) == 0xffu) break;
        if (__any_sync(0xffffffffu, gtime() - t0 > 500000000LL)) { alive = false; break; }
        __nanosleep(256);
    }
    __threadfence();                                    // acquire side of the words' release
    asm volatile("fence.proxy.async;" ::: "memory");    // the loads that follow are TMA (async proxy) reads
    if (!alive) return false;
    int n = 1;
    while (n < 4 && ((got >> (8 * n)) & 0xffu) == 0xffu) ++n;
    ready_lo = rt; ready_hi = rt + n;
    return true;
}

template <int STAGES>
__device__ __forceinline__ void dx_producer(uint8_t* smem, uint64_t* full, uint64_t* empty, uint32_t& it, const CUtensorMap* tm_dz,
                                            const CUtensorMap* tm_w, const DxShape& d, int rank, int lane, int dbg,
                                            const unsigned int* row_flags, unsigned int flag_target, int flag_slots) {
    const int n_tiles = (d.K + kDxBN - 1) / kDxBN, total = ((d.B + 255) / 256) * n_tiles;
    const int n_kb1 = (d.C + BK - 1) / BK, n_kb = n_kb1 * d.planes;
    int ready_lo = 0, ready_hi = 0;
    for (int tile = d.first; tile < total; tile += d.stride) {
        const int mt = tile / n_tiles, nt = tile % n_tiles;
        if (row_flags && !wait_row_tiles(row_flags, flag_target, mt, (d.B + 255) / 256, flag_slots, lane, ready_lo, ready_hi)) {
            ready_lo = 0; ready_hi = 1 << 30;           // the forward never published (it cannot happen inside one kernel): go on
        }
        for (int kb = 0; kb < n_kb; ++kb, ++it) {
            const int pl = kb >= n_kb1 ? 1 : 0, kc = kb - pl * n_kb1;
            const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* st = smem + s * kDx2StageBytes;
            if (elect_one()) {
                if ((dbg & 1) && it >= STAGES) {        // developer probe: MMA rate without operand traffic (stale tiles)
                    if (rank == 0) mbar_arrive(&full[s]);
                } else {
                    if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * kDx2StageBytes);     // both CTAs' bytes land on the leader's barrier
                    tma_load_3d_2sm(st, tm_dz, &full[s], kc * BK, mt * 256 + rank * 128, pl);            // 64 classes x 128 rows
#pragma unroll
                    for (int j = 0; j < 2; ++j)                                                         // 64 features x 64 classes
                        tma_load_2d_2sm(st + kABytes + j * kBoxBytes, tm_w, &full[s], nt * kDxBN + rank * 128 + j * 64, kc * BK);
                }
            }
            __syncwarp();
        }
    }
}

template <int STAGES>
__device__ __forceinline__ void dx_mma(uint8_t* smem, uint64_t* full, uint64_t* empty, uint64_t* t_full, uint64_t* t_empty,
                                       uint32_t& it, uint32_t& tc, uint32_t tbase, const DxShape& d, int lane, long long* trace) {
    const int n_tiles = (d.K + kDxBN - 1) / kDxBN, total = ((d.B + 255) / 256) * n_tiles;
    const int n_kb1 = (d.C + BK - 1) / BK, n_kb = n_kb1 * d.planes;
    const uint32_t idesc = make_idesc_bf16(256, kDxBN, false, true);
    uint32_t tn = 0;
    for (int tile = d.first; tile < total; tile += d.stride, ++tc, ++tn) {
        const uint32_t buf = tc & 1, tph = (tc >> 1) & 1;
        mbar_wait(&t_empty[buf], tph ^ 1);
        tc_fence_after_sync();
        if (trace && tn < 8 && lane == 0) { trace[blockIdx.x * 64 + tn * 8 + 0] = gtime(); trace[blockIdx.x * 64 + tn * 8 + 7] = clock64(); }
        for (int kb = 0; kb < n_kb; ++kb, ++it) {
            const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(&full[s], ph);
            tc_fence_after_sync();
            if (trace && blockIdx.x == 0 && tn >= 1 && tn < 3 && lane == 0) trace[256 * 64 + (tn - 1) * 32 + kb] = clock64();
            const uint32_t a_addr = smem_u32(smem + s * kDx2StageBytes);
            const uint64_t adesc = make_desc_kmajor(a_addr), bdesc = make_desc_mnmajor(a_addr + kABytes, kBoxBytes);
            const int rem = d.C - (kb >= n_kb1 ? kb - n_kb1 : kb) * BK;
            const int ksteps = rem >= BK ? BK / 16 : (rem + 15) / 16;
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                    if (k < ksteps) umma2_bf16(tbase + buf * kDxBN, adesc + 2 * k, bdesc + 128 * k, idesc, (kb | k) != 0);
                umma2_commit_mc(&empty[s]);
                if (kb == n_kb - 1) umma2_commit_mc(&t_full[buf]);
            }
            __syncwarp();
        }
        if (trace && tn < 8 && lane == 0) trace[blockIdx.x * 64 + tn * 8 + 2] = gtime();
    }
}

// warp (q, hh): lane quadrant q, column half hh.  A thread pulls its 128 accumulator columns into registers with one
// TMEM sweep and hands the buffer back to the MMA warp at once; conversion, staging and the TMA stores run after
// that.  Every 64-column box has its own staging buffer, so nothing waits for a store issued in the same tile
// (with two buffers per warp and four boxes per tile the drain waited ~0.5 us per box for the store two boxes
// back to finish reading, 4.5 us per tile against 5.3 us of MMAs).
template <bool F32OUT>
__device__ __forceinline__ void dx_epilogue(uint8_t* my_stage, uint64_t* t_full, uint64_t* t_empty, uint32_t& tc, uint32_t tbase,
                                            const CUtensorMap* tm_dx, const DxShape& d, int rank, int warp, int lane,
                                            const float* gscale, const double* den, long long* trace, int dbg) {
    const int n_tiles = (d.K + kDxBN - 1) / kDxBN, total = ((d.B + 255) / 256) * n_tiles;
    const int q = warp & 3, hh = (warp - kEpiWarp0) >> 2;
    constexpr int kBoxes = kDxBN / 64 / kDx2EW;                  // boxes (and staging buffers) per warp: my_stage holds that many
    const float gs = (float)((gscale ? (double)gscale[0] : 1.0) / den[0]);
    uint32_t tn = 0;
    for (int tile = d.first; tile < total; tile += d.stride, ++tc, ++tn) {
        const int mt = tile / n_tiles, nt = tile % n_tiles;
        const uint32_t buf = tc & 1, tph = (tc >> 1) & 1;
        mbar_wait_warp(&t_full[buf], tph);
        tc_fence_after_sync();
        const bool tr = trace && tn < 8 && threadIdx.x == kEpiWarp0 * 32;
        if (tr) trace[blockIdx.x * 64 + tn * 8 + 3] = gtime();
        const uint32_t t_row = tmem_addr(tbase, q * 32, buf * kDxBN + hh * kBoxes * 64);
        uint32_t v[2 * kBoxes][32];
        if (!(dbg & 2)) {
#pragma unroll
            for (int g = 0; g < 2 * kBoxes; ++g) tmem_ld32(t_row + g * 32, v[g]);
            tmem_ld_wait();
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {
            mbar_arrive_leader(&t_empty[buf]);
            tma_store_wait_read<0>();                            // last tile's stores have read the staging buffers
        }
        __syncwarp();
        if (F32OUT) {
            // fp32 dX: the warp's 128 columns are four boxes of 32 floats; two staging buffers, two rounds
#pragma unroll
            for (int rnd = 0; rnd < 2; ++rnd) {
                if (rnd) { if (lane == 0) tma_store_wait_read<0>(); __syncwarp(); }
#pragma unroll
                for (int b = 0; b < kBoxes; ++b) {
                    const int g = rnd * kBoxes + b;
                    uint8_t* sbuf = my_stage + b * kStageWarpBytes;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        float4 o = make_float4(gs * __uint_as_float(v[g][4 * c]), gs * __uint_as_float(v[g][4 * c + 1]),
                                               gs * __uint_as_float(v[g][4 * c + 2]), gs * __uint_as_float(v[g][4 * c + 3]));
                        *reinterpret_cast<float4*>(sbuf + lane * 128 + ((c ^ (lane & 7)) << 4)) = o;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(tm_dx, sbuf, nt * kDxBN + hh * kBoxes * 64 + g * 32, mt * 256 + rank * 128 + q * 32);
                        tma_store_commit();
                    }
                }
            }
        } else if (!(dbg & 2)) {
#pragma unroll
            for (int b = 0; b < kBoxes; ++b) {
                uint8_t* sbuf = my_stage + b * kStageWarpBytes;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float g[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) g[i] = gs * __uint_as_float(v[2 * b + h][i]);
#pragma unroll
                    for (int cch = 0; cch < 4; ++cch) stage_chunk(sbuf, lane, h * 4 + cch, g + cch * 8);
                }
                // (developer probe: staging without the store gives 5.3 instead of 6.3 us per tile; coalesced LSU stores of the
                //  staged box instead of TMA give the same 6.3 -- it is the output traffic itself, not the TMA unit)
                if (dbg & 8) continue;
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(tm_dx, sbuf, nt * kDxBN + (hh * kBoxes + b) * 64, mt * 256 + rank * 128 + q * 32);
                    tma_store_commit();
                }
            }
        }
        if (tr) trace[blockIdx.x * 64 + tn * 8 + 4] = gtime();
    }
    if (lane == 0) tma_store_wait_read<0>();      // staging buffers read; the stores land before the grid ends
}

// =================================================================================================
// forward: logits in TMEM -> loss, statistics, dZ
// =================================================================================================
struct FwdArgs {
    const float* bias; const int64_t* y; const float* class_w;
    int B, K, C; int variant; float t_uniform; int64_t ignore_index; const double* den; int valid_cols;
    double* loss_num; float* loss;
    float* row_ptgt; float* row_pmax; float* row_pmax_all; int32_t* row_pred;
    int write_dz;
    int dz_planes;             // 1, or 2: plane 1 receives the bf16 rounding residual of plane 0
    int den_from_prev;         // launched programmatically behind the kernel that writes den: wait for it before den is read
    double* partials; unsigned int* ticket;
    uint2* xch_stats;          // [row block][2][CL][128][6] (value, epoch) words: per-row statistics exchanged between the units of a group
    unsigned int* xch_epoch;   // epochs used so far on this workspace; only grows, so stale words never match
    unsigned int* tile_flags;  // nullable: [row tile][unit * 2 + cta][epilogue warp] "grad-logits landed" words (OSB_FWD_PUBLISH_ROW_TILES)
    unsigned int* flag_epoch;  // fused forward + dX: launches so far on this workspace; a published word holds 0x80000000 + epoch + 1
    long long* trace;      // optional [gridDim.x][64] clock stamps (developer tool, NULL in production)
};

constexpr int kStatFloats = 6;          // m, s, zy, sumz, mv, argmax  (per row, per column slice)

// The dX GEMM behind the forward loops of the same kernel (head_fwd_kernel<..., FUSE = true>): its operands and output.
struct DxFuse {
    CUtensorMap tm_dz;         // dZ as the K-major A operand (loads)
    CUtensorMap tm_w;          // W as the MN-major B operand
    CUtensorMap tm_dx;         // dX (stores)
    const float* gscale;       // upstream gradient (device scalar) or NULL = 1
    int dx_f32;                // element type of dX
    int fwd_units;             // units [0, fwd_units) run the forward; every unit of the grid runs dX tiles
};

// One CTA owns 128 rows x BN classes of the logits (BN = 128 or 256 TMEM columns, double-buffered so the
// epilogue of tile t overlaps the MMAs of tile t+1).  Heads wider than BN run as a group of CL = ceil(C/BN)
// units (CL <= 4) that share the rows, split the classes and exchange per-row softmax statistics through L2.
// EW epilogue warps per TMEM lane quadrant each own a slice of the CTA's columns.
// PAIR: a unit is a 2-CTA cluster issuing M = 256 MMAs (cta_group::2): CTA r of the pair stages rows [128 r, +128)
// of a 256-row X tile and half of the unit's W rows, i.e. 32 KB instead of 48 KB per k-block and CTA for the same
// flops — the single-CTA kernel is bound by L2 -> SM operand traffic (~11.5 TB/s), not by the tensor pipe.
// Each CTA still drains its own 128 x BN accumulator, so the epilogue is the same code in both modes.
template <int BN, int CL, int STAGES, int EW, bool PAIR>
struct FwdCfg {
    static constexpr int kThreads = (2 + 4 * EW) * 32;
    static constexpr int kEpiThreads = 4 * EW * 32;
    static constexpr int kWRows = PAIR ? BN / 2 : BN;                              // W rows this CTA stages per k-block
    static constexpr int kStageBytes = kABytes + kWRows * BK * 2;
    static constexpr int kTmemCols = 2 * BN;
    static constexpr int kTileRows = PAIR ? 2 * BM : BM;                           // rows per tile of a CTA pair / a CTA
    static constexpr int kPeers = CL > 1 ? CL - 1 : 1;
    static constexpr int kOffDz = STAGES * kStageBytes;                            // [4 EW warps][2] boxes of 32 rows x 128 B (grad-logits staging)
    static constexpr int kOffX = kOffDz + 4 * EW * 2 * kStageWarpBytes;            // [2][128][8] (value, epoch) words between the two owners of a row
    static constexpr int kOffBias = kOffX + 2 * BM * 8 * 8;                        // [BN] floats
    static_assert(EW == 2, "the epilogue pairs exactly two threads per accumulator row");
    static constexpr int kGroups = BN / 32 / EW;                                   // 32-column groups per epilogue thread
    static constexpr int kOffBars = kOffBias + BN * 4;
    static constexpr int kNumBars = 2 * STAGES + 4;
    static constexpr int kSmemBytes = kOffBars + kNumBars * 8 + 16 + 1024;  // + tmem slot + alignment slack
    static_assert(kSmemBytes <= 232448, "exceeds 227 KB of shared memory");
};

__device__ __forceinline__ float max32(const uint32_t (&z)[32]) {
    float t[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = fmaxf(__uint_as_float(z[i]), __uint_as_float(z[i + 16]));
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = fmaxf(t[i], t[i + 8]);
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = fmaxf(t[i], t[i + 4]);
    return fmaxf(fmaxf(t[0], t[1]), fmaxf(t[2], t[3]));
}

// v[idx] for a per-thread idx in [0, 32): registers cannot be indexed dynamically, so a 5-level select tree on the bits of
// idx (31 selects) replaces 32 compare + select pairs.
__device__ __forceinline__ float select32(const uint32_t (&v)[32], int idx) {
    uint32_t t[16];
    const bool b0 = idx & 1, b1 = idx & 2, b2 = idx & 4, b3 = idx & 8, b4 = idx & 16;
#pragma unroll
    for (int j = 0; j < 16; ++j) t[j] = b0 ? v[2 * j + 1] : v[2 * j];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = b1 ? t[2 * j + 1] : t[2 * j];
#pragma unroll
    for (int j = 0; j < 4; ++j) t[j] = b2 ? t[2 * j + 1] : t[2 * j];
#pragma unroll
    for (int j = 0; j < 2; ++j) t[j] = b3 ? t[2 * j + 1] : t[2 * j];
    return __uint_as_float(b4 ? t[1] : t[0]);
}

// 8 fp32 -> 8 bf16 -> one 16-byte global store
__device__ __forceinline__ void store8_bf16(__nv_bfloat16* dst, const float* f) {
    __nv_bfloat162 p0 = __floats2bfloat162_rn(f[0], f[1]);
    __nv_bfloat162 p1 = __floats2bfloat162_rn(f[2], f[3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(f[4], f[5]);
    __nv_bfloat162 p3 = __floats2bfloat162_rn(f[6], f[7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
    u.z = *reinterpret_cast<uint32_t*>(&p2); u.w = *reinterpret_cast<uint32_t*>(&p3);
    *reinterpret_cast<uint4*>(dst) = u;
}

// 10 warps get 168 registers each: the register file is handed out in 4-warp granules, so 320 threads are budgeted as 384
// (a 200-register build fails to launch).
// FUSE: the grid covers every SM; the units beyond fz.fwd_units skip the forward, and EVERY unit goes on with tiles of the dX
// GEMM on the same pipeline state (stage ring, TMEM buffers, barriers): the MMA warp issues its first dX tiles while the
// epilogue warps are still in the forward's last epilogue, and the SMs the forward leaves idle work from the start.  dZ
// still travels through global memory; the row-tile words (value 0x80000000 + launch epoch + 1) say which rows have landed.
template <int BN, int CL, int STAGES, int EW, bool PAIR, bool HAS_BIAS, bool WANT_PRED, bool PLANES2, bool FUSE = false>
__global__ void __launch_bounds__((2 + 4 * EW) * 32, 1)
head_fwd_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                const __grid_constant__ CUtensorMap tm_dz, const FwdArgs a, const __grid_constant__ DxFuse fz) {
    using Cfg = FwdCfg<BN, CL, STAGES, EW, PAIR>;
    static_assert(!FUSE || (PAIR && BN == kDxBN && Cfg::kStageBytes == kDx2StageBytes && EW == kDx2EW && !WANT_PRED),
                  "the fused dX phase reuses the forward's stages, staging buffers and TMEM columns as they are");
    constexpr int TM = Cfg::kTileRows;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    uint8_t* sDz = smem + Cfg::kOffDz;
    uint2* sX = reinterpret_cast<uint2*>(smem + Cfg::kOffX);
    float* sBias = reinterpret_cast<float*>(smem + Cfg::kOffBias);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBars);
    uint64_t* full = bars;                uint64_t* empty = full + STAGES;
    uint64_t* t_full = empty + STAGES;    uint64_t* t_empty = t_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (a.trace && threadIdx.x == 0) a.trace[(384 + blockIdx.x / 32) * 64 + (blockIdx.x % 32) * 2] = gtime();
    // A unit is one CTA, or one 2-CTA cluster (PAIR; cr = this CTA's half of the unit's rows).  CL consecutive units
    // form a group that shares the rows and splits the classes.  The group is not a hardware cluster: clusters of 4
    // strand ~16 of the 148 SMs (GPC packing); it only needs co-residency, which the cooperative launch gives.
    const int cr = PAIR ? (int)cluster_ctarank_() : 0;
    const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int n_units = FUSE ? fz.fwd_units : (PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x);
    const bool in_fwd = !FUSE || unit < n_units;
    const uint32_t rank = (CL > 1) ? (uint32_t)(unit % CL) : 0u;
    const int cluster_id = unit / CL, n_clusters = n_units / CL;
    const int xid = PAIR ? cluster_id * 2 + cr : cluster_id;      // slot of this CTA's rows in the exchange buffers
    const int n_tiles = in_fwd ? (a.B + TM - 1) / TM : 0;         // (a unit outside the forward has no forward tiles)
    const uint32_t pub_value = FUSE ? (0x80000000u + *a.flag_epoch + 1u) : 1u;
    DxShape dxs;
    if (FUSE) { dxs.B = a.B; dxs.K = a.K; dxs.C = a.C; dxs.planes = a.dz_planes; dxs.first = unit; dxs.stride = (int)(gridDim.x >> 1); }
    const int n_kb = (a.K + BK - 1) / BK;
    const int n0 = (int)rank * BN;                        // class range of this unit
    int ncols = a.C - n0;
    ncols = ncols < 0 ? 0 : (ncols > BN ? BN : ncols);
    // UMMA N: multiple of 16 (of 32 for a pair, whose CTAs each stage N / 2 rows of W)
    const int ncolsN = PAIR ? ((ncols + 31) & ~31) : ((ncols + 15) & ~15);
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], (PAIR ? 2 : 1) * 4 * EW); }
        fence_barrier_init();
        prefetch_tensormap(&tm_x); prefetch_tensormap(&tm_w); prefetch_tensormap(&tm_dz);
        if (FUSE) { prefetch_tensormap(&fz.tm_dz); prefetch_tensormap(&fz.tm_w); prefetch_tensormap(&fz.tm_dx); }
    }
    if (HAS_BIAS) {
        for (int i = threadIdx.x; i < BN; i += Cfg::kThreads) sBias[i] = (i < ncols) ? a.bias[n0 + i] : 0.f;
    }
    for (int i = threadIdx.x; i < 2 * BM * 8; i += Cfg::kThreads) sX[i] = make_uint2(0u, 0u);     // epoch 0 = nothing published
    if (warp == 1) { if (PAIR) tmem_alloc2(tmem_slot, Cfg::kTmemCols); else tmem_alloc(tmem_slot, Cfg::kTmemCols); }
    if (PAIR && !FUSE && a.tile_flags && warp >= kEpiWarp0 && lane == 0) {
        // every word is cleared by the thread that sets it later, and before this CTA lets the dependent kernel start
        for (int tile = cluster_id; tile < n_tiles; tile += n_clusters)
            a.tile_flags[(size_t)tile * kRowFlagWords + (rank * 2 + cr) * 8 + (warp - kEpiWarp0)] = 0u;
        __threadfence();
    }
    tc_fence_before_sync();
    if (PAIR) cluster_sync_(); else __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = *tmem_slot;
    pdl_trigger();              // the backward GEMM behind this kernel may set itself up on SMs this grid leaves idle or has left

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        // (issue loops run warp-converged with one elected lane issuing, see the MMA issuer below)
        uint32_t it = 0;
        if (ncols > 0) {
            // the first tile has nobody to prefetch it: ask for all of its feature rows at once (the units of the group
            // share the k-blocks), so that HBM latency is paid once and not once per pipeline refill
            if (cluster_id < n_tiles && elect_one()) {
                for (int kb = (int)rank; kb < n_kb; kb += CL) tma_prefetch_l2_2d(&tm_x, kb * BK, cluster_id * TM + cr * BM);
            }
            __syncwarp();
            for (int tile = cluster_id; tile < n_tiles; tile += n_clusters) {
                for (int kb = 0; kb < n_kb; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = smem + s * Cfg::kStageBytes;
                    if (elect_one()) {
                        // The feature rows come from HBM and the pipeline is only STAGES deep: pull the NEXT tile's rows
                        // into L2 while this tile is computed (each unit of the group takes every CL-th k-block).
                        if ((kb % CL) == (int)rank && tile + n_clusters < n_tiles)
                            tma_prefetch_l2_2d(&tm_x, kb * BK, (tile + n_clusters) * TM + cr * BM);
                        if (PAIR) {
                            // both CTAs' bytes are counted on the leader's barrier (the MMA issuer waits there)
                            if (cr == 0) mbar_arrive_expect_tx(&full[s], 2 * Cfg::kStageBytes);
                            tma_load_2d_2sm(st, &tm_x, &full[s], kb * BK, tile * TM + cr * BM);
                            tma_load_2d_2sm(st + kABytes, &tm_w, &full[s], kb * BK, n0 + cr * (ncolsN >> 1));
                        } else {
                            mbar_arrive_expect_tx(&full[s], Cfg::kStageBytes);
                            tma_load_2d(st, &tm_x, &full[s], kb * BK, tile * BM);
                            tma_load_2d(st + kABytes, &tm_w, &full[s], kb * BK, n0);
                        }
                    }
                    __syncwarp();
                }
            }
        }
        if constexpr (FUSE) dx_producer<STAGES>(smem, full, empty, it, &fz.tm_dz, &fz.tm_w, dxs, cr, lane, 0, a.tile_flags, pub_value, CL * 2);
    } else if (warp == 1) {
        // ------------------------------ MMA issuer --------------------------------
        // The whole warp runs the issue loop converged and ONE elected lane issues.  With the loop inside an
        // `if (lane == 0)` region ptxas keeps the descriptors in vector registers and moves five of them to uniform
        // registers (ELECT + R2UR.BROADCAST) before every MMA: 143..182 cycles per MMA measured, more than the
        // 128 cycles the tensor pipe needs (tools/mma_rate_probe.cu).
        uint32_t it = 0, tc = 0;
        if (ncols > 0 && cr == 0) {
            const uint32_t idesc = make_idesc_bf16(TM, ncolsN, false, false);
            for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, ++tc) {
                const uint32_t buf = tc & 1, tph = (tc >> 1) & 1;
                mbar_wait(&t_empty[buf], tph ^ 1);
                tc_fence_after_sync();
                if (a.trace && tc < 4 && lane == 0) { a.trace[blockIdx.x * 64 + tc * 8 + 0] = gtime(); a.trace[blockIdx.x * 64 + tc * 8 + 7] = clock64(); }
                for (int kb = 0; kb < n_kb; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after_sync();
                    if (a.trace && tc < 4 && kb == 0 && lane == 0) a.trace[blockIdx.x * 64 + tc * 8 + 1] = gtime();
                    if (a.trace && blockIdx.x == 0 && tc < 4 && kb < 32 && lane == 0) a.trace[256 * 64 + tc * 32 + kb] = clock64();
                    const uint32_t a_addr = smem_u32(smem + s * Cfg::kStageBytes);
                    const uint64_t adesc = make_desc_kmajor(a_addr), bdesc = make_desc_kmajor(a_addr + kABytes);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            // +32 bytes along K = +2 in the descriptor's 16-byte address field
                            if (PAIR) umma2_bf16(tbase + buf * BN, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                            else umma_bf16(tbase + buf * BN, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                        }
                        if (PAIR) umma2_commit_mc(&empty[s]); else umma_commit(&empty[s]);
                        if (kb == n_kb - 1) { if (PAIR) umma2_commit_mc(&t_full[buf]); else umma_commit(&t_full[buf]); }
                    }
                    __syncwarp();
                }
                if (a.trace && tc < 4 && lane == 0) a.trace[blockIdx.x * 64 + tc * 8 + 2] = gtime();
            }
        }
        if constexpr (FUSE) {
            if (cr == 0) dx_mma<STAGES>(smem, full, empty, t_full, t_empty, it, tc, tbase, dxs, lane, nullptr);
            __syncwarp();
        }
    } else {
        // ------------------------------ epilogue ----------------------------------
        // warp (q, h): TMEM lane quadrant q = warp % 4 (hardware rule), column slice h; one thread per row.
        // A thread reads its kGroups * 32 logits from TMEM ONCE, keeps them in registers and hands the accumulator back to
        // the MMA warp at once; everything after that (max, exp, statistics exchange, grad-logits) runs out of registers.
        // (Three TMEM sweeps with exp(z - max) parked in TMEM made the epilogue, not the MMAs, the pace of the kernel.)
        constexpr int G = Cfg::kGroups;
        const int q = warp & 3, h = (warp - kEpiWarp0) >> 2;
        const int r_in_tile = q * 32 + lane;
        const float s_uniform = a.t_uniform * (float)a.C;
        const int c_base = h * G * 32;                      // this warp's columns: [c_base, c_base + G * 32) of the unit's BN
        uint8_t* my_stage = sDz + (warp - kEpiWarp0) * 2 * kStageWarpBytes;
        uint32_t n_stores = 0;
        int pub_tile = -1;                                   // tile whose grad-logits are stored but not yet published
        auto publish = [&]() {
            if (PAIR && a.tile_flags && pub_tile >= 0 && lane == 0) {
                tma_store_wait<0>();                         // in global memory, not merely read out of shared memory
                __threadfence();
                asm volatile("st.release.gpu.global.u32 [%0], %1;"
                             :: "l"(a.tile_flags + (size_t)pub_tile * kRowFlagWords + (rank * 2 + cr) * 8 + (warp - kEpiWarp0)), "r"(pub_value) : "memory");
            }
            pub_tile = -1;
        };
        const uint32_t ep_base = (CL > 1) ? *a.xch_epoch : 0u;
        double loss_acc = 0.0;
        uint32_t it = 0;
        for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, ++it) {
            const uint32_t buf = it & 1, tph = (it >> 1) & 1;
            const int row = tile * TM + cr * BM + r_in_tile;
            const bool row_ok = row < a.B;
            const int64_t y = row_ok ? a.y[row] : -1;
            RowMode rm = classify_row(a.variant, y, a.C, a.ignore_index, a.class_w, s_uniform);
            if (!row_ok) { rm.mode = kRowIgnored; rm.omega = 0.f; }
            int c_end = c_base + G * 32;
            c_end = c_end > ncols ? ncols : c_end;
            const int y_local = (y >= n0 + c_base && y < n0 + c_end) ? (int)(y - n0) : -1;   // label column, if in my slice
            if (ncols > 0) {
                mbar_wait_warp(&t_full[buf], tph);
                tc_fence_after_sync();
            }
            const bool tr = a.trace && it < 4 && threadIdx.x == kEpiWarp0 * 32;
            if (tr) a.trace[blockIdx.x * 64 + it * 8 + 3] = gtime();
            const uint32_t t_row = tmem_addr(tbase, q * 32, buf * BN);

            // ---- the one TMEM sweep; the accumulator buffer goes back to the MMA warp right after it ----
            uint32_t v[G][32];
#pragma unroll
            for (int g = 0; g < G; ++g)
                if (c_base + g * 32 < ncols) tmem_ld32(t_row + c_base + g * 32, v[g]);
            tmem_ld_wait();
            if (tr && blockIdx.x == 0) a.trace[256 * 64 + 128 + it * 8 + 0] = gtime();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) { if (PAIR) mbar_arrive_leader(&t_empty[buf]); else mbar_arrive(&t_empty[buf]); }

            // ---- row max (and arg-max / valid-column max) over my slice; no transcendental ----
            float m = -INFINITY, mv = -INFINITY;
            int am = 0x7fffffff;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int c0 = c_base + g * 32;
                if (c0 < ncols) {
                    if (HAS_BIAS || c0 + 32 > ncols) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            float z = __uint_as_float(v[g][i]);
                            if (HAS_BIAS) z += sBias[c0 + i];
                            if (c0 + i >= ncols) z = -INFINITY;
                            v[g][i] = __float_as_uint(z);
                        }
                    }
                    const float gm = max32(v[g]);
                    if (WANT_PRED) {
                        if (gm > m) {                       // first arg-max (numpy tie rule): groups ascend, strict '>'
#pragma unroll
                            for (int i = 31; i >= 0; --i) if (__uint_as_float(v[g][i]) == gm) am = n0 + c0 + i;
                        }
                    }
                    m = fmaxf(m, gm);
                    const int gcol = n0 + c0;
                    if (gcol + 32 <= a.valid_cols) mv = fmaxf(mv, gm);
                    else if (gcol < a.valid_cols) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) if (gcol + i < a.valid_cols) mv = fmaxf(mv, __uint_as_float(v[g][i]));
                    }
                }
            }
            if (tr && blockIdx.x == 0) a.trace[256 * 64 + 128 + it * 8 + 1] = gtime();
            // ---- e = exp(z - m_slice) in place; partial sums with 4-way ILP.  The label's logit comes out of a select
            //      tree and the sum of the logits (uniform-target rows) runs unconditionally: per-element predicates cost
            //      more instructions than the arithmetic they guard. ----
            float s = 0.f, sumz = 0.f, zy = 0.f;
            {
                const float mb = m * kLog2e;
                float s4[4] = {0.f, 0.f, 0.f, 0.f}, z4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const int c0 = c_base + g * 32;
                    if (c0 < ncols) {
                        const int yl = y_local - c0;
                        const float zsel = select32(v[g], yl & 31);
                        if (yl >= 0 && yl < 32) zy = zsel;
                        if (c0 + 32 > ncols) {          // tail group: masked columns hold -inf and must not enter the sum
#pragma unroll
                            for (int i = 0; i < 32; ++i) if (c0 + i < ncols) z4[i & 3] += __uint_as_float(v[g][i]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) z4[i & 3] += __uint_as_float(v[g][i]);
                        }
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float e = fast_exp2(fmaf(__uint_as_float(v[g][i]), kLog2e, -mb));
                            s4[i & 3] += e;
                            v[g][i] = __float_as_uint(e);
                        }
                    }
                }
                s = (s4[0] + s4[1]) + (s4[2] + s4[3]);
                sumz = (z4[0] + z4[1]) + (z4[2] + z4[3]);
            }
            if (tr) a.trace[blockIdx.x * 64 + it * 8 + 4] = gtime();
            const float m_slice = m, zy_own = zy;
            // ---- the two owners of a row (column halves h = 0, 1) meet through (value, epoch) words in shared memory, the
            //      units of a group through the same kind of words in L2: pairwise waits only, no CTA-wide barrier (with
            //      bar.sync every warp waited for the slowest warp AND for the L2 round trip of the polling threads) ----
            const uint32_t eps = it + 1;                                   // shared-memory epoch (this launch only)
            uint2* sx = sX + ((it & 1) * BM + r_in_tile) * 8;              // [0..5] h=1 -> h=0 statistics, [6..7] h=0 -> h=1 row max / sum
            float m_cta = m;                                               // max over this unit's columns (arg-max fold at the end)
            if (h == 1) {
                sts_ll(sx + 0, m, eps); sts_ll(sx + 1, s, eps); sts_ll(sx + 2, zy, eps);
                sts_ll(sx + 3, sumz, eps); sts_ll(sx + 4, mv, eps); sts_ll(sx + 5, __int_as_float(am), eps);
                uint2 g0, g1;
                do { g0 = lds_ll(sx + 6); g1 = lds_ll(sx + 7); if (g0.y != eps || g1.y != eps) __nanosleep(32); } while (g0.y != eps || g1.y != eps);
                m = __uint_as_float(g0.x); s = __uint_as_float(g1.x);
            } else {
                {
                    uint2 w[kStatFloats];
                    bool ok;
                    do {
                        ok = true;
#pragma unroll
                        for (int j = 0; j < kStatFloats; ++j) { w[j] = lds_ll(sx + j); ok = ok && (w[j].y == eps); }
                    } while (!ok);
                    const float pm = __uint_as_float(w[0].x);
                    if (pm > m) am = (int)w[5].x;                          // the partner owns the higher columns: ties stay here
                    const float mc = fmaxf(m, pm);
                    s = s * fast_exp2((m - mc) * kLog2e) + __uint_as_float(w[1].x) * fast_exp2((pm - mc) * kLog2e);
                    m = mc; m_cta = mc;
                    zy += __uint_as_float(w[2].x); sumz += __uint_as_float(w[3].x); mv = fmaxf(mv, __uint_as_float(w[4].x));
                }
                if (tr && blockIdx.x == 0) a.trace[256 * 64 + 128 + it * 8 + 2] = gtime();
                if (CL > 1) {
                    // Through L2, not DSMEM: with the tensor pipe busy on the next tile, remote shared-memory stores and
                    // arrives to a peer SM took ~8 us per tile (measured).  One thread per row publishes (value, epoch)
                    // words and polls the peers' words of the same row: one L2 round trip, no flag, no fence.  Everybody
                    // needs the row max and sum of the whole head now; the other four statistics only feed the per-row
                    // outputs, which unit 0 collects AFTER the grad-logits are out (registers are free again by then).
                    const uint32_t ep = ep_base + it + 1;
                    uint2* xs = a.xch_stats + ((size_t)(xid * 2 + (it & 1)) * CL) * BM * kStatFloats;
                    uint2* dst = xs + ((size_t)rank * BM + r_in_tile) * kStatFloats;
                    st_ll(dst + 0, m, ep); st_ll(dst + 1, s, ep);
                    if (rank != 0) {
                        st_ll(dst + 2, zy, ep); st_ll(dst + 3, sumz, ep); st_ll(dst + 4, mv, ep); st_ll(dst + 5, __int_as_float(am), ep);
                    }
                    uint2 pv[Cfg::kPeers][2];
                    bool ok;
                    do {
                        ok = true;
#pragma unroll
                        for (int d = 1; d < CL; ++d) {
                            const uint2* src = xs + ((size_t)((rank + d) % CL) * BM + r_in_tile) * kStatFloats;
                            pv[d - 1][0] = ld_ll(src); pv[d - 1][1] = ld_ll(src + 1);
                        }
#pragma unroll
                        for (int d = 1; d < CL; ++d) ok = ok && (pv[d - 1][0].y == ep) && (pv[d - 1][1].y == ep);
                    } while (!ok);
                    float mg = m;
#pragma unroll
                    for (int d = 1; d < CL; ++d) mg = fmaxf(mg, __uint_as_float(pv[d - 1][0].x));
                    s *= fast_exp2((m - mg) * kLog2e);
#pragma unroll
                    for (int d = 1; d < CL; ++d)
                        s += __uint_as_float(pv[d - 1][1].x) * fast_exp2((__uint_as_float(pv[d - 1][0].x) - mg) * kLog2e);
                    m = mg;
                }
                sts_ll(sx + 6, m, eps); sts_ll(sx + 7, s, eps);
            }
            if (tr) a.trace[blockIdx.x * 64 + it * 8 + 5] = gtime();
            // ---- denominator-free grad-logits  dz = S p - T = coef * e - T  -> bf16 -> swizzled smem -> TMA store ----
            // (a thread owns a row, so direct stores would scatter 16-byte pieces over 32 rows per instruction; the TMA
            // store writes whole 128-byte lines and clips the columns >= C and the rows >= B itself)
            if (a.write_dz) {
                // Row-tile hand-over: the PREVIOUS tile's stores were issued a whole tile ago and have long landed in global
                // memory (waiting for them costs nothing here; waiting right behind the stores would stall the warp for
                // their full latency): publish that tile to the dW GEMM that may already run on the SMs this grid leaves idle
                publish();
                const bool live = (rm.mode == kRowOneHot || rm.mode == kRowUniform);
                const float coef = live ? rm.omega / s : 0.f;
                const float coef_e = coef * fast_exp2((m_slice - m) * kLog2e);   // e was taken relative to the slice max
                const float t_all = (rm.mode == kRowUniform) ? a.t_uniform : 0.f;
                const float t_hot = (rm.mode == kRowOneHot) ? rm.omega : 0.f;
                const float gy = fmaf(coef, fast_exp2((zy_own - m) * kLog2e), -t_all) - t_hot;   // label column
#pragma unroll
                for (int b = 0; b < (G + 1) / 2; ++b) {         // boxes of 64 columns
                    const int c0 = c_base + b * 64;
                    if (c0 < ncols) {
                        const int cl = y_local - c0;
                        const bool patch = y_local >= 0 && cl >= 0 && cl < 64;
                        // plane 0, and (two planes) plane 1 = what the bf16 rounding of plane 0 dropped: straight-line code,
                        // the second box behind a warp-uniform branch
                        auto emit = [&](auto lo_tag) {
                            constexpr bool LO = decltype(lo_tag)::value;
                            uint8_t* sbuf = my_stage + (n_stores & 1) * kStageWarpBytes;
                            if (lane == 0) tma_store_wait_read<1>();        // the store that read this buffer two boxes ago is done
                            __syncwarp();
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                const int g = 2 * b + hh;
                                if (g < G && c0 + hh * 32 < ncols) {
#pragma unroll
                                    for (int cch = 0; cch < 4; ++cch) {
                                        float gl[8];
#pragma unroll
                                        for (int i = 0; i < 8; ++i) {
                                            const float gv = fmaf(coef_e, __uint_as_float(v[g < G ? g : 0][cch * 8 + i]), -t_all);
                                            gl[i] = LO ? gv - __bfloat162float(__float2bfloat16_rn(gv)) : gv;
                                        }
                                        stage_chunk(sbuf, lane, hh * 4 + cch, gl);
                                    }
                                }
                            }
                            // the label column carries the -T term of a one-hot row: patched in the staged row by its owner
                            if (patch) {
                                const float gp = LO ? gy - __bfloat162float(__float2bfloat16_rn(gy)) : gy;
                                *reinterpret_cast<__nv_bfloat16*>(sbuf + lane * 128 + (((cl >> 3) ^ (lane & 7)) << 4) + (cl & 7) * 2) = __float2bfloat16_rn(gp);
                            }
                            fence_proxy_async_smem();
                            __syncwarp();
                            if (lane == 0) {
                                tma_store_3d(&tm_dz, sbuf, n0 + c0, tile * TM + cr * BM + q * 32, LO ? 1 : 0);
                                tma_store_commit();
                            }
                            ++n_stores;
                        };
                        emit(std::false_type{});
                        if (PLANES2) emit(std::true_type{});
                    }
                }
                pub_tile = tile;
            }
            if (tr && blockIdx.x == 0) a.trace[256 * 64 + 128 + it * 8 + 3] = gtime();
            // ---- per-row outputs (unit 0 of the group): loss, p[target], max p over the valid columns, max p, arg-max ----
            if (rank == 0 && h == 0) {
                if (CL > 1) {
                    const uint32_t ep = ep_base + it + 1;
                    const uint2* xs = a.xch_stats + ((size_t)(xid * 2 + (it & 1)) * CL) * BM * kStatFloats;
                    // all peers' words in flight at once (one L2 round trip; the logits' registers are free again)
                    uint2 w[Cfg::kPeers][kStatFloats];
                    bool ok;
                    do {
                        ok = true;
#pragma unroll
                        for (int d = 1; d < CL; ++d) {
                            const uint2* src = xs + ((size_t)d * BM + r_in_tile) * kStatFloats;
#pragma unroll
                            for (int j = 0; j < kStatFloats; ++j) if (j != 1) w[d - 1][j] = ld_ll(src + j);
                        }
#pragma unroll
                        for (int d = 1; d < CL; ++d)
#pragma unroll
                            for (int j = 0; j < kStatFloats; ++j) if (j != 1) ok = ok && (w[d - 1][j].y == ep);
                    } while (!ok);
                    float m_am = m_cta;
#pragma unroll
                    for (int d = 1; d < CL; ++d) {          // units in column order, so ties keep the smaller column index
                        const float pm = __uint_as_float(w[d - 1][0].x);
                        const int pam = (int)w[d - 1][5].x;
                        if (pm > m_am || (pm == m_am && pam < am)) { am = pam; m_am = pm; }
                        zy += __uint_as_float(w[d - 1][2].x); sumz += __uint_as_float(w[d - 1][3].x);
                        mv = fmaxf(mv, __uint_as_float(w[d - 1][4].x));
                    }
                }
                if (row_ok) {
                    const float lse = m + fast_log2(s) * kLn2;
                    const bool is_col = (y >= 0 && y < a.C);
                    float rowloss = 0.f;
                    if (rm.mode == kRowOneHot) rowloss = rm.omega * (lse - zy);
                    else if (rm.mode == kRowUniform) rowloss = a.t_uniform * ((float)a.C * lse - sumz);
                    else if (rm.mode == kRowBad) rowloss = __int_as_float(0x7fc00000);
                    loss_acc += (double)rowloss;
                    if (a.row_ptgt) a.row_ptgt[row] = is_col ? fast_exp2((zy - lse) * kLog2e) : 0.f;
                    if (a.row_pmax) a.row_pmax[row] = fast_exp2((mv - lse) * kLog2e);
                    if (WANT_PRED) {
                        if (a.row_pmax_all) a.row_pmax_all[row] = fast_exp2((m - lse) * kLog2e);
                        if (a.row_pred) a.row_pred[row] = am;
                    }
                }
            }
            if (tr) a.trace[blockIdx.x * 64 + it * 8 + 6] = gtime();
        }
        // ---- deterministic loss reduction without a CTA barrier: only the warps that accumulated row losses (unit 0 of a
        //      group, column half 0) take part.  Each writes its partial and takes a ticket; the LAST warp of the grid sums
        //      all partials in index order (lane l takes l, l + 32, ...; fixed shuffle tree on top) ----
        const bool trx = a.trace && blockIdx.x == 0 && threadIdx.x == kEpiWarp0 * 32;
        if (trx) a.trace[256 * 64 + 192 + 0] = gtime();
        if (FUSE) publish();                              // the dX tiles of every pair wait for these rows: publish them now
        if (in_fwd && rank == 0 && h == 0) {
            const unsigned n_part = (unsigned)n_clusters * (PAIR ? 2u : 1u) * 4u;
            const double wl = warp_sum(loss_acc);
            unsigned last = 0;
            if (lane == 0) {
                a.partials[xid * 4 + q] = wl;
                __threadfence();
                last = (atomicAdd(a.ticket, 1u) == n_part - 1) ? 1u : 0u;
            }
            last = __shfl_sync(0xffffffffu, last, 0);
            if (trx) a.trace[256 * 64 + 192 + 2] = gtime();
            if (last) {
                __threadfence();
                double tot = 0.0;
                for (unsigned b = lane; b < n_part; b += 32) tot += __ldcg(&a.partials[b]);
                tot = warp_sum(tot);
                if (lane == 0) {
                    a.loss_num[0] = tot;
                    a.loss[0] = a.den ? (float)(tot / a.den[0]) : (float)tot;
                    *a.ticket = 0u;
                    if (CL > 1) *a.xch_epoch = ep_base + (uint32_t)((n_tiles + n_clusters - 1) / n_clusters) + 1u;
                    if (FUSE) *a.flag_epoch = pub_value - 0x80000000u;      // every CTA read the old value when it started
                }
            }
        }
        publish();                                        // the last tile of this CTA (after the loss reduction: this one does wait)
        if (lane == 0) tma_store_wait_read<0>();          // the staging buffers have been read; the data lands before the grid ends
        if constexpr (FUSE) {
            __syncwarp();
            // (the warp keeps ITS staging buffers: the wait above covers only the stores this warp issued)
            if (fz.dx_f32) dx_epilogue<true>(my_stage, t_full, t_empty, it, tbase, &fz.tm_dx, dxs, cr, warp, lane, fz.gscale, a.den, nullptr, 0);
            else dx_epilogue<false>(my_stage, t_full, t_empty, it, tbase, &fz.tm_dx, dxs, cr, warp, lane, fz.gscale, a.den, nullptr, 0);
        }
        if (trx) a.trace[256 * 64 + 192 + 1] = gtime();
    }
    if (a.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0) a.trace[256 * 64 + 200 + warp] = gtime();
    tc_fence_before_sync();
    if (PAIR) cluster_sync_(); else __syncthreads();
    if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[256 * 64 + 192 + 3] = gtime();
    if (warp == 1) { if (PAIR) tmem_dealloc2(tbase, Cfg::kTmemCols); else tmem_dealloc(tbase, Cfg::kTmemCols); }
    if (a.trace && threadIdx.x == 0) a.trace[(384 + blockIdx.x / 32) * 64 + (blockIdx.x % 32) * 2 + 1] = gtime();
}

template <bool F32OUT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kDx2Threads, 1)
head_dx2_kernel(const __grid_constant__ CUtensorMap tm_dz, const __grid_constant__ CUtensorMap tm_w,
                const __grid_constant__ CUtensorMap tm_dx, int B, int K, int C, int planes,
                const float* __restrict__ gscale, const double* __restrict__ den,
                float4* __restrict__ zero_dst, unsigned long long zero_n4, long long* trace, int dbg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    uint8_t* sStage = smem + kDx2OffStage;
    if (zero_dst) {
        // dW for the GEMM behind this kernel, whose batch splits both ADD their tile at the destination: every CTA clears a
        // slice on its way in (nothing in front of this launch touches dW, so this runs under the forward's tail)
        const unsigned long long per = (zero_n4 + gridDim.x - 1) / gridDim.x;
        const unsigned long long lo = per * blockIdx.x, hi = lo + per < zero_n4 ? lo + per : zero_n4;
        for (unsigned long long i = lo + threadIdx.x; i < hi; i += kDx2Threads) zero_dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDx2OffBars);
    uint64_t* full = bars; uint64_t* empty = full + kDx2Stages;
    uint64_t* t_full = empty + kDx2Stages; uint64_t* t_empty = t_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank_();
    if (threadIdx.x == 0) {
        for (int i = 0; i < kDx2Stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 2 * 4 * kDx2EW); }     // epilogue warps x 2 CTAs
        fence_barrier_init();
        prefetch_tensormap(&tm_dz); prefetch_tensormap(&tm_w); prefetch_tensormap(&tm_dx);
    }
    if (warp == 1) tmem_alloc2(tmem_slot, 512);
    tc_fence_before_sync();
    cluster_sync_();
    tc_fence_after_sync();
    pdl_wait();                 // dZ comes from the kernel in front of this one
    pdl_trigger();
    const uint32_t tbase = *tmem_slot;
    DxShape d;
    d.B = B; d.K = K; d.C = C; d.planes = planes; d.first = (int)(blockIdx.x >> 1); d.stride = (int)(gridDim.x >> 1);
    if (warp == 0) {
        uint32_t it = 0;
        dx_producer<kDx2Stages>(smem, full, empty, it, &tm_dz, &tm_w, d, rank, lane, dbg, nullptr, 0u, 0);
    } else if (warp == 1) {
        if (rank == 0) {
            uint32_t it = 0, tc = 0;
            dx_mma<kDx2Stages>(smem, full, empty, t_full, t_empty, it, tc, tbase, d, lane, trace);
        }
        __syncwarp();
    } else {
        uint32_t tc = 0;
        dx_epilogue<F32OUT>(sStage + (warp - kEpiWarp0) * 2 * kStageWarpBytes, t_full, t_empty, tc, tbase, &tm_dx, d, rank, warp, lane,
                            gscale, den, trace, dbg);
    }
    tc_fence_before_sync();
    cluster_sync_();
    if (warp == 1) tmem_dealloc2(tbase, 512);
}

// =================================================================================================
// dW partials = dZ^T X over a slice of the batch: both operands MN-major views of row-major tiles
// =================================================================================================
constexpr int kDwBN = 256;
constexpr int kDwStages = 4;
constexpr int kDwStageBytes = 2 * kBoxBytes + 4 * kBoxBytes;    // 48 KB: dZ 2 boxes + X 4 boxes, 64 rows each
constexpr int kDwOffBars = kDwStages * kDwStageBytes;
constexpr int kDwSmemBytes = kDwOffBars + (2 * kDwStages + 1) * 8 + 16 + 1024;

__global__ void __launch_bounds__(kThreads, 1)
head_dw_kernel(const __grid_constant__ CUtensorMap tm_dz, const __grid_constant__ CUtensorMap tm_x,
               int B, int K, int C, int planes, int splits, float* __restrict__ partials) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDwOffBars);
    uint64_t* full = bars; uint64_t* empty = full + kDwStages; uint64_t* t_full = empty + kDwStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kDwStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(t_full, 1);
        fence_barrier_init();
        prefetch_tensormap(&tm_dz); prefetch_tensormap(&tm_x);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    pdl_wait();                 // dZ comes from the forward kernel in front of this one
    pdl_trigger();              // the split-K reduce behind it may be scheduled
    const uint32_t tbase = *tmem_slot;
    const int n_tiles = (K + kDwBN - 1) / kDwBN;
    const int unit = blockIdx.x;
    const int sp = unit % splits, tile = unit / splits;
    const int mt = tile / n_tiles, nt = tile % n_tiles;
    // the reduction runs over the batch rows of every dZ plane (plane 1 = rounding residual, same X rows)
    const int kb_rows = (B + BK - 1) / BK, kb_total = kb_rows * planes;
    const int kb0 = (int)((int64_t)kb_total * sp / splits), kb1 = (int)((int64_t)kb_total * (sp + 1) / splits);

    // warps 0 and 1 run their issue loops converged, one elected lane issues (see head_fwd_kernel)
    if (warp == 0) {
        uint32_t it = 0;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
            const uint32_t s = it % kDwStages, ph = (it / kDwStages) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* st = smem + s * kDwStageBytes;
            const int pl = kb >= kb_rows ? 1 : 0, kr = kb - pl * kb_rows;
            if (elect_one()) {
                mbar_arrive_expect_tx(&full[s], kDwStageBytes);
#pragma unroll
                for (int j = 0; j < 2; ++j) tma_load_3d(st + j * kBoxBytes, &tm_dz, &full[s], mt * BM + j * 64, kr * BK, pl);
#pragma unroll
                for (int j = 0; j < 4; ++j) tma_load_2d(st + (2 + j) * kBoxBytes, &tm_x, &full[s], nt * kDwBN + j * 64, kr * BK);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        const uint32_t idesc = make_idesc_bf16(BM, kDwBN, true, true);
        uint32_t it = 0;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
            const uint32_t s = it % kDwStages, ph = (it / kDwStages) & 1;
            mbar_wait(&full[s], ph);
            tc_fence_after_sync();
            const uint32_t a_addr = smem_u32(smem + s * kDwStageBytes);
            const uint64_t adesc = make_desc_mnmajor(a_addr, kBoxBytes), bdesc = make_desc_mnmajor(a_addr + 2 * kBoxBytes, kBoxBytes);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) umma_bf16(tbase, adesc + 128 * k, bdesc + 128 * k, idesc, (it | k) != 0);
                umma_commit(&empty[s]);
                if (kb == kb1 - 1) umma_commit(t_full);
            }
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        const int cls = mt * BM + q * 32 + lane;
        mbar_wait(t_full, 0);
        tc_fence_after_sync();
        float* dst = partials + ((size_t)sp * C + (size_t)(cls < C ? cls : 0)) * K + (size_t)nt * kDwBN;
        const uint32_t t_row = tmem_addr(tbase, q * 32, 0);
#pragma unroll 1
        for (int c0 = 0; c0 < kDwBN; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(t_row + c0, v);
            tmem_ld_wait();
            if (cls < C) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    if (nt * kDwBN + c0 + i < K) {
                        float4 o = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]),
                                               __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
                        *reinterpret_cast<float4*>(dst + c0 + i) = o;
                    }
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tbase, 256);
}

// =================================================================================================
// dW partials with CTA pairs (cta_group::2): one 2-CTA cluster per 256 classes x 256 features x batch slice, M = 256
// MMAs.  CTA r stages classes [128 r, +128) of the dZ block and features [128 r, +128) of the X block (32 KB per
// k-block instead of 48 KB): the single-CTA kernel needs 96 B/clk of operand reads plus 96 B/clk of TMA fill from a
// shared memory that delivers 128 B/clk; the pair needs 64 + 64.
// =================================================================================================
constexpr int kDw2Stages = 6;
constexpr int kDw2StageBytes = 4 * kBoxBytes;                   // 32 KB: dZ 2 boxes + X 2 boxes, 64 batch rows each
constexpr int kDw2OffBars = kDw2Stages * kDw2StageBytes;
constexpr int kDw2SmemBytes = kDw2OffBars + (2 * kDw2Stages + 1) * 8 + 16 + 1024;
static_assert(kDw2SmemBytes <= 232448, "dw2 kernel shared memory");

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
head_dw2_kernel(const __grid_constant__ CUtensorMap tm_dz, const __grid_constant__ CUtensorMap tm_x,
                const __grid_constant__ CUtensorMap tm_part, const __grid_constant__ CUtensorMap tm_dw,
                int B, int K, int C, int planes, int splits, int direct, const float* __restrict__ gscale,
                const double* __restrict__ den, unsigned int* __restrict__ tile_flags,
                const unsigned int* row_flags, int row_flag_slots, long long* trace) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDw2OffBars);
    uint64_t* full = bars; uint64_t* empty = full + kDw2Stages; uint64_t* t_full = empty + kDw2Stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank_();
    if (trace && threadIdx.x == 0) trace[blockIdx.x * 8 + 0] = gtime();
    if (threadIdx.x == 0) {
        for (int i = 0; i < kDw2Stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(t_full, 1);
        fence_barrier_init();
        prefetch_tensormap(&tm_dz); prefetch_tensormap(&tm_x); prefetch_tensormap(direct ? &tm_dw : &tm_part);
    }
    if (warp == 1) tmem_alloc2(tmem_slot, 256);
    tc_fence_before_sync();
    cluster_sync_();
    tc_fence_after_sync();
    // dZ comes from the forward kernel in front of this one: wait for all of it -- or (row_flags) run under it on the SMs it
    // leaves idle and take the rows of dZ as the forward publishes them, 256-row tile by tile
    if (!row_flags) pdl_wait();
    pdl_trigger();              // the split-K reduce behind it may be scheduled
    const uint32_t tbase = *tmem_slot;
    const int n_tiles = (K + kDwBN - 1) / kDwBN;
    const int unit = blockIdx.x >> 1;
    // split-major: the first CTA pairs (the ones that become resident first) take the first rows of the batch
    const int tiles_all = ((C + 255) / 256) * n_tiles;
    const int sp = unit / tiles_all, tile = unit % tiles_all;
    const int mt = tile / n_tiles, nt = tile % n_tiles;
    const int kb_rows = (B + BK - 1) / BK, kb_total = kb_rows * planes;     // batch rows of every dZ plane
    const int kb0 = (int)((int64_t)kb_total * sp / splits), kb1 = (int)((int64_t)kb_total * (sp + 1) / splits);

    // warps 0 and 1 run their issue loops converged, one elected lane issues (see head_fwd_kernel)
    if (warp == 0) {
        uint32_t it = 0;
        int ready_lo = 0, ready_hi = 0;                     // row tiles [ready_lo, ready_hi) of dZ are known to have landed
        const int n_row_tiles = (B + kRowFlagTileRows - 1) / kRowFlagTileRows;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
            const uint32_t s = it % kDw2Stages, ph = (it / kDw2Stages) & 1;
            const int pl = kb >= kb_rows ? 1 : 0, kr = kb - pl * kb_rows;
            if (row_flags && !wait_row_tiles(row_flags, 1u, (kr * BK) / kRowFlagTileRows, n_row_tiles, row_flag_slots, lane, ready_lo, ready_hi)) {
                pdl_wait();                                     // not behind a publishing forward after all: plain dependency
                ready_lo = 0; ready_hi = 1 << 30;
            }
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* st = smem + s * kDw2StageBytes;
            if (elect_one()) {
                if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * kDw2StageBytes);     // both CTAs' bytes land on the leader's barrier
#pragma unroll
                for (int j = 0; j < 2; ++j) tma_load_3d_2sm(st + j * kBoxBytes, &tm_dz, &full[s], mt * 256 + rank * 128 + j * 64, kr * BK, pl);
#pragma unroll
                for (int j = 0; j < 2; ++j) tma_load_2d_2sm(st + (2 + j) * kBoxBytes, &tm_x, &full[s], nt * kDwBN + rank * 128 + j * 64, kr * BK);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        if (rank == 0) {
            const uint32_t idesc = make_idesc_bf16(256, kDwBN, true, true);
            uint32_t it = 0;
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const uint32_t s = it % kDw2Stages, ph = (it / kDw2Stages) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after_sync();
                const uint32_t a_addr = smem_u32(smem + s * kDw2StageBytes);
                const uint64_t adesc = make_desc_mnmajor(a_addr, kBoxBytes), bdesc = make_desc_mnmajor(a_addr + 2 * kBoxBytes, kBoxBytes);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) umma2_bf16(tbase, adesc + 128 * k, bdesc + 128 * k, idesc, (it | k) != 0);
                    umma2_commit_mc(&empty[s]);
                    if (kb == kb1 - 1) umma2_commit_mc(t_full);
                }
                __syncwarp();
            }
        }
    } else {
        const int q = warp & 3;
        mbar_wait_warp(t_full, 0);
        tc_fence_after_sync();
        if (row_flags) pdl_wait();      // den may come from the kernels in front; by now they are done or nearly so
        if (trace && threadIdx.x == kEpiWarp0 * 32) trace[blockIdx.x * 8 + 1] = gtime();
        // Every MMA has completed, so the operand stages are free: the warp stages its 32 classes x 256 features (fp32) there
        // as eight 128B-swizzled boxes and hands them to TMA stores, which write whole lines of the partial matrix and clip
        // the classes >= C (direct stores put one 16-byte piece per class row and instruction: 5.2 us per CTA, measured).
        uint8_t* my_stage = smem + q * (8 * kStageWarpBytes);
        const uint32_t t_row = tmem_addr(tbase, q * 32, 0);
        // direct mode (at most two splits): the scaled tile goes straight into dW -- split 0 stores, split 1 ADDS at the
        // destination (TMA reduce) once split 0's rows have landed; a + b is the same whichever order the hardware took,
        // so the result is as deterministic as the separate fixed-order reduce kernel it replaces, one pass over dW cheaper
        const float gs = direct ? (float)((gscale ? (double)gscale[0] : 1.0) / den[0]) : 1.f;
#pragma unroll 1
        for (int g = 0; g < kDwBN / 32; g += 2) {
            uint32_t v[2][32];
            tmem_ld32(t_row + g * 32, v[0]);
            tmem_ld32(t_row + g * 32 + 32, v[1]);
            tmem_ld_wait();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint8_t* sbuf = my_stage + (g + h) * kStageWarpBytes;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<float4*>(sbuf + lane * 128 + ((c ^ (lane & 7)) << 4)) =
                        make_float4(gs * __uint_as_float(v[h][4 * c]), gs * __uint_as_float(v[h][4 * c + 1]),
                                    gs * __uint_as_float(v[h][4 * c + 2]), gs * __uint_as_float(v[h][4 * c + 3]));
            }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            const int row0 = mt * 256 + rank * 128 + q * 32;
            if (!direct) {
#pragma unroll
                for (int g = 0; g < kDwBN / 32; ++g)
                    tma_store_3d(&tm_part, my_stage + g * kStageWarpBytes, nt * kDwBN + g * 32, row0, sp);
                tma_store_commit();
                tma_store_wait_read<0>();
            } else if (direct == 2) {
                // dW was cleared by the kernel in front: both splits add at the destination, nobody waits for anybody
                // (0 + a + b is the same bits in either order)
#pragma unroll
                for (int g = 0; g < kDwBN / 32; ++g)
                    tma_reduce_add_2d(&tm_dw, my_stage + g * kStageWarpBytes, nt * kDwBN + g * 32, row0);
                tma_store_commit();
                tma_store_wait_read<0>();
            } else {
                unsigned int* flag = tile_flags + ((size_t)tile * 2 + rank) * 4 + q;
                if (sp == 0) {
#pragma unroll
                    for (int g = 0; g < kDwBN / 32; ++g)
                        tma_store_2d(&tm_dw, my_stage + g * kStageWarpBytes, nt * kDwBN + g * 32, row0);
                    tma_store_commit();
                    if (splits > 1) {
                        tma_store_wait<0>();                      // the rows are in global memory, not merely read out of shared
                        __threadfence();
                        asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(flag), "r"(1u) : "memory");
                    } else {
                        tma_store_wait_read<0>();
                    }
                } else {
                    unsigned int f;
                    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(flag) : "memory"); } while (f == 0u);
                    asm volatile("fence.proxy.async;" ::: "memory");
#pragma unroll
                    for (int g = 0; g < kDwBN / 32; ++g)
                        tma_reduce_add_2d(&tm_dw, my_stage + g * kStageWarpBytes, nt * kDwBN + g * 32, row0);
                    tma_store_commit();
                    *flag = 0u;                                   // consumed: the workspace is left reusable
                    tma_store_wait_read<0>();
                }
            }
        }
        __syncwarp();
    }
    if (trace && threadIdx.x == kEpiWarp0 * 32) trace[blockIdx.x * 8 + 2] = gtime();
    tc_fence_before_sync();
    cluster_sync_();
    if (warp == 1) tmem_dealloc2(tbase, 256);
    if (trace && threadIdx.x == 0) trace[blockIdx.x * 8 + 3] = gtime();
}

// dW = g * sum_s partials[s]  (fixed order => deterministic)
__global__ void __launch_bounds__(256)
dw_reduce_kernel(const float4* partials, int splits, size_t n4, const float* __restrict__ gscale,
                 const double* __restrict__ den, float4* dw) {
    pdl_wait();                 // the partials come from the dW GEMM in front of this kernel
    const float gs = den ? (float)((gscale ? (double)gscale[0] : 1.0) / den[0]) : 1.f;      // den == NULL: plain sum
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 acc = partials[i];
        for (int s = 1; s < splits; ++s) {
            const float4 p = partials[(size_t)s * n4 + i];
            acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
        }
        acc.x *= gs; acc.y *= gs; acc.z *= gs; acc.w *= gs;
        dw[i] = acc;
    }
}

// db = g * column sums of dZ (bf16 [planes][B, lddz]); one block per 64 columns, fixed summation order.  den == NULL:
// un-scaled sums (input of the cross-GPU all-reduce, which applies g / den itself).
__global__ void __launch_bounds__(256)
db_kernel(const __nv_bfloat16* __restrict__ dz, int64_t lddz, int B, int C, int planes, const float* __restrict__ gscale,
          const double* __restrict__ den, float* __restrict__ db) {
    __shared__ float sh[4][64];
    const int c = blockIdx.x * 64 + (threadIdx.x & 63);
    const int part = threadIdx.x >> 6;
    float acc = 0.f;
    if (c < C) {
        for (int r = part; r < B; r += 4) acc += __bfloat162float(dz[(int64_t)r * lddz + c]);
        if (planes > 1) {
            const __nv_bfloat16* lo = dz + (int64_t)B * lddz;
            float acc2 = 0.f;
            for (int r = part; r < B; r += 4) acc2 += __bfloat162float(lo[(int64_t)r * lddz + c]);
            acc += acc2;
        }
    }
    sh[part][threadIdx.x & 63] = acc;
    __syncthreads();
    if (threadIdx.x < 64 && c < C) {
        const float gs = den ? (float)((gscale ? (double)gscale[0] : 1.0) / den[0]) : 1.f;
        db[c] = gs * (((sh[0][threadIdx.x] + sh[1][threadIdx.x]) + sh[2][threadIdx.x]) + sh[3][threadIdx.x]);
    }
}

// =================================================================================================
// forward for small batches: split-K over a cluster of 8 CTAs, reduce-scatter through distributed shared memory
// =================================================================================================
// With B <= 2048 rows the persistent kernel above keeps at most 8 CTA pairs busy, and each of them walks all of K
// (one 256-row tile: ~22 us for K = 2048, whatever the tensor pipe could do).  Here a cluster of 8 CTAs owns one
// 128-row tile: CTA r multiplies k-blocks [n_kb r / 8, n_kb (r + 1) / 8) into a 128 x C partial in its TMEM, the
// partials are reduce-scattered through distributed shared memory (CTA r ends up with rows [16 r, +16) of the tile, the 8
// partials of a row summed in rank order: deterministic), and 8 threads per row run the same softmax / loss /
// grad-logits epilogue as head_fwd_kernel.  Every operand byte is loaded by exactly one CTA of the cluster.
constexpr int kSkS = 8;                                          // CTAs per cluster = K slices = row owners
constexpr int kSkRows = BM / kSkS;                               // rows of the tile a CTA owns after the reduce-scatter
constexpr int kSkLanes = 128 / kSkRows;                          // epilogue threads per owned row
constexpr int kSkStages = 4;
constexpr int kSkBN = 256;                                       // widest head of this path = TMEM columns
constexpr int kSkStageBytes = kABytes + kSkBN * BK * 2;          // 48 KB: X 128 rows + W up to 256 rows, 128 B each
constexpr int kSkOffBars = kSkStages * kSkStageBytes;            // the stages double as the exchange buffer (128 KB at most)
constexpr int kSkOffLoss = kSkOffBars + (2 * kSkStages + 1) * 8 + 16;
constexpr int kSkSmemBytes = kSkOffLoss + kSkRows * 8 + 1024;
static_assert(kSkS * (kSkBN / 4) * kSkRows * 16 <= kSkStages * kSkStageBytes, "exchange buffer must fit the operand stages");
static_assert(kSkSmemBytes <= 232448, "split-K forward shared memory");
static_assert(kSkLanes == 8, "the row epilogue reduces over 8 lanes");

__device__ __forceinline__ void cluster_sync_relaxed() {
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_peer_v4(uint32_t raddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(raddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
template <int W> __device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
template <int W> __device__ __forceinline__ float group_max(float v) {
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

template <bool HAS_BIAS>
__global__ void __cluster_dims__(kSkS, 1, 1) __launch_bounds__(kThreads, 1)
head_fwd_splitk_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, const FwdArgs a,
                       __nv_bfloat16* __restrict__ dz, int64_t lddz) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSkOffBars);
    uint64_t* full = bars; uint64_t* empty = full + kSkStages; uint64_t* t_full = empty + kSkStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_full + 1);
    double* sLoss = reinterpret_cast<double*>(smem + kSkOffLoss);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank_();
    const int tile = (int)(blockIdx.x / kSkS);
    long long* trace = a.trace ? a.trace + (size_t)blockIdx.x * 64 : nullptr;     // developer hook: phase stamps of this CTA
    auto stamp = [&](int slot) { if (trace && threadIdx.x == kEpiWarp0 * 32) trace[slot] = gtime(); };
    stamp(0);
    if (trace && threadIdx.x == kEpiWarp0 * 32) trace[20] = clock64();
    const int n_kb = (a.K + BK - 1) / BK;
    const int kb0 = n_kb * rank / kSkS, kb1 = n_kb * (rank + 1) / kSkS;
    const bool has_k = kb1 > kb0;
    const int ncolsN = (a.C + 15) & ~15;            // UMMA N = rows of the W box
    const int ncolsP = (a.C + 31) & ~31;            // columns drained and exchanged (whole 32-column TMEM loads)
    const int nc4 = ncolsP >> 2;                    // float4 groups per row; a multiple of 8
    if (threadIdx.x == 0) {
        for (int i = 0; i < kSkStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(t_full, 1);
        fence_barrier_init();
        prefetch_tensormap(&tm_x); prefetch_tensormap(&tm_w);
    }
    if (warp == 1) tmem_alloc(tmem_slot, kSkBN);
    tc_fence_before_sync();
    __syncthreads();                                                  // nothing remote is touched before the cluster barrier below
    tc_fence_after_sync();
    const uint32_t tbase = *tmem_slot;
    stamp(1);
    pdl_trigger();              // the dW GEMM behind this kernel may set itself up; it waits for dZ (pdl_wait) before it loads
    // The epilogue threads have nothing to do until the MMAs complete: the label of the row a thread will own after the
    // reduce-scatter and its class weight (two dependent global round trips) are fetched now, under the operand loads.
    const int te = (int)threadIdx.x - kEpiWarp0 * 32;
    const int row16 = te / kSkLanes, sub = te % kSkLanes;
    const int row = tile * BM + rank * kSkRows + row16;
    const bool row_ok = te >= 0 && row < a.B;
    int64_t y = -1;
    RowMode rm; rm.mode = kRowIgnored; rm.omega = 0.f;
    if (te >= 0 && row_ok) {
        y = a.y[row];
        rm = classify_row(a.variant, y, a.C, a.ignore_index, a.class_w, a.t_uniform * (float)a.C);
    }

    if (warp == 0) {
        // ------------------------------ TMA producer: this CTA's k-blocks of the X tile and of W ----------------------
        const uint32_t stage_tx = (uint32_t)(kABytes + ncolsN * BK * 2);
        uint32_t it = 0;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
            const uint32_t s = it % kSkStages, ph = (it / kSkStages) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* st = smem + s * kSkStageBytes;
            if (elect_one()) {
                mbar_arrive_expect_tx(&full[s], stage_tx);
                tma_load_2d(st, &tm_x, &full[s], kb * BK, tile * BM);
                tma_load_2d(st + kABytes, &tm_w, &full[s], kb * BK, 0);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer (warp-converged loop, one elected lane issues) ----------------------
        const uint32_t idesc = make_idesc_bf16(BM, ncolsN, false, false);
        uint32_t it = 0;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
            const uint32_t s = it % kSkStages, ph = (it / kSkStages) & 1;
            mbar_wait(&full[s], ph);
            tc_fence_after_sync();
            const uint32_t a_addr = smem_u32(smem + s * kSkStageBytes);
            const uint64_t adesc = make_desc_kmajor(a_addr), bdesc = make_desc_kmajor(a_addr + kABytes);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) umma_bf16(tbase, adesc + 2 * k, bdesc + 2 * k, idesc, (it | (uint32_t)k) != 0u);
                umma_commit(&empty[s]);
                if (kb == kb1 - 1) umma_commit(t_full);
            }
            __syncwarp();
        }
    } else if (has_k) {
        mbar_wait_warp(t_full, 0);
        tc_fence_after_sync();
    }
    stamp(2);
    // every MMA of every CTA of the cluster has completed (the epilogue warps arrive after their t_full): the operand
    // stages of all 8 CTAs are free and become the exchange buffer  [source rank][float4 group][owned row] x 16 bytes.
    // Execution ordering only (the reads it protects are complete): the relaxed arrive spares the MEMBAR.ALL.GPU that
    // the release form compiles to.
    cluster_sync_relaxed();
    stamp(3);

    if (warp >= kEpiWarp0) {
        const int q = warp & 3;
        const int r = q * 32 + lane;                                  // row of the partial tile = TMEM lane
        const uint32_t peer = map_to_peer(smem_u32(smem), (uint32_t)(r / kSkRows));
        const uint32_t dst0 = peer + (uint32_t)(((rank * nc4) * kSkRows + (r % kSkRows)) * 16);
        const uint32_t t_row = tmem_addr(tbase, q * 32, 0);
#pragma unroll 1
        for (int c0 = 0; c0 < ncolsP; c0 += 32) {
            uint32_t v[32];
            if (has_k) {
                tmem_ld32(t_row + c0, v);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0u;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
                st_peer_v4(dst0 + (uint32_t)((((c0 >> 2) + i) * kSkRows) * 16), v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
    }
    stamp(4);
    tc_fence_before_sync();
    cluster_sync_();                                                  // partials delivered (release / acquire at cluster scope)
    stamp(5);

    if (warp >= kEpiWarp0) {
        const float4* rbuf = reinterpret_cast<const float4*>(smem);
        constexpr int NJ = kSkBN / 32;                                // float4 groups per thread, at most
        const int nj = nc4 / kSkLanes;
        // thread (row, sub) owns the float4 groups sub, sub + 8, ...: logits = the 8 partials summed in rank order
        float z[NJ][4];
        float m = -INFINITY, mv = -INFINITY;
        int am = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            if (j < nj) {
                const int c4 = sub + kSkLanes * j;
                float4 acc = rbuf[(size_t)(c4 * kSkRows + row16)];
#pragma unroll
                for (int src = 1; src < kSkS; ++src) {
                    const float4 p = rbuf[(size_t)((src * nc4 + c4) * kSkRows + row16)];
                    acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
                }
                z[j][0] = acc.x; z[j][1] = acc.y; z[j][2] = acc.z; z[j][3] = acc.w;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int col = 4 * c4 + e;
                    float v = z[j][e];
                    if (HAS_BIAS) v += (col < a.C) ? a.bias[col] : 0.f;
                    if (col >= a.C) v = -INFINITY;
                    z[j][e] = v;
                    if (v > m) { m = v; am = col; }               // columns ascend inside a thread: first arg-max
                    if (col < a.valid_cols) mv = fmaxf(mv, v);
                }
            }
        }
        stamp(6);
        group_argmax<kSkLanes>(m, am);                                // ties go to the smaller column (numpy)
        mv = group_max<kSkLanes>(mv);
        float s = 0.f, sumz = 0.f, zy = 0.f;
        const float mb = m * kLog2e;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            if (j < nj) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int col = 4 * (sub + kSkLanes * j) + e;
                    const float v = z[j][e];
                    if (col < a.C) sumz += v;
                    if ((int64_t)col == y) zy = v;
                    const float ex = fast_exp2(fmaf(v, kLog2e, -mb));     // masked columns: exp2(-inf) = 0
                    s += ex;
                    z[j][e] = ex;
                }
            }
        }
        s = group_sum<kSkLanes>(s); sumz = group_sum<kSkLanes>(sumz); zy = group_sum<kSkLanes>(zy);
        // ---- denominator-free grad-logits  dz = S p - T = coef * e - T  (bf16, 8-byte stores: 4 columns); with two planes
        //      the second one receives what the bf16 rounding of the first dropped ----
        if (a.write_dz && row_ok) {
            const bool live = (rm.mode == kRowOneHot || rm.mode == kRowUniform);
            const float coef = live ? rm.omega / s : 0.f;
            const float t_all = (rm.mode == kRowUniform) ? a.t_uniform : 0.f;
            const float t_hot = (rm.mode == kRowOneHot) ? rm.omega : 0.f;
            __nv_bfloat16* drow = dz + (int64_t)row * lddz;
            __nv_bfloat16* drow_lo = a.dz_planes > 1 ? drow + (int64_t)a.B * lddz : nullptr;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                if (j < nj) {
                    const int col0 = 4 * (sub + kSkLanes * j);
                    if (col0 < a.C) {
                        float g[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            g[e] = fmaf(coef, z[j][e], -t_all) - (((int64_t)(col0 + e) == y) ? t_hot : 0.f);
                        if (col0 + 4 <= a.C) {
                            __nv_bfloat162 p0 = __floats2bfloat162_rn(g[0], g[1]), p1 = __floats2bfloat162_rn(g[2], g[3]);
                            uint2 u;
                            u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
                            *reinterpret_cast<uint2*>(drow + col0) = u;
                            if (drow_lo) {
                                const float2 h0 = __bfloat1622float2(p0), h1 = __bfloat1622float2(p1);
                                __nv_bfloat162 q0 = __floats2bfloat162_rn(g[0] - h0.x, g[1] - h0.y), q1 = __floats2bfloat162_rn(g[2] - h1.x, g[3] - h1.y);
                                u.x = *reinterpret_cast<uint32_t*>(&q0); u.y = *reinterpret_cast<uint32_t*>(&q1);
                                *reinterpret_cast<uint2*>(drow_lo + col0) = u;
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                if (col0 + e < a.C) {
                                    const __nv_bfloat16 hi = __float2bfloat16_rn(g[e]);
                                    drow[col0 + e] = hi;
                                    if (drow_lo) drow_lo[col0 + e] = __float2bfloat16_rn(g[e] - __bfloat162float(hi));
                                }
                            }
                        }
                    }
                }
            }
        }
        stamp(7);
        // ---- per-row outputs and the row's loss term ----
        if (sub == 0) {
            double rl = 0.0;
            if (row_ok) {
                const float lse = m + fast_log2(s) * kLn2;
                const bool is_col = (y >= 0 && y < a.C);
                float rowloss = 0.f;
                if (rm.mode == kRowOneHot) rowloss = rm.omega * (lse - zy);
                else if (rm.mode == kRowUniform) rowloss = a.t_uniform * ((float)a.C * lse - sumz);
                else if (rm.mode == kRowBad) rowloss = __int_as_float(0x7fc00000);
                rl = (double)rowloss;
                if (a.row_ptgt) a.row_ptgt[row] = is_col ? fast_exp2((zy - lse) * kLog2e) : 0.f;
                if (a.row_pmax) a.row_pmax[row] = fast_exp2((mv - lse) * kLog2e);
                if (a.row_pmax_all) a.row_pmax_all[row] = fast_exp2((m - lse) * kLog2e);
                if (a.row_pred) a.row_pred[row] = am;
            }
            sLoss[row16] = rl;
        }
        epi_bar_sync();
        // ---- deterministic loss: rows in order inside the CTA, CTAs in index order by the last one out ----
        if (warp == kEpiWarp0) {
            unsigned last = 0;
            if (lane == 0) {
                double t = 0.0;
#pragma unroll
                for (int i = 0; i < kSkRows; ++i) t += sLoss[i];
                a.partials[blockIdx.x] = t;
                __threadfence();
                last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
            }
            last = __shfl_sync(kFull, last, 0);
            if (last) {
                __threadfence();
                double tot = 0.0;
                for (unsigned b = lane; b < gridDim.x; b += 32) tot += __ldcg(&a.partials[b]);
                tot = warp_sum(tot);
                if (lane == 0) {
                    a.loss_num[0] = tot;
                    // Launched with programmatic stream serialization behind the kernel that writes den
                    // (OSB_FWD_DEN_FROM_PREV_KERNEL), this grid has run under it: wait for it only now
                    if (a.den_from_prev) pdl_wait();
                    a.loss[0] = a.den ? (float)(tot / a.den[0]) : (float)tot;
                    *a.ticket = 0u;
                }
            }
        }
    }
    stamp(8);
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tbase, kSkBN);
    stamp(9);
    if (trace && threadIdx.x == kEpiWarp0 * 32) trace[21] = clock64();
}

// ---- host --------------------------------------------------------------------------------------
// the dW GEMM runs on CTA pairs (256 classes per pair) when the head is wider than one single-CTA tile
static bool dw_pairs(int C) { return C > BM; }

// split-K factor of the dW GEMM: output tiles x splits fill the SMs (pairs) once
static int dw_splits(int B, int K, int C) {
    const bool pair = dw_pairs(C);
    const int mt = pair ? (C + 255) / 256 : (C + BM - 1) / BM, nt = (K + kDwBN - 1) / kDwBN;
    const int kb_total = (B + BK - 1) / BK;
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    if (pair) sms /= 2;
    int s = sms / (mt * nt);
    if (s < 1) s = 1;
    if (s > 16) s = 16;
    if (s > kb_total) s = kb_total < 1 ? 1 : kb_total;
    // (two splits for small batches too, so that the GEMM writes dW itself, were measured at the BG configuration, 1024 x 2048
    //  -> 152: 33.2 -> 32.6 us per step on one GPU, 56.6 us unchanged on two: not adopted)
    return s;
}

// once per kernel (per process): the attribute is sticky
template <typename KernelT>
static int set_smem(KernelT kernel, int bytes) {
    static bool done = false;          // one instance per KernelT instantiation... and per kernel pointer below
    static const void* last = nullptr;
    if (done && last == reinterpret_cast<const void*>(kernel)) return OSB_OK;
    OSB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done = true; last = reinterpret_cast<const void*>(kernel);
    return OSB_OK;
}

template <int BN, int CL, int STAGES, int EW, bool PAIR, bool HAS_BIAS, bool WANT_PRED, bool PLANES2, bool FUSE = false>
static int launch_fwd_k(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& tdz, const FwdArgs& a,
                        int max_ctas, cudaStream_t st, const DxFuse* fuse = nullptr) {
    using Cfg = FwdCfg<BN, CL, STAGES, EW, PAIR>;
    constexpr int kCtasPerGroup = CL * (PAIR ? 2 : 1);
    const int n_tiles = (a.B + Cfg::kTileRows - 1) / Cfg::kTileRows;
    auto k = head_fwd_kernel<BN, CL, STAGES, EW, PAIR, HAS_BIAS, WANT_PRED, PLANES2, FUSE>;
    int rc = set_smem(k, Cfg::kSmemBytes); if (rc) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(Cfg::kThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = st;
    // persistent kernel, one CTA per SM, launched cooperatively when units wait on each other (CL > 1): every CTA of a
    // group must be resident while its peers spin on its statistics flags
    cudaLaunchAttribute attr[2];
    int na = 0;
    // (developer switch for Nsight Compute, which fails to replay a launch that is both cooperative and clustered: the
    // grid never exceeds the SM count, so on an otherwise idle GPU all CTAs are resident without the attribute too)
    static const bool no_coop = getenv("OSB_NO_COOP_LAUNCH") != nullptr;
    // (FUSE: every CTA must have read the launch epoch of the row-tile words before the forward's last warp advances it,
    //  i.e. all CTAs start together -- the same guarantee)
    if ((CL > 1 || FUSE) && !(no_coop && PAIR)) { attr[na].id = cudaLaunchAttributeCooperative; attr[na].val.cooperative = 1; ++na; }
    if (PAIR) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1; ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    int groups = cta_limit(max_ctas) / kCtasPerGroup;
    if (groups * (PAIR ? 2 : 1) > kMaxXid) groups = kMaxXid / (PAIR ? 2 : 1);
    if (groups < 1) groups = 1;
    // every group runs ceil(n_tiles / groups) rounds: use the fewest groups that keep that round count (the others
    // would only add L2 traffic contention and power)
    const int rounds = (n_tiles + groups - 1) / groups;
    groups = (n_tiles + rounds - 1) / rounds;
    cfg.gridDim = dim3(groups * kCtasPerGroup);
    DxFuse fz;
    if (FUSE) {
        // every SM: the units beyond the forward's groups start with dX tiles at once
        fz = *fuse;
        fz.fwd_units = groups * CL;
        cfg.gridDim = dim3((unsigned)(cta_limit(max_ctas) / 2 * 2));
        if (cfg.gridDim.x < (unsigned)(groups * kCtasPerGroup)) cfg.gridDim = dim3(groups * kCtasPerGroup);
    } else {
        memset(&fz, 0, sizeof(fz));
    }
    OSB_CUDA(cudaLaunchKernelEx(&cfg, k, tx, tw, tdz, a, fz));
    return OSB_OK;
}

// the fused forward + dX kernel: 256-column units only, no per-row predictions (a training step does not read them)
template <int CL>
static int launch_fwd_dx(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& tdz, const FwdArgs& a,
                         int max_ctas, cudaStream_t st, const DxFuse& fuse) {
    const bool two = a.dz_planes > 1;
#define OSB_FWDX(BIAS_) (two ? launch_fwd_k<256, CL, 4, 2, true, BIAS_, false, true, true>(tx, tw, tdz, a, max_ctas, st, &fuse) \
                             : launch_fwd_k<256, CL, 4, 2, true, BIAS_, false, false, true>(tx, tw, tdz, a, max_ctas, st, &fuse))
    return a.bias ? OSB_FWDX(true) : OSB_FWDX(false);
#undef OSB_FWDX
}

template <int BN, int CL, int STAGES>
static int launch_fwd(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& tdz, const FwdArgs& a,
                      int max_ctas, cudaStream_t st) {
    const bool pred = a.row_pred || a.row_pmax_all;
    const bool two = a.write_dz && a.dz_planes > 1;
#define OSB_FWD(BIAS_, PRED_) (two ? launch_fwd_k<BN, CL, STAGES, 2, true, BIAS_, PRED_, true>(tx, tw, tdz, a, max_ctas, st) \
                                   : launch_fwd_k<BN, CL, STAGES, 2, true, BIAS_, PRED_, false>(tx, tw, tdz, a, max_ctas, st))
    if (a.bias) return pred ? OSB_FWD(true, true) : OSB_FWD(true, false);
    return pred ? OSB_FWD(false, true) : OSB_FWD(false, false);
#undef OSB_FWD
}

// rows of the W box one CTA of a pair stages per k-block
static int fwd_box_rows(int C) { return (C <= 128 ? 128 : 256) / 2; }

}  // namespace head
}  // namespace osb

using namespace osb;
using namespace osb::head;

// developer hook (osb_debug.h): 0 sends small batches through the persistent kernels too (A/B reference of the split-K forward)
static int g_head_splitk = 1;
extern "C" void osb_debug_set_head_splitk(int on) { g_head_splitk = on; }
// the split-K forward takes batches whose 128-row tiles x 8 CTAs fit the GPU at once (B <= 2048), heads of one tile
// (C <= 256) and at least one k-block per CTA of the cluster
// clusters of 8 CTAs (one CTA per SM: ~193 KB of shared memory) that the device keeps resident at once; measured on
// B200: a second wave of clusters costs more than the persistent kernel, so larger batches stay there
static int splitk_max_clusters() {
    static int cached = -1;
    if (cached >= 0) return cached;
    int n = 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kSkS * 32); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSkSmemBytes;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = kSkS; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    if (cudaFuncSetAttribute(head_fwd_splitk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSkSmemBytes) != cudaSuccess ||
        cudaOccupancyMaxActiveClusters(&n, head_fwd_splitk_kernel<false>, &cfg) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    cached = n;
    return cached;
}
extern "C" int osb_debug_head_splitk_max_clusters(void) { return splitk_max_clusters(); }
static bool fwd_use_splitk(int64_t B, int K, int C, int max_ctas) {
    if (!g_head_splitk || C > kSkBN || K < BK * kSkS) return false;
    const int64_t tiles = (B + BM - 1) / BM;
    return tiles * kSkS <= cta_limit(max_ctas) && tiles <= splitk_max_clusters();
}
extern "C" int osb_head_forward_is_splitk(int64_t B, int K, int C) {
    if (osb::check_device() != OSB_OK) return 0;
    return fwd_use_splitk(B, K, C, 0) ? 1 : 0;
}
// Row-tile hand-over between the persistent forward and the dW GEMM behind it (OSB_FWD_PUBLISH_ROW_TILES /
// OSB_BWD_DZ_FROM_PREV_FORWARD): both sides decide from the shape alone.  A forward that is bounded by max_ctas may run
// as the persistent kernel where this says split-K; it then publishes nothing and the backward does not ask (safe side).
static bool row_flags_usable(int64_t B, int K, int C) {
    static const bool off = getenv("OSB_NO_ROW_FLAGS") != nullptr;               // developer A/B switch
    return !off && C > BM && !fwd_use_splitk(B, K, C, 0) && (B + kRowFlagTileRows - 1) / kRowFlagTileRows <= kMaxRowFlagTiles;
}
static long long* g_trace = nullptr;
static int g_dbg = 0;
extern "C" void osb_debug_set_head_dbg(int v) { g_dbg = v; }
// developer hook (osb_debug.h): device buffer of [grid][64] int64 clock stamps, or NULL
extern "C" void osb_debug_set_head_trace(void* p) { g_trace = reinterpret_cast<long long*>(p); }
// developer hook (osb_debug.h): a separate stamp buffer for the dW GEMM ([grid][8]: entry, mainloop done, epilogue done, exit), so that one
// replay of a whole step can be laid out on a time line (tools/timeline.py); NULL = the dW GEMM stamps into the head trace
static long long* g_trace_dw = nullptr;
extern "C" void osb_debug_set_dw_trace(void* p) { g_trace_dw = reinterpret_cast<long long*>(p); }

// launch configuration with programmatic stream serialization: the kernel's prologue may overlap the tail of the kernel in
// front of it on the stream; the kernel itself calls pdl_wait before it touches anything that kernel produces
static cudaLaunchConfig_t pdl_config(dim3 grid, dim3 block, size_t smem, cudaStream_t st, cudaLaunchAttribute* attr) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    attr->id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr->val.programmaticStreamSerializationAllowed = 1;
    // developer switch for A/B measurements: OSB_NO_PDL=1 launches in plain stream order
    static const bool no_pdl = getenv("OSB_NO_PDL") != nullptr;
    cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
    return cfg;
}

extern "C" size_t osb_head_workspace_bytes(int64_t B, int K, int C) {
    if (B < 0 || K <= 0 || C <= 0) return 0;
    return kWsOffDw + align_up((size_t)dw_splits((int)B, K, C) * C * K * sizeof(float), 256) + 256;
}

#define OSB_TMAP(call, what)                                                              \
    do {                                                                                  \
        int rc__ = (call);                                                                \
        if (rc__ != 0) {                                                                  \
            osb::set_error("cuTensorMapEncodeTiled failed for %s (code %d)", what, rc__); \
            return OSB_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

static int check_head_shapes(const char* fn, int64_t B, int K, int C, int64_t ldx) {
    OSB_REQUIRE(B >= 1 && B < (1LL << 30), OSB_ERR_ARG, "%s: bad batch B=%lld", fn, (long long)B);
    OSB_REQUIRE(K >= 64 && K % 64 == 0, OSB_ERR_SHAPE, "%s: in_features K=%d must be a positive multiple of 64", fn, K);
    OSB_REQUIRE(C >= 2 && C <= 1024, OSB_ERR_SHAPE, "%s: head width C=%d outside [2,1024]", fn, C);
    OSB_REQUIRE(ldx >= K && ldx % 8 == 0, OSB_ERR_ARG, "%s: ldx=%lld must be >= K and a multiple of 8", fn, (long long)ldx);
    return OSB_OK;
}

// what osb_head_forward_dx adds to a forward call
struct FuseDxRequest { const float* gscale; void* dx; int64_t lddx; int dx_dtype; };

static int head_forward_impl(const void* x, int64_t ldx, const void* w, const float* bias, const int64_t* y,
                             const float* class_w, int64_t B, int K, int C, int variant, double unk_weight,
                             int64_t ignore_index, const double* den, int valid_cols,
                             double* loss_num, float* loss,
                             float* row_ptgt, float* row_pmax, float* row_pmax_all, int32_t* row_pred,
                             void* dz, int64_t lddz, int dz_planes, int flags, int max_ctas,
                             void* ws, size_t ws_bytes, void* stream, const FuseDxRequest* fuse, bool* fused_out) {
    OSB_REQUIRE(x && w && y && loss_num && loss && ws, OSB_ERR_ARG, "osb_head_forward: null pointer");
    int rc = check_head_shapes("osb_head_forward", B, K, C, ldx);
    if (rc) return rc;
    OSB_REQUIRE(variant >= OSB_ENTROPIC && variant <= OSB_GARBAGE, OSB_ERR_ARG, "osb_head_forward: bad variant %d", variant);
    OSB_REQUIRE(variant != OSB_GARBAGE || class_w, OSB_ERR_ARG, "osb_head_forward: garbage variant needs class_w");
    OSB_REQUIRE(valid_cols >= 1 && valid_cols <= C, OSB_ERR_ARG, "osb_head_forward: valid_cols=%d outside [1,%d]", valid_cols, C);
    OSB_REQUIRE(!dz || (lddz >= C && lddz % 8 == 0), OSB_ERR_ARG, "osb_head_forward: lddz=%lld must be >= C and a multiple of 8", (long long)lddz);
    OSB_REQUIRE(!dz || dz_planes == 1 || dz_planes == 2, OSB_ERR_ARG, "osb_head_forward: dz_planes=%d must be 1 or 2", dz_planes);
    OSB_REQUIRE(max_ctas >= 0 && (flags & ~(OSB_FWD_DEN_FROM_PREV_KERNEL | OSB_FWD_PUBLISH_ROW_TILES)) == 0, OSB_ERR_ARG,
                "osb_head_forward: bad flags / max_ctas");
    OSB_REQUIRE(!(flags & OSB_FWD_DEN_FROM_PREV_KERNEL) || den, OSB_ERR_ARG, "osb_head_forward: OSB_FWD_DEN_FROM_PREV_KERNEL without den");
    OSB_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)w % 16 == 0) && ((uintptr_t)dz % 16 == 0), OSB_ERR_ARG,
                "osb_head_forward: x, w and dz must be 16-byte aligned");
    OSB_REQUIRE(ws_bytes >= osb_head_workspace_bytes(B, K, C), OSB_ERR_WORKSPACE, "osb_head_forward: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap tx, tw;
    OSB_TMAP(make_tmap_bf16(&tx, x, (uint64_t)B, (uint64_t)K, (uint64_t)ldx, BM), "x");
    const bool splitk = fwd_use_splitk(B, K, C, max_ctas);
    OSB_TMAP(make_tmap_bf16(&tw, w, (uint64_t)C, (uint64_t)K, (uint64_t)K, splitk ? (uint32_t)((C + 15) & ~15) : (uint32_t)fwd_box_rows(C)), "w");
    CUtensorMap tdz = tx;       // placeholder when no grad-logits are written
    if (dz && !splitk)
        OSB_TMAP(make_tmap_bf16_3d(&tdz, dz, (uint64_t)dz_planes, (uint64_t)B, (uint64_t)C, (uint64_t)lddz, (uint64_t)B * (uint64_t)lddz, 32), "dz (store)");
    FwdArgs a;
    a.bias = bias; a.y = y; a.class_w = class_w; a.B = (int)B; a.K = K; a.C = C; a.variant = variant;
    a.t_uniform = (float)(unk_weight / (double)C);
    a.ignore_index = ignore_index; a.den = den; a.valid_cols = valid_cols;
    a.loss_num = loss_num; a.loss = loss;
    a.row_ptgt = row_ptgt; a.row_pmax = row_pmax; a.row_pmax_all = row_pmax_all; a.row_pred = row_pred;
    a.write_dz = dz ? 1 : 0;
    a.dz_planes = dz ? dz_planes : 1;
    a.den_from_prev = 0;
    a.ticket = reinterpret_cast<unsigned int*>(ws);
    a.partials = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + kWsOffPartials);
    a.xch_epoch = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(ws) + kWsOffEpoch);
    a.xch_stats = reinterpret_cast<uint2*>(reinterpret_cast<char*>(ws) + kWsOffStats);
    a.trace = g_trace;
    a.tile_flags = ((flags & OSB_FWD_PUBLISH_ROW_TILES) && dz && row_flags_usable(B, K, C))
                       ? reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(ws) + kWsOffRowFlags) : nullptr;
    a.flag_epoch = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(ws) + kWsOffFlagEpoch);
    if (fused_out) *fused_out = false;
    // forward + dX as ONE kernel: persistent 256-column units (heads wider than 128 classes, batches the split-K kernel
    // does not take), grad-logits and a denominator present, row-tile words for the whole batch
    static const bool no_fuse = getenv("OSB_NO_FUSED_DX") != nullptr;            // developer A/B switch
    if (fuse && !no_fuse && !splitk && C > BM && dz && den && !row_pred && !row_pmax_all &&
        (B + kRowFlagTileRows - 1) / kRowFlagTileRows <= kMaxRowFlagTiles) {
        DxFuse fz;
        memset(&fz, 0, sizeof(fz));
        OSB_TMAP(make_tmap_bf16_3d(&fz.tm_dz, dz, (uint64_t)dz_planes, (uint64_t)B, (uint64_t)C, (uint64_t)lddz, (uint64_t)B * (uint64_t)lddz, BM), "dz (K-major)");
        OSB_TMAP(make_tmap_bf16(&fz.tm_w, w, (uint64_t)C, (uint64_t)K, (uint64_t)K, 64), "w (MN-major)");
        if (fuse->dx_dtype == OSB_F32) OSB_TMAP(make_tmap_f32_2d(&fz.tm_dx, fuse->dx, (uint64_t)B, (uint64_t)K, (uint64_t)fuse->lddx, 32), "dx (fp32)");
        else OSB_TMAP(make_tmap_bf16(&fz.tm_dx, fuse->dx, (uint64_t)B, (uint64_t)K, (uint64_t)fuse->lddx, 32), "dx");
        fz.gscale = fuse->gscale;
        fz.dx_f32 = fuse->dx_dtype == OSB_F32 ? 1 : 0;
        a.tile_flags = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(ws) + kWsOffRowFlags);
        if (fused_out) *fused_out = true;
        if (C <= 256) return launch_fwd_dx<1>(tx, tw, tdz, a, max_ctas, st, fz);
        if (C <= 512) return launch_fwd_dx<2>(tx, tw, tdz, a, max_ctas, st, fz);
        if (C <= 768) return launch_fwd_dx<3>(tx, tw, tdz, a, max_ctas, st, fz);
        return launch_fwd_dx<4>(tx, tw, tdz, a, max_ctas, st, fz);
    }
    if (splitk) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)((B + BM - 1) / BM) * kSkS);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = kSkSmemBytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        static const bool no_pdl = getenv("OSB_NO_PDL") != nullptr;
        a.den_from_prev = ((flags & OSB_FWD_DEN_FROM_PREV_KERNEL) && !no_pdl) ? 1 : 0;
        cfg.numAttrs = a.den_from_prev;                 // the cluster shape is a compile-time attribute of the kernel
        __nv_bfloat16* dzp = reinterpret_cast<__nv_bfloat16*>(dz);
        if (bias) {
            rc = set_smem(head_fwd_splitk_kernel<true>, kSkSmemBytes); if (rc) return rc;
            OSB_CUDA(cudaLaunchKernelEx(&cfg, head_fwd_splitk_kernel<true>, tx, tw, a, dzp, lddz));
        } else {
            rc = set_smem(head_fwd_splitk_kernel<false>, kSkSmemBytes); if (rc) return rc;
            OSB_CUDA(cudaLaunchKernelEx(&cfg, head_fwd_splitk_kernel<false>, tx, tw, a, dzp, lddz));
        }
        return OSB_OK;
    }
    // the persistent kernel is launched in plain stream order: den (if any) is complete when it starts
    if (C <= 128) return launch_fwd<128, 1, 6>(tx, tw, tdz, a, max_ctas, st);
    if (C <= 256) return launch_fwd<256, 1, 4>(tx, tw, tdz, a, max_ctas, st);
    if (C <= 512) return launch_fwd<256, 2, 4>(tx, tw, tdz, a, max_ctas, st);
    if (C <= 768) return launch_fwd<256, 3, 4>(tx, tw, tdz, a, max_ctas, st);
    return launch_fwd<256, 4, 4>(tx, tw, tdz, a, max_ctas, st);
}

extern "C" int osb_head_forward(const void* x, int64_t ldx, const void* w, const float* bias, const int64_t* y,
                                const float* class_w, int64_t B, int K, int C, int variant, double unk_weight,
                                int64_t ignore_index, const double* den, int valid_cols,
                                double* loss_num, float* loss,
                                float* row_ptgt, float* row_pmax, float* row_pmax_all, int32_t* row_pred,
                                void* dz, int64_t lddz, int dz_planes, int flags, int max_ctas,
                                void* ws, size_t ws_bytes, void* stream) {
    OSB_ENTRY_GUARD();
    return head_forward_impl(x, ldx, w, bias, y, class_w, B, K, C, variant, unk_weight, ignore_index, den, valid_cols, loss_num, loss,
                             row_ptgt, row_pmax, row_pmax_all, row_pred, dz, lddz, dz_planes, flags, max_ctas, ws, ws_bytes, stream,
                             nullptr, nullptr);
}

// forward + dX of a training step in one call; ONE kernel where the shape allows (osb_head_forward_dx_is_fused), else the
// forward kernel and the dX GEMM one after the other -- the same bits either way
extern "C" int osb_head_forward_dx(const void* x, int64_t ldx, const void* w, const float* bias, const int64_t* y,
                                   const float* class_w, int64_t B, int K, int C, int variant, double unk_weight,
                                   int64_t ignore_index, const double* den, int valid_cols,
                                   double* loss_num, float* loss, float* row_ptgt, float* row_pmax,
                                   void* dz, int64_t lddz, int dz_planes, const float* gscale,
                                   void* dx, int64_t lddx, int dx_dtype, int flags, int max_ctas,
                                   void* ws, size_t ws_bytes, void* stream) {
    OSB_ENTRY_GUARD();
    OSB_REQUIRE(dz && dx && den, OSB_ERR_ARG, "osb_head_forward_dx: dz, dx and den are required");
    OSB_REQUIRE(dx_dtype == OSB_BF16 || dx_dtype == OSB_F32, OSB_ERR_ARG, "osb_head_forward_dx: bad dx_dtype %d", dx_dtype);
    OSB_REQUIRE(lddx >= K && lddx % 8 == 0 && (uintptr_t)dx % 16 == 0, OSB_ERR_ARG, "osb_head_forward_dx: bad dx / lddx=%lld", (long long)lddx);
    FuseDxRequest rq{gscale, dx, lddx, dx_dtype};
    bool fused = false;
    int rc = head_forward_impl(x, ldx, w, bias, y, class_w, B, K, C, variant, unk_weight, ignore_index, den, valid_cols, loss_num, loss,
                               row_ptgt, row_pmax, nullptr, nullptr, dz, lddz, dz_planes, flags, max_ctas, ws, ws_bytes, stream, &rq, &fused);
    if (rc || fused) return rc;
    return osb_head_backward(dz, lddz, dz_planes, den, x, ldx, w, B, K, C, gscale, dx, lddx, dx_dtype, nullptr, nullptr, max_ctas, 0,
                             ws, ws_bytes, stream);
}
extern "C" int osb_head_forward_dx_is_fused(int64_t B, int K, int C) {
    if (osb::check_device() != OSB_OK) return 0;
    static const bool no_fuse = getenv("OSB_NO_FUSED_DX") != nullptr;
    return (!no_fuse && !fwd_use_splitk(B, K, C, 0) && C > BM && (B + kRowFlagTileRows - 1) / kRowFlagTileRows <= kMaxRowFlagTiles) ? 1 : 0;
}

// the dW GEMM (split-K partials into the workspace); returns the split count through *splits_out
// direct_dw != NULL (pair kernel, at most two splits): the scaled result goes straight into dW, no partials, no reduce kernel;
// *direct_out says whether that happened
static int run_dw_gemm(const void* dz, int64_t lddz, int planes, const void* x, int64_t ldx, int64_t B, int K, int C,
                       void* ws, cudaStream_t st, int* splits_out, bool row_flags, float* direct_dw = nullptr,
                       const float* gscale = nullptr, const double* den = nullptr, bool* direct_out = nullptr,
                       bool prezeroed = false) {
    CUtensorMap tdz, tx;
    OSB_TMAP(make_tmap_bf16_3d(&tdz, dz, (uint64_t)planes, (uint64_t)B, (uint64_t)C, (uint64_t)lddz, (uint64_t)B * (uint64_t)lddz, 64), "dz (MN-major)");
    OSB_TMAP(make_tmap_bf16(&tx, x, (uint64_t)B, (uint64_t)K, (uint64_t)ldx, 64), "x (MN-major)");
    const int splits = dw_splits((int)B, K, C);
    *splits_out = splits;
    float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kWsOffDw);
    cudaLaunchAttribute pdl_attr;
    int rc;
    if (dw_pairs(C)) {
        rc = set_smem(head_dw2_kernel, kDw2SmemBytes);
        if (rc) return rc;
        const int tiles = ((C + 255) / 256) * ((K + kDwBN - 1) / kDwBN);
        const int units = tiles * splits;
        static const bool no_direct = getenv("OSB_NO_DIRECT_DW") != nullptr;          // developer A/B switch
        const bool direct = direct_dw != nullptr && splits <= 2 && tiles <= kMaxDwFlagTiles && !no_direct;
        if (direct_out) *direct_out = direct;
        CUtensorMap tpart, tdw;
        OSB_TMAP(make_tmap_f32_3d(&tpart, partials, (uint64_t)splits, (uint64_t)C, (uint64_t)K, (uint64_t)K, 32), "dW partials (store)");
        tdw = tpart;
        if (direct) OSB_TMAP(make_tmap_f32_2d(&tdw, direct_dw, (uint64_t)C, (uint64_t)K, (uint64_t)K, 32), "dW (direct store / reduce)");
        unsigned int* flags = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(ws) + kWsOffDwFlags);
        cudaLaunchConfig_t cfg = pdl_config(dim3(units * 2), dim3(kThreads), kDw2SmemBytes, st, &pdl_attr);
        // the forward's units: ceil(C / 256) class slices x 2 CTAs publish one word per epilogue warp and row tile
        const unsigned int* rflags = (row_flags && row_flags_usable(B, K, C))
                                         ? reinterpret_cast<const unsigned int*>(reinterpret_cast<const char*>(ws) + kWsOffRowFlags) : nullptr;
        OSB_CUDA(cudaLaunchKernelEx(&cfg, head_dw2_kernel, tdz, tx, tpart, tdw, (int)B, K, C, planes, splits,
                                    direct ? ((prezeroed && splits == 2) ? 2 : 1) : 0, gscale, den,
                                    flags, rflags, ((C + 255) / 256) * 2, g_trace_dw ? g_trace_dw : g_trace));
    } else {
        rc = set_smem(head_dw_kernel, kDwSmemBytes);
        if (rc) return rc;
        const int units = ((C + BM - 1) / BM) * ((K + kDwBN - 1) / kDwBN) * splits;
        cudaLaunchConfig_t cfg = pdl_config(dim3(units), dim3(kThreads), kDwSmemBytes, st, &pdl_attr);
        OSB_CUDA(cudaLaunchKernelEx(&cfg, head_dw_kernel, tdz, tx, (int)B, K, C, planes, splits, partials));
    }
    return OSB_OK;
}

static int check_backward_args(const char* fn, const void* dz, int64_t lddz, int planes, const void* x, int64_t ldx,
                               int64_t B, int K, int C, const void* ws, size_t ws_bytes) {
    OSB_REQUIRE(dz && x && ws, OSB_ERR_ARG, "%s: null pointer", fn);
    int rc = check_head_shapes(fn, B, K, C, ldx);
    if (rc) return rc;
    OSB_REQUIRE(lddz >= C && lddz % 8 == 0, OSB_ERR_ARG, "%s: lddz=%lld must be >= C and a multiple of 8", fn, (long long)lddz);
    OSB_REQUIRE(planes == 1 || planes == 2, OSB_ERR_ARG, "%s: dz_planes=%d must be 1 or 2", fn, planes);
    OSB_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)dz % 16 == 0) && ((uintptr_t)ws % 16 == 0), OSB_ERR_ARG,
                "%s: x, dz and the workspace must be 16-byte aligned", fn);
    OSB_REQUIRE(ws_bytes >= osb_head_workspace_bytes(B, K, C), OSB_ERR_WORKSPACE, "%s: workspace too small", fn);
    return OSB_OK;
}

extern "C" int osb_head_backward(const void* dz, int64_t lddz, int dz_planes, const double* den, const void* x, int64_t ldx,
                                 const void* w, int64_t B, int K, int C, const float* gscale,
                                 void* dx, int64_t lddx, int dx_dtype, float* dw, float* db, int max_ctas, int flags,
                                 void* ws, size_t ws_bytes, void* stream) {
    OSB_ENTRY_GUARD();
    int rc = check_backward_args("osb_head_backward", dz, lddz, dz_planes, x, ldx, B, K, C, ws, ws_bytes);
    if (rc) return rc;
    OSB_REQUIRE(den, OSB_ERR_ARG, "osb_head_backward: null den");
    OSB_REQUIRE(!dx || w, OSB_ERR_ARG, "osb_head_backward: dx needs w");
    OSB_REQUIRE((flags & ~OSB_BWD_DZ_FROM_PREV_FORWARD) == 0, OSB_ERR_ARG, "osb_head_backward: bad flags %d", flags);
    OSB_REQUIRE(dx_dtype == OSB_BF16 || dx_dtype == OSB_F32, OSB_ERR_ARG, "osb_head_backward: bad dx_dtype %d", dx_dtype);
    OSB_REQUIRE(!dx || (lddx >= K && lddx % 8 == 0), OSB_ERR_ARG, "osb_head_backward: bad lddx=%lld", (long long)lddx);
    OSB_REQUIRE(((uintptr_t)w % 16 == 0) && ((uintptr_t)dx % 16 == 0) && ((uintptr_t)dw % 16 == 0) && max_ctas >= 0, OSB_ERR_ARG,
                "osb_head_backward: w, dx and dw must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    // dX and dW in one call, and a dW GEMM of two batch splits that writes dW itself: the dX kernel clears dW on its way in
    // and both splits add their tile at the destination without waiting for each other
    static const bool no_prezero = getenv("OSB_NO_DW_PREZERO") != nullptr;           // developer A/B switch
    const bool prezero = dx && dw && !no_prezero && osb_head_dw_is_direct(B, K, C) && dw_splits((int)B, K, C) == 2;
    float4* zero_dst = prezero ? reinterpret_cast<float4*>(dw) : nullptr;
    const unsigned long long zero_n4 = prezero ? (unsigned long long)C * (unsigned long long)K / 4ull : 0ull;
    if (dx) {
        CUtensorMap tdz, tw, tdx;
        OSB_TMAP(make_tmap_bf16_3d(&tdz, dz, (uint64_t)dz_planes, (uint64_t)B, (uint64_t)C, (uint64_t)lddz, (uint64_t)B * (uint64_t)lddz, BM), "dz (K-major)");
        OSB_TMAP(make_tmap_bf16(&tw, w, (uint64_t)C, (uint64_t)K, (uint64_t)K, 64), "w (MN-major)");
        const int total2 = (int)((B + 255) / 256) * ((K + kDxBN - 1) / kDxBN);
        int pairs = cta_limit(max_ctas) / 2;
        if (pairs > total2) pairs = total2;
        if (pairs < 1) pairs = 1;
        cudaLaunchAttribute pdl_attr;
        cudaLaunchConfig_t cfg = pdl_config(dim3(pairs * 2), dim3(kDx2Threads), kDx2SmemBytes, st, &pdl_attr);
        if (dx_dtype == OSB_F32) {
            OSB_TMAP(make_tmap_f32_2d(&tdx, dx, (uint64_t)B, (uint64_t)K, (uint64_t)lddx, 32), "dx (fp32)");
            rc = set_smem(head_dx2_kernel<true>, kDx2SmemBytes);
            if (rc) return rc;
            OSB_CUDA(cudaLaunchKernelEx(&cfg, head_dx2_kernel<true>, tdz, tw, tdx, (int)B, K, C, dz_planes, gscale, den, zero_dst, zero_n4, g_trace, g_dbg));
        } else {
            OSB_TMAP(make_tmap_bf16(&tdx, dx, (uint64_t)B, (uint64_t)K, (uint64_t)lddx, 32), "dx");
            rc = set_smem(head_dx2_kernel<false>, kDx2SmemBytes);
            if (rc) return rc;
            OSB_CUDA(cudaLaunchKernelEx(&cfg, head_dx2_kernel<false>, tdz, tw, tdx, (int)B, K, C, dz_planes, gscale, den, zero_dst, zero_n4, g_trace, g_dbg));
        }
    }
    if (dw) {
        int splits = 1;
        bool direct = false;
        // the dW GEMM is directly behind the forward only when this call launches no dX in front of it
        rc = run_dw_gemm(dz, lddz, dz_planes, x, ldx, B, K, C, ws, st, &splits, (flags & OSB_BWD_DZ_FROM_PREV_FORWARD) && !dx,
                         dw, gscale, den, &direct, prezero);
        if (rc) return rc;
      if (!direct) {
        const float* partials = reinterpret_cast<const float*>(reinterpret_cast<const char*>(ws) + kWsOffDw);
        const size_t n4 = (size_t)C * K / 4;
        int rgrid = (int)((n4 + 255) / 256);
        const int cap = sm_count() * 8;
        if (rgrid > cap) rgrid = cap;
        cudaLaunchAttribute pdl_attr;
        cudaLaunchConfig_t cfg = pdl_config(dim3(rgrid), dim3(256), 0, st, &pdl_attr);
        OSB_CUDA(cudaLaunchKernelEx(&cfg, dw_reduce_kernel, reinterpret_cast<const float4*>(partials), splits, n4, gscale, den,
                                    reinterpret_cast<float4*>(dw)));
      }
    }
    if (db) {
        db_kernel<<<(C + 63) / 64, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(dz), lddz, (int)B, C, dz_planes, gscale, den, db);
        OSB_LAUNCH_CHECK();
    }
    return OSB_OK;
}

// ---- multi-GPU: the dW GEMM alone (split-K partials stay in the workspace), for the peer-memory all-reduce ----
extern "C" size_t osb_head_dw_partials_offset(void) { return kWsOffDw; }
extern "C" int osb_head_dw_is_direct(int64_t B, int K, int C) {
    if (osb::check_device() != OSB_OK || !dw_pairs(C)) return 0;
    static const bool no_direct = getenv("OSB_NO_DIRECT_DW") != nullptr;
    const int tiles = ((C + 255) / 256) * ((K + kDwBN - 1) / kDwBN);
    return (dw_splits((int)B, K, C) <= 2 && tiles <= kMaxDwFlagTiles && !no_direct) ? 1 : 0;
}
extern "C" int osb_head_dw_splits(int64_t B, int K, int C) { return dw_splits((int)B, K, C); }
extern "C" int osb_head_backward_dw_partials(const void* dz, int64_t lddz, int dz_planes, const void* x, int64_t ldx,
                                             int64_t B, int K, int C, float* db_partial, int local_reduce, int flags,
                                             void* ws, size_t ws_bytes, void* stream) {
    OSB_ENTRY_GUARD();
    int rc = check_backward_args("osb_head_backward_dw_partials", dz, lddz, dz_planes, x, ldx, B, K, C, ws, ws_bytes);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    int splits = 1;
    OSB_REQUIRE((flags & ~OSB_BWD_DZ_FROM_PREV_FORWARD) == 0, OSB_ERR_ARG, "osb_head_backward_dw_partials: bad flags %d", flags);
    rc = run_dw_gemm(dz, lddz, dz_planes, x, ldx, B, K, C, ws, st, &splits, (flags & OSB_BWD_DZ_FROM_PREV_FORWARD) != 0);
    if (rc) return rc;
    if (local_reduce && splits > 1) {
        // sum the splits into split 0 in place (every thread reads all splits of its elements before it writes them): the
        // cross-GPU all-reduce then pulls ONE matrix per rank over NVLink instead of `splits`
        float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kWsOffDw);
        const size_t n4 = (size_t)C * K / 4;
        int rgrid = (int)((n4 + 255) / 256);
        const int cap = sm_count() * 8;
        if (rgrid > cap) rgrid = cap;
        cudaLaunchAttribute pdl_attr;
        cudaLaunchConfig_t cfg = pdl_config(dim3(rgrid), dim3(256), 0, st, &pdl_attr);
        OSB_CUDA(cudaLaunchKernelEx(&cfg, dw_reduce_kernel, reinterpret_cast<const float4*>(partials), splits, n4,
                                    (const float*)nullptr, (const double*)nullptr, reinterpret_cast<float4*>(partials)));
    }
    if (db_partial) {
        db_kernel<<<(C + 63) / 64, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(dz), lddz, (int)B, C, dz_planes, nullptr, nullptr, db_partial);
        OSB_LAUNCH_CHECK();
    }
    return OSB_OK;
}
