// Fused open-set head on tcgen05 tensor cores (sm_100a).
//
//   head_fwd_kernel : Z = X W^T (+b) accumulated in TMEM, and an epilogue that never writes Z: online
//                     log-softmax statistics, entropic / softmax / garbage loss, confidence statistics,
//                     arg-max, and the un-scaled grad-logits dZ (bf16, TMA store).  Heads wider than the
//                     512 TMEM columns (C <= 1024) run as a 2-CTA cluster that splits the classes and
//                     exchanges the per-row softmax statistics through distributed shared memory.
//   head_dx_kernel  : dX = g * dZ W          (persistent, double-buffered TMEM accumulators)
//   head_dw_kernel  : dW = g * dZ^T X        (split over the batch, deterministic partial reduce)
//
// All operands are 128B-swizzled TMA tiles; one tile layout serves as K-major or MN-major operand
// (tc05.cuh).  Warp roles per CTA (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM
// allocator, warps 2..5 = epilogue (TMEM lane quadrant = warp % 4, one thread per accumulator row).
//
// Reference arithmetic: model.py:23-26,37-38 (linear), losses.py:16-29 / train.py:343,347 (losses),
// train.py:177 + metrics.py:33-40 + util.py:109-111 (the per-row statistics), autograd of those.
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
constexpr int kMaxClusters = 148;
constexpr size_t kXchFlagBytes = (size_t)kMaxClusters * 2 * 4 * 4 * sizeof(unsigned int);
constexpr size_t kXchStatBytes = (size_t)kMaxClusters * 2 * 4 * 128 * 6 * sizeof(float);
// workspace: [ticket][loss partials][exchange flags][exchange statistics][dW partials]
constexpr size_t kWsOffPartials = kHeadTicketBytes;
constexpr size_t kWsOffFlags = kWsOffPartials + sizeof(double) * kMaxPartials;
constexpr size_t kWsOffStats = kWsOffFlags + kXchFlagBytes;
constexpr size_t kWsOffDw = (kWsOffStats + kXchStatBytes + 255) / 256 * 256;

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
    return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

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
__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
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

// =================================================================================================
// forward: logits in TMEM -> loss, statistics, dZ
// =================================================================================================
struct FwdArgs {
    const float* bias; const int64_t* y; const float* class_w;
    int B, K, C; int variant; float t_uniform; int64_t ignore_index; const double* den; int valid_cols;
    double* loss_num; float* loss;
    float* row_ptgt; float* row_pmax; float* row_pmax_all; int32_t* row_pred;
    int write_dz;
    double* partials; unsigned int* ticket;
    float* xch_stats;          // [clusters][2][CL][128][6]  per-row statistics exchanged between the CTAs of a cluster
    unsigned int* xch_flags;   // [clusters][2][CL][4]       release/acquire flags (tile counter), zero between launches
    long long* trace;      // optional [gridDim.x][64] clock stamps (developer tool, NULL in production)
};

constexpr int kStatFloats = 6;          // m, s, zy, sumz, mv, argmax  (per row, per column slice)

// One CTA owns 128 rows x BN classes of the logits (BN = 128 or 256 TMEM columns, double-buffered so the
// epilogue of tile t overlaps the MMAs of tile t+1).  Heads wider than BN run as a cluster of CL = ceil(C/BN)
// CTAs (CL <= 4) that share the rows, split the classes and exchange per-row softmax statistics through DSMEM.
// EW epilogue warps per TMEM lane quadrant each own a slice of the CTA's columns.
template <int BN, int CL, int STAGES, int EW>
struct FwdCfg {
    static constexpr int kThreads = (2 + 4 * EW) * 32;
    static constexpr int kEpiThreads = 4 * EW * 32;
    static constexpr int kStageBytes = kABytes + BN * BK * 2;
    static constexpr int kTmemCols = 2 * BN;
    static constexpr int kPeers = CL > 1 ? CL - 1 : 1;
    static constexpr int kOffLx = STAGES * kStageBytes;                            // [2][EW][128][6] floats
    static constexpr int kOffPx = kOffLx + 2 * EW * BM * kStatFloats * 4;          // [2][CL-1][128][6] floats
    static constexpr int kOffBias = kOffPx + 2 * kPeers * BM * kStatFloats * 4;    // [BN] floats
    static constexpr int kOffBars = kOffBias + BN * 4;
    static constexpr int kNumBars = 2 * STAGES + 4 + 2;
    static constexpr int kSmemBytes = kOffBars + kNumBars * 8 + 16 + 1024;  // + tmem slot + alignment slack
    static_assert(kSmemBytes <= 232448, "exceeds 227 KB of shared memory");
};

__device__ __forceinline__ float max32(const float (&z)[32]) {
    float t[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = fmaxf(z[i], z[i + 16]);
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = fmaxf(t[i], t[i + 8]);
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = fmaxf(t[i], t[i + 4]);
    return fmaxf(fmaxf(t[0], t[1]), fmaxf(t[2], t[3]));
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

template <int BN, int CL, int STAGES, int EW, bool HAS_BIAS, bool WANT_PRED>
__global__ void __launch_bounds__((2 + 4 * EW) * 32, 1)
head_fwd_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                __nv_bfloat16* __restrict__ dz, int64_t lddz, const FwdArgs a) {
    using Cfg = FwdCfg<BN, CL, STAGES, EW>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    float* sLx = reinterpret_cast<float*>(smem + Cfg::kOffLx);
    float* sPx = reinterpret_cast<float*>(smem + Cfg::kOffPx);
    float* sBias = reinterpret_cast<float*>(smem + Cfg::kOffBias);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBars);
    uint64_t* full = bars;                uint64_t* empty = full + STAGES;
    uint64_t* t_full = empty + STAGES;    uint64_t* t_empty = t_full + 2;
    uint64_t* x_bar = t_empty + 2;        // [2] statistics exchange with the peer CTAs (cluster only)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // CL consecutive CTAs form a group that shares the rows and splits the classes.  Not a hardware cluster: clusters of
    // 4 strand ~16 of the 148 SMs (GPC packing); the group only needs co-residency, which the cooperative launch gives.
    const uint32_t rank = (CL > 1) ? (blockIdx.x % CL) : 0u;
    const int cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;
    const int n_tiles = (a.B + BM - 1) / BM;
    const int n_kb = (a.K + BK - 1) / BK;
    const int n0 = (int)rank * BN;                        // class range of this CTA
    int ncols = a.C - n0;
    ncols = ncols < 0 ? 0 : (ncols > BN ? BN : ncols);
    const int ncols16 = (ncols + 15) & ~15;               // UMMA N granularity
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4 * EW); }
        mbar_init(&x_bar[0], 4 * Cfg::kPeers); mbar_init(&x_bar[1], 4 * Cfg::kPeers);
        fence_barrier_init();
        prefetch_tensormap(&tm_x); prefetch_tensormap(&tm_w);
    }
    if (HAS_BIAS) {
        for (int i = threadIdx.x; i < BN; i += Cfg::kThreads) sBias[i] = (i < ncols) ? a.bias[n0 + i] : 0.f;
    }
    if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = *tmem_slot;

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0 && ncols > 0) {
            uint32_t it = 0;
            for (int tile = cluster_id; tile < n_tiles; tile += n_clusters) {
                for (int kb = 0; kb < n_kb; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full[s], Cfg::kStageBytes);
                    uint8_t* st = smem + s * Cfg::kStageBytes;
                    tma_load_2d(st, &tm_x, &full[s], kb * BK, tile * BM);
                    tma_load_2d(st + kABytes, &tm_w, &full[s], kb * BK, n0);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------ MMA issuer --------------------------------
        if (lane == 0 && ncols > 0) {
            const uint32_t idesc = make_idesc_bf16(BM, ncols16, false, false);
            uint32_t it = 0, tc = 0;
            for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, ++tc) {
                const uint32_t buf = tc & 1, tph = (tc >> 1) & 1;
                mbar_wait(&t_empty[buf], tph ^ 1);
                tc_fence_after_sync();
                if (a.trace && tc < 4) a.trace[blockIdx.x * 64 + tc * 8 + 0] = gtime();
                for (int kb = 0; kb < n_kb; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after_sync();
                    if (a.trace && tc < 4 && kb == 0) a.trace[blockIdx.x * 64 + tc * 8 + 1] = gtime();
                    const uint32_t a_addr = smem_u32(smem + s * Cfg::kStageBytes);
                    const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_bf16(tbase + buf * BN, make_desc_kmajor(a_addr + k * 32), make_desc_kmajor(b_addr + k * 32),
                                  idesc, (kb | k) != 0);
                    umma_commit(&empty[s]);
                }
                umma_commit(&t_full[buf]);
                if (a.trace && tc < 4) a.trace[blockIdx.x * 64 + tc * 8 + 2] = gtime();
            }
        }
        __syncwarp();
    } else {
        // ------------------------------ epilogue ----------------------------------
        // warp (q, h): TMEM lane quadrant q = warp % 4 (hardware rule), column slice h; one thread per row.
        const int q = warp & 3, h = (warp - kEpiWarp0) >> 2;
        const int r_in_tile = q * 32 + lane;
        const float inv_den = (float)(1.0 / a.den[0]);
        const float s_uniform = a.t_uniform * (float)a.C;
        // this warp's slice of the CTA's columns, in whole 32-column groups
        const int n32 = (ncols + 31) >> 5, g_per = (n32 + EW - 1) / EW;
        const int c_lo = h * g_per * 32;
        int c_hi = (h + 1) * g_per * 32;
        c_hi = c_hi > ncols ? ncols : c_hi;
        const bool dz_vec = (lddz % 8 == 0);
        double loss_acc = 0.0;
        uint32_t it = 0;
        for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, ++it) {
            const uint32_t buf = it & 1, tph = (it >> 1) & 1;
            const int row = tile * BM + r_in_tile;
            const bool row_ok = row < a.B;
            const int64_t y = row_ok ? a.y[row] : -1;
            RowMode rm = classify_row(a.variant, y, a.C, a.ignore_index, a.class_w, s_uniform);
            if (!row_ok) { rm.mode = kRowIgnored; rm.omega = 0.f; }
            const int y_local = (y >= n0 + c_lo && y < n0 + c_hi) ? (int)(y - n0) : -1;   // label column, if in my slice
            if (ncols > 0) {
                mbar_wait(&t_full[buf], tph);
                tc_fence_after_sync();
            }
            const bool tr = a.trace && it < 4 && threadIdx.x == kEpiWarp0 * 32;
            if (tr) a.trace[blockIdx.x * 64 + it * 8 + 3] = gtime();
            const uint32_t t_row = tmem_addr(tbase, q * 32, buf * BN);

            // ---- sweep 0: row max (and arg-max / valid-column max) over my slice; no transcendental ----
            float m = -INFINITY, mv = -INFINITY;
            int am = 0x7fffffff;
            for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(t_row + c0, v);
                tmem_ld_wait();
                float z[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    z[i] = __uint_as_float(v[i]);
                    if (HAS_BIAS) z[i] += sBias[c0 + i];
                }
                if (c0 + 32 > ncols) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) if (c0 + i >= ncols) z[i] = -INFINITY;
                }
                const float gm = max32(z);
                if (WANT_PRED) {
                    if (gm > m) {                       // first arg-max (numpy tie rule): groups ascend, strict '>'
#pragma unroll
                        for (int i = 31; i >= 0; --i) if (z[i] == gm) am = n0 + c0 + i;
                    }
                }
                m = fmaxf(m, gm);
                const int gcol = n0 + c0;
                if (gcol + 32 <= a.valid_cols) mv = fmaxf(mv, gm);
                else if (gcol < a.valid_cols) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) if (gcol + i < a.valid_cols) mv = fmaxf(mv, z[i]);
                }
            }
            // ---- sweep 1: e = exp(z - m_slice) written back to TMEM; partial sums with 4-way ILP ----
            float s = 0.f, sumz = 0.f, zy = 0.f;
            {
                const float mb = m * kLog2e;
                float s4[4] = {0.f, 0.f, 0.f, 0.f}, z4[4] = {0.f, 0.f, 0.f, 0.f};
                for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(t_row + c0, v);
                    tmem_ld_wait();
                    const bool tail = c0 + 32 > ncols;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float z = __uint_as_float(v[i]);
                        if (HAS_BIAS) z += sBias[c0 + i];
                        if (tail && c0 + i >= ncols) z = -INFINITY;
                        if (i == y_local - c0) zy = z;
                        if (rm.mode == kRowUniform) z4[i & 3] += (tail && c0 + i >= ncols) ? 0.f : z;
                        const float e = fast_exp2(fmaf(z, kLog2e, -mb));
                        s4[i & 3] += e;
                        v[i] = __float_as_uint(e);
                    }
                    if (a.write_dz) tmem_st32(t_row + c0, v);
                }
                s = (s4[0] + s4[1]) + (s4[2] + s4[3]);
                sumz = (z4[0] + z4[1]) + (z4[2] + z4[3]);
                if (a.write_dz) tmem_st_wait();
            }
            if (tr) a.trace[blockIdx.x * 64 + it * 8 + 4] = gtime();
            const float m_slice = m;
            // ---- combine the column slices of this CTA (shared memory), then the CTAs of the cluster (DSMEM) ----
            {
                float* lx = sLx + (((it & 1) * EW + h) * BM + r_in_tile) * kStatFloats;
                lx[0] = m; lx[1] = s; lx[2] = zy; lx[3] = sumz; lx[4] = mv; lx[5] = __int_as_float(am);
                asm volatile("bar.sync 1, %0;" :: "n"(Cfg::kEpiThreads) : "memory");
                m = -INFINITY; mv = -INFINITY; am = 0x7fffffff;
                float pm[EW];
#pragma unroll
                for (int e = 0; e < EW; ++e) {
                    const float* o = sLx + (((it & 1) * EW + e) * BM + r_in_tile) * kStatFloats;
                    pm[e] = o[0];
                    const int oam = __float_as_int(o[5]);
                    if (pm[e] > m || (pm[e] == m && oam < am)) am = oam;
                    m = fmaxf(m, pm[e]);
                    mv = fmaxf(mv, o[4]);
                }
                s = 0.f; zy = 0.f; sumz = 0.f;
#pragma unroll
                for (int e = 0; e < EW; ++e) {
                    const float* o = sLx + (((it & 1) * EW + e) * BM + r_in_tile) * kStatFloats;
                    s += o[1] * fast_exp2((pm[e] - m) * kLog2e);
                    zy += o[2]; sumz += o[3];
                }
            }
            if (CL > 1) {
                // Through L2, not DSMEM: with the tensor pipe busy on the next tile, remote shared-memory stores
                // and arrives to a peer SM took ~8 us per tile (measured); a global round trip takes ~1.5 us.
                const uint32_t par = it & 1;
                float* xs = a.xch_stats + ((size_t)(cluster_id * 2 + par) * CL) * BM * kStatFloats;
                unsigned int* xf = a.xch_flags + (size_t)(cluster_id * 2 + par) * CL * 4;
                if (h == 0) {
                    float2* dst = reinterpret_cast<float2*>(xs + ((size_t)rank * BM + r_in_tile) * kStatFloats);
                    __stcg(dst + 0, make_float2(m, s)); __stcg(dst + 1, make_float2(zy, sumz));
                    __stcg(dst + 2, make_float2(mv, __int_as_float(am)));
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) st_release_gpu(xf + rank * 4 + q, it + 1);
                }
                if (lane < CL && lane != (int)rank) {
                    while (ld_acquire_gpu(xf + lane * 4 + q) < it + 1) { }
                }
                __syncwarp();
                float mg = m;
                float2 pv[Cfg::kPeers][3];
#pragma unroll
                for (int d = 1; d < CL; ++d) {
                    const uint32_t peer = (rank + d) % CL;
                    const float2* src = reinterpret_cast<const float2*>(xs + ((size_t)peer * BM + r_in_tile) * kStatFloats);
                    pv[d - 1][0] = __ldcg(src); pv[d - 1][1] = __ldcg(src + 1); pv[d - 1][2] = __ldcg(src + 2);
                    mg = fmaxf(mg, pv[d - 1][0].x);
                }
                s *= fast_exp2((m - mg) * kLog2e);
                // ties between CTAs resolve to the smaller column index, i.e. scan peers in rank order
#pragma unroll
                for (int d = 1; d < CL; ++d) {
                    const float pm = pv[d - 1][0].x;
                    const int pam = __float_as_int(pv[d - 1][2].y);
                    s += pv[d - 1][0].y * fast_exp2((pm - mg) * kLog2e);
                    if (pm > m || (pm == m && pam < am)) { am = pam; m = pm; }
                    zy += pv[d - 1][1].x; sumz += pv[d - 1][1].y; mv = fmaxf(mv, pv[d - 1][2].x);
                }
                m = mg;
            }
            if (tr) a.trace[blockIdx.x * 64 + it * 8 + 5] = gtime();
            const float lse = m + fast_log2(s) * kLn2;
            const bool is_col = (y >= 0 && y < a.C);
            if (rank == 0 && h == 0 && row_ok) {
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
            // ---- sweep 2: un-scaled grad-logits  dz = (S p - T) / den = coef * e - T/den  -> bf16 -> global ----
            if (a.write_dz) {
                const bool live = (rm.mode == kRowOneHot || rm.mode == kRowUniform);
                const float coef = live ? inv_den * rm.omega / s : 0.f;
                const float coef_e = coef * fast_exp2((m_slice - m) * kLog2e);   // e was taken relative to the slice max
                const float t_all = (rm.mode == kRowUniform) ? inv_den * a.t_uniform : 0.f;
                const float t_hot = (rm.mode == kRowOneHot) ? inv_den * rm.omega : 0.f;
                const float gy = fmaf(coef, fast_exp2((zy - m) * kLog2e), -t_all) - t_hot;   // label column
                __nv_bfloat16* drow = dz + (int64_t)row * lddz + n0;
                for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(t_row + c0, v);
                    tmem_ld_wait();
                    float g[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        g[i] = fmaf(coef_e, __uint_as_float(v[i]), -t_all);
                        if (i == y_local - c0) g[i] = gy;
                    }
                    if (row_ok) {
                        if (dz_vec && c0 + 32 <= ncols) {
#pragma unroll
                            for (int cch = 0; cch < 4; ++cch) store8_bf16(drow + c0 + cch * 8, g + cch * 8);
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) if (c0 + i < ncols) drow[c0 + i] = __float2bfloat16_rn(g[i]);
                        }
                    }
                }
            }
            if (tr) a.trace[blockIdx.x * 64 + it * 8 + 6] = gtime();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[buf]);
        }
        // ---- deterministic loss reduction: quadrants in order, CTAs in order ----
        __shared__ double s_loss[4];
        double wl = warp_sum(loss_acc);
        if (lane == 0 && h == 0) s_loss[q] = wl;
        asm volatile("bar.sync 1, %0;" :: "n"(Cfg::kEpiThreads) : "memory");
        if (warp == kEpiWarp0 && lane == 0) {
            a.partials[blockIdx.x] = ((s_loss[0] + s_loss[1]) + s_loss[2]) + s_loss[3];
            __threadfence();
            const unsigned t = atomicAdd(a.ticket, 1u);
            if (t == gridDim.x - 1) {
                __threadfence();
                double tot = 0.0;
                for (unsigned b = 0; b < gridDim.x; ++b) tot += __ldcg(&a.partials[b]);
                a.loss_num[0] = tot;
                a.loss[0] = (float)(tot / a.den[0]);
                *a.ticket = 0u;
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (CL > 1 && threadIdx.x < 8) a.xch_flags[((size_t)(cluster_id * 2 + (threadIdx.x >> 2)) * CL + rank) * 4 + (threadIdx.x & 3)] = 0u;
    if (warp == 1) tmem_dealloc(tbase, Cfg::kTmemCols);
}

// =================================================================================================
// dX = g * dZ W : A = dZ tile (K-major over classes), B = W tile viewed MN-major (N = feature index)
// =================================================================================================
constexpr int kDxBN = 256;
constexpr int kDxStages = 4;
constexpr int kDxStageBytes = kABytes + 4 * kBoxBytes;          // 48 KB
constexpr int kDxOffStage = kDxStages * kDxStageBytes;
constexpr int kDxOffBars = kDxOffStage + 4 * 2 * kStageWarpBytes;
constexpr int kDxSmemBytes = kDxOffBars + (2 * kDxStages + 4) * 8 + 16 + 1024;
static_assert(kDxSmemBytes <= 232448, "dx kernel shared memory");

__global__ void __launch_bounds__(kThreads, 1)
head_dx_kernel(const __grid_constant__ CUtensorMap tm_dz, const __grid_constant__ CUtensorMap tm_w,
               const __grid_constant__ CUtensorMap tm_dx, int B, int K, int C, const float* __restrict__ gscale) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    uint8_t* sStage = smem + kDxOffStage;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDxOffBars);
    uint64_t* full = bars; uint64_t* empty = full + kDxStages;
    uint64_t* t_full = empty + kDxStages; uint64_t* t_empty = t_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kDxStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
        fence_barrier_init();
        prefetch_tensormap(&tm_dz); prefetch_tensormap(&tm_w); prefetch_tensormap(&tm_dx);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = *tmem_slot;
    const int m_tiles = (B + BM - 1) / BM, n_tiles = (K + kDxBN - 1) / kDxBN;
    const int total = m_tiles * n_tiles;
    const int n_kb = (C + BK - 1) / BK;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const int mt = tile / n_tiles, nt = tile % n_tiles;
                for (int kb = 0; kb < n_kb; ++kb, ++it) {
                    const uint32_t s = it % kDxStages, ph = (it / kDxStages) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full[s], kDxStageBytes);
                    uint8_t* st = smem + s * kDxStageBytes;
                    tma_load_2d(st, &tm_dz, &full[s], kb * BK, mt * BM);                        // 64 classes x 128 rows
#pragma unroll
                    for (int j = 0; j < 4; ++j)                                                  // 64 features x 64 classes
                        tma_load_2d(st + kABytes + j * kBoxBytes, &tm_w, &full[s], nt * kDxBN + j * 64, kb * BK);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(BM, kDxBN, false, true);
            uint32_t it = 0, tc = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tc) {
                const uint32_t buf = tc & 1, tph = (tc >> 1) & 1;
                mbar_wait(&t_empty[buf], tph ^ 1);
                tc_fence_after_sync();
                for (int kb = 0; kb < n_kb; ++kb, ++it) {
                    const uint32_t s = it % kDxStages, ph = (it / kDxStages) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after_sync();
                    const uint32_t a_addr = smem_u32(smem + s * kDxStageBytes);
                    const uint32_t b_addr = a_addr + kABytes;
                    const int rem = C - kb * BK;
                    const int ksteps = rem >= BK ? BK / 16 : (rem + 15) / 16;
                    for (int k = 0; k < ksteps; ++k)
                        umma_bf16(tbase + buf * kDxBN, make_desc_kmajor(a_addr + k * 32),
                                  make_desc_mnmajor(b_addr + k * 2048, kBoxBytes), idesc, (kb | k) != 0);
                    umma_commit(&empty[s]);
                }
                umma_commit(&t_full[buf]);
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        uint8_t* my_stage = sStage + q * 2 * kStageWarpBytes;
        const float gs = gscale ? gscale[0] : 1.f;
        uint32_t tc = 0, n_stores = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tc) {
            const int mt = tile / n_tiles, nt = tile % n_tiles;
            const uint32_t buf = tc & 1, tph = (tc >> 1) & 1;
            mbar_wait(&t_full[buf], tph);
            tc_fence_after_sync();
            const uint32_t t_row = tmem_addr(tbase, q * 32, buf * kDxBN);
#pragma unroll 1
            for (int c0 = 0; c0 < kDxBN; c0 += 64) {
                uint8_t* sbuf = my_stage + (n_stores & 1) * kStageWarpBytes;
                if (lane == 0) tma_store_wait_read<1>();
                __syncwarp();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v[32];
                    tmem_ld32(t_row + c0 + h * 32, v);
                    tmem_ld_wait();
                    float g[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) g[i] = gs * __uint_as_float(v[i]);
#pragma unroll
                    for (int cch = 0; cch < 4; ++cch) stage_chunk(sbuf, lane, h * 4 + cch, g + cch * 8);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tm_dx, sbuf, nt * kDxBN + c0, mt * BM + q * 32);
                    tma_store_commit();
                }
                ++n_stores;
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[buf]);
        }
        if (lane == 0) tma_store_wait<0>();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tbase, 512);
}

// =================================================================================================
// dX with CTA pairs (cta_group::2): one cluster of 2 CTAs per 256 rows x 256 features, M = 256 MMAs.
// CTA r of the pair loads rows [128 r, +128) of the dZ tile and features [128 r, +128) of the W tile (32 KB per stage
// instead of 48 KB), the leader issues the MMAs, each CTA drains its own 128 x 256 accumulator from its TMEM.
// =================================================================================================
constexpr int kDx2Stages = 6;
constexpr int kDx2StageBytes = kABytes + 2 * kBoxBytes;         // 32 KB per CTA
constexpr int kDx2OffStage = kDx2Stages * kDx2StageBytes;
constexpr int kDx2OffBars = kDx2OffStage + 4 * 2 * kStageWarpBytes;
constexpr int kDx2SmemBytes = kDx2OffBars + (2 * kDx2Stages + 4) * 8 + 16 + 1024;
static_assert(kDx2SmemBytes <= 232448, "dx2 kernel shared memory");

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
head_dx2_kernel(const __grid_constant__ CUtensorMap tm_dz, const __grid_constant__ CUtensorMap tm_w,
                const __grid_constant__ CUtensorMap tm_dx, int B, int K, int C, const float* __restrict__ gscale) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    uint8_t* sStage = smem + kDx2OffStage;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDx2OffBars);
    uint64_t* full = bars; uint64_t* empty = full + kDx2Stages;
    uint64_t* t_full = empty + kDx2Stages; uint64_t* t_empty = t_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank_();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kDx2Stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 8); }     // 4 epilogue warps x 2 CTAs
        fence_barrier_init();
        prefetch_tensormap(&tm_dz); prefetch_tensormap(&tm_w); prefetch_tensormap(&tm_dx);
    }
    if (warp == 1) tmem_alloc2(tmem_slot, 512);
    tc_fence_before_sync();
    cluster_sync_();
    tc_fence_after_sync();
    const uint32_t tbase = *tmem_slot;
    const int m_tiles = (B + 255) / 256, n_tiles = (K + kDxBN - 1) / kDxBN;
    const int total = m_tiles * n_tiles;
    const int n_kb = (C + BK - 1) / BK;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = pair; tile < total; tile += n_pairs) {
                const int mt = tile / n_tiles, nt = tile % n_tiles;
                for (int kb = 0; kb < n_kb; ++kb, ++it) {
                    const uint32_t s = it % kDx2Stages, ph = (it / kDx2Stages) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * kDx2StageBytes);     // both CTAs' bytes land on the leader's barrier
                    uint8_t* st = smem + s * kDx2StageBytes;
                    tma_load_2d_2sm(st, &tm_dz, &full[s], kb * BK, mt * 256 + rank * 128);               // 64 classes x 128 rows
#pragma unroll
                    for (int j = 0; j < 2; ++j)                                                         // 64 features x 64 classes
                        tma_load_2d_2sm(st + kABytes + j * kBoxBytes, &tm_w, &full[s], nt * kDxBN + rank * 128 + j * 64, kb * BK);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = make_idesc_bf16(256, kDxBN, false, true);
            uint32_t it = 0, tc = 0;
            for (int tile = pair; tile < total; tile += n_pairs, ++tc) {
                const uint32_t buf = tc & 1, tph = (tc >> 1) & 1;
                mbar_wait(&t_empty[buf], tph ^ 1);
                tc_fence_after_sync();
                for (int kb = 0; kb < n_kb; ++kb, ++it) {
                    const uint32_t s = it % kDx2Stages, ph = (it / kDx2Stages) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after_sync();
                    const uint32_t a_addr = smem_u32(smem + s * kDx2StageBytes);
                    const uint32_t b_addr = a_addr + kABytes;
                    const int rem = C - kb * BK;
                    const int ksteps = rem >= BK ? BK / 16 : (rem + 15) / 16;
                    for (int k = 0; k < ksteps; ++k)
                        umma2_bf16(tbase + buf * kDxBN, make_desc_kmajor(a_addr + k * 32),
                                   make_desc_mnmajor(b_addr + k * 2048, kBoxBytes), idesc, (kb | k) != 0);
                    umma2_commit_mc(&empty[s]);
                }
                umma2_commit_mc(&t_full[buf]);
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        uint8_t* my_stage = sStage + q * 2 * kStageWarpBytes;
        const float gs = gscale ? gscale[0] : 1.f;
        uint32_t tc = 0, n_stores = 0;
        for (int tile = pair; tile < total; tile += n_pairs, ++tc) {
            const int mt = tile / n_tiles, nt = tile % n_tiles;
            const uint32_t buf = tc & 1, tph = (tc >> 1) & 1;
            mbar_wait(&t_full[buf], tph);
            tc_fence_after_sync();
            const uint32_t t_row = tmem_addr(tbase, q * 32, buf * kDxBN);
#pragma unroll 1
            for (int c0 = 0; c0 < kDxBN; c0 += 64) {
                uint8_t* sbuf = my_stage + (n_stores & 1) * kStageWarpBytes;
                if (lane == 0) tma_store_wait_read<1>();
                __syncwarp();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v[32];
                    tmem_ld32(t_row + c0 + h * 32, v);
                    tmem_ld_wait();
                    float g[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) g[i] = gs * __uint_as_float(v[i]);
#pragma unroll
                    for (int cch = 0; cch < 4; ++cch) stage_chunk(sbuf, lane, h * 4 + cch, g + cch * 8);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tm_dx, sbuf, nt * kDxBN + c0, mt * 256 + rank * 128 + q * 32);
                    tma_store_commit();
                }
                ++n_stores;
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&t_empty[buf]);
        }
        if (lane == 0) tma_store_wait<0>();
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
               int B, int K, int C, int splits, float* __restrict__ partials) {
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
    const uint32_t tbase = *tmem_slot;
    const int n_tiles = (K + kDwBN - 1) / kDwBN;
    const int unit = blockIdx.x;
    const int sp = unit % splits, tile = unit / splits;
    const int mt = tile / n_tiles, nt = tile % n_tiles;
    const int kb_total = (B + BK - 1) / BK;
    const int kb0 = (int)((int64_t)kb_total * sp / splits), kb1 = (int)((int64_t)kb_total * (sp + 1) / splits);

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const uint32_t s = it % kDwStages, ph = (it / kDwStages) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&full[s], kDwStageBytes);
                uint8_t* st = smem + s * kDwStageBytes;
#pragma unroll
                for (int j = 0; j < 2; ++j) tma_load_2d(st + j * kBoxBytes, &tm_dz, &full[s], mt * BM + j * 64, kb * BK);
#pragma unroll
                for (int j = 0; j < 4; ++j) tma_load_2d(st + (2 + j) * kBoxBytes, &tm_x, &full[s], nt * kDwBN + j * 64, kb * BK);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(BM, kDwBN, true, true);
            uint32_t it = 0;
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const uint32_t s = it % kDwStages, ph = (it / kDwStages) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after_sync();
                const uint32_t a_addr = smem_u32(smem + s * kDwStageBytes);
                const uint32_t b_addr = a_addr + 2 * kBoxBytes;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                    umma_bf16(tbase, make_desc_mnmajor(a_addr + k * 2048, kBoxBytes),
                              make_desc_mnmajor(b_addr + k * 2048, kBoxBytes), idesc, (it | k) != 0);
                umma_commit(&empty[s]);
            }
            umma_commit(t_full);
        }
        __syncwarp();
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

// dW = g * sum_s partials[s]  (fixed order => deterministic)
__global__ void __launch_bounds__(256)
dw_reduce_kernel(const float4* __restrict__ partials, int splits, size_t n4, const float* __restrict__ gscale,
                 float4* __restrict__ dw) {
    const float gs = gscale ? gscale[0] : 1.f;
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

// db = g * column sums of dZ (bf16 [B, lddz]); one block per 64 columns, fixed summation order
__global__ void __launch_bounds__(256)
db_kernel(const __nv_bfloat16* __restrict__ dz, int64_t lddz, int B, int C, const float* __restrict__ gscale,
          float* __restrict__ db) {
    __shared__ float sh[4][64];
    const int c = blockIdx.x * 64 + (threadIdx.x & 63);
    const int part = threadIdx.x >> 6;
    float acc = 0.f;
    if (c < C)
        for (int r = part; r < B; r += 4) acc += __bfloat162float(dz[(int64_t)r * lddz + c]);
    sh[part][threadIdx.x & 63] = acc;
    __syncthreads();
    if (threadIdx.x < 64 && c < C) {
        const float gs = gscale ? gscale[0] : 1.f;
        db[c] = gs * (((sh[0][threadIdx.x] + sh[1][threadIdx.x]) + sh[2][threadIdx.x]) + sh[3][threadIdx.x]);
    }
}

// ---- host --------------------------------------------------------------------------------------
static int dw_splits(int B, int K, int C) {
    const int mt = (C + BM - 1) / BM, nt = (K + kDwBN - 1) / kDwBN;
    const int kb_total = (B + BK - 1) / BK;
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    int s = sms / (mt * nt);
    if (s < 1) s = 1;
    if (s > 16) s = 16;
    if (s > kb_total) s = kb_total < 1 ? 1 : kb_total;
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

template <int BN, int CL, int STAGES, int EW, bool HAS_BIAS, bool WANT_PRED>
static int launch_fwd_k(const CUtensorMap& tx, const CUtensorMap& tw, __nv_bfloat16* dz, int64_t lddz, const FwdArgs& a,
                        cudaStream_t st) {
    using Cfg = FwdCfg<BN, CL, STAGES, EW>;
    const int n_tiles = (a.B + BM - 1) / BM;
    auto k = head_fwd_kernel<BN, CL, STAGES, EW, HAS_BIAS, WANT_PRED>;
    int rc = set_smem(k, Cfg::kSmemBytes); if (rc) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sm_count() / CL * CL);
    cfg.blockDim = dim3(Cfg::kThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = st;
    // persistent kernel, one CTA per SM, launched cooperatively when CTAs wait on each other (CL > 1): every CTA of a
    // group must be resident while its peers spin on its statistics flags
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = CL > 1 ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int groups = cta_limit() / CL;
    if (groups > n_tiles) groups = n_tiles;
    if (groups < 1) groups = 1;
    cfg.gridDim = dim3(groups * CL);
    OSB_CUDA(cudaLaunchKernelEx(&cfg, k, tx, tw, dz, lddz, a));
    return OSB_OK;
}

template <int BN, int CL, int STAGES, int EW>
static int launch_fwd(const CUtensorMap& tx, const CUtensorMap& tw, __nv_bfloat16* dz, int64_t lddz, const FwdArgs& a,
                      cudaStream_t st) {
    const bool pred = a.row_pred || a.row_pmax_all;
    if (a.bias) return pred ? launch_fwd_k<BN, CL, STAGES, EW, true, true>(tx, tw, dz, lddz, a, st)
                            : launch_fwd_k<BN, CL, STAGES, EW, true, false>(tx, tw, dz, lddz, a, st);
    return pred ? launch_fwd_k<BN, CL, STAGES, EW, false, true>(tx, tw, dz, lddz, a, st)
                : launch_fwd_k<BN, CL, STAGES, EW, false, false>(tx, tw, dz, lddz, a, st);
}

static int fwd_box_rows(int C) { return C <= 128 ? 128 : 256; }

}  // namespace head
}  // namespace osb

using namespace osb;
using namespace osb::head;

static int g_head_pairs = 1;          // 1: GEMMs run on CTA pairs (cta_group::2) where a pair kernel exists; 0: single-CTA kernels
// developer hook (not part of the public header): A/B switch between the single-CTA and the CTA-pair kernels
extern "C" void osb_debug_set_head_pairs(int on) { g_head_pairs = on; }
static thread_local bool g_dw_partials_only = false;   // set by osb_head_backward_dw_partials around its call
static long long* g_trace = nullptr;
// developer hook (not part of the public header): device buffer of [grid][64] int64 clock stamps, or NULL
extern "C" void osb_debug_set_head_trace(void* p) { g_trace = reinterpret_cast<long long*>(p); }

extern "C" size_t osb_head_workspace_bytes(int64_t B, int K, int C) {
    if (B < 0 || K <= 0 || C <= 0) return 0;
    const int s = dw_splits((int)B, K, C);
    return kWsOffDw + align_up((size_t)s * C * K * sizeof(float), 256) + 256;
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

extern "C" int osb_head_forward(const void* x, int64_t ldx, const void* w, const float* bias, const int64_t* y,
                                const float* class_w, int64_t B, int K, int C, int variant, double unk_weight,
                                int64_t ignore_index, const double* den, int valid_cols,
                                double* loss_num, float* loss,
                                float* row_ptgt, float* row_pmax, float* row_pmax_all, int32_t* row_pred,
                                void* dz, int64_t lddz, void* ws, size_t ws_bytes, void* stream) {
    OSB_ENTRY_GUARD();
    OSB_REQUIRE(x && w && y && den && loss_num && loss && ws, OSB_ERR_ARG, "osb_head_forward: null pointer");
    int rc = check_head_shapes("osb_head_forward", B, K, C, ldx);
    if (rc) return rc;
    OSB_REQUIRE(variant >= OSB_ENTROPIC && variant <= OSB_GARBAGE, OSB_ERR_ARG, "osb_head_forward: bad variant %d", variant);
    OSB_REQUIRE(variant != OSB_GARBAGE || class_w, OSB_ERR_ARG, "osb_head_forward: garbage variant needs class_w");
    OSB_REQUIRE(valid_cols >= 1 && valid_cols <= C, OSB_ERR_ARG, "osb_head_forward: valid_cols=%d outside [1,%d]", valid_cols, C);
    OSB_REQUIRE(!dz || (lddz >= C && lddz % 8 == 0), OSB_ERR_ARG, "osb_head_forward: lddz=%lld must be >= C and a multiple of 8", (long long)lddz);
    OSB_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)w % 16 == 0) && ((uintptr_t)dz % 16 == 0), OSB_ERR_ARG,
                "osb_head_forward: x, w and dz must be 16-byte aligned");
    OSB_REQUIRE(ws_bytes >= osb_head_workspace_bytes(B, K, C), OSB_ERR_WORKSPACE, "osb_head_forward: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap tx, tw;
    OSB_TMAP(make_tmap_bf16(&tx, x, (uint64_t)B, (uint64_t)K, (uint64_t)ldx, BM), "x");
    OSB_TMAP(make_tmap_bf16(&tw, w, (uint64_t)C, (uint64_t)K, (uint64_t)K, fwd_box_rows(C)), "w");
    __nv_bfloat16* dzp = reinterpret_cast<__nv_bfloat16*>(dz);
    FwdArgs a;
    a.bias = bias; a.y = y; a.class_w = class_w; a.B = (int)B; a.K = K; a.C = C; a.variant = variant;
    a.t_uniform = (float)(unk_weight / (double)C);
    a.ignore_index = ignore_index; a.den = den; a.valid_cols = valid_cols;
    a.loss_num = loss_num; a.loss = loss;
    a.row_ptgt = row_ptgt; a.row_pmax = row_pmax; a.row_pmax_all = row_pmax_all; a.row_pred = row_pred;
    a.write_dz = dz ? 1 : 0;
    a.ticket = reinterpret_cast<unsigned int*>(ws);
    a.partials = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + kWsOffPartials);
    a.xch_flags = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(ws) + kWsOffFlags);
    a.xch_stats = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kWsOffStats);
    a.trace = g_trace;
    if (C <= 128) return launch_fwd<128, 1, 6, 2>(tx, tw, dzp, lddz, a, st);
    if (C <= 256) return launch_fwd<256, 1, 4, 2>(tx, tw, dzp, lddz, a, st);
    if (C <= 512) return launch_fwd<256, 2, 4, 2>(tx, tw, dzp, lddz, a, st);
    if (C <= 768) return launch_fwd<256, 3, 4, 2>(tx, tw, dzp, lddz, a, st);
    return launch_fwd<256, 4, 4, 2>(tx, tw, dzp, lddz, a, st);
}

extern "C" int osb_head_backward(const void* dz, int64_t lddz, const void* x, int64_t ldx, const void* w,
                                 int64_t B, int K, int C, const float* gscale,
                                 void* dx, int64_t lddx, float* dw, float* db,
                                 void* ws, size_t ws_bytes, void* stream) {
    OSB_ENTRY_GUARD();
    OSB_REQUIRE(dz && x && w && ws, OSB_ERR_ARG, "osb_head_backward: null pointer");
    int rc = check_head_shapes("osb_head_backward", B, K, C, ldx);
    if (rc) return rc;
    OSB_REQUIRE(lddz >= C && lddz % 8 == 0, OSB_ERR_ARG, "osb_head_backward: lddz=%lld must be >= C and a multiple of 8", (long long)lddz);
    OSB_REQUIRE(!dx || (lddx >= K && lddx % 8 == 0), OSB_ERR_ARG, "osb_head_backward: bad lddx=%lld", (long long)lddx);
    OSB_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)w % 16 == 0) && ((uintptr_t)dz % 16 == 0) &&
                ((uintptr_t)dx % 16 == 0) && ((uintptr_t)dw % 16 == 0) && ((uintptr_t)ws % 16 == 0), OSB_ERR_ARG,
                "osb_head_backward: x, w, dz, dx and dw must be 16-byte aligned");
    OSB_REQUIRE(ws_bytes >= osb_head_workspace_bytes(B, K, C), OSB_ERR_WORKSPACE, "osb_head_backward: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (dx) {
        CUtensorMap tdz, tw, tdx;
        OSB_TMAP(make_tmap_bf16(&tdz, dz, (uint64_t)B, (uint64_t)C, (uint64_t)lddz, BM), "dz (K-major)");
        OSB_TMAP(make_tmap_bf16(&tw, w, (uint64_t)C, (uint64_t)K, (uint64_t)K, 64), "w (MN-major)");
        OSB_TMAP(make_tmap_bf16(&tdx, dx, (uint64_t)B, (uint64_t)K, (uint64_t)lddx, 32), "dx");
        if (g_head_pairs) {
            rc = set_smem(head_dx2_kernel, kDx2SmemBytes);
            if (rc) return rc;
            const int total2 = (int)((B + 255) / 256) * ((K + kDxBN - 1) / kDxBN);
            int pairs = cta_limit() / 2;
            if (pairs > total2) pairs = total2;
            head_dx2_kernel<<<pairs * 2, kThreads, kDx2SmemBytes, st>>>(tdz, tw, tdx, (int)B, K, C, gscale);
            OSB_LAUNCH_CHECK();
        } else {
            rc = set_smem(head_dx_kernel, kDxSmemBytes);
            if (rc) return rc;
            const int total = (int)((B + BM - 1) / BM) * ((K + kDxBN - 1) / kDxBN);
            int grid = cta_limit();
            if (grid > total) grid = total;
            head_dx_kernel<<<grid, kThreads, kDxSmemBytes, st>>>(tdz, tw, tdx, (int)B, K, C, gscale);
            OSB_LAUNCH_CHECK();
        }
    }
    if (dw || g_dw_partials_only) {
        CUtensorMap tdz, tx;
        OSB_TMAP(make_tmap_bf16(&tdz, dz, (uint64_t)B, (uint64_t)C, (uint64_t)lddz, 64), "dz (MN-major)");
        OSB_TMAP(make_tmap_bf16(&tx, x, (uint64_t)B, (uint64_t)K, (uint64_t)ldx, 64), "x (MN-major)");
        rc = set_smem(head_dw_kernel, kDwSmemBytes);
        if (rc) return rc;
        const int splits = dw_splits((int)B, K, C);
        const int units = ((C + BM - 1) / BM) * ((K + kDwBN - 1) / kDwBN) * splits;
        float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kWsOffDw);
        head_dw_kernel<<<units, kThreads, kDwSmemBytes, st>>>(tdz, tx, (int)B, K, C, splits, partials);
        OSB_LAUNCH_CHECK();
        if (dw) {
            const size_t n4 = (size_t)C * K / 4;
            int rgrid = (int)((n4 + 255) / 256);
            const int cap = sm_count() * 8;
            if (rgrid > cap) rgrid = cap;
            dw_reduce_kernel<<<rgrid, 256, 0, st>>>(reinterpret_cast<const float4*>(partials), splits, n4, gscale,
                                                    reinterpret_cast<float4*>(dw));
            OSB_LAUNCH_CHECK();
        }
    }
    if (db) {
        db_kernel<<<(C + 63) / 64, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(dz), lddz, (int)B, C, gscale, db);
        OSB_LAUNCH_CHECK();
    }
    return OSB_OK;
}

// ---- multi-GPU: the dW GEMM alone (split-K partials stay in the workspace), for the peer-memory all-reduce ----
extern "C" size_t osb_head_dw_partials_offset(void) { return kWsOffDw; }
extern "C" int osb_head_dw_splits(int64_t B, int K, int C) { return dw_splits((int)B, K, C); }
extern "C" int osb_head_backward_dw_partials(const void* dz, int64_t lddz, const void* x, int64_t ldx, int64_t B, int K, int C,
                                             void* ws, size_t ws_bytes, void* stream) {
    g_dw_partials_only = true;
    // w is not read by the dW GEMM; x doubles as a valid non-null placeholder for the argument checks
    const int rc = osb_head_backward(dz, lddz, x, ldx, x, B, K, C, nullptr, nullptr, K, nullptr, nullptr, ws, ws_bytes, stream);
    g_dw_partials_only = false;
    return rc;
}
