// metrics.confidence (metrics.py:8-42) as one streaming kernel.
// Contiguous score matrices stream through shared memory in whole-row chunks (cp.async.bulk + mbarrier, as OSCR
// stage A); other layouts use the direct-load kernel, where known rows need a single gathered element s[i, y_i]
// and only rows labelled `unknown_class` are read in full.
// Counts are exact integers; the two sums are accumulated in double in a fixed order.
#include "osb_common.cuh"
#include "rowscan.cuh"
#include "rowstream.cuh"

namespace osb {

constexpr int kConfThreads = 256;
constexpr int kConfMaxGrid = 1024;
constexpr size_t kConfTicketBytes = 128;

template <int LPR>
__global__ void __launch_bounds__(kConfThreads)
confidence_kernel(const float* __restrict__ scores, int64_t ld, const int64_t* __restrict__ y, int64_t N, int C,
                  float one_plus_offset, int64_t unknown_class, int valid_cols, bool vec4,
                  double* __restrict__ sums, int64_t* __restrict__ counts, double* partials, unsigned int* ticket) {
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LPR;
    const int64_t warps_total = (int64_t)gridDim.x * (kConfThreads / 32);
    const int64_t warp_global = (int64_t)blockIdx.x * (kConfThreads / 32) + warp;

    double kn_sum = 0.0, neg_sum = 0.0;
    long long kn_cnt = 0, neg_cnt = 0, bad_cnt = 0;
    for (int64_t base = warp_global * RPW; base < N; base += warps_total * RPW) {
        const int64_t row = base + lane / LPR;
        const bool valid = row < N;
        const int64_t lab = valid ? y[row] : 0;
        const bool unknown = valid && (lab == unknown_class);                 // metrics.py:25
        const bool known = valid && (lab >= 0) && !unknown;                   // metrics.py:26
        const bool bad = known && (lab >= C);
        float m; int am;
        group_row_argmax<LPR>(scores + row * ld, valid_cols, unknown, vec4, sub, m, am);
        if (sub == 0) {
            if (known && !bad) { kn_sum += (double)scores[row * ld + lab]; }  // metrics.py:33
            if (unknown) { neg_sum += (double)(one_plus_offset - m); }        // metrics.py:36-40 (float32 per sample)
            kn_cnt += known; neg_cnt += unknown; bad_cnt += bad;
        }
    }
    __shared__ double sh[5][kConfThreads / 32];
    double v[5] = {kn_sum, neg_sum, (double)kn_cnt, (double)neg_cnt, (double)bad_cnt};
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        v[s] = warp_sum(v[s]);
        if (lane == 0) sh[s][warp] = v[s];
    }
    __syncthreads();
    double mine[5] = {0, 0, 0, 0, 0};
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < 5; ++s)
            for (int w = 0; w < kConfThreads / 32; ++w) mine[s] += sh[s][w];
    }
    double tot[5];
    if (last_block_reduce<5>(mine, partials, ticket, tot)) {
        sums[0] = tot[0]; sums[1] = tot[1];
        counts[0] = (int64_t)tot[2]; counts[1] = (int64_t)tot[3]; counts[2] = (int64_t)tot[4];
    }
}


// ---- bulk-copy variant -----------------------------------------------------------------------
constexpr int kConfBulkThreads = 256;
constexpr int kConfBulkWarps = kConfBulkThreads / 32;
constexpr int kConfCtasPerSm = 2;
constexpr size_t kConfStageBudget = 104 * 1024;
constexpr int kConfMaxStages = 8;

template <int LPR, bool ALIGNED>
__global__ void __launch_bounds__(kConfBulkThreads, kConfCtasPerSm)
confidence_bulk_kernel(const float* __restrict__ scores, const int64_t* __restrict__ y, int64_t N, int C, int R, int S,
                       float one_plus_offset, int64_t unknown_class, int valid_cols,
                       double* __restrict__ sums, int64_t* __restrict__ counts, double* partials, unsigned int* ticket) {
    constexpr int RPW = 32 / LPR;
    extern __shared__ __align__(128) uint8_t conf_smem[];
    const uint32_t stage_bytes = (uint32_t)R * (uint32_t)C * 4u;
    const uint32_t lab_bytes = (uint32_t)R * 8u;
    uint8_t* stage = conf_smem;
    uint8_t* labs = conf_smem + (size_t)S * stage_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(conf_smem + (((size_t)S * (stage_bytes + lab_bytes) + 15) & ~(size_t)15));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, sub = lane % LPR;
    const int n_chunks = (int)((N + R - 1) / R);
    const int n_full = (int)(N / R);
    const int G = (int)gridDim.x;
    const int my_row = warp * RPW + lane / LPR;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) tc::mbar_init(&full[s], 1);
        tc::fence_barrier_init();
    }
    __syncthreads();
    auto issue = [&](int chunk, int s) {
        if (chunk >= n_full) return;
        const size_t row0 = (size_t)chunk * (size_t)R;
        tc::mbar_arrive_expect_tx(&full[s], stage_bytes + lab_bytes);
        bulk_g2s(stage + (size_t)s * stage_bytes, scores + row0 * (size_t)C, stage_bytes, &full[s]);
        bulk_g2s(labs + (size_t)s * lab_bytes, y + row0, lab_bytes, &full[s]);
    };
    if (tid == 0)
        for (int k = 0; k < S; ++k) issue((int)blockIdx.x + k * G, k);

    double kn_sum = 0.0, neg_sum = 0.0;
    long long kn_cnt = 0, neg_cnt = 0, bad_cnt = 0;
    int s = 0;
    uint32_t phase = 0;
    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += G) {
        const size_t row0 = (size_t)chunk * (size_t)R;
        const int rows = chunk < n_full ? R : (int)(N - (int64_t)row0);
        const float* sbase = reinterpret_cast<const float*>(stage + (size_t)s * stage_bytes);
        const int64_t* slab = reinterpret_cast<const int64_t*>(labs + (size_t)s * lab_bytes);
        if (chunk < n_full) {
            tc::mbar_wait(&full[s], phase);
        } else {
            float* wb = const_cast<float*>(sbase);
            for (int i = tid; i < rows * C; i += kConfBulkThreads) wb[i] = ld_stream_f1(scores + row0 * C + i);
            int64_t* wl = const_cast<int64_t*>(slab);
            for (int i = tid; i < rows; i += kConfBulkThreads) wl[i] = y[row0 + i];
            __syncthreads();
        }
        for (int r = my_row; r - lane / LPR < rows; r += kConfBulkWarps * RPW) {
            const bool valid = r < rows;
            const int64_t lab = valid ? slab[r] : 0;
            const bool unknown = valid && (lab == unknown_class);                 // metrics.py:25
            const bool known = valid && (lab >= 0) && !unknown;                   // metrics.py:26
            const bool bad = known && (lab >= C);
            const float* row = sbase + (size_t)r * C;
            // max over the first valid_cols columns (metrics.py:39): whole float4 groups before column valid_cols, then
            // the (<= 3) elements of its own group
            float m_lo, m_all; int head, jb;
            smem_row_split_max<LPR, ALIGNED>(row, C, valid_cols, unknown, sub, m_lo, m_all, head, jb);
            if (sub == 0) {
                if (unknown) {
                    float m = valid_cols >= C ? m_all : m_lo;
                    if (valid_cols < C) {
                        const int c0 = valid_cols < head ? valid_cols : head + 4 * jb;
#pragma unroll 1
                        for (int c = c0; c < valid_cols; ++c) m = fmaxf(m, row[c]);
                    }
                    neg_sum += (double)(one_plus_offset - m);                    // metrics.py:36-40 (float32 per sample)
                }
                if (known && !bad) kn_sum += (double)row[lab];                    // metrics.py:33
                kn_cnt += known; neg_cnt += unknown; bad_cnt += bad;
            }
        }
        __syncthreads();                       // every warp is done with stage s
        if (tid == 0) issue(chunk + S * G, s);
        if (++s == S) { s = 0; phase ^= 1u; }
    }
    __shared__ double sh[5][kConfBulkWarps];
    double v[5] = {kn_sum, neg_sum, (double)kn_cnt, (double)neg_cnt, (double)bad_cnt};
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        v[q] = warp_sum(v[q]);
        if (lane == 0) sh[q][warp] = v[q];
    }
    __syncthreads();
    double mine[5] = {0, 0, 0, 0, 0};
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < 5; ++q)
            for (int w = 0; w < kConfBulkWarps; ++w) mine[q] += sh[q][w];
    }
    double tot[5];
    if (last_block_reduce<5>(mine, partials, ticket, tot)) {
        sums[0] = tot[0]; sums[1] = tot[1];
        counts[0] = (int64_t)tot[2]; counts[1] = (int64_t)tot[3]; counts[2] = (int64_t)tot[4];
    }
}

template <int LPR, bool ALIGNED>
static int launch_conf_bulk(const StreamPlan& pl, const float* scores, const int64_t* y, int64_t N, int C,
                            float one_plus_offset, int64_t unknown_class, int valid_cols,
                            double* sums, int64_t* counts, double* partials, unsigned int* ticket, cudaStream_t st) {
    auto k = confidence_bulk_kernel<LPR, ALIGNED>;
    const size_t smem = (((size_t)pl.S * pl.stage_bytes + 15) & ~(size_t)15) + kConfMaxStages * sizeof(uint64_t);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        OSB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    const int64_t n_chunks = (N + pl.R - 1) / pl.R;
    int cap = sm_count() * kConfCtasPerSm;
    if (cap > kConfMaxGrid) cap = kConfMaxGrid;
    const int grid = (int)(n_chunks < 1 ? 1 : (n_chunks > cap ? cap : n_chunks));
    k<<<grid, kConfBulkThreads, smem, st>>>(scores, y, N, C, pl.R, pl.S, one_plus_offset, unknown_class, valid_cols,
                                           sums, counts, partials, ticket);
    OSB_LAUNCH_CHECK();
    return OSB_OK;
}


// ---- streaming variant: per-row statistics of the fused head instead of a score matrix ----------
// One block, fixed summation order; adds one validation batch to the running totals (train.py:172-192 without
// the all_scores tensor).
__global__ void __launch_bounds__(1024)
confidence_from_stats_kernel(const float* __restrict__ p_target, const float* __restrict__ p_max, const int64_t* __restrict__ y,
                             int64_t N, int C, float one_plus_offset, int64_t unknown_class, int accumulate,
                             double* __restrict__ sums, int64_t* __restrict__ counts) {
    double kn_sum = 0.0, neg_sum = 0.0;
    long long kn_cnt = 0, neg_cnt = 0, bad_cnt = 0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
        const int64_t lab = y[i];
        const bool unknown = lab == unknown_class;                           // metrics.py:25
        const bool known = lab >= 0 && !unknown;                             // metrics.py:26
        const bool bad = known && lab >= C;
        if (known && !bad) kn_sum += (double)p_target[i];                    // metrics.py:33
        if (unknown) neg_sum += (double)(one_plus_offset - p_max[i]);        // metrics.py:36-40
        kn_cnt += known; neg_cnt += unknown; bad_cnt += bad;
    }
    __shared__ double sh[5][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double v[5] = {kn_sum, neg_sum, (double)kn_cnt, (double)neg_cnt, (double)bad_cnt};
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        v[q] = warp_sum(v[q]);
        if (lane == 0) sh[q][warp] = v[q];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot[5] = {0, 0, 0, 0, 0};
        for (int q = 0; q < 5; ++q)
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot[q] += sh[q][w];
        sums[0] = (accumulate ? sums[0] : 0.0) + tot[0];
        sums[1] = (accumulate ? sums[1] : 0.0) + tot[1];
        counts[0] = (accumulate ? counts[0] : 0) + (int64_t)tot[2];
        counts[1] = (accumulate ? counts[1] : 0) + (int64_t)tot[3];
        counts[2] = (accumulate ? counts[2] : 0) + (int64_t)tot[4];
    }
}

}  // namespace osb

using namespace osb;

extern "C" size_t osb_confidence_workspace_bytes(int64_t N) {
    (void)N;
    return kConfTicketBytes + sizeof(double) * 5 * kConfMaxGrid;
}

extern "C" int osb_confidence(const float* scores, int64_t ld, const int64_t* y, int64_t N, int C,
                              float one_plus_offset, int64_t unknown_class, int valid_cols,
                              double* sums, int64_t* counts, void* ws, size_t ws_bytes, void* stream) {
    OSB_ENTRY_GUARD();
    OSB_REQUIRE(scores && y && sums && counts && ws, OSB_ERR_ARG, "osb_confidence: null pointer");
    OSB_REQUIRE(N >= 0 && C > 0 && ld >= C, OSB_ERR_ARG, "osb_confidence: bad sizes N=%lld C=%d ld=%lld",
                (long long)N, C, (long long)ld);
    OSB_REQUIRE(valid_cols >= 1 && valid_cols <= C, OSB_ERR_ARG, "osb_confidence: valid_cols=%d outside [1,%d]", valid_cols, C);
    OSB_REQUIRE(ws_bytes >= osb_confidence_workspace_bytes(N), OSB_ERR_WORKSPACE, "osb_confidence: workspace too small");
    auto* ticket = reinterpret_cast<unsigned int*>(ws);
    auto* partials = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + kConfTicketBytes);
    cudaStream_t st = (cudaStream_t)stream;
    if (ld == C && N >= 1 && N < (1LL << 30) && reinterpret_cast<uintptr_t>(scores) % 16 == 0 && reinterpret_cast<uintptr_t>(y) % 16 == 0) {
        const bool aligned4 = (C % 4 == 0);
        const StreamPlan pl = plan_row_stream(C, sizeof(int64_t), kConfBulkWarps, aligned4, kConfStageBudget, 32768, kConfMaxStages);
        if (pl.ok) {
#define OSB_CONF(LPR_, AL_) return launch_conf_bulk<LPR_, AL_>(pl, scores, y, N, C, one_plus_offset, unknown_class, valid_cols, sums, counts, partials, ticket, st)
            if (pl.lpr == 4) OSB_CONF(4, true);
            if (pl.lpr == 8) { if (aligned4) OSB_CONF(8, true); OSB_CONF(8, false); }
            if (aligned4) OSB_CONF(32, true);
            OSB_CONF(32, false);
#undef OSB_CONF
        }
    }
    const bool vec4 = rows_are_vec4(scores, ld);
    const int lpr = valid_cols <= 256 ? 8 : 32;
    const int rpw = 32 / lpr;
    int64_t want = (N + (int64_t)rpw * (kConfThreads / 32) - 1) / ((int64_t)rpw * (kConfThreads / 32));
    int cap = sm_count() * 8;
    if (cap > kConfMaxGrid) cap = kConfMaxGrid;
    int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    if (lpr == 8)
        confidence_kernel<8><<<grid, kConfThreads, 0, st>>>(scores, ld, y, N, C, one_plus_offset, unknown_class,
                                                            valid_cols, vec4, sums, counts, partials, ticket);
    else
        confidence_kernel<32><<<grid, kConfThreads, 0, st>>>(scores, ld, y, N, C, one_plus_offset, unknown_class,
                                                             valid_cols, vec4, sums, counts, partials, ticket);
    OSB_LAUNCH_CHECK();
    return OSB_OK;
}

extern "C" int osb_confidence_from_stats(const float* p_target, const float* p_max, const int64_t* y, int64_t N, int C,
                                         float one_plus_offset, int64_t unknown_class, int accumulate,
                                         double* sums, int64_t* counts, void* stream) {
    OSB_ENTRY_GUARD();
    OSB_REQUIRE(sums && counts && ((p_target && p_max && y) || N == 0), OSB_ERR_ARG, "osb_confidence_from_stats: null pointer");
    OSB_REQUIRE(N >= 0 && C > 0, OSB_ERR_ARG, "osb_confidence_from_stats: bad sizes N=%lld C=%d", (long long)N, C);
    confidence_from_stats_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(p_target, p_max, y, N, C, one_plus_offset, unknown_class,
                                                                      accumulate, sums, counts);
    OSB_LAUNCH_CHECK();
    return OSB_OK;
}
