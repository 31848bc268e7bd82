// metrics.confidence (metrics.py:8-42) as one streaming kernel.
// Known rows need a single gathered element s[i, y_i]; only rows labelled `unknown_class` are read in
// full (max over the first valid_cols columns), rows in neither group are not touched at all.
// Counts are exact integers; the two sums are accumulated in double in a fixed order.
#include "osb_common.cuh"
#include "rowscan.cuh"

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
    const bool vec4 = rows_are_vec4(scores, ld);
    const int lpr = valid_cols <= 256 ? 8 : 32;
    const int rpw = 32 / lpr;
    int64_t want = (N + (int64_t)rpw * (kConfThreads / 32) - 1) / ((int64_t)rpw * (kConfThreads / 32));
    int cap = sm_count() * 8;
    if (cap > kConfMaxGrid) cap = kConfMaxGrid;
    int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    cudaStream_t st = (cudaStream_t)stream;
    if (lpr == 8)
        confidence_kernel<8><<<grid, kConfThreads, 0, st>>>(scores, ld, y, N, C, one_plus_offset, unknown_class,
                                                            valid_cols, vec4, sums, counts, partials, ticket);
    else
        confidence_kernel<32><<<grid, kConfThreads, 0, st>>>(scores, ld, y, N, C, one_plus_offset, unknown_class,
                                                             valid_cols, vec4, sums, counts, partials, ticket);
    OSB_LAUNCH_CHECK();
    return OSB_OK;
}
