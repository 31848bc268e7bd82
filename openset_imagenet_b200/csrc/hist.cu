// Score / feature-norm histograms (util.get_histogram, util.py:202-228) on the GPU.
//
//   osb_hist_metrics : one pass over scores [N, C] (and, optionally, features [N, K]): per row the value the
//                      reference histograms -- scores[i, gt[i]] for a known row, max_c scores[i, :valid_cols] for a row
//                      labelled unk_label (metric 'score'), or ||features[i]|| for both groups (metric 'norm') -- plus
//                      the counts and the min / max of each group (np.histogram derives its bin range from them).
//   osb_histogram    : counts of values in the bins given by their edges, with np.histogram's rules: bin j holds
//                      edges[j] <= x < edges[j+1], the last bin also holds x == edges[-1], values outside the range
//                      and NaNs are dropped.  Integer counts, so the result is bit-exact for any summation order.
#include "osb_common.cuh"

namespace osb {
namespace hist {

constexpr int kMaxEdges = 2049;      // 2048 bins: edges (fp64) and block counters fit the 48 KB of static shared memory
constexpr int kThreads = 256;

// order-preserving image of a float in an unsigned integer (atomicMin / atomicMax work on it)
__device__ __forceinline__ unsigned int f2ord(float f) {
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// one warp per row
__global__ void __launch_bounds__(kThreads)
metrics_kernel(const float* __restrict__ scores, int64_t ld, const float* __restrict__ features, int64_t ldf, int K,
               const int64_t* __restrict__ gt, int64_t N, int C, int valid_cols, int64_t unk_label, int use_norm,
               float* __restrict__ kn_metric, float* __restrict__ unk_metric, int64_t* __restrict__ counts,
               unsigned int* __restrict__ minmax) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float kNaN = __int_as_float(0x7fc00000);
    unsigned int kmin = 0xffffffffu, kmax = 0u, umin = 0xffffffffu, umax = 0u;
    long long nk = 0, nu = 0;
    for (int64_t i = warp0; i < N; i += n_warps) {
        const int64_t y = gt[i];
        const bool kn = y >= 0, unk = (y == unk_label);
        float kv = kNaN, uv = kNaN;
        if (use_norm) {
            if (kn || unk) {
                // fp64 accumulation in a fixed order: lane-strided partial sums, then a fixed shuffle tree
                double acc = 0.0;
                const float* f = features + i * ldf;
                for (int c = lane; c < K; c += 32) { const double v = (double)f[c]; acc = fma(v, v, acc); }
                acc = warp_sum(acc);
                const float nrm = (float)sqrt(acc);
                if (kn) kv = nrm;
                if (unk) uv = nrm;
            }
        } else {
            const float* s = scores + i * ld;
            if (kn && y < C) kv = s[y];
            if (unk) {
                float m = -INFINITY;
                for (int c = lane; c < valid_cols; c += 32) m = fmaxf(m, s[c]);
                uv = warp_max(m);
            }
        }
        if (lane == 0) {
            kn_metric[i] = kv; unk_metric[i] = uv;
            if (kv == kv) { const unsigned int o = f2ord(kv); kmin = min(kmin, o); kmax = max(kmax, o); ++nk; }
            if (uv == uv) { const unsigned int o = f2ord(uv); umin = min(umin, o); umax = max(umax, o); ++nu; }
        }
    }
    if (lane == 0) {
        if (nk) { atomicMin(&minmax[0], kmin); atomicMax(&minmax[1], kmax); atomicAdd(reinterpret_cast<unsigned long long*>(&counts[0]), (unsigned long long)nk); }
        if (nu) { atomicMin(&minmax[2], umin); atomicMax(&minmax[3], umax); atomicAdd(reinterpret_cast<unsigned long long*>(&counts[1]), (unsigned long long)nu); }
    }
}

__global__ void init_kernel(int64_t* counts, unsigned int* minmax) {
    counts[0] = 0; counts[1] = 0;
    minmax[0] = 0xffffffffu; minmax[1] = 0u; minmax[2] = 0xffffffffu; minmax[3] = 0u;
}

__global__ void __launch_bounds__(kThreads)
histogram_kernel(const float* __restrict__ values, int64_t n, const double* __restrict__ edges, int n_edges,
                 unsigned long long* __restrict__ counts) {
    __shared__ double s_edges[kMaxEdges];
    __shared__ unsigned int s_cnt[kMaxEdges - 1];
    const int n_bins = n_edges - 1;
    for (int i = threadIdx.x; i < n_edges; i += blockDim.x) s_edges[i] = edges[i];
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x) s_cnt[i] = 0u;
    __syncthreads();
    const double lo = s_edges[0], hi = s_edges[n_bins];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = (double)values[i];
        if (!(x >= lo && x <= hi)) continue;              // NaN and out-of-range values are dropped (np.histogram)
        // last edge e with e <= x  (np.searchsorted(edges, x, 'right') - 1); x == edges[-1] goes to the last bin
        int a = 0, b = n_bins;                            // invariant: edges[a] <= x, and x < edges[b] or b == n_bins
        while (b - a > 1) {
            const int mid = (a + b) >> 1;
            if (s_edges[mid] <= x) a = mid; else b = mid;
        }
        atomicAdd(&s_cnt[a], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x)
        if (s_cnt[i]) atomicAdd(&counts[i], (unsigned long long)s_cnt[i]);
}

}  // namespace hist
}  // namespace osb

using namespace osb;

extern "C" int osb_hist_metrics(const float* scores, int64_t ld, const float* features, int64_t ldf, int K,
                                const int64_t* gt, int64_t N, int C, int valid_cols, int64_t unk_label,
                                float* kn_metric, float* unk_metric, int64_t* counts, float* minmax, void* stream) {
    OSB_ENTRY_GUARD();
    OSB_REQUIRE(gt && kn_metric && unk_metric && counts && minmax, OSB_ERR_ARG, "osb_hist_metrics: null pointer");
    OSB_REQUIRE((scores != nullptr) != (features != nullptr), OSB_ERR_ARG,
                "osb_hist_metrics: pass scores (metric 'score') or features (metric 'norm'), not both");
    OSB_REQUIRE(N >= 0, OSB_ERR_ARG, "osb_hist_metrics: bad N=%lld", (long long)N);
    if (scores) {
        OSB_REQUIRE(C >= 1 && ld >= C, OSB_ERR_ARG, "osb_hist_metrics: bad C=%d ld=%lld", C, (long long)ld);
        OSB_REQUIRE(valid_cols >= 1 && valid_cols <= C, OSB_ERR_ARG, "osb_hist_metrics: valid_cols=%d outside [1,%d]", valid_cols, C);
    } else {
        OSB_REQUIRE(K >= 1 && ldf >= K, OSB_ERR_ARG, "osb_hist_metrics: bad K=%d ldf=%lld", K, (long long)ldf);
    }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int* mm = reinterpret_cast<unsigned int*>(minmax);
    hist::init_kernel<<<1, 1, 0, st>>>(counts, mm);
    OSB_LAUNCH_CHECK();
    if (N > 0) {
        int64_t want = (N * 32 + hist::kThreads - 1) / hist::kThreads;
        const int cap = sm_count() * 8;
        const int grid = (int)(want > cap ? cap : want);
        hist::metrics_kernel<<<grid, hist::kThreads, 0, st>>>(scores, ld, features, ldf, K, gt, N, C, valid_cols, unk_label,
                                                             features ? 1 : 0, kn_metric, unk_metric, counts, mm);
        OSB_LAUNCH_CHECK();
    }
    return OSB_OK;
}

extern "C" int osb_histogram(const float* values, int64_t n, const double* edges, int n_edges, int64_t* counts, void* stream) {
    OSB_ENTRY_GUARD();
    OSB_REQUIRE(edges && counts && (values || n == 0), OSB_ERR_ARG, "osb_histogram: null pointer");
    OSB_REQUIRE(n >= 0, OSB_ERR_ARG, "osb_histogram: bad n=%lld", (long long)n);
    OSB_REQUIRE(n_edges >= 2 && n_edges <= hist::kMaxEdges, OSB_ERR_SHAPE, "osb_histogram: n_edges=%d outside [2,%d]", n_edges, hist::kMaxEdges);
    cudaStream_t st = (cudaStream_t)stream;
    OSB_CUDA(cudaMemsetAsync(counts, 0, sizeof(int64_t) * (size_t)(n_edges - 1), st));
    if (n > 0) {
        int64_t want = (n + hist::kThreads * 8 - 1) / (hist::kThreads * 8);
        const int cap = sm_count() * 4;
        const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
        hist::histogram_kernel<<<grid, hist::kThreads, 0, st>>>(values, n, edges, n_edges, reinterpret_cast<unsigned long long*>(counts));
        OSB_LAUNCH_CHECK();
    }
    return OSB_OK;
}
