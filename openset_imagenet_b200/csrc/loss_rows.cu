// Row-wise open-set losses on materialised logits, and the label bookkeeping shared with the fused
// head.  One warp owns one row; a row of up to 1024 columns stays in registers between the three
// sweeps (max/argmax, sum-exp, grad), wider rows are re-read through L1.
//
// Reference arithmetic: losses.py:16-29 (entropic), train.py:343 (softmax, ignore_index=-1),
// train.py:346-347 (garbage, class weights).  Closed forms in SURVEY.md §9.1-9.2.
#include "osb_common.cuh"
#include "row_modes.cuh"

namespace osb {

// ---------------------------------------------------------------------------------------------
// label statistics
// ---------------------------------------------------------------------------------------------
constexpr int kLabelThreads = 256;

__global__ void __launch_bounds__(kLabelThreads)
label_stats_kernel(const int64_t* __restrict__ y, const float* __restrict__ class_w, int64_t B, int C,
                   int variant, int64_t ignore_index, double* __restrict__ den_out,
                   int64_t* __restrict__ counts_out, double* partials, unsigned int* ticket) {
    // A forward kernel launched with programmatic stream serialization (OSB_FWD_DEN_FROM_PREV_KERNEL) may start
    // now; it waits for this grid (griddepcontrol.wait) before it reads the denominator.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    double den = 0.0;
    long long c_onehot = 0, c_uniform = 0, c_ignored = 0, c_bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
        RowMode rm = classify_row(variant, y[i], C, ignore_index, class_w, 1.0f);
        c_onehot += rm.mode == kRowOneHot;
        c_uniform += rm.mode == kRowUniform;
        c_ignored += rm.mode == kRowIgnored;
        c_bad += rm.mode == kRowBad;
        if (variant == OSB_ENTROPIC) den += 1.0;                                  // mean over ALL rows (losses.py:14,29)
        else if (rm.mode == kRowOneHot) den += (double)rm.omega;
    }
    // block reduce (fixed tree => deterministic)
    __shared__ double sh[5][kLabelThreads / 32];
    double v[5] = {den, (double)c_onehot, (double)c_uniform, (double)c_ignored, (double)c_bad};
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        v[s] = warp_sum(v[s]);
        if ((threadIdx.x & 31) == 0) sh[s][threadIdx.x >> 5] = v[s];
    }
    __syncthreads();
    double mine[5] = {0, 0, 0, 0, 0};
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < 5; ++s)
            for (int w = 0; w < kLabelThreads / 32; ++w) mine[s] += sh[s][w];
    }
    double tot[5];
    if (last_block_reduce<5>(mine, partials, ticket, tot)) {
        den_out[0] = tot[0];
        counts_out[0] = (int64_t)tot[1];
        counts_out[1] = (int64_t)tot[2];
        counts_out[2] = (int64_t)tot[3];
        counts_out[3] = (int64_t)tot[4];
    }
}

// ---------------------------------------------------------------------------------------------
// loss from logits
// ---------------------------------------------------------------------------------------------
constexpr int kLossWarps = 8;

struct LossArgs {
    const void* logits; int64_t ld; const int64_t* y; const float* class_w;
    int64_t B; int C; int variant; float t_uniform; int64_t ignore_index;
    const double* den; int valid_cols;
    double* loss_num; float* loss; float* lse;
    float* row_ptgt; float* row_pmax; float* row_pmax_all; int32_t* row_pred;
    void* dlogits; int64_t ldd; const float* gscale;
    double* partials; unsigned int* ticket;
};

// VPT > 0: the row lives in VPT registers per lane (columns lane + 32k).  VPT == 0: re-read.
template <typename T, int VPT>
__global__ void __launch_bounds__(kLossWarps * 32)
loss_rows_kernel(const LossArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = a.C;
    const T* __restrict__ logits = reinterpret_cast<const T*>(a.logits);
    T* __restrict__ dlogits = reinterpret_cast<T*>(a.dlogits);
    const float inv_den = (float)(1.0 / a.den[0]);
    const float gs = a.gscale ? a.gscale[0] : 1.0f;
    const int nk = (C + 31) / 32;

    double my_loss = 0.0;   // lane 0 of each warp accumulates its rows in visiting order
    for (int64_t row = (int64_t)blockIdx.x * kLossWarps + warp; row < a.B; row += (int64_t)gridDim.x * kLossWarps) {
        const T* zr = logits + row * a.ld;
        const int64_t y = a.y[row];
        const RowMode rm = classify_row(a.variant, y, C, a.ignore_index, a.class_w, a.t_uniform * (float)C);

        float v[VPT > 0 ? VPT : 1] = {};
        auto col = [&](int k) -> float {
            if constexpr (VPT > 0) return v[k];
            else { int c = lane + 32 * k; return c < C ? to_float<T>(zr[c]) : -INFINITY; }
        };
        if constexpr (VPT > 0) {
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                int c = lane + 32 * k;
                v[k] = c < C ? to_float<T>(zr[c]) : -INFINITY;
            }
        }
        // sweep 1: max / first argmax over all columns, max over the first valid_cols columns
        float m = -INFINITY, mv = -INFINITY; int am = 0x7fffffff;
        auto sweep1 = [&](int k) {
            int c = lane + 32 * k; float z = col(k);
            if (z > m) { m = z; am = c; }
            if (c < a.valid_cols) mv = fmaxf(mv, z);
        };
        if constexpr (VPT > 0) {
#pragma unroll
            for (int k = 0; k < VPT; ++k) sweep1(k);
        } else { for (int k = 0; k < nk; ++k) sweep1(k); }
        group_argmax<32>(m, am);
        mv = warp_max(mv);
        // sweep 2: sum exp, sum z, label column
        float s = 0.f, sumz = 0.f, zy = 0.f;
        const float mb = m * kLog2e;
        auto sweep2 = [&](int k) {
            int c = lane + 32 * k; float z = col(k);
            if (c < C) { s += fast_exp2(fmaf(z, kLog2e, -mb)); sumz += z; if (c == y) zy = z; }
        };
        if constexpr (VPT > 0) {
#pragma unroll
            for (int k = 0; k < VPT; ++k) sweep2(k);
        } else { for (int k = 0; k < nk; ++k) sweep2(k); }
        s = warp_sum(s); sumz = warp_sum(sumz); zy = warp_sum(zy);
        const float lse = m + fast_log2(s) * kLn2;

        float rowloss = 0.f;
        if (rm.mode == kRowOneHot) rowloss = rm.omega * (lse - zy);
        else if (rm.mode == kRowUniform) rowloss = a.t_uniform * ((float)C * lse - sumz);
        else if (rm.mode == kRowBad) rowloss = __int_as_float(0x7fc00000);
        if (lane == 0) {
            my_loss += (double)rowloss;
            if (a.lse) a.lse[row] = lse;
            const bool is_col = (y >= 0 && y < C);
            if (a.row_ptgt) a.row_ptgt[row] = is_col ? fast_exp2((zy - lse) * kLog2e) : 0.f;
            if (a.row_pmax) a.row_pmax[row] = fast_exp2((mv - lse) * kLog2e);
            if (a.row_pmax_all) a.row_pmax_all[row] = fast_exp2((m - lse) * kLog2e);
            if (a.row_pred) a.row_pred[row] = am;
        }
        // sweep 3: grad-logits  dz = gs/den * (S p - T)
        if (dlogits) {
            T* dr = dlogits + row * a.ldd;
            const bool live = (rm.mode == kRowOneHot || rm.mode == kRowUniform);
            const float coef = live ? gs * inv_den * rm.omega / s : 0.f;             // S = omega
            const float t_all = (rm.mode == kRowUniform) ? gs * inv_den * a.t_uniform : 0.f;
            const float t_hot = (rm.mode == kRowOneHot) ? gs * inv_den * rm.omega : 0.f;
            auto sweep3 = [&](int k) {
                int c = lane + 32 * k;
                if (c < C) {
                    float z = col(k);
                    float g = coef * fast_exp2(fmaf(z, kLog2e, -mb)) - t_all - ((c == y) ? t_hot : 0.f);
                    dr[c] = from_float<T>(live ? g : 0.f);
                }
            };
            if constexpr (VPT > 0) {
#pragma unroll
                for (int k = 0; k < VPT; ++k) sweep3(k);
            } else { for (int k = 0; k < nk; ++k) sweep3(k); }
        }
    }
    // deterministic reduction: warps in order inside the block, blocks in order at the end
    __shared__ double sh[kLossWarps];
    if (lane == 0) sh[warp] = my_loss;
    __syncthreads();
    double mine[1] = {0.0};
    if (threadIdx.x == 0)
        for (int w = 0; w < kLossWarps; ++w) mine[0] += sh[w];
    double tot[1];
    if (last_block_reduce<1>(mine, a.partials, a.ticket, tot)) {
        a.loss_num[0] = tot[0];
        a.loss[0] = (float)(tot[0] / a.den[0]);
    }
}

template <typename T>
static int launch_loss(const LossArgs& a, int grid, cudaStream_t st) {
    const int C = a.C;
    dim3 block(kLossWarps * 32);
    if (C <= 128) loss_rows_kernel<T, 4><<<grid, block, 0, st>>>(a);
    else if (C <= 256) loss_rows_kernel<T, 8><<<grid, block, 0, st>>>(a);
    else if (C <= 512) loss_rows_kernel<T, 16><<<grid, block, 0, st>>>(a);
    else if (C <= 1024) loss_rows_kernel<T, 32><<<grid, block, 0, st>>>(a);
    else loss_rows_kernel<T, 0><<<grid, block, 0, st>>>(a);
    OSB_LAUNCH_CHECK();
    return OSB_OK;
}

constexpr int kMaxReduceGrid = 1024;           // partial slots reserved in the workspace
constexpr size_t kTicketBytes = 128;

}  // namespace osb

using namespace osb;

extern "C" size_t osb_loss_workspace_bytes(int64_t B) {
    (void)B;
    return kTicketBytes + sizeof(double) * 5 * kMaxReduceGrid;
}

extern "C" int osb_label_stats(const int64_t* y, const float* class_w, int64_t B, int C, int variant,
                               int64_t ignore_index, double* den, int64_t* counts,
                               void* ws, size_t ws_bytes, int max_ctas, void* stream) {
    OSB_ENTRY_GUARD();
    OSB_REQUIRE(y && den && counts && ws, OSB_ERR_ARG, "osb_label_stats: null pointer");
    OSB_REQUIRE(B >= 0 && C > 0, OSB_ERR_ARG, "osb_label_stats: bad sizes B=%lld C=%d", (long long)B, C);
    OSB_REQUIRE(variant >= OSB_ENTROPIC && variant <= OSB_GARBAGE, OSB_ERR_ARG, "osb_label_stats: bad variant %d", variant);
    OSB_REQUIRE(variant != OSB_GARBAGE || class_w, OSB_ERR_ARG, "osb_label_stats: garbage variant needs class_w");
    OSB_REQUIRE(ws_bytes >= osb_loss_workspace_bytes(B), OSB_ERR_WORKSPACE, "osb_label_stats: workspace too small");
    int grid = (int)((B + kLabelThreads - 1) / kLabelThreads);
    grid = grid < 1 ? 1 : (grid > kMaxReduceGrid ? kMaxReduceGrid : grid);
    // a caller that runs the statistics beside a GEMM kernel keeps them on a few SMs (the sums are integers and the
    // row count: the result does not depend on the grid)
    if (grid > cta_limit(max_ctas)) grid = cta_limit(max_ctas);
    auto* ticket = reinterpret_cast<unsigned int*>(ws);
    auto* partials = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + kTicketBytes);
    label_stats_kernel<<<grid, kLabelThreads, 0, (cudaStream_t)stream>>>(y, class_w, B, C, variant, ignore_index,
                                                                      den, counts, partials, ticket);
    OSB_LAUNCH_CHECK();
    return OSB_OK;
}

extern "C" int osb_loss_from_logits(const void* logits, int dtype, int64_t ld, const int64_t* y,
                                    const float* class_w, int64_t B, int C, int variant, double unk_weight,
                                    int64_t ignore_index, const double* den, int valid_cols,
                                    double* loss_num, float* loss, float* lse,
                                    float* row_ptgt, float* row_pmax, float* row_pmax_all, int32_t* row_pred,
                                    void* dlogits, int64_t ldd, const float* gscale,
                                    void* ws, size_t ws_bytes, void* stream) {
    OSB_ENTRY_GUARD();
    OSB_REQUIRE(logits && y && den && loss_num && loss && ws, OSB_ERR_ARG, "osb_loss_from_logits: null pointer");
    OSB_REQUIRE(B >= 0 && C > 0 && ld >= C, OSB_ERR_ARG, "osb_loss_from_logits: bad sizes B=%lld C=%d ld=%lld",
                (long long)B, C, (long long)ld);
    OSB_REQUIRE(dtype == OSB_F32 || dtype == OSB_BF16, OSB_ERR_ARG, "osb_loss_from_logits: bad dtype %d", dtype);
    OSB_REQUIRE(variant >= OSB_ENTROPIC && variant <= OSB_GARBAGE, OSB_ERR_ARG, "osb_loss_from_logits: bad variant %d", variant);
    OSB_REQUIRE(variant != OSB_GARBAGE || class_w, OSB_ERR_ARG, "osb_loss_from_logits: garbage variant needs class_w");
    OSB_REQUIRE(!dlogits || ldd >= C, OSB_ERR_ARG, "osb_loss_from_logits: ldd < C");
    OSB_REQUIRE(valid_cols >= 0 && valid_cols <= C, OSB_ERR_ARG, "osb_loss_from_logits: valid_cols out of range");
    OSB_REQUIRE(ws_bytes >= osb_loss_workspace_bytes(B), OSB_ERR_WORKSPACE, "osb_loss_from_logits: workspace too small");
    LossArgs a;
    a.logits = logits; a.ld = ld; a.y = y; a.class_w = class_w; a.B = B; a.C = C; a.variant = variant;
    a.t_uniform = (float)(unk_weight / (double)C);     // float32(w/C): the reference's `ones` buffer, losses.py:12-13
    a.ignore_index = ignore_index; a.den = den; a.valid_cols = valid_cols;
    a.loss_num = loss_num; a.loss = loss; a.lse = lse;
    a.row_ptgt = row_ptgt; a.row_pmax = row_pmax; a.row_pmax_all = row_pmax_all; a.row_pred = row_pred;
    a.dlogits = dlogits; a.ldd = ldd; a.gscale = gscale;
    a.ticket = reinterpret_cast<unsigned int*>(ws);
    a.partials = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + kTicketBytes);
    int64_t want = (B + kLossWarps - 1) / kLossWarps;
    int cap = sm_count() * 8;
    if (cap > kMaxReduceGrid) cap = kMaxReduceGrid;
    int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    return dtype == OSB_F32 ? launch_loss<float>(a, grid, (cudaStream_t)stream)
                            : launch_loss<__nv_bfloat16>(a, grid, (cudaStream_t)stream);
}
