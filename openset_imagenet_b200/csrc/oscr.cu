// OSCR (util.py:90-122) as streaming + sort + scan kernels; bit-exact on the reference's integer counts.
//
//   stage A  rowreduce : scores [N,C] -> three key segments (known target score, correct-known target
//                        score, unknown row max) as order-preserving uint32 keys          (HBM-bound)
//   stage B  sort      : LSD radix sort, 8-bit digits; warp-level match ranking inside a 4096-key tile
//                        and a single-pass chained scan (decoupled look-back) across tiles
//   stage C  thresholds: unique known keys (np.unique, util.py:113) by a chained scan of head flags
//   stage D  counts    : #correct-known > tau and #unk > tau (strict, util.py:114,117) by upper-bound
//                        searches in the sorted segments; counts are additive across shards
//   lookup             : first argmin |fpr - q| (script/plot_all.py:346-350)
#include "osb_common.cuh"
#include "rowscan.cuh"

namespace osb {

// ---------------------------------------------------------------------------------------------
// stage A
// ---------------------------------------------------------------------------------------------
constexpr int kRowThreads = 256;

template <typename L> __device__ __forceinline__ int64_t label_as_int(L v);
template <> __device__ __forceinline__ int64_t label_as_int<int64_t>(int64_t v) { return v; }
template <> __device__ __forceinline__ int64_t label_as_int<int32_t>(int32_t v) { return v; }
template <> __device__ __forceinline__ int64_t label_as_int<float>(float v) { return (int64_t)v; }   // astype(int): truncation

template <int LPR, typename L>
__global__ void __launch_bounds__(kRowThreads)
oscr_rowreduce_kernel(const float* __restrict__ scores, int64_t ld, const L* __restrict__ gt, int64_t N, int C,
                      int64_t unk_label, bool vec4, uint32_t* __restrict__ keys, unsigned long long* __restrict__ totals) {
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LPR;
    const int64_t warps_total = (int64_t)gridDim.x * (kRowThreads / 32);
    const int64_t warp_global = (int64_t)blockIdx.x * (kRowThreads / 32) + warp;
    unsigned n_kn = 0, n_ok = 0, n_unk = 0, n_bad = 0;
    for (int64_t base = warp_global * RPW; base < N; base += warps_total * RPW) {
        const int64_t row = base + lane / LPR;
        const bool valid = row < N;
        const int64_t lab = valid ? label_as_int<L>(gt[row]) : -1;
        const bool bad = valid && lab >= C;                    // reference: IndexError at util.py:111
        const bool known = valid && lab >= 0 && !bad;          // util.py:101
        const bool unk = valid && lab == unk_label && !bad;    // util.py:102
        float m; int am;
        group_row_argmax<LPR>(scores + row * ld, C, known || unk, vec4, sub, m, am);   // util.py:109-110
        if (sub == 0 && valid) {
            uint32_t k_kn = kAbsentKey, k_ok = kAbsentKey, k_unk = kAbsentKey;
            if (known) {
                const float ts = scores[row * ld + lab];        // util.py:111
                k_kn = float_to_key(ts);
                if (am == (int)lab) k_ok = k_kn;                // pred_class == gt, util.py:114
            }
            if (unk) k_unk = float_to_key(m);
            keys[row] = k_kn;
            keys[N + row] = k_ok;
            keys[2 * N + row] = k_unk;
            n_kn += known; n_ok += (k_ok != kAbsentKey); n_unk += unk; n_bad += bad;
        }
    }
    __shared__ unsigned sh[4];
    if (threadIdx.x < 4) sh[threadIdx.x] = 0;
    __syncthreads();
    unsigned v[4] = {n_kn, n_ok, n_unk, n_bad};
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        unsigned x = v[s];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
        if (lane == 0 && x) atomicAdd(&sh[s], x);
    }
    __syncthreads();
    if (threadIdx.x < 4 && sh[threadIdx.x]) atomicAdd(&totals[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------
// stage B: radix sort
// ---------------------------------------------------------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortKpt = 16;
constexpr int kSortTile = kSortThreads * kSortKpt;     // 4096 keys
constexpr int kRadix = 256;
constexpr uint32_t kFlagLocal = 1u << 30, kFlagIncl = 2u << 30, kValMask = (1u << 30) - 1;

__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ int64_t seg_length(const int64_t* seg_len, int seg, int64_t n) {
    if (!seg_len) return n;
    int64_t l = seg_len[seg];
    return l < 0 ? 0 : (l > n ? n : l);
}

// All four digit histograms of every segment in one read of the keys.  Warp-aggregated shared-memory
// atomics (match.any) so that runs of equal keys (absent-key padding) do not serialise.
__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const uint32_t* __restrict__ keys, int64_t n, const int64_t* __restrict__ seg_len,
                  uint32_t* __restrict__ ghist /* [nseg][4][256] */) {
    const int seg = blockIdx.y;
    const int64_t len = seg_length(seg_len, seg, n);
    __shared__ uint32_t h[4][kRadix];
    for (int i = threadIdx.x; i < 4 * kRadix; i += kSortThreads) (&h[0][0])[i] = 0;
    __syncthreads();
    const uint32_t* k = keys + (int64_t)seg * n;
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * kSortThreads;
    // round the trip count up so that every lane of a warp executes the match together
    const int64_t trips = (len + stride - 1) / stride;
    int64_t i = (int64_t)blockIdx.x * kSortThreads + threadIdx.x;
    for (int64_t t = 0; t < trips; ++t, i += stride) {
        const bool valid = i < len;
        const uint32_t key = valid ? k[i] : 0u;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const uint32_t d = (key >> (8 * p)) & 0xffu;
            const unsigned peers = __match_any_sync(kFull, valid ? d : (0x100u + lane));
            if (valid && lane == __ffs(peers) - 1) atomicAdd(&h[p][d], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    uint32_t* g = ghist + (size_t)seg * 4 * kRadix;
    for (int j = threadIdx.x; j < 4 * kRadix; j += kSortThreads) {
        uint32_t c = (&h[0][0])[j];
        if (c) atomicAdd(&g[j], c);
    }
}

// One LSD pass.  Tile ids come from an atomic ticket so that every tile a block waits on is already
// resident or finished (forward progress of the look-back).
__global__ void __launch_bounds__(kSortThreads)
radix_pass_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int64_t n,
                  const int64_t* __restrict__ seg_len, int pass, int ntiles,
                  const uint32_t* __restrict__ ghist, uint32_t* __restrict__ status /* [nseg][ntiles][256] */,
                  uint32_t* __restrict__ tickets /* [nseg] */) {
    const int seg = blockIdx.y;
    const int64_t len = seg_length(seg_len, seg, n);
    __shared__ uint32_t s_keys[kSortTile];
    __shared__ uint32_t s_whist[kSortThreads / 32][kRadix];
    __shared__ uint32_t s_start[kRadix];      // first local slot of each digit
    __shared__ uint32_t s_goff[kRadix];       // global index of local slot 0 of each digit run, minus s_start
    __shared__ uint32_t s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(&tickets[seg], 1u);
    for (int i = tid; i < (kSortThreads / 32) * kRadix; i += kSortThreads) (&s_whist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t base = (int64_t)tile * kSortTile;
    if (base >= len) return;
    const int count = (int)((len - base) < kSortTile ? (len - base) : kSortTile);
    const uint32_t* src = in + (int64_t)seg * n + base;
    uint32_t* dst = out + (int64_t)seg * n;
    const int shift = 8 * pass;

    // ---- rank keys inside the warp's 512-key slice, in index order (stable) ----
    uint32_t key[kSortKpt];
    uint32_t rnk[kSortKpt];
    const int wbase = warp * (kSortKpt * 32);
#pragma unroll
    for (int r = 0; r < kSortKpt; ++r) {
        const int idx = wbase + r * 32 + lane;
        key[r] = idx < count ? src[idx] : 0xffffffffu;
    }
#pragma unroll
    for (int r = 0; r < kSortKpt; ++r) {
        const int idx = wbase + r * 32 + lane;
        const bool valid = idx < count;
        const uint32_t d = (key[r] >> shift) & 0xffu;
        const unsigned peers = __match_any_sync(kFull, valid ? d : (0x100u + lane));
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (valid && lane == leader) { old = s_whist[warp][d]; s_whist[warp][d] = old + __popc(peers); }
        old = __shfl_sync(kFull, old, leader);
        rnk[r] = old + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
    // ---- thread d owns digit d: exclusive scan over warps, tile count, look-back ----
    uint32_t tile_cnt = 0;
#pragma unroll
    for (int w = 0; w < kSortThreads / 32; ++w) {
        uint32_t c = s_whist[w][tid];
        s_whist[w][tid] = tile_cnt;
        tile_cnt += c;
    }
    uint32_t* my_status = status + ((size_t)seg * ntiles + tile) * kRadix + tid;
    if (tile == 0) st_relaxed(my_status, tile_cnt | kFlagIncl);
    else st_relaxed(my_status, tile_cnt | kFlagLocal);
    // block-wide exclusive scans over the 256 digits: local starts and global digit bases
    const uint32_t gcount = ghist[((size_t)seg * 4 + pass) * kRadix + tid];
    uint32_t inc_l = tile_cnt, inc_g = gcount;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t a = __shfl_up_sync(kFull, inc_l, o), b = __shfl_up_sync(kFull, inc_g, o);
        if (lane >= o) { inc_l += a; inc_g += b; }
    }
    __shared__ uint32_t s_wsum_l[kSortThreads / 32], s_wsum_g[kSortThreads / 32];
    if (lane == 31) { s_wsum_l[warp] = inc_l; s_wsum_g[warp] = inc_g; }
    __syncthreads();
    uint32_t off_l = 0, off_g = 0;
#pragma unroll
    for (int w = 0; w < kSortThreads / 32; ++w)
        if (w < warp) { off_l += s_wsum_l[w]; off_g += s_wsum_g[w]; }
    const uint32_t start_l = off_l + inc_l - tile_cnt;     // exclusive
    const uint32_t base_g = off_g + inc_g - gcount;        // exclusive
    uint32_t excl = 0;
    if (tile > 0) {
        for (int64_t t = (int64_t)tile - 1; t >= 0; --t) {
            const uint32_t* p = status + ((size_t)seg * ntiles + t) * kRadix + tid;
            uint32_t v;
            do { v = ld_relaxed(p); } while ((v & ~kValMask) == 0u);
            excl += v & kValMask;
            if (v & kFlagIncl) break;
        }
        st_relaxed(my_status, (excl + tile_cnt) | kFlagIncl);
    }
    s_start[tid] = start_l;
    s_goff[tid] = base_g + excl - start_l;
    __syncthreads();
    // ---- scatter to the tile-local sorted order, then stream out runs of equal digits ----
#pragma unroll
    for (int r = 0; r < kSortKpt; ++r) {
        const int idx = wbase + r * 32 + lane;
        if (idx < count) {
            const uint32_t d = (key[r] >> shift) & 0xffu;
            s_keys[s_start[d] + s_whist[warp][d] + rnk[r]] = key[r];
        }
    }
    __syncthreads();
    for (int i = tid; i < count; i += kSortThreads) {
        const uint32_t k = s_keys[i];
        const uint32_t d = (k >> shift) & 0xffu;
        dst[s_goff[d] + (uint32_t)i] = k;
    }
}

// ---------------------------------------------------------------------------------------------
// stage C/D: unique thresholds by a chained scan of head flags, counts by upper-bound search
// ---------------------------------------------------------------------------------------------
constexpr int kUniqThreads = 256;
constexpr int kUniqIpt = 16;
constexpr int kUniqTile = kUniqThreads * kUniqIpt;

__device__ __forceinline__ int64_t upper_bound_u32(const uint32_t* __restrict__ a, int64_t n, uint32_t key) {
    int64_t lo = 0, hi = n;            // first index with a[idx] > key
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) <= key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

struct CountArgs {          // optional fused counting (single-GPU path)
    const uint32_t* sorted_correct; const int64_t* n_correct;
    const uint32_t* sorted_unk;     const int64_t* n_unk;
    int64_t* ccr_cnt; int64_t* fpr_cnt; double* ccr; double* fpr;
};

__device__ __forceinline__ void emit_counts(const CountArgs& c, int64_t j, uint32_t key, int64_t n_known) {
    const int64_t nc = c.n_correct[0], nu = c.n_unk[0];
    const int64_t cc = nc - upper_bound_u32(c.sorted_correct, nc, key);
    const int64_t fc = nu - upper_bound_u32(c.sorted_unk, nu, key);
    c.ccr_cnt[j] = cc;
    c.fpr_cnt[j] = fc;
    if (c.ccr) c.ccr[j] = (double)cc / (double)n_known;     // util.py:114
    if (c.fpr) c.fpr[j] = (double)fc / (double)nu;          // util.py:117 (0/0 -> NaN like numpy)
}

__global__ void __launch_bounds__(kUniqThreads)
unique_kernel(const uint32_t* __restrict__ sorted, int64_t n, const int64_t* __restrict__ n_known_p,
              uint32_t* __restrict__ tau_keys, float* __restrict__ taus, int64_t* __restrict__ n_unique,
              uint32_t* __restrict__ status /* [ntiles] */, uint32_t* __restrict__ ticket, CountArgs cnt) {
    int64_t n_known = n_known_p[0];
    n_known = n_known < 0 ? 0 : (n_known > n ? n : n_known);
    __shared__ uint32_t s_tile, s_excl;
    __shared__ uint32_t s_wsum[kUniqThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t base = (int64_t)tile * kUniqTile;
    if (n_known == 0) { if (tile == 0 && tid == 0) n_unique[0] = 0; return; }
    if (base >= n_known) return;
    const int64_t first = base + (int64_t)tid * kUniqIpt;
    uint32_t k[kUniqIpt];
    uint32_t prev = (first > 0 && first - 1 < n_known) ? sorted[first - 1] : 0u;
    uint32_t heads = 0;     // bit i set when element first+i starts a run
#pragma unroll
    for (int i = 0; i < kUniqIpt; ++i) {
        const int64_t idx = first + i;
        const bool valid = idx < n_known;
        k[i] = valid ? sorted[idx] : 0u;
        if (valid && (idx == 0 || k[i] != prev)) heads |= 1u << i;
        prev = k[i];
    }
    const uint32_t mine = __popc(heads);
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t a = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += a; }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    uint32_t off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kUniqThreads / 32; ++w) { if (w < warp) off += s_wsum[w]; total += s_wsum[w]; }
    if (tid == 0) {
        uint32_t excl = 0;
        if (tile == 0) st_relaxed(status + tile, total | kFlagIncl);
        else {
            st_relaxed(status + tile, total | kFlagLocal);
            for (int64_t t = (int64_t)tile - 1; t >= 0; --t) {
                uint32_t v;
                do { v = ld_relaxed(status + t); } while ((v & ~kValMask) == 0u);
                excl += v & kValMask;
                if (v & kFlagIncl) break;
            }
            st_relaxed(status + tile, (excl + total) | kFlagIncl);
        }
        s_excl = excl;
        if (base + kUniqTile >= n_known) n_unique[0] = (int64_t)excl + total;   // tile holding the last known key
    }
    __syncthreads();
    int64_t j = (int64_t)s_excl + off + inc - mine;
#pragma unroll
    for (int i = 0; i < kUniqIpt; ++i) {
        if (heads & (1u << i)) {
            tau_keys[j] = k[i];
            if (taus) taus[j] = key_to_float(k[i]);
            if (cnt.sorted_correct) emit_counts(cnt, j, k[i], n_known);
            ++j;
        }
    }
}

__global__ void __launch_bounds__(256)
counts_kernel(const uint32_t* __restrict__ tau_keys, const int64_t* __restrict__ n_unique, CountArgs cnt) {
    const int64_t nu = n_unique[0];
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nu; j += (int64_t)gridDim.x * blockDim.x)
        emit_counts(cnt, j, tau_keys[j], 1);
}

__global__ void __launch_bounds__(256)
ratios_kernel(const int64_t* __restrict__ ccr_cnt, const int64_t* __restrict__ fpr_cnt,
              const int64_t* __restrict__ n_unique, const int64_t* __restrict__ n_known,
              const int64_t* __restrict__ n_unk, double* __restrict__ ccr, double* __restrict__ fpr) {
    const int64_t nu = n_unique[0];
    const double dk = (double)n_known[0], du = (double)n_unk[0];
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nu; j += (int64_t)gridDim.x * blockDim.x) {
        ccr[j] = (double)ccr_cnt[j] / dk;
        fpr[j] = (double)fpr_cnt[j] / du;
    }
}

// ---------------------------------------------------------------------------------------------
// CCR@FPR: first argmin |fpr - q| (numpy argmin: NaN counts as the minimum)
// ---------------------------------------------------------------------------------------------
constexpr int kMaxQueries = 16;
struct Queries { double q[kMaxQueries]; };

__device__ __forceinline__ bool better(double av, int64_t ai, double bv, int64_t bi) {
    const bool an = isnan(av), bn = isnan(bv);
    if (an != bn) return an;
    if (an || av == bv) return ai < bi;
    return av < bv;
}

__global__ void __launch_bounds__(1024)
ccr_at_fpr_kernel(const double* __restrict__ ccr, const double* __restrict__ fpr, const int64_t* __restrict__ n_unique,
                  Queries qs, int64_t* __restrict__ out_idx, double* __restrict__ out_fpr, double* __restrict__ out_ccr) {
    const int64_t nthr = n_unique[0] > 0 ? n_unique[0] - 1 : 0;      // the largest threshold is dropped, util.py:113
    const double q = qs.q[blockIdx.x];
    double bv = INFINITY; int64_t bi = INT64_MAX;
    for (int64_t j = threadIdx.x; j < nthr; j += blockDim.x) {
        const double d = fabs(fpr[j] - q);
        if (better(d, j, bv, bi)) { bv = d; bi = j; }
    }
    __shared__ double sv[32]; __shared__ int64_t si[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(kFull, bv, o); int64_t oi = __shfl_xor_sync(kFull, bi, o);
        if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
        bv = sv[lane]; bi = si[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(kFull, bv, o); int64_t oi = __shfl_xor_sync(kFull, bi, o);
            if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) {
            const bool any = nthr > 0 && bi != INT64_MAX;
            out_idx[blockIdx.x] = any ? bi : -1;
            out_fpr[blockIdx.x] = any ? fpr[bi] : __longlong_as_double(0x7ff8000000000000LL);
            out_ccr[blockIdx.x] = any ? ccr[bi] : __longlong_as_double(0x7ff8000000000000LL);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int ntiles_for(int64_t n) { return (int)((n + kSortTile - 1) / kSortTile) + 1; }

struct SortWs {
    uint32_t* ghist;      // [nseg][4][256]
    uint32_t* tickets;    // [4][nseg]
    uint32_t* status;     // [4][nseg][ntiles][256]
    size_t zero_bytes;    // everything above is one memset
};
static size_t sort_ws_bytes(int64_t n, int nseg) {
    size_t nt = (size_t)ntiles_for(n);
    return align_up(sizeof(uint32_t) * ((size_t)nseg * 4 * kRadix + 4 * (size_t)nseg + 4 * (size_t)nseg * nt * kRadix), 256);
}
static SortWs carve_sort_ws(void* ws, int64_t n, int nseg) {
    SortWs s;
    size_t nt = (size_t)ntiles_for(n);
    s.ghist = reinterpret_cast<uint32_t*>(ws);
    s.tickets = s.ghist + (size_t)nseg * 4 * kRadix;
    s.status = s.tickets + 4 * (size_t)nseg;
    s.zero_bytes = sizeof(uint32_t) * ((size_t)nseg * 4 * kRadix + 4 * (size_t)nseg + 4 * (size_t)nseg * nt * kRadix);
    return s;
}

static int run_sort(uint32_t* keys, uint32_t* alt, int64_t n, int nseg, const int64_t* seg_len, void* ws, cudaStream_t st) {
    if (n == 0) return OSB_OK;
    SortWs s = carve_sort_ws(ws, n, nseg);
    const int nt = ntiles_for(n);
    OSB_CUDA(cudaMemsetAsync(ws, 0, s.zero_bytes, st));
    int hgrid = (int)((n + (int64_t)kSortThreads * 16 - 1) / ((int64_t)kSortThreads * 16));
    hgrid = hgrid < 1 ? 1 : (hgrid > 2 * sm_count() ? 2 * sm_count() : hgrid);
    radix_hist_kernel<<<dim3(hgrid, nseg), kSortThreads, 0, st>>>(keys, n, seg_len, s.ghist);
    OSB_LAUNCH_CHECK();
    uint32_t* a = keys; uint32_t* b = alt;
    for (int pass = 0; pass < 4; ++pass) {
        radix_pass_kernel<<<dim3(nt, nseg), kSortThreads, 0, st>>>(a, b, n, seg_len, pass, nt, s.ghist,
                                                                    s.status + (size_t)pass * nseg * nt * kRadix,
                                                                    s.tickets + (size_t)pass * nseg);
        OSB_LAUNCH_CHECK();
        uint32_t* t = a; a = b; b = t;
    }
    return OSB_OK;      // four passes: result is back in `keys`
}

static size_t uniq_ws_bytes(int64_t n) {
    return align_up(sizeof(uint32_t) * (16 + (size_t)((n + kUniqTile - 1) / kUniqTile + 1)), 256);
}

static int run_unique(const uint32_t* sorted, int64_t n, const int64_t* n_known, uint32_t* tau_keys, float* taus,
                      int64_t* n_unique, void* ws, const CountArgs& cnt, cudaStream_t st) {
    const size_t bytes = uniq_ws_bytes(n);
    OSB_CUDA(cudaMemsetAsync(ws, 0, bytes, st));
    uint32_t* ticket = reinterpret_cast<uint32_t*>(ws);
    uint32_t* status = ticket + 16;
    int grid = (int)((n + kUniqTile - 1) / kUniqTile);
    if (grid < 1) grid = 1;
    unique_kernel<<<grid, kUniqThreads, 0, st>>>(sorted, n, n_known, tau_keys, taus, n_unique, status, ticket, cnt);
    OSB_LAUNCH_CHECK();
    return OSB_OK;
}

template <typename L>
static int launch_rowreduce(const float* scores, int64_t ld, const L* gt, int64_t N, int C, int64_t unk_label,
                            uint32_t* keys, int64_t* totals, cudaStream_t st) {
    const bool vec4 = rows_are_vec4(scores, ld);
    const int lpr = C <= 256 ? 8 : 32;
    const int rows_per_block = (32 / lpr) * (kRowThreads / 32);
    int64_t want = (N + rows_per_block - 1) / rows_per_block;
    const int cap = sm_count() * 8;
    const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    auto* tot = reinterpret_cast<unsigned long long*>(totals);
    if (lpr == 8) oscr_rowreduce_kernel<8, L><<<grid, kRowThreads, 0, st>>>(scores, ld, gt, N, C, unk_label, vec4, keys, tot);
    else oscr_rowreduce_kernel<32, L><<<grid, kRowThreads, 0, st>>>(scores, ld, gt, N, C, unk_label, vec4, keys, tot);
    OSB_LAUNCH_CHECK();
    return OSB_OK;
}

static int run_rowreduce(const float* scores, int64_t ld, const void* gt, int gt_type, int64_t N, int C,
                         int64_t unk_label, uint32_t* keys, int64_t* totals, cudaStream_t st) {
    OSB_CUDA(cudaMemsetAsync(totals, 0, 4 * sizeof(int64_t), st));
    if (N == 0) return OSB_OK;
    switch (gt_type) {
        case OSB_LABEL_I64: return launch_rowreduce(scores, ld, reinterpret_cast<const int64_t*>(gt), N, C, unk_label, keys, totals, st);
        case OSB_LABEL_F32: return launch_rowreduce(scores, ld, reinterpret_cast<const float*>(gt), N, C, unk_label, keys, totals, st);
        case OSB_LABEL_I32: return launch_rowreduce(scores, ld, reinterpret_cast<const int32_t*>(gt), N, C, unk_label, keys, totals, st);
    }
    set_error("osb_oscr: bad label type %d", gt_type);
    return OSB_ERR_ARG;
}

}  // namespace osb

using namespace osb;

// workspace of the composed entry: [keys 3N][alt 3N][tau_keys N][totals 4 x i64 (padded)][sort ws][uniq ws]
extern "C" size_t osb_oscr_workspace_bytes(int64_t N) {
    if (N < 0) return 0;
    size_t n = (size_t)N;
    return align_up(7 * n * sizeof(uint32_t), 256) + 256 + sort_ws_bytes(N, 3) + uniq_ws_bytes(N);
}

extern "C" int osb_oscr_rowreduce(const float* scores, int64_t ld, const void* gt, int gt_type, int64_t N, int C,
                                  int64_t unk_label, uint32_t* keys, int64_t* totals, void* stream) {
    OSB_ENTRY_GUARD();
    OSB_REQUIRE((scores && gt && keys) || N == 0, OSB_ERR_ARG, "osb_oscr_rowreduce: null pointer");
    OSB_REQUIRE(totals, OSB_ERR_ARG, "osb_oscr_rowreduce: null totals");
    OSB_REQUIRE(N >= 0 && N < (1LL << 30) && C > 0 && ld >= C, OSB_ERR_ARG, "osb_oscr_rowreduce: bad sizes N=%lld C=%d ld=%lld",
                (long long)N, C, (long long)ld);
    return run_rowreduce(scores, ld, gt, gt_type, N, C, unk_label, keys, totals, (cudaStream_t)stream);
}

extern "C" int osb_oscr_sort(uint32_t* keys, uint32_t* keys_alt, int64_t n, int nseg, const int64_t* seg_len,
                             void* ws, size_t ws_bytes, void* stream) {
    OSB_ENTRY_GUARD();
    OSB_REQUIRE((keys && keys_alt && ws) || n == 0, OSB_ERR_ARG, "osb_oscr_sort: null pointer");
    OSB_REQUIRE(n >= 0 && n < (1LL << 30) && nseg >= 1 && nseg <= 64, OSB_ERR_ARG, "osb_oscr_sort: bad sizes n=%lld nseg=%d", (long long)n, nseg);
    OSB_REQUIRE(ws_bytes >= sort_ws_bytes(n, nseg), OSB_ERR_WORKSPACE, "osb_oscr_sort: workspace too small (%zu < %zu)",
                ws_bytes, sort_ws_bytes(n, nseg));
    return run_sort(keys, keys_alt, n, nseg, seg_len, ws, (cudaStream_t)stream);
}

extern "C" size_t osb_oscr_sort_workspace_bytes(int64_t n, int nseg) { return n < 0 ? 0 : sort_ws_bytes(n, nseg); }
extern "C" size_t osb_oscr_thresholds_workspace_bytes(int64_t n) { return n < 0 ? 0 : uniq_ws_bytes(n); }

extern "C" int osb_oscr_thresholds(const uint32_t* sorted_known, int64_t n, const int64_t* n_known,
                                   uint32_t* tau_keys, float* taus, int64_t* n_unique,
                                   void* ws, size_t ws_bytes, void* stream) {
    OSB_ENTRY_GUARD();
    OSB_REQUIRE(n_known && n_unique && ws && ((sorted_known && tau_keys) || n == 0), OSB_ERR_ARG, "osb_oscr_thresholds: null pointer");
    OSB_REQUIRE(n >= 0 && n < (1LL << 30), OSB_ERR_ARG, "osb_oscr_thresholds: bad n=%lld", (long long)n);
    OSB_REQUIRE(ws_bytes >= uniq_ws_bytes(n), OSB_ERR_WORKSPACE, "osb_oscr_thresholds: workspace too small");
    CountArgs none = {};
    return run_unique(sorted_known, n, n_known, tau_keys, taus, n_unique, ws, none, (cudaStream_t)stream);
}

extern "C" int osb_oscr_counts(const uint32_t* tau_keys, const int64_t* n_unique,
                               const uint32_t* sorted_correct, const int64_t* n_correct,
                               const uint32_t* sorted_unk, const int64_t* n_unk, int64_t n,
                               int64_t* ccr_cnt, int64_t* fpr_cnt, void* stream) {
    OSB_ENTRY_GUARD();
    OSB_REQUIRE(n_unique && n_correct && n_unk, OSB_ERR_ARG, "osb_oscr_counts: null pointer");
    if (n == 0) return OSB_OK;
    OSB_REQUIRE(tau_keys && sorted_correct && sorted_unk && ccr_cnt && fpr_cnt, OSB_ERR_ARG, "osb_oscr_counts: null pointer");
    CountArgs c = {sorted_correct, n_correct, sorted_unk, n_unk, ccr_cnt, fpr_cnt, nullptr, nullptr};
    int grid = (int)((n + 255) / 256);
    const int cap = sm_count() * 8;
    grid = grid > cap ? cap : grid;
    counts_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(tau_keys, n_unique, c);
    OSB_LAUNCH_CHECK();
    return OSB_OK;
}

extern "C" int osb_oscr_ratios(const int64_t* ccr_cnt, const int64_t* fpr_cnt, const int64_t* n_unique,
                               const int64_t* n_known, const int64_t* n_unk, int64_t n,
                               double* ccr, double* fpr, void* stream) {
    OSB_ENTRY_GUARD();
    OSB_REQUIRE(n_unique && n_known && n_unk, OSB_ERR_ARG, "osb_oscr_ratios: null pointer");
    if (n == 0) return OSB_OK;
    OSB_REQUIRE(ccr_cnt && fpr_cnt && ccr && fpr, OSB_ERR_ARG, "osb_oscr_ratios: null pointer");
    int grid = (int)((n + 255) / 256);
    const int cap = sm_count() * 8;
    grid = grid > cap ? cap : grid;
    ratios_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(ccr_cnt, fpr_cnt, n_unique, n_known, n_unk, ccr, fpr);
    OSB_LAUNCH_CHECK();
    return OSB_OK;
}

extern "C" int osb_oscr(const float* scores, int64_t ld, const void* gt, int gt_type, int64_t N, int C,
                        int64_t unk_label, float* taus, int64_t* ccr_cnt, int64_t* fpr_cnt,
                        double* ccr, double* fpr, int64_t* meta, void* ws, size_t ws_bytes, void* stream) {
    OSB_ENTRY_GUARD();
    OSB_REQUIRE(meta && ws, OSB_ERR_ARG, "osb_oscr: null pointer");
    OSB_REQUIRE(N >= 0 && N < (1LL << 30) && C > 0 && ld >= C, OSB_ERR_ARG, "osb_oscr: bad sizes N=%lld C=%d ld=%lld",
                (long long)N, C, (long long)ld);
    OSB_REQUIRE(N == 0 || (scores && gt && taus && ccr_cnt && fpr_cnt && ccr && fpr), OSB_ERR_ARG, "osb_oscr: null pointer");
    OSB_REQUIRE(ws_bytes >= osb_oscr_workspace_bytes(N), OSB_ERR_WORKSPACE, "osb_oscr: workspace too small (%zu < %zu)",
                ws_bytes, osb_oscr_workspace_bytes(N));
    cudaStream_t st = (cudaStream_t)stream;
    char* p = reinterpret_cast<char*>(ws);
    uint32_t* keys = reinterpret_cast<uint32_t*>(p);
    uint32_t* alt = keys + 3 * (size_t)N;
    uint32_t* tau_keys = alt + 3 * (size_t)N;
    p += align_up(7 * (size_t)N * sizeof(uint32_t), 256);
    int64_t* totals = reinterpret_cast<int64_t*>(p);         // n_known, n_correct, n_unk, n_bad
    p += 256;
    void* sort_ws = p;
    p += sort_ws_bytes(N, 3);
    void* uniq_ws = p;
    int rc = run_rowreduce(scores, ld, gt, gt_type, N, C, unk_label, keys, totals, st);
    if (rc != OSB_OK) return rc;
    rc = run_sort(keys, alt, N, 3, nullptr, sort_ws, st);
    if (rc != OSB_OK) return rc;
    CountArgs c = {keys + (size_t)N, totals + 1, keys + 2 * (size_t)N, totals + 2, ccr_cnt, fpr_cnt, ccr, fpr};
    if (N == 0) { OSB_CUDA(cudaMemsetAsync(meta, 0, 5 * sizeof(int64_t), st)); return OSB_OK; }
    rc = run_unique(keys, N, totals, tau_keys, taus, meta, uniq_ws, c, st);
    if (rc != OSB_OK) return rc;
    OSB_CUDA(cudaMemcpyAsync(meta + 1, totals, 4 * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    return OSB_OK;
}

extern "C" int osb_ccr_at_fpr(const double* ccr, const double* fpr, const int64_t* n_unique, int64_t n,
                              const double* queries_host, int nq, int64_t* out_idx, double* out_fpr,
                              double* out_ccr, void* stream) {
    OSB_ENTRY_GUARD();
    (void)n;
    OSB_REQUIRE(n_unique && queries_host && out_idx && out_fpr && out_ccr, OSB_ERR_ARG, "osb_ccr_at_fpr: null pointer");
    OSB_REQUIRE(nq >= 1 && nq <= kMaxQueries, OSB_ERR_ARG, "osb_ccr_at_fpr: nq=%d outside [1,%d]", nq, kMaxQueries);
    Queries qs;
    for (int i = 0; i < nq; ++i) qs.q[i] = queries_host[i];
    ccr_at_fpr_kernel<<<nq, 1024, 0, (cudaStream_t)stream>>>(ccr, fpr, n_unique, qs, out_idx, out_fpr, out_ccr);
    OSB_LAUNCH_CHECK();
    return OSB_OK;
}
