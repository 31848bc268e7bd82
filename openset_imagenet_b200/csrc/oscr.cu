// OSCR (util.py:90-122) as two kernels; bit-exact on the reference's integer counts.
//
//   stage A  oscr_records_kernel : the one pass over the N x C float32 scores (HBM-bound).  Persistent CTAs
//            stream chunks of whole rows into shared memory with 1-D bulk copies (cp.async.bulk + mbarrier,
//            S stages in flight), reduce every row to first arg-max / row max / target score
//            (util.py:109-111) and append one 8-byte RECORD per known row and per `unk_label` row to two
//            compacted arrays (block-aggregated slot allocation, no per-row global atomics):
//                record = key << 32 | type,   key  = order-preserving uint32 image of the float32 score,
//                                             type = 0 known & wrong, 1 known & correct, 2 unk, 3 skip (padding)
//   stage B  oscr_curve_kernel  : ONE cooperative kernel (all CTAs co-resident, grid barriers in L2) that
//            LSD-radix-sorts the records by key (4 passes of 8 bits, stable: warp match.any ranking inside a
//            tile, a [digit][tile] count matrix for the cross-tile offsets) and then scans the sorted records:
//            a run of equal keys that contains a known record is a threshold tau (np.unique, util.py:113);
//            ccr_cnt[tau] = #correct-known records with key > tau, fpr_cnt[tau] = #unk records with key > tau
//            (strict '>', util.py:114,117) come from one prefix sum.  The logical sort input is
//            rec_a ++ rec_b with the unk records FIRST: the sort is stable, so inside a run of equal keys every
//            unk record precedes every known record and "the run contains a known record" is simply "its last
//            record is known".
//   lookup   ccr_at_fpr_kernel  : first argmin |fpr - q| (script/plot_all.py:346-350)
//
// Counts at a fixed threshold are additive across row shards, which is what the multi-GPU path uses
// (dist.py): every rank sorts its own unk records plus ALL ranks' known records (remote ones demoted to type 0).
#include "osb_common.cuh"
#include "rowscan.cuh"
#include "tc05.cuh"
#include "rowstream.cuh"

namespace osb {

using tc::smem_u32;

constexpr uint64_t kTypeWrong = 0, kTypeCorrect = 1, kTypeUnk = 2;   // 3 = skip (padding record)
__device__ __forceinline__ uint64_t make_rec(uint32_t key, uint64_t type) { return ((uint64_t)key << 32) | type; }

template <typename L> __device__ __forceinline__ int64_t label_as_int(L v);
template <> __device__ __forceinline__ int64_t label_as_int<int64_t>(int64_t v) { return v; }
template <> __device__ __forceinline__ int64_t label_as_int<int32_t>(int32_t v) { return v; }
template <> __device__ __forceinline__ int64_t label_as_int<float>(float v) { return (int64_t)v; }   // astype(int): truncation

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---------------------------------------------------------------------------------------------
// stage A
// ---------------------------------------------------------------------------------------------
constexpr int kRecThreads = 256;          // two CTAs per SM: one computes while the other waits for its next chunk
constexpr int kRecWarps = kRecThreads / 32;
constexpr int kStageCap = 1024;          // staged records per type per CTA
constexpr int kRecCtasPerSm = 2;
constexpr size_t kRecStageBudget = 92 * 1024;   // bulk-copy stages per CTA (2 x (92 KB + 16.5 KB staging) fits one SM)
constexpr int kMaxStages = 8;

// Optional by-product of stage A: the two sums of metrics.confidence (metrics.py:33,36-40) over the same rows, for callers
// that want the confidence of the evaluation set together with its OSCR curve (script/plot_all.py:363-372 computes both
// from the same score matrix).  sums_out == nullptr switches it off.
struct ConfAcc {
    double* sums_out;            // [2]: sum over known rows of s[i, y_i];  sum over unk rows of float32(1 + offset) - max s[i, :valid_cols]
    double* partials;            // [2 * grid] block partials (fixed-order final sum: bit-reproducible)
    unsigned int* ticket;        // zero between calls
    float one_plus_offset;
    int valid_cols;
};

struct RecShared {
    uint64_t recK[kStageCap];
    uint64_t recU[kStageCap];
    uint64_t full[kMaxStages];           // mbarriers, one per stage
    unsigned nK, nU, baseK, baseU;
    unsigned n_ok, n_bad;
};

// Block-wide: move the staged records of one type to the global array (slot range from ONE atomic).
__device__ __forceinline__ void flush_records(uint64_t* s_rec, unsigned* s_n, unsigned* s_base,
                                              uint64_t* __restrict__ g_rec, unsigned long long* g_count) {
    const unsigned n = *s_n;                 // uniform: read after a block barrier
    if (n == 0) return;
    if (threadIdx.x == 0) *s_base = (unsigned)atomicAdd(g_count, (unsigned long long)n);
    __syncthreads();
    const unsigned base = *s_base;
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) g_rec[base + i] = s_rec[i];
    __syncthreads();
    if (threadIdx.x == 0) *s_n = 0;
    __syncthreads();
}

// TMA = true : rows are contiguous (ld == C) and the base is 16-byte aligned; chunks of R rows arrive by bulk copy.
// TMA = false: any ld / alignment / width; rows are read with 128-bit (or scalar) non-allocating loads.
template <typename L, int LPR, bool TMA, bool ALIGNED>
__global__ void __launch_bounds__(kRecThreads, kRecCtasPerSm)
oscr_records_kernel(const float* __restrict__ scores, int64_t ld, const L* __restrict__ gt, int64_t N, int C,
                    int R, int S, int64_t unk_label, bool vec4,
                    uint64_t* __restrict__ recU, uint64_t* __restrict__ recK, unsigned long long* __restrict__ totals, int dbg,
                    const ConfAcc conf) {
    constexpr int RPW = 32 / LPR;
    extern __shared__ __align__(128) uint8_t rec_smem[];
    const uint32_t stage_bytes = TMA ? (uint32_t)R * (uint32_t)C * 4u : 0u;
    const uint32_t lab_bytes = TMA ? (uint32_t)R * (uint32_t)sizeof(L) : 0u;
    uint8_t* stage = rec_smem;
    uint8_t* labs = rec_smem + (size_t)S * stage_bytes;
    RecShared& sh = *reinterpret_cast<RecShared*>(rec_smem + (((size_t)S * (stage_bytes + lab_bytes) + 15) & ~(size_t)15));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, sub = lane % LPR;
    const int n_chunks = (int)((N + R - 1) / R);               // N < 2^30
    const int n_full = (int)(N / R);                           // chunks [0, n_full) hold R rows
    const int G = (int)gridDim.x;
    const int my_row = warp * RPW + lane / LPR;                // this lane group's first row inside a chunk

    if (tid == 0) {
        sh.nK = sh.nU = sh.n_ok = sh.n_bad = 0;
        if (TMA) {
            for (int s = 0; s < S; ++s) tc::mbar_init(&sh.full[s], 1);
            tc::fence_barrier_init();
        }
    }
    __syncthreads();

    auto issue = [&](int chunk, int s) {   // thread 0: start the bulk copy of a full chunk into stage s
        if (chunk >= n_full) return;       // past the end, or the partial last chunk (copied by the consumers)
        const size_t row0 = (size_t)chunk * (size_t)R;
        tc::mbar_arrive_expect_tx(&sh.full[s], stage_bytes + lab_bytes);
        bulk_g2s(stage + (size_t)s * stage_bytes, scores + row0 * (size_t)C, stage_bytes, &sh.full[s]);
        bulk_g2s(labs + (size_t)s * lab_bytes, gt + row0, lab_bytes, &sh.full[s]);
    };
    if (TMA && tid == 0)
        for (int k = 0; k < S; ++k) issue((int)blockIdx.x + k * G, k);

    unsigned n_ok = 0, n_bad = 0;
    double c_kn = 0.0, c_neg = 0.0;                            // confidence sums of this thread's rows (lanes with sub == 0)
    const bool want_conf = conf.sums_out != nullptr;
    const bool conf_cut = want_conf && conf.valid_cols < C;    // the negatives' maximum skips trailing columns (garbage class)
    int s = 0;
    uint32_t phase = 0;
    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += G) {
        const size_t row0 = (size_t)chunk * (size_t)R;
        const int rows = chunk < n_full ? R : (int)(N - (int64_t)row0);
        const float* sbase = reinterpret_cast<const float*>(stage + (size_t)s * stage_bytes);
        const L* slab = reinterpret_cast<const L*>(labs + (size_t)s * lab_bytes);
        if (TMA) {
            if (chunk < n_full) {
                tc::mbar_wait(&sh.full[s], phase);
            } else {
                float* wb = const_cast<float*>(sbase);
                for (int i = tid; i < rows * C; i += kRecThreads) wb[i] = ld_stream_f1(scores + row0 * C + i);
                L* wl = const_cast<L*>(slab);
                for (int i = tid; i < rows; i += kRecThreads) wl[i] = gt[row0 + i];
                __syncthreads();
            }
        }
        for (int r = my_row; r - lane / LPR < rows && !(dbg & 2); r += kRecWarps * RPW) {     // warp-uniform trip count
            const bool valid = r < rows;
            const int64_t l64 = valid ? label_as_int<L>(TMA ? slab[r] : gt[row0 + r]) : -1;
            const bool bad = valid && l64 >= C;                    // reference: IndexError at util.py:111
            const bool known = valid && l64 >= 0 && !bad;          // util.py:101
            const bool unk = valid && l64 == unk_label && !bad;    // util.py:102
            const int lab = known ? (int)l64 : -1;
            float m, ts = 0.f;
            bool ok = false;
            if (TMA) {
                const float* row = sbase + r * C;
                float m_lo; int head, jb;
                // the split column: the label of a known row; for an unk row whose confidence maximum stops at valid_cols,
                // that column (m_lo is then the maximum over the whole groups before it)
                const int split = known ? lab : ((unk && conf_cut) ? conf.valid_cols : -1);
                smem_row_split_max<LPR, ALIGNED>(row, C, split, known || unk, sub, m_lo, m, head, jb);
                if (unk && want_conf && sub == 0) {
                    float mv = m;
                    if (conf_cut) {
                        mv = m_lo;
                        if (ALIGNED) {
                            const float4 g4 = reinterpret_cast<const float4*>(row)[split >> 2];
                            const int q = split & 3;
                            if (q > 0) mv = fmaxf(mv, g4.x);
                            if (q > 1) mv = fmaxf(mv, g4.y);
                            if (q > 2) mv = fmaxf(mv, g4.z);
                        } else {
                            const int c0 = split < head ? split : head + 4 * jb;
#pragma unroll 1
                            for (int c = c0; c < split; ++c) mv = fmaxf(mv, row[c]);
                        }
                    }
                    c_neg += (double)(conf.one_plus_offset - mv);          // float32 subtraction as torch does it (metrics.py:37-40)
                }
                if (known && sub == 0) {
                    ts = row[lab];                                 // util.py:111
                    ok = (ts > m_lo) && (ts == m);                 // first arg-max == gt (util.py:109,114) ...
                    if (ALIGNED) {                                 // ... and inside lab's own 4-element group
                        const float4 g4 = reinterpret_cast<const float4*>(row)[lab >> 2];
                        const int q = lab & 3;
                        ok = ok && (q < 1 || g4.x < ts) && (q < 2 || g4.y < ts) && (q < 3 || g4.z < ts);
                    } else {
                        const int c0 = lab < head ? lab : head + 4 * jb;
#pragma unroll 1
                        for (int c = c0; c < lab; ++c) ok = ok && (row[c] < ts);
                    }
                }
            } else {
                int am;
                group_row_argmax<LPR>(scores + (row0 + r) * ld, C, known || unk, vec4, sub, m, am);      // util.py:109-110
                if (am == 0x7fffffff) am = 0;                      // a row of -inf: numpy's argmax is 0
                if (known && sub == 0) { ts = scores[(row0 + r) * ld + lab]; ok = (am == lab); }
                if (want_conf) {                                   // (this variant is the rare strided / unaligned fallback)
                    float mv = m; int amv;
                    if (conf_cut) group_row_argmax<LPR>(scores + (row0 + r) * ld, conf.valid_cols, unk, vec4, sub, mv, amv);
                    if (unk && sub == 0) c_neg += (double)(conf.one_plus_offset - mv);
                }
            }
            if (want_conf && known && sub == 0) c_kn += (double)ts;       // metrics.py:33
            if (sub == 0 && !(dbg & 1)) {
                if (known) {
                    const unsigned p = atomicAdd(&sh.nK, 1u);
                    sh.recK[p] = make_rec(float_to_key(ts), ok ? kTypeCorrect : kTypeWrong);
                    n_ok += ok;
                }
                if (unk) {
                    const unsigned p = atomicAdd(&sh.nU, 1u);
                    sh.recU[p] = make_rec(float_to_key(m), kTypeUnk);
                }
                n_bad += bad;
            }
        }
        // One barrier does three jobs: every warp is done with stage s; the staged counts are final; and the
        // flush decision is block-uniform (the predicate is monotone in the counts, and the thread that staged the
        // last record evaluates it on the final counts).
        const bool full = __syncthreads_or((sh.nK + (unsigned)R > (unsigned)kStageCap) || (sh.nU + (unsigned)R > (unsigned)kStageCap));
        if (TMA) {
            if (tid == 0) issue(chunk + S * G, s);
            if (++s == S) { s = 0; phase ^= 1u; }
        }
        if (full) {
            flush_records(sh.recK, &sh.nK, &sh.baseK, recK, totals + 0);
            flush_records(sh.recU, &sh.nU, &sh.baseU, recU, totals + 2);
        }
    }
    __syncthreads();
    flush_records(sh.recK, &sh.nK, &sh.baseK, recK, totals + 0);
    flush_records(sh.recU, &sh.nU, &sh.baseU, recU, totals + 2);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { n_ok += __shfl_xor_sync(kFull, n_ok, o); n_bad += __shfl_xor_sync(kFull, n_bad, o); }
    if (lane == 0) { if (n_ok) atomicAdd(&sh.n_ok, n_ok); if (n_bad) atomicAdd(&sh.n_bad, n_bad); }
    __syncthreads();
    if (tid == 0) {
        if (sh.n_ok) atomicAdd(totals + 1, (unsigned long long)sh.n_ok);
        if (sh.n_bad) atomicAdd(totals + 3, (unsigned long long)sh.n_bad);
    }
    if (want_conf) {
        // fixed tree inside the warp, warps in order, blocks in order by the last one out: bit-reproducible for a given grid
        __shared__ double s_conf[2][kRecWarps];
        c_kn = warp_sum(c_kn); c_neg = warp_sum(c_neg);
        if (lane == 0) { s_conf[0][warp] = c_kn; s_conf[1][warp] = c_neg; }
        __syncthreads();
        double mine[2] = {0.0, 0.0};
        if (tid == 0)
            for (int w = 0; w < kRecWarps; ++w) { mine[0] += s_conf[0][w]; mine[1] += s_conf[1][w]; }
        double tot[2];
        if (last_block_reduce<2>(mine, conf.partials, conf.ticket, tot)) { conf.sums_out[0] = tot[0]; conf.sums_out[1] = tot[1]; }
    }
}

// Compact-input variant: the fused head already reduced every row to (p_target, p_max_all, pred).
constexpr int kStatThreads = 256;
template <typename L>
__global__ void __launch_bounds__(kStatThreads)
oscr_records_from_stats_kernel(const float* __restrict__ p_target, const float* __restrict__ p_max, const int32_t* __restrict__ pred,
                               const L* __restrict__ gt, int64_t N, int C, int64_t unk_label,
                               uint64_t* __restrict__ recU, uint64_t* __restrict__ recK, unsigned long long* __restrict__ totals) {
    __shared__ RecShared sh;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) sh.nK = sh.nU = sh.n_ok = sh.n_bad = 0;
    __syncthreads();
    unsigned n_ok = 0, n_bad = 0;
    const int64_t n_chunks = (N + kStatThreads - 1) / kStatThreads;
    for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int64_t row = chunk * kStatThreads + tid;
        if (row < N) {
            const int64_t lab = label_as_int<L>(gt[row]);
            const bool bad = lab >= C;
            if (lab >= 0 && !bad) {
                const bool ok = (pred[row] == (int)lab);
                const unsigned p = atomicAdd(&sh.nK, 1u);
                sh.recK[p] = make_rec(float_to_key(p_target[row]), ok ? kTypeCorrect : kTypeWrong);
                n_ok += ok;
            }
            if (lab == unk_label && !bad) {
                const unsigned p = atomicAdd(&sh.nU, 1u);
                sh.recU[p] = make_rec(float_to_key(p_max[row]), kTypeUnk);
            }
            n_bad += bad;
        }
        if (__syncthreads_or((sh.nK + kStatThreads > (unsigned)kStageCap) || (sh.nU + kStatThreads > (unsigned)kStageCap))) {
            flush_records(sh.recK, &sh.nK, &sh.baseK, recK, totals + 0);
            flush_records(sh.recU, &sh.nU, &sh.baseU, recU, totals + 2);
        }
    }
    __syncthreads();
    flush_records(sh.recK, &sh.nK, &sh.baseK, recK, totals + 0);
    flush_records(sh.recU, &sh.nU, &sh.baseU, recU, totals + 2);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { n_ok += __shfl_xor_sync(kFull, n_ok, o); n_bad += __shfl_xor_sync(kFull, n_bad, o); }
    if (lane == 0) { if (n_ok) atomicAdd(&sh.n_ok, n_ok); if (n_bad) atomicAdd(&sh.n_bad, n_bad); }
    __syncthreads();
    if (tid == 0) {
        if (sh.n_ok) atomicAdd(totals + 1, (unsigned long long)sh.n_ok);
        if (sh.n_bad) atomicAdd(totals + 3, (unsigned long long)sh.n_bad);
    }
}

// ---------------------------------------------------------------------------------------------
// stage B: cooperative sort + scan
// ---------------------------------------------------------------------------------------------
constexpr int kCvThreads = 1024;
constexpr int kCvWarps = kCvThreads / 32;
constexpr int kCvKpt = 8;                        // records per thread per tile, at most
constexpr int kCvTile = kCvThreads * kCvKpt;     // 8192
constexpr int kRadix = 256;
constexpr int kScanRows = 4;                     // thresholds, correct, unk, known
constexpr int kGroup = 16;                       // tiles per group of the two-level cross-tile prefix (more when T > 1024)
constexpr int kMaxGroups = 64;                   // fixed-size group-sum table: its zero-state sits at a fixed workspace offset
constexpr int kTraceSlots = 32;
constexpr int kQuarters = kCvThreads / kRadix;   // 4
constexpr int kWarpsPerQuarter = kCvWarps / kQuarters;
static_assert(kCvWarps == 32, "the second-level scans take one lane per warp");
static_assert(kCvThreads % kRadix == 0 && kCvWarps % kQuarters == 0, "block shape");

struct CurveParams {
    const uint64_t* rec_a; const int64_t* n_a;   // logical sort input = rec_a[0..n_a) ++ rec_b[0..n_b)
    const uint64_t* rec_b; const int64_t* n_b;
    int64_t cap;                                 // n_a + n_b <= cap
    uint64_t* buf0; uint64_t* buf1;              // ping-pong, cap records each
    uint32_t* gsum;                              // [4][64][256]   digit counts per group of tiles; zero at entry and at exit
    uint32_t* matrix;                            // [tpad][256]    digit counts per tile, current pass (digit contiguous: coalesced)
    uint32_t* scanmat;                           // [4][tpad]      per-tile scan totals
    unsigned int* sync;                          // [0] barrier counter, [1] exit ticket; zero at entry and at exit
    int tpad;
    float* taus; int64_t* ccr_cnt; int64_t* fpr_cnt; double* ccr; double* fpr;
    int64_t* meta;                               // [5] n_unique, n_known, n_correct, n_unk, n_bad
    unsigned long long* totals_in;               // nullable: stage-A totals (n_bad -> meta[4]); zeroed at exit
    long long* trace;                            // nullable developer hook: [grid][kTraceSlots] SM clock stamps
};

struct CurveShared {
    uint64_t rec[kCvTile];                       // tile of records; in the emit phase: staged threshold keys | correct counts
    uint32_t stage_u[kCvTile];                   // emit phase: staged unk counts
    uint32_t whist[kCvWarps][kRadix];
    uint32_t qsum[kQuarters][kRadix], qpre[kQuarters][kRadix];
    uint32_t start[kRadix];
    uint32_t goff[kRadix];
    uint32_t part[kRadix];
    uint32_t wsum_l[kCvWarps], wsum_g[kCvWarps];
    uint32_t wscan[3][kCvWarps];
    uint32_t tile_tot[kScanRows];
    uint32_t pre[kScanRows], tot[kScanRows];
};

// All CTAs are co-resident (cooperative launch).  Release: every write of this CTA (ordered before by bar.sync) is
// published by the gpu-scope release reduction; acquire: the polling load orders the reads that follow bar.sync.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& target, unsigned int G) {
    target += G;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(counter), "r"(1u) : "memory");
        while (ld_acquire_u32(counter) < target) { }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kCvThreads, 1)
oscr_curve_kernel(const CurveParams p) {
    extern __shared__ __align__(16) uint8_t cv_smem[];
    CurveShared& sh = *reinterpret_cast<CurveShared*>(cv_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned int G = gridDim.x, g = blockIdx.x;
    unsigned int bar_target = 0;
    long long* trace = p.trace ? p.trace + (size_t)g * kTraceSlots : nullptr;
    auto stamp = [&](int slot) { if (trace && tid == 0) trace[slot] = clock64(); };

    int64_t nA = p.n_a ? p.n_a[0] : 0, nB = p.n_b ? p.n_b[0] : 0;
    nA = nA < 0 ? 0 : (nA > p.cap ? p.cap : nA);
    nB = nB < 0 ? 0 : (nB > p.cap - nA ? p.cap - nA : nB);
    const int64_t n = nA + nB;
    // tile size: spread the records over the grid, 512..8192 records per tile
    int kpt = (int)((n + (int64_t)G * kCvThreads - 1) / ((int64_t)G * kCvThreads));
    kpt = kpt < 1 ? 1 : (kpt > kCvKpt ? kCvKpt : kpt);
    const int tile_sz = kpt * kCvThreads;
    const int64_t T = (n + tile_sz - 1) / tile_sz;
    const int64_t grp = T > (int64_t)kGroup * kMaxGroups ? (T + kMaxGroups - 1) / kMaxGroups : kGroup;   // tiles per group
    const int64_t NG = (T + grp - 1) / grp;          // <= kMaxGroups
    const bool single = T <= (int64_t)G;            // at most one tile per CTA: records stay in registers across the barrier
    const size_t tpad = (size_t)p.tpad;

    uint64_t rec[kCvKpt];
    uint32_t rnk[kCvKpt];
    uint32_t tile_cnt = 0;                          // thread d < 256: records of this tile with digit d
    const int wbase = warp * (kpt * 32);

#pragma unroll 1
    for (int pass = 0; pass < 4; ++pass) {
        const uint64_t* src = (pass & 1) ? p.buf0 : p.buf1;     // pass 0 reads rec_a ++ rec_b instead
        uint64_t* dst = (pass & 1) ? p.buf1 : p.buf0;
        const int shift = 8 * pass;                 // digit of the 32-bit key = high word of the record
        uint32_t* gs = p.gsum + (size_t)pass * kMaxGroups * kRadix;
        stamp(pass * 6 + 0);

        // rank the records of tile t inside their warp slice, in index order (stable); per-warp digit counts
        auto load_rank = [&](int64_t t) {
            for (int i = tid; i < kCvWarps * kRadix; i += kCvThreads) (&sh.whist[0][0])[i] = 0;
            __syncthreads();
            if (pass == 1) stamp(28);
            const int64_t base = t * tile_sz;
            const int count = (int)((n - base) < tile_sz ? (n - base) : tile_sz);
#pragma unroll
            for (int r = 0; r < kCvKpt; ++r) {              // predicated, not branched: every load is in flight at once
                const int idx = wbase + r * 32 + lane;
                const int64_t gi = base + idx;
                const uint64_t* ptr = pass == 0 ? (gi < nA ? p.rec_a + gi : p.rec_b + (gi - nA)) : src + gi;
                rec[r] = (r < kpt && idx < count) ? __ldcg(ptr) : ~0ull;
            }
#pragma unroll
            for (int r = 0; r < kCvKpt; ++r) {
                if (r < kpt) {
                    const bool valid = (wbase + r * 32 + lane) < count;
                    const uint32_t d = ((uint32_t)(rec[r] >> 32) >> shift) & 0xffu;
                    // lanes holding the same digit: 8 ballots (match.any costs one round per DISTINCT value in the warp)
                    unsigned peers = __ballot_sync(kFull, valid);
#pragma unroll
                    for (int b = 0; b < 8; ++b) {
                        const bool bit = (d & (1u << b)) != 0u;
                        const unsigned m = __ballot_sync(kFull, bit);
                        peers &= m ^ (bit ? 0u : 0xffffffffu);
                    }
                    if (!valid) peers = 1u << lane;
                    const int leader = __ffs(peers) - 1;
                    uint32_t old = 0;
                    if (valid && lane == leader) { old = sh.whist[warp][d]; sh.whist[warp][d] = old + __popc(peers); }
                    old = __shfl_sync(kFull, old, leader);
                    rnk[r] = old + __popc(peers & ((1u << lane) - 1u));
                    __syncwarp();
                    if (pass == 1 && r == 0) stamp(29);
                }
            }
            if (pass == 1) stamp(30);
            __syncthreads();
            if (pass == 1) stamp(31);
            // exclusive scan of the per-warp digit counts over the warps: quarter q of the block takes warps
            // [q*W, (q+1)*W) of digit tid & 255, then the quarters are chained through qpre
            {
                const int d = tid & (kRadix - 1), q = tid >> 8;
                uint32_t run = 0;
#pragma unroll
                for (int w = 0; w < kWarpsPerQuarter; ++w) {
                    const uint32_t c = sh.whist[q * kWarpsPerQuarter + w][d];
                    sh.whist[q * kWarpsPerQuarter + w][d] = run;
                    run += c;
                }
                sh.qsum[q][d] = run;
            }
            __syncthreads();
            {
                const int d = tid & (kRadix - 1), q = tid >> 8;
                uint32_t pre = 0, tot = 0;
#pragma unroll
                for (int qq = 0; qq < kQuarters; ++qq) { const uint32_t v = sh.qsum[qq][d]; tot += v; pre += qq < q ? v : 0u; }
                sh.qpre[q][d] = pre;
                tile_cnt = tid < kRadix ? tot : 0u;
            }
            return count;
        };

        // ---- phase 1: digit counts of every tile -> matrix, and added into the tile group's sums ----
        int my_count = 0;
        for (int64_t t = g; t < T; t += G) {
            my_count = load_rank(t);
            if (tid < kRadix) {
                p.matrix[(size_t)t * kRadix + tid] = tile_cnt;
                if (tile_cnt) atomicAdd(&gs[(size_t)(t / grp) * kRadix + tid], tile_cnt);
            }
            if (!single) __syncthreads();
        }
        stamp(pass * 6 + 1);
        grid_barrier(p.sync, bar_target, G);
        stamp(pass * 6 + 2);

        // ---- phase 2: offsets, then scatter ----
        for (int64_t t = g; t < T; t += G) {
            if (!single) my_count = load_rank(t);
            // digit d = tid & 255.  First half of the block: sums over whole groups (prefix and grand total);
            // second half: the tiles of t's own group that precede t.
            const int d = tid & (kRadix - 1);
            const int64_t tg = t / grp;
            uint32_t before = 0, total = 0;
            if (tid < kRadix) {
                const uint32_t* col = gs + d;
                if (NG <= 16) {                       // common case: all loads issued at once
                    uint32_t v[16];
#pragma unroll
                    for (int gg = 0; gg < 16; ++gg) v[gg] = gg < NG ? __ldcg(col + gg * kRadix) : 0u;
#pragma unroll
                    for (int gg = 0; gg < 16; ++gg) { total += v[gg]; before += gg < tg ? v[gg] : 0u; }
                } else {
#pragma unroll 4
                    for (int64_t gg = 0; gg < NG; ++gg) {
                        const uint32_t v = __ldcg(col + gg * kRadix);
                        total += v;
                        before += gg < tg ? v : 0u;
                    }
                }
            } else if (tid < 2 * kRadix) {
                const uint32_t* col = p.matrix + d;
                uint32_t acc = 0;
                if (grp == kGroup) {
                    uint32_t v[kGroup];
#pragma unroll
                    for (int q = 0; q < kGroup; ++q) v[q] = (tg * kGroup + q) < t ? __ldcg(col + (size_t)(tg * kGroup + q) * kRadix) : 0u;
#pragma unroll
                    for (int q = 0; q < kGroup; ++q) acc += v[q];
                } else {
#pragma unroll 8
                    for (int64_t tt = tg * grp; tt < t; ++tt) acc += __ldcg(col + (size_t)tt * kRadix);
                }
                sh.part[d] = acc;
            }
            // exclusive scans over the digits: global bases (from the totals) and tile-local starts
            uint32_t inc_g = total, inc_l = tile_cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t a = __shfl_up_sync(kFull, inc_g, o), b = __shfl_up_sync(kFull, inc_l, o);
                if (lane >= o) { inc_g += a; inc_l += b; }
            }
            if (lane == 31) { sh.wsum_g[warp] = inc_g; sh.wsum_l[warp] = inc_l; }
            __syncthreads();
            if (tid < kRadix) {
                uint32_t off_g = 0, off_l = 0;
#pragma unroll
                for (int w = 0; w < kRadix / 32; ++w) if (w < warp) { off_g += sh.wsum_g[w]; off_l += sh.wsum_l[w]; }
                const uint32_t base_g = off_g + inc_g - total;
                const uint32_t start_l = off_l + inc_l - tile_cnt;
                sh.start[tid] = start_l;
                sh.goff[tid] = base_g + before + sh.part[tid] - start_l;
            }
            __syncthreads();
            if (t == g) stamp(pass * 6 + 3);
#pragma unroll
            for (int r = 0; r < kCvKpt; ++r) {
                if (r < kpt) {
                    const int idx = wbase + r * 32 + lane;
                    if (idx < my_count) {
                        const uint32_t dd = ((uint32_t)(rec[r] >> 32) >> shift) & 0xffu;
                        sh.rec[sh.start[dd] + sh.qpre[warp / kWarpsPerQuarter][dd] + sh.whist[warp][dd] + rnk[r]] = rec[r];
                    }
                }
            }
            __syncthreads();
            for (int i = tid; i < my_count; i += kCvThreads) {
                const uint64_t v = sh.rec[i];
                const uint32_t dd = ((uint32_t)(v >> 32) >> shift) & 0xffu;
                dst[sh.goff[dd] + (uint32_t)i] = v;
            }
            if (!single) __syncthreads();
        }
        stamp(pass * 6 + 4);
        grid_barrier(p.sync, bar_target, G);
        stamp(pass * 6 + 5);
    }

    // ---- scan of the sorted records (buf1): thresholds and counts ----
    const uint64_t* sorted = p.buf1;
    uint32_t eE = 0, eC = 0, eU = 0;       // this thread's exclusive prefixes inside its tile
    auto scan_tile = [&](int64_t t) {
        // coalesced load of the tile (+ the first record of the next tile) into shared memory, then each thread
        // takes kpt CONSECUTIVE records so that prefixes inside a thread are sequential
        const int64_t tbase = t * tile_sz;
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kCvKpt; ++r) {                  // all loads in flight, then the shared-memory stores
            const int64_t gi = tbase + r * kCvThreads + tid;
            rec[r] = (r < kpt && gi < n) ? __ldcg(sorted + gi) : 0ull;
        }
#pragma unroll
        for (int r = 0; r < kCvKpt; ++r)
            if (r < kpt) sh.rec[r * kCvThreads + tid] = rec[r];
        __shared__ uint64_t s_next;
        if (tid == 0 && tbase + tile_sz < n) s_next = __ldcg(sorted + tbase + tile_sz);
        __syncthreads();
        const int64_t i0 = tbase + (int64_t)tid * kpt;
        uint64_t nxt = (i0 < n) ? sh.rec[tid * kpt] : 0ull;
        uint32_t cE = 0, cC = 0, cU = 0, cK = 0;
#pragma unroll
        for (int r = 0; r < kCvKpt; ++r) {
            if (r < kpt) {
                const int64_t i = i0 + r;
                const uint64_t v = nxt;
                rec[r] = v;
                if (i < n) {
                    const bool last = (i + 1 >= n);
                    if (!last) nxt = (tid * kpt + r + 1 < tile_sz) ? sh.rec[tid * kpt + r + 1] : s_next;
                    const uint32_t type = (uint32_t)v & 3u;
                    const bool known = type < 2u;
                    const bool end = known && (last || (uint32_t)(nxt >> 32) != (uint32_t)(v >> 32));
                    rnk[r] = end ? 1u : 0u;
                    cE += end; cC += (type == 1u); cU += (type == 2u); cK += known;
                } else {
                    rnk[r] = 0u;
                }
            }
        }
        // block-exclusive scan of (cE, cC, cU); block total of cK
        uint32_t iE = cE, iC = cC, iU = cU, sK = cK;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t a = __shfl_up_sync(kFull, iE, o), b = __shfl_up_sync(kFull, iC, o), c = __shfl_up_sync(kFull, iU, o);
            if (lane >= o) { iE += a; iC += b; iU += c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sK += __shfl_xor_sync(kFull, sK, o);
        __syncthreads();                  // previous users of wscan / tile_tot are done
        if (lane == 31) { sh.wscan[0][warp] = iE; sh.wscan[1][warp] = iC; sh.wscan[2][warp] = iU; }
        if (lane == 0) sh.wsum_l[warp] = sK;
        __syncthreads();
        // second level: warps 0..3 scan the per-warp totals of one quantity each with shuffles (kCvWarps == 32 lanes)
        if (warp < 4) {
            const uint32_t v = warp < 3 ? sh.wscan[warp][lane] : sh.wsum_l[lane];
            uint32_t inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t a = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += a; }
            if (warp < 3) sh.wscan[warp][lane] = inc - v;          // exclusive prefix of the warp totals
            if (lane == 31) sh.tile_tot[warp] = inc;               // tile totals: thresholds, correct, unk, known
        }
        __syncthreads();
        eE = sh.wscan[0][warp] + iE - cE; eC = sh.wscan[1][warp] + iC - cC; eU = sh.wscan[2][warp] + iU - cU;
    };
    stamp(24);
    for (int64_t t = g; t < T; t += G) {
        scan_tile(t);
        if (tid < kScanRows) p.scanmat[(size_t)tid * tpad + (size_t)t] = sh.tile_tot[tid];
    }
    stamp(25);
    grid_barrier(p.sync, bar_target, G);
    stamp(26);
    // totals over all tiles, and the prefix of this CTA's first tile, from one sweep (warp x sums row x)
    if (warp < kScanRows) {
        const uint32_t* row = p.scanmat + (size_t)warp * tpad;
        uint32_t acc = 0, pre = 0;
        for (int64_t tt = lane; tt < T; tt += 32) { const uint32_t v = __ldcg(row + tt); acc += v; pre += tt < (int64_t)g ? v : 0u; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { acc += __shfl_xor_sync(kFull, acc, o); pre += __shfl_xor_sync(kFull, pre, o); }
        if (lane == 0) { sh.tot[warp] = acc; sh.pre[warp] = pre; }
    }
    __syncthreads();
    const int64_t n_unique = sh.tot[0], n_correct = sh.tot[1], n_unk = sh.tot[2], n_known = sh.tot[3];
    // count / total as a correctly rounded float64 quotient without a division per threshold: with r = RN(1/b) and
    // q = RN(a*r), q' = fma(fma(-q, b, a), r, q) is RN(a/b) (Markstein) for these operands (integers below 2^30: no
    // overflow, no subnormals, and b's significand is never all ones).  0/0 stays NaN like numpy.
    const double d_known = (double)n_known, d_unk = (double)n_unk;
    const double r_known = 1.0 / d_known, r_unk = 1.0 / d_unk;
    auto exact_div = [](double a, double b, double r) {
        const double q = a * r;
        return fma(fma(-q, b, a), r, q);
    };
    uint32_t wrun = sh.pre[warp < kScanRows ? warp : 0];   // warp x < 4: sum of row x over tiles [0, t_seen)
    int64_t t_seen = g;
    for (int64_t t = g; t < T; t += G) {
        if (!single) scan_tile(t);
        if (warp < kScanRows && t > (int64_t)g) {
            const uint32_t* row = p.scanmat + (size_t)warp * tpad;
            uint32_t acc = 0;
            for (int64_t tt = t_seen + lane; tt < t; tt += 32) acc += __ldcg(row + tt);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
            wrun += acc;
            t_seen = t;
            if (lane == 0) sh.pre[warp] = wrun;
        }
        __syncthreads();
        // Stage this tile's thresholds in shared memory at their tile-local index, then write the five output arrays
        // with fully coalesced stores (a thread's own thresholds are a few slots apart: direct stores waste sectors).
        uint32_t* s_key = reinterpret_cast<uint32_t*>(sh.rec);
        uint32_t* s_cc = s_key + kCvTile;
        uint32_t jl = eE;
        int64_t cc = (int64_t)sh.pre[1] + eC, uu = (int64_t)sh.pre[2] + eU;   // records with key <= current, so far
#pragma unroll
        for (int r = 0; r < kCvKpt; ++r) {
            if (r < kpt) {
                const uint32_t type = (uint32_t)rec[r] & 3u;
                const bool inb = (t * tile_sz + (int64_t)tid * kpt + r) < n;
                cc += (inb && type == 1u); uu += (inb && type == 2u);
                if (rnk[r]) {
                    s_key[jl] = (uint32_t)(rec[r] >> 32);
                    s_cc[jl] = (uint32_t)(n_correct - cc);                      // strict '>' (util.py:114,117)
                    sh.stage_u[jl] = (uint32_t)(n_unk - uu);
                    ++jl;
                }
            }
        }
        __syncthreads();
        const int64_t j0 = (int64_t)sh.pre[0];
        const int tile_e = (int)sh.tile_tot[0];
        for (int i = tid; i < tile_e; i += kCvThreads) {
            const int64_t j = j0 + i;
            const int64_t c_gt = s_cc[i], u_gt = sh.stage_u[i];
            if (p.taus) p.taus[j] = key_to_float(s_key[i]);
            if (p.ccr_cnt) p.ccr_cnt[j] = c_gt;
            if (p.fpr_cnt) p.fpr_cnt[j] = u_gt;
            if (p.ccr) p.ccr[j] = exact_div((double)c_gt, d_known, r_known);   // util.py:114
            if (p.fpr) p.fpr[j] = exact_div((double)u_gt, d_unk, r_unk);       // util.py:117 (0/0 -> NaN like numpy)
        }
        __syncthreads();
    }
    if (g == 0 && tid == 0) {
        p.meta[0] = n_unique; p.meta[1] = n_known; p.meta[2] = n_correct; p.meta[3] = n_unk;
        p.meta[4] = p.totals_in ? (int64_t)p.totals_in[3] : 0;
    }
    stamp(27);
    // ---- leave the workspace reusable: the last CTA out zeroes the counters ----
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned int tk = atomicAdd(p.sync + 1, 1u);
        sh.tile_tot[0] = (tk == G - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (sh.tile_tot[0]) {
        const int used = (int)NG * kRadix;                       // entries touched per pass: groups [0, NG) x 256 digits
        for (int i = tid; i < 4 * used; i += kCvThreads) p.gsum[(size_t)(i / used) * kMaxGroups * kRadix + (i % used)] = 0u;
        if (p.totals_in && tid < 4) p.totals_in[tid] = 0ull;
        if (tid == 0) { p.sync[0] = 0u; p.sync[1] = 0u; }
    }
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
static int g_rec_debug = 0;      // developer hook: bit 0 = do not emit records, bit 1 = skip the row loop (timing experiments only)
struct RecPlan { bool tma; bool aligned; int lpr; int R; int S; size_t smem; bool vec4; int grid; };

template <typename L>
static RecPlan plan_records(const float* scores, int64_t ld, const L* gt, int64_t N, int C) {
    RecPlan pl = {};
    const size_t row_bytes = (size_t)C * 4;
    const bool base_ok = (reinterpret_cast<uintptr_t>(scores) % 16 == 0) && (reinterpret_cast<uintptr_t>(gt) % 16 == 0);
    pl.tma = base_ok && ld == C;
    pl.aligned = (C % 4 == 0);
    // lanes per row: 4 for narrow aligned rows (8 float4 per lane at C = 128), 8 up to 256 columns, a whole warp above
    pl.lpr = (C <= 128 && pl.aligned && pl.tma) ? 4 : (C <= 256 ? 8 : 32);
    if (pl.tma) {
        // rows per chunk: one sweep of the 8 warps when it fits (<= 32 KB per stage), else fewer rows
        int R = kRecWarps * (32 / pl.lpr);                       // 64, 32 or 8
        while (R > 8 && (size_t)R * row_bytes > 32768) R >>= 1;
        if ((size_t)R * row_bytes > kRecStageBudget / 2) pl.tma = false;
        pl.R = R;
    }
    if (pl.tma) {
        const size_t per_stage = (size_t)pl.R * row_bytes + (size_t)pl.R * sizeof(L);
        int S = (int)(kRecStageBudget / per_stage);
        pl.S = S > kMaxStages ? kMaxStages : S;
        if (pl.S < 2) pl.tma = false;
    }
    if (pl.tma) {
        pl.vec4 = true;                                          // unused by the shared-memory path
        pl.smem = (((size_t)pl.S * ((size_t)pl.R * row_bytes + (size_t)pl.R * sizeof(L)) + 15) & ~(size_t)15) + sizeof(RecShared);
    } else {
        pl.lpr = C <= 256 ? 8 : 32;
        pl.R = 64; pl.S = 0;
        pl.vec4 = rows_are_vec4(scores, ld);
        pl.smem = sizeof(RecShared) + 16;
    }
    const int64_t n_chunks = (N + pl.R - 1) / pl.R;
    const int cap = sm_count() * kRecCtasPerSm;
    pl.grid = (int)(n_chunks < 1 ? 1 : (n_chunks > cap ? cap : n_chunks));
    return pl;
}

template <typename L, int LPR, bool TMA, bool ALIGNED>
static int launch_records_k(const RecPlan& pl, const float* scores, int64_t ld, const L* gt, int64_t N, int C, int64_t unk_label,
                            uint64_t* recU, uint64_t* recK, unsigned long long* totals, const ConfAcc& conf, cudaStream_t st) {
    auto k = oscr_records_kernel<L, LPR, TMA, ALIGNED>;
    static size_t smem_set = 0;                                  // per instantiation
    if (pl.smem > smem_set) {
        OSB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        smem_set = pl.smem;
    }
    k<<<pl.grid, kRecThreads, pl.smem, st>>>(scores, ld, gt, N, C, pl.R, pl.S, unk_label, pl.vec4, recU, recK, totals, g_rec_debug, conf);
    OSB_LAUNCH_CHECK();
    return OSB_OK;
}

template <typename L>
static int launch_records(const float* scores, int64_t ld, const L* gt, int64_t N, int C, int64_t unk_label,
                          uint64_t* recU, uint64_t* recK, unsigned long long* totals, const ConfAcc& conf, cudaStream_t st) {
    const RecPlan pl = plan_records(scores, ld, gt, N, C);
#define OSB_REC(LPR_, TMA_, AL_) return launch_records_k<L, LPR_, TMA_, AL_>(pl, scores, ld, gt, N, C, unk_label, recU, recK, totals, conf, st)
    if (pl.tma) {
        if (pl.lpr == 4) OSB_REC(4, true, true);
        if (pl.lpr == 8) { if (pl.aligned) OSB_REC(8, true, true); OSB_REC(8, true, false); }
        if (pl.aligned) OSB_REC(32, true, true);
        OSB_REC(32, true, false);
    }
    if (pl.lpr == 8) OSB_REC(8, false, false);
    OSB_REC(32, false, false);
#undef OSB_REC
}

static int run_records(const float* scores, int64_t ld, const void* gt, int gt_type, int64_t N, int C, int64_t unk_label,
                       uint64_t* recU, uint64_t* recK, unsigned long long* totals, cudaStream_t st, const ConfAcc& conf = ConfAcc{}) {
    if (N == 0) {
        if (conf.sums_out) OSB_CUDA(cudaMemsetAsync(conf.sums_out, 0, 2 * sizeof(double), st));
        return OSB_OK;
    }
    switch (gt_type) {
        case OSB_LABEL_I64: return launch_records(scores, ld, reinterpret_cast<const int64_t*>(gt), N, C, unk_label, recU, recK, totals, conf, st);
        case OSB_LABEL_F32: return launch_records(scores, ld, reinterpret_cast<const float*>(gt), N, C, unk_label, recU, recK, totals, conf, st);
        case OSB_LABEL_I32: return launch_records(scores, ld, reinterpret_cast<const int32_t*>(gt), N, C, unk_label, recU, recK, totals, conf, st);
    }
    set_error("osb_oscr: bad label type %d", gt_type);
    return OSB_ERR_ARG;
}

// ---- curve workspace: [sync 256 B][group sums][matrix][scanmat][buf0][buf1]; the zero-state (sync) is at offset 0 ----
static int curve_grid(int64_t cap) {
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    int64_t want = (cap + kCvThreads - 1) / kCvThreads;
    return (int)(want < 1 ? 1 : (want > sms ? sms : want));
}
static int curve_tpad(int64_t cap) {
    // tiles never exceed max(grid, ceil(cap / 8192)); padded to a multiple of 4
    int64_t t = (cap + kCvTile - 1) / kCvTile;
    const int g = curve_grid(cap);
    if (t < g) t = g;
    return (int)((t + 4) & ~(int64_t)3);
}
struct CurveWs { unsigned int* sync; uint32_t* gsum; uint32_t* matrix; uint32_t* scanmat; uint64_t* buf0; uint64_t* buf1; size_t bytes; };
static CurveWs carve_curve_ws(void* ws, int64_t cap) {
    CurveWs c;
    char* p = reinterpret_cast<char*>(ws);
    const size_t tpad = (size_t)curve_tpad(cap);
    c.sync = reinterpret_cast<unsigned int*>(p); p += 256;
    c.gsum = reinterpret_cast<uint32_t*>(p); p += 4 * kRadix * kMaxGroups * sizeof(uint32_t);      // fixed size
    c.matrix = reinterpret_cast<uint32_t*>(p); p += align_up(kRadix * tpad * sizeof(uint32_t), 256);
    c.scanmat = reinterpret_cast<uint32_t*>(p); p += align_up(kScanRows * tpad * sizeof(uint32_t), 256);
    c.buf0 = reinterpret_cast<uint64_t*>(p); p += align_up((size_t)(cap > 0 ? cap : 1) * sizeof(uint64_t), 256);
    c.buf1 = reinterpret_cast<uint64_t*>(p); p += align_up((size_t)(cap > 0 ? cap : 1) * sizeof(uint64_t), 256);
    c.bytes = (size_t)(p - reinterpret_cast<char*>(ws));
    return c;
}
static size_t curve_ws_bytes(int64_t cap) { return carve_curve_ws(nullptr, cap).bytes; }

static long long* g_curve_trace = nullptr;

static int run_curve(const uint64_t* rec_a, const int64_t* n_a, const uint64_t* rec_b, const int64_t* n_b, int64_t cap,
                     float* taus, int64_t* ccr_cnt, int64_t* fpr_cnt, double* ccr, double* fpr, int64_t* meta,
                     unsigned long long* totals_in, void* ws, cudaStream_t st) {
    const CurveWs c = carve_curve_ws(ws, cap);
    CurveParams p;
    p.rec_a = rec_a; p.n_a = n_a; p.rec_b = rec_b; p.n_b = n_b; p.cap = cap;
    p.buf0 = c.buf0; p.buf1 = c.buf1; p.gsum = c.gsum; p.matrix = c.matrix; p.scanmat = c.scanmat; p.sync = c.sync;
    p.tpad = curve_tpad(cap);
    p.taus = taus; p.ccr_cnt = ccr_cnt; p.fpr_cnt = fpr_cnt; p.ccr = ccr; p.fpr = fpr; p.meta = meta; p.totals_in = totals_in;
    p.trace = g_curve_trace;
    const size_t smem = sizeof(CurveShared);
    static bool attr_set = false;
    if (!attr_set) {
        OSB_CUDA(cudaFuncSetAttribute(oscr_curve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(curve_grid(cap));
    cfg.blockDim = dim3(kCvThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;       // every CTA resident at once: the grid barriers cannot deadlock
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    OSB_CUDA(cudaLaunchKernelEx(&cfg, oscr_curve_kernel, p));
    return OSB_OK;
}

}  // namespace osb

using namespace osb;

// developer hook (not part of the public header): device buffer of [grid][32] int64 SM clock stamps, or NULL
extern "C" void osb_debug_set_curve_trace(void* p) { g_curve_trace = reinterpret_cast<long long*>(p); }
extern "C" void osb_debug_set_records_mode(int m) { g_rec_debug = m; }

// workspace of the composed entry: [totals 256 B][curve ws][rec_unk N][rec_known N].  Everything that must be zero
// between calls (totals, barrier counters, digit histograms) sits at FIXED offsets at the front, so one workspace
// serves calls of any size.
extern "C" size_t osb_oscr_curve_workspace_bytes(int64_t cap) { return cap < 0 ? 0 : curve_ws_bytes(cap); }
constexpr size_t kConfAreaBytes = 128 + 2 * sizeof(double) * 1024;      // ticket + block partials of the optional confidence sums
extern "C" size_t osb_oscr_workspace_bytes(int64_t N) {
    if (N < 0) return 0;
    return 256 + curve_ws_bytes(N) + 2 * align_up((size_t)(N > 0 ? N : 1) * sizeof(uint64_t), 256) + kConfAreaBytes;
}

static int check_oscr_sizes(const char* fn, int64_t N, int C, int64_t ld) {
    OSB_REQUIRE(N >= 0 && N < (1LL << 30) && C > 0 && ld >= C, OSB_ERR_ARG, "%s: bad sizes N=%lld C=%d ld=%lld", fn,
                (long long)N, C, (long long)ld);
    return OSB_OK;
}

extern "C" int osb_oscr_records(const float* scores, int64_t ld, const void* gt, int gt_type, int64_t N, int C,
                                int64_t unk_label, uint64_t* rec_unk, uint64_t* rec_known, int64_t* totals, void* stream) {
    OSB_ENTRY_GUARD();
    OSB_REQUIRE((scores && gt && rec_unk && rec_known) || N == 0, OSB_ERR_ARG, "osb_oscr_records: null pointer");
    OSB_REQUIRE(totals, OSB_ERR_ARG, "osb_oscr_records: null totals");
    int rc = check_oscr_sizes("osb_oscr_records", N, C, ld);
    if (rc) return rc;
    OSB_CUDA(cudaMemsetAsync(totals, 0, 4 * sizeof(int64_t), (cudaStream_t)stream));
    return run_records(scores, ld, gt, gt_type, N, C, unk_label, rec_unk, rec_known,
                       reinterpret_cast<unsigned long long*>(totals), (cudaStream_t)stream);
}

extern "C" int osb_oscr_records_from_stats(const float* p_target, const float* p_max_all, const int32_t* pred,
                                           const void* gt, int gt_type, int64_t N, int C, int64_t unk_label,
                                           uint64_t* rec_unk, uint64_t* rec_known, int64_t* totals, void* stream) {
    OSB_ENTRY_GUARD();
    OSB_REQUIRE((p_target && p_max_all && pred && gt && rec_unk && rec_known) || N == 0, OSB_ERR_ARG, "osb_oscr_records_from_stats: null pointer");
    OSB_REQUIRE(totals, OSB_ERR_ARG, "osb_oscr_records_from_stats: null totals");
    int rc = check_oscr_sizes("osb_oscr_records_from_stats", N, C, C);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    OSB_CUDA(cudaMemsetAsync(totals, 0, 4 * sizeof(int64_t), st));
    if (N == 0) return OSB_OK;
    auto* tot = reinterpret_cast<unsigned long long*>(totals);
    const int64_t n_chunks = (N + kStatThreads - 1) / kStatThreads;
    const int cap = sm_count() * 4;
    const int grid = (int)(n_chunks > cap ? cap : n_chunks);
    switch (gt_type) {
        case OSB_LABEL_I64:
            oscr_records_from_stats_kernel<<<grid, kStatThreads, 0, st>>>(p_target, p_max_all, pred, reinterpret_cast<const int64_t*>(gt), N, C, unk_label, rec_unk, rec_known, tot);
            break;
        case OSB_LABEL_F32:
            oscr_records_from_stats_kernel<<<grid, kStatThreads, 0, st>>>(p_target, p_max_all, pred, reinterpret_cast<const float*>(gt), N, C, unk_label, rec_unk, rec_known, tot);
            break;
        case OSB_LABEL_I32:
            oscr_records_from_stats_kernel<<<grid, kStatThreads, 0, st>>>(p_target, p_max_all, pred, reinterpret_cast<const int32_t*>(gt), N, C, unk_label, rec_unk, rec_known, tot);
            break;
        default:
            set_error("osb_oscr_records_from_stats: bad label type %d", gt_type);
            return OSB_ERR_ARG;
    }
    OSB_LAUNCH_CHECK();
    return OSB_OK;
}

extern "C" int osb_oscr_curve(const uint64_t* rec_a, const int64_t* n_a, const uint64_t* rec_b, const int64_t* n_b, int64_t cap,
                              float* taus, int64_t* ccr_cnt, int64_t* fpr_cnt, double* ccr, double* fpr, int64_t* meta,
                              void* ws, size_t ws_bytes, void* stream) {
    OSB_ENTRY_GUARD();
    OSB_REQUIRE(meta && ws, OSB_ERR_ARG, "osb_oscr_curve: null pointer");
    OSB_REQUIRE(cap >= 0 && cap < (1LL << 30), OSB_ERR_ARG, "osb_oscr_curve: bad cap=%lld", (long long)cap);
    OSB_REQUIRE((!rec_a || n_a) && (!rec_b || n_b), OSB_ERR_ARG, "osb_oscr_curve: record array without its count");
    OSB_REQUIRE(cap == 0 || (ccr_cnt && fpr_cnt) || (ccr && fpr), OSB_ERR_ARG, "osb_oscr_curve: needs the count outputs or the ratio outputs");
    OSB_REQUIRE(ws_bytes >= curve_ws_bytes(cap), OSB_ERR_WORKSPACE, "osb_oscr_curve: workspace too small (%zu < %zu)", ws_bytes, curve_ws_bytes(cap));
    return run_curve(rec_a, rec_a ? n_a : nullptr, rec_b, rec_b ? n_b : nullptr, cap, taus, ccr_cnt, fpr_cnt, ccr, fpr, meta, nullptr,
                     ws, (cudaStream_t)stream);
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

static int oscr_impl(const char* fn, const float* scores, int64_t ld, const void* gt, int gt_type, int64_t N, int C,
                     int64_t unk_label, float* taus, int64_t* ccr_cnt, int64_t* fpr_cnt,
                     double* ccr, double* fpr, int64_t* meta, double* conf_sums, float one_plus_offset, int valid_cols,
                     void* ws, size_t ws_bytes, void* stream) {
    OSB_REQUIRE(meta && ws, OSB_ERR_ARG, "%s: null pointer", fn);
    int rc = check_oscr_sizes(fn, N, C, ld);
    if (rc) return rc;
    OSB_REQUIRE(N == 0 || (scores && gt && ccr && fpr), OSB_ERR_ARG, "%s: null pointer", fn);
    OSB_REQUIRE(ws_bytes >= osb_oscr_workspace_bytes(N), OSB_ERR_WORKSPACE, "%s: workspace too small (%zu < %zu)", fn,
                ws_bytes, osb_oscr_workspace_bytes(N));
    cudaStream_t st = (cudaStream_t)stream;
    char* p = reinterpret_cast<char*>(ws);
    auto* totals = reinterpret_cast<unsigned long long*>(p);     // n_known, n_correct, n_unk, n_bad: zero between calls
    p += 256;
    void* curve_ws = p;
    p += curve_ws_bytes(N);
    const size_t rec_bytes = align_up((size_t)(N > 0 ? N : 1) * sizeof(uint64_t), 256);
    uint64_t* rec_unk = reinterpret_cast<uint64_t*>(p); p += rec_bytes;
    uint64_t* rec_known = reinterpret_cast<uint64_t*>(p); p += rec_bytes;
    ConfAcc conf = {};
    if (conf_sums) {
        conf.sums_out = conf_sums;
        conf.ticket = reinterpret_cast<unsigned int*>(p);
        conf.partials = reinterpret_cast<double*>(p + 128);
        conf.one_plus_offset = one_plus_offset;
        conf.valid_cols = valid_cols;
    }
    rc = run_records(scores, ld, gt, gt_type, N, C, unk_label, rec_unk, rec_known, totals, st, conf);
    if (rc != OSB_OK) return rc;
    return run_curve(rec_unk, reinterpret_cast<const int64_t*>(totals + 2), rec_known, reinterpret_cast<const int64_t*>(totals + 0), N,
                     taus, ccr_cnt, fpr_cnt, ccr, fpr, meta, totals, curve_ws, st);
}

extern "C" int osb_oscr(const float* scores, int64_t ld, const void* gt, int gt_type, int64_t N, int C,
                        int64_t unk_label, float* taus, int64_t* ccr_cnt, int64_t* fpr_cnt,
                        double* ccr, double* fpr, int64_t* meta, void* ws, size_t ws_bytes, void* stream) {
    OSB_ENTRY_GUARD();
    return oscr_impl("osb_oscr", scores, ld, gt, gt_type, N, C, unk_label, taus, ccr_cnt, fpr_cnt, ccr, fpr, meta, nullptr, 0.f, C,
                     ws, ws_bytes, stream);
}

extern "C" int osb_oscr_with_confidence(const float* scores, int64_t ld, const void* gt, int gt_type, int64_t N, int C,
                                        int64_t unk_label, float one_plus_offset, int valid_cols,
                                        float* taus, int64_t* ccr_cnt, int64_t* fpr_cnt, double* ccr, double* fpr,
                                        int64_t* meta, double* conf_sums, void* ws, size_t ws_bytes, void* stream) {
    OSB_ENTRY_GUARD();
    OSB_REQUIRE(conf_sums, OSB_ERR_ARG, "osb_oscr_with_confidence: null conf_sums");
    OSB_REQUIRE(unk_label < 0, OSB_ERR_ARG, "osb_oscr_with_confidence: unk_label must be negative (a row is then known or unk, never both)");
    OSB_REQUIRE(valid_cols >= 1 && valid_cols <= C, OSB_ERR_ARG, "osb_oscr_with_confidence: valid_cols=%d outside [1,%d]", valid_cols, C);
    return oscr_impl("osb_oscr_with_confidence", scores, ld, gt, gt_type, N, C, unk_label, taus, ccr_cnt, fpr_cnt, ccr, fpr, meta,
                     conf_sums, one_plus_offset, valid_cols, ws, ws_bytes, stream);
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
