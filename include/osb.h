/*
 * osb.h — C ABI of the B200-native open-set head + evaluation path ("osb" = open-set B200).
 *
 * One shared library (openset_imagenet_b200/csrc/libosb.so, sm_100a only) exports exactly the
 * functions declared here, plus the developer hooks of osb_debug.h (timing probes, A/B switches; nothing in the
 * product path or in bench.py calls those).  Plain pointers and sizes only; no C++/torch types.  The reference
 * (AIML-IfI/openset-imagenet) is pure Python and has no FFI of its own, so each entry cites the
 * reference call site (file:line, relative to the reference root) whose arithmetic it replaces;
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its comment says host;
 *   - the caller owns all memory, including workspaces (size from the matching *_workspace_bytes);
 *     the library never allocates or frees device memory and keeps no mutable state between calls
 *     (per-thread caches of encoded TMA descriptors and a cached device query are the only statics);  a workspace
 *     must be zero-filled ONCE by its owner before first use, every entry leaves it reusable, and one
 *     workspace must not be shared by calls that may run concurrently on different streams;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is
 *     enqueued on it, nothing synchronises, every entry is CUDA-graph capturable;
 *   - every function returns an int status: OSB_OK or a negative OSB_ERR_*;  no exception crosses
 *     the ABI;  osb_last_error() gives a thread-local message for the last failure;
 *   - there is NO CPU fallback: without a compute-capability-10.x device every compute entry
 *     returns OSB_ERR_ARCH.
 */
#ifndef OSB_H_
#define OSB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSB_ABI_VERSION 4

/* status codes */
#define OSB_OK             0
#define OSB_ERR_ARG       -1   /* null pointer, negative size, misaligned pointer, bad enum        */
#define OSB_ERR_SHAPE     -2   /* shape outside what the kernels support (message says which)      */
#define OSB_ERR_WORKSPACE -3   /* workspace too small                                              */
#define OSB_ERR_CUDA      -4   /* a CUDA runtime / driver call failed (message has the error name) */
#define OSB_ERR_ARCH      -5   /* no CUDA device, or device is not sm_100                           */

/* loss variants: losses.py:7-29 (entropic), train.py:343 (softmax), train.py:346-347 (garbage) */
#define OSB_ENTROPIC 0
#define OSB_SOFTMAX  1
#define OSB_GARBAGE  2

/* element types of logits / dlogits (osb_loss_from_logits) and of dx (osb_head_backward) */
#define OSB_F32  0
#define OSB_BF16 1

/* osb_head_forward flags */
#define OSB_FWD_DEN_FROM_PREV_KERNEL 1   /* see osb_head_forward */
#define OSB_FWD_PUBLISH_ROW_TILES 2      /* see osb_head_forward */
#define OSB_BWD_DZ_FROM_PREV_FORWARD 1   /* see osb_head_backward */

/* label types accepted by the OSCR row reduce (`gt` is float32 in the reference's npz files,
 * train.py:212, and cast with astype(int) at util.py:100) */
#define OSB_LABEL_I64 0
#define OSB_LABEL_F32 1
#define OSB_LABEL_I32 2

/* ---- library ------------------------------------------------------------------------------ */
int         osb_abi_version(void);
const char* osb_last_error(void);            /* host string, valid until the next failing call on this thread */
int         osb_device_check(void);          /* OSB_OK iff the current device is compute capability 10.x */
int         osb_sm_count(void);              /* SM count of the current device (148 on B200), <0 on error */

/* ---- labels -> denominator and counts ------------------------------------------------------ *
 * Replaces the host-synchronising label bookkeeping of losses.py:18-26 and the implicit reductions
 * inside torch.nn.CrossEntropyLoss (train.py:343,347).
 *   den[0]    (double)  loss denominator of THIS shard: B (entropic), #rows with y != ignore_index
 *                       (softmax), sum_i class_w[y_i] over non-ignored rows (garbage);
 *   counts[4] (int64)   {rows with a one-hot target, rows with the uniform negative target,
 *                        ignored rows, rows whose label is out of range}.
 * Out-of-range labels (reference: IndexError / device assert) never fault: they are counted in
 * counts[3], treated as ignored rows, and poison the loss with NaN in the consumers below.
 * max_ctas > 0 bounds the grid (a caller that runs the statistics beside a GEMM keeps them on a few SMs; the result
 * does not depend on the grid), 0 = default.
 */
int osb_label_stats(const int64_t* y, const float* class_w, int64_t B, int C, int variant,
                    int64_t ignore_index, double* den, int64_t* counts,
                    void* ws, size_t ws_bytes, int max_ctas, void* stream);   /* ws: osb_loss_workspace_bytes(B) */

/* ---- loss from materialised logits --------------------------------------------------------- *
 * Drop-in arithmetic of EntropicOpensetLoss.__call__ (losses.py:16-29) and of the two
 * CrossEntropyLoss configurations (train.py:343,347) when the caller already holds logits [B,C].
 *   logits   [B, ld] of `dtype`; y [B] int64; class_w [C] float32 (garbage only, else NULL)
 *   den      device double[1]: denominator to divide by (osb_label_stats output, or its all-reduced
 *            sum when the batch is sharded across GPUs)
 *   loss_num device double[1] out: sum of un-normalised row losses of this shard
 *   loss     device float[1]  out: loss_num / den  (NaN if den == 0 or any label was out of range)
 *   lse      device float[B]  out, nullable: row log-sum-exp (saved for backward)
 *   row_ptgt / row_pmax / row_pmax_all / row_pred : nullable per-row outputs
 *            softmax prob of the label column (0 when the label is not a column), max prob over the
 *            first `valid_cols` columns (metrics.py:39), max prob over all columns and first argmax
 *            over all columns (util.py:109-110).
 *   dlogits  nullable [B, ldd] of `dtype`: gscale * d loss / d logits; gscale is a device float[1]
 *            (upstream gradient, NULL = 1) so backward needs no host sync.
 */
size_t osb_loss_workspace_bytes(int64_t B);
int osb_loss_from_logits(const void* logits, int dtype, int64_t ld, const int64_t* y,
                         const float* class_w, int64_t B, int C, int variant, double unk_weight,
                         int64_t ignore_index, const double* den, int valid_cols,
                         double* loss_num, float* loss, float* lse,
                         float* row_ptgt, float* row_pmax, float* row_pmax_all, int32_t* row_pred,
                         void* dlogits, int64_t ldd, const float* gscale,
                         void* ws, size_t ws_bytes, void* stream);

/* ---- fused head: linear -> softmax statistics -> loss -> grad-logits -> dX, dW, db ---------- *
 * Replaces model.py:20-26,37-38 (resnet_base.fc and the logits linear), losses.py:16-29 / train.py:343,347 (loss),
 * train.py:177 (softmax for confidence/OSCR, reduced to per-row statistics) and the autograd
 * backward of all of them (train.py:138).  tcgen05/TMEM GEMMs fed by TMA; the logits live only in
 * TMEM/registers.
 *   x [B, K] bf16 row-major (ldx elements), w [C, K] bf16 row-major (nn.Linear layout), bias [C]
 *   float32 or NULL.  K % 64 == 0, C <= 1024.
 * forward : loss_num[0] (double) = sum of the un-normalised row losses; loss[0] (float) = loss_num / den[0], or
 *           (float)loss_num when den is NULL (a batch-sharded caller divides by the global denominator itself);
 *           row_* as in osb_loss_from_logits.
 *           dz (nullable) [dz_planes][B, lddz] bf16 = DENOMINATOR-FREE grad-logits  omega_i (p_i - t_i)
 *           (lddz % 8 == 0, lddz >= C, planes B*lddz elements apart): the forward kernel does not depend on the label
 *           statistics, and backward divides by den.  dz_planes = 1: one bf16 rounding of dZ (2^-9 per element,
 *           ~1e-3 norm-wise in dW / dX).  dz_planes = 2 ("exact"): plane 1 holds the bf16 rounding RESIDUAL of plane 0,
 *           and the backward GEMMs accumulate both planes in fp32 (2^-17 per element) at twice the MMA count -
 *           free for the launch-latency-bound protocol heads (C <= 256).
 *           flags & OSB_FWD_DEN_FROM_PREV_KERNEL: the caller promises that the kernel directly in front of this
 *           call on `stream` is the one that writes den (osb_label_stats, or osb_peer_sum after it) and that ITS
 *           predecessor has completed the inputs: small batches are then launched with programmatic stream
 *           serialization, run under that kernel, and wait for it only before den is read.
 *           flags & OSB_FWD_PUBLISH_ROW_TILES: the persistent forward (heads wider than 128 classes, batches the
 *           split-K kernel does not take) clears a word per (256-row tile, unit, epilogue warp) in the workspace when it
 *           starts and sets it when that warp's grad-logits of the tile have landed in global memory; it costs nothing
 *           and is the producer side of OSB_BWD_DZ_FROM_PREV_FORWARD below.
 *           max_ctas > 0 bounds the persistent grids (forward, dX) so that a collective running beside them finds
 *           free SMs; 0 = one CTA per SM.
 * backward: g = gscale[0] (device, NULL = 1) / den[0];  dx (nullable) [B, lddx] of dx_dtype (OSB_BF16 or OSB_F32)
 *           = g * dz @ w ; dw [C, K] float32 = g * dz^T @ x ; db (nullable) [C] float32 = g * column sums of dz.
 *           flags & OSB_BWD_DZ_FROM_PREV_FORWARD (dw without dx, and osb_head_backward_dw_partials): the caller promises
 *           that the kernel directly in front of this call on `stream` is osb_head_forward with
 *           OSB_FWD_PUBLISH_ROW_TILES on the same B, K, C, dz and workspace.  The dW GEMM then does not wait for that
 *           kernel to finish: its first CTA pairs become resident on the SMs the forward leaves idle (64 of 74 SM pairs
 *           work at 16384 x 2048 -> 1000) and consume the rows of dz tile by tile as they are published.  The result is
 *           the same bits; a caller that breaks the promise gets the plain dependency back after a 0.5 s time-out.
 *           Opt-in: at that shape the chip is power-limited under a full load, and the overlap measured 5 % SLOWER than
 *           the plain launch order; it is kept for shapes and power budgets where idle SMs are free.
 */
size_t osb_head_workspace_bytes(int64_t B, int K, int C);
int osb_head_forward(const void* x, int64_t ldx, const void* w, const float* bias, const int64_t* y,
                     const float* class_w, int64_t B, int K, int C, int variant, double unk_weight,
                     int64_t ignore_index, const double* den /* nullable */, int valid_cols,
                     double* loss_num, float* loss,
                     float* row_ptgt, float* row_pmax, float* row_pmax_all, int32_t* row_pred,
                     void* dz, int64_t lddz, int dz_planes, int flags, int max_ctas,
                     void* ws, size_t ws_bytes, void* stream);
int osb_head_backward(const void* dz, int64_t lddz, int dz_planes, const double* den, const void* x, int64_t ldx,
                      const void* w, int64_t B, int K, int C, const float* gscale,
                      void* dx, int64_t lddx, int dx_dtype, float* dw, float* db, int max_ctas, int flags,
                      void* ws, size_t ws_bytes, void* stream);
/* Forward and dX of a training step from ONE call: loss / row outputs / dz as osb_head_forward (no per-row predictions),
 * dx = gscale[0] / den[0] * dz @ w as osb_head_backward.  Where osb_head_forward_dx_is_fused(B, K, C) says 1 (heads wider
 * than 128 classes on the persistent forward) it is ONE kernel: the grid covers every SM, the CTA pairs the forward
 * leaves idle start with dX tiles at once (taking the rows of dz as the forward publishes them, 256-row tile by tile),
 * and every other pair goes on with dX tiles on the same pipeline state the moment its last forward MMA is issued, so
 * its first dX tiles run under the forward's last epilogue.  Elsewhere the two kernels run one after the other.  The
 * same bits either way.  den must not be NULL (it scales dx). */
int osb_head_forward_dx(const void* x, int64_t ldx, const void* w, const float* bias, const int64_t* y,
                        const float* class_w, int64_t B, int K, int C, int variant, double unk_weight,
                        int64_t ignore_index, const double* den, int valid_cols,
                        double* loss_num, float* loss, float* row_ptgt, float* row_pmax,
                        void* dz, int64_t lddz, int dz_planes, const float* gscale,
                        void* dx, int64_t lddx, int dx_dtype, int flags, int max_ctas,
                        void* ws, size_t ws_bytes, void* stream);
int osb_head_forward_dx_is_fused(int64_t B, int K, int C);
/* 1 when osb_head_forward runs this shape as the split-K cluster kernel (small batches, C <= 256), else 0 */
int osb_head_forward_is_splitk(int64_t B, int K, int C);
/* 1 when osb_head_backward writes dW straight from the GEMM (at most two batch splits: the first stores, the second adds
 * at the destination with a TMA reduce), 0 when split-K partials are summed by a separate fixed-order reduce kernel.
 * Both are deterministic; the answer only changes the launch count of a step. */
int osb_head_dw_is_direct(int64_t B, int K, int C);

/* ---- multi-GPU: all-reduce of dW over NVLink peer memory, fused with the split-K reduction ----- *
 * The reference has no collective at all (script/train_all.py:96-118 runs independent jobs); this is what a
 * batch-sharded head needs instead of `reduce locally, then NCCL`.
 * osb_head_backward_dw_partials runs only the dW GEMM: `osb_head_dw_splits(B,K,C)` fp32 partial matrices [C,K]
 * (UN-scaled: sums of dz * x) are left at byte offset osb_head_dw_partials_offset() of the workspace, consecutive
 * splits C*K elements apart; db_partial (nullable) [C] float32 receives the un-scaled column sums of dz.
 * local_reduce != 0 sums the splits into the first one on this GPU (pass splits = 1 to the all-reduce then): worth it
 * when there are many splits (small batches), so that ONE matrix per rank crosses NVLink.
 * osb_head_dw_allreduce (one launch per rank): rank r sums slice r of ALL ranks' partials in a fixed order
 * (bit-identical result on every rank), scales by gscale[0] / den[0] and stores it into every rank's dW; `n_extra`
 * floats per rank are summed into extra_out: the first n_extra - n_plain of them (db) are scaled like dW, the last
 * n_plain (loss numerators, counts) are plain sums.  den is the GLOBAL denominator (device double; constant for the
 * entropic loss, osb_peer_sum output otherwise).  The four pointer tables are HOST arrays of `world` DEVICE pointers
 * to peer-mapped memory (torch.distributed._symmetric_memory): partial buffers, dW buffers, extra vectors, flag areas
 * of osb_head_dw_allreduce_flag_bytes(world) bytes (zeroed once).  The epoch lives in the flag area (device memory)
 * and advances by one per launch, so the launch is CUDA-graph capturable and every replay is a new epoch; all ranks
 * must launch the same number of times.  A rank that waits ~2 s for a peer gives up: it stores nothing, sets every
 * plain extra to NaN and writes the epoch to the error word (osb_head_dw_allreduce_error reads it back, host sync).
 * mc_partials / mc_dw (both or neither): MULTICAST addresses (NVSwitch, cuMulticast / torch symmetric memory's
 * multicast_ptr) of the same partial and dW buffers.  With them the owner of a slice fetches the cross-rank sum with
 * multimem.ld_reduce (added inside the switch: 1/world of the traffic) and broadcasts the result with multimem.st; the
 * sum order is then the switch's (still one reduction per element, so all ranks stay bit-identical).
 * max_ctas bounds the grid (default 32) so the kernel overlaps dX. */
size_t osb_head_dw_partials_offset(void);
int osb_head_dw_splits(int64_t B, int K, int C);
int osb_head_backward_dw_partials(const void* dz, int64_t lddz, int dz_planes, const void* x, int64_t ldx,
                                  int64_t B, int K, int C, float* db_partial, int local_reduce, int flags,
                                  void* ws, size_t ws_bytes, void* stream);
size_t osb_head_dw_allreduce_flag_bytes(int world);
int osb_head_dw_allreduce(const void* const* peer_partials, void* const* peer_dw, const void* const* peer_extra,
                          void* const* peer_flags, int rank, int world, int splits, int64_t n, int64_t split_stride,
                          float* extra_out, int n_extra, int n_plain, const float* gscale, const double* den,
                          const void* mc_partials /* nullable */, void* mc_dw /* nullable */, int max_ctas, void* stream);
int osb_head_dw_allreduce_error(const void* flags_local, int world, uint32_t* epoch_out);   /* host sync; 0 = no failure */

/* Sum of up to 8 doubles over the ranks through peer memory (the loss denominator of the softmax / garbage variants,
 * train.py:343,347, for a batch-sharded step): every rank stores its values as (32 bits, epoch) words into every
 * peer's slot area and sums what arrives in rank order (bit-identical everywhere).  One CTA, ~5 us, graph capturable
 * (the epoch lives in the slot area: osb_peer_sum_bytes(world) bytes per rank, peer-mapped, zeroed once).
 * peer_slots: HOST array of `world` DEVICE pointers.  Times out like the all-reduce (out = NaN). */
size_t osb_peer_sum_bytes(int world);
int osb_peer_sum(const double* in, double* out, int n, void* const* peer_slots, int rank, int world, void* stream);

/* ---- confidence (metrics.py:8-42) ---------------------------------------------------------- *
 *   scores [N, ld] float32, y [N] int64.
 *   sums[2]   (double) {sum over known rows of s[i,y_i],
 *                       sum over rows with y == unknown_class of float32(1+offset) - max_c s[i,:valid_cols]}
 *   counts[3] (int64)  {kn_count, neg_count, known rows whose label is >= C (reference: IndexError)}
 * Counts are exact; sums are accumulated in double in a fixed order (bit-reproducible run to run).
 * The division and the empty-group rule (metrics.py:29-34) are left to the caller.
 */
size_t osb_confidence_workspace_bytes(int64_t N);
int osb_confidence(const float* scores, int64_t ld, const int64_t* y, int64_t N, int C,
                   float one_plus_offset, int64_t unknown_class, int valid_cols,
                   double* sums, int64_t* counts, void* ws, size_t ws_bytes, void* stream);

/* Streaming variant for the validation loop (train.py:172-192): the rows of one batch already reduced by the
 * fused head to p_target[i] = s[i, y_i] and p_max[i] = max_c s[i, :valid_cols] (osb_head_forward row_ptgt /
 * row_pmax), so no [N, C] score tensor exists.  accumulate != 0 adds this batch to sums[2] / counts[3]
 * (same meaning as above), accumulate == 0 overwrites them.  One block, fixed summation order. */
int osb_confidence_from_stats(const float* p_target, const float* p_max, const int64_t* y, int64_t N, int C,
                              float one_plus_offset, int64_t unknown_class, int accumulate,
                              double* sums, int64_t* counts, void* stream);

/* ---- OSCR (util.py:90-122) ----------------------------------------------------------------- *
 * RECORD = (uint64) key << 32 | type.  key = order-preserving uint32 image of a float32 score
 * (-0.0 == +0.0); type 0 = known row, arg-max != label; 1 = known row, arg-max == label (util.py:114);
 * 2 = row labelled unk_label (util.py:102); 3 = padding, ignored.
 * Stage A  osb_oscr_records: the one pass over scores [N, ld] float32 (util.py:100-111).  Appends one
 *   record per known row (key = target score) to rec_known and one per unk row (key = row max) to
 *   rec_unk, both compacted, in no particular order, capacity N each.
 *   totals[4] int64 (zeroed by the call) {n_known, n_correct, n_unk, n_bad (known label >= C)};
 *   totals[0] / totals[2] are the record counts of rec_known / rec_unk.
 *   osb_oscr_records_from_stats is the compact-input variant: rows already reduced by the fused head to
 *   (p_target, p_max_all, pred) (osb_head_forward row_* outputs).
 * Stage B  osb_oscr_curve: stable LSD radix sort of rec_a[0..n_a) ++ rec_b[0..n_b) by key and one scan.
 *   Every run of equal keys holding a known record is a threshold (np.unique, util.py:113): taus[j]
 *   float32 ascending, ccr_cnt[j] / fpr_cnt[j] int64 = records of type 1 / type 2 with key > tau
 *   (strict, util.py:114,117), optional ccr[j] / fpr[j] = counts / n_known / n_unk as float64.
 *   Pass the unk records as rec_a and the known records as rec_b (ties must list unk records first).
 *   n_a, n_b: DEVICE int64 scalars; cap >= n_a + n_b sizes the workspace and the outputs.
 *   meta[5] int64 {n_unique, n_known, n_correct, n_unk, 0}.  The reference drops the largest threshold
 *   (util.py:113): n_thresholds = max(n_unique - 1, 0), left to the caller.
 *   Counts at a fixed threshold are additive across row shards: a shard passes its own unk records plus
 *   ALL shards' known records (remote ones demoted to type 0), then sums ccr_cnt / fpr_cnt over shards.
 *   ONE cooperative launch (all CTAs co-resident; grid barriers through L2).
 * osb_oscr = A + B on one GPU (two launches); meta[4] = n_bad.  taus / ccr_cnt / fpr_cnt may be NULL when only the
 *   reference's own outputs (ccr, fpr) are wanted.  osb_oscr_ratios: counts -> float64 ratios after a cross-shard sum.
 */
size_t osb_oscr_workspace_bytes(int64_t N);
size_t osb_oscr_curve_workspace_bytes(int64_t cap);
int osb_oscr_records(const float* scores, int64_t ld, const void* gt, int gt_type, int64_t N, int C,
                     int64_t unk_label, uint64_t* rec_unk, uint64_t* rec_known, int64_t* totals, void* stream);
int osb_oscr_records_from_stats(const float* p_target, const float* p_max_all, const int32_t* pred,
                                const void* gt, int gt_type, int64_t N, int C, int64_t unk_label,
                                uint64_t* rec_unk, uint64_t* rec_known, int64_t* totals, void* stream);
int osb_oscr_curve(const uint64_t* rec_a, const int64_t* n_a, const uint64_t* rec_b, const int64_t* n_b, int64_t cap,
                   float* taus, int64_t* ccr_cnt, int64_t* fpr_cnt, double* ccr /* nullable */, double* fpr /* nullable */,
                   int64_t* meta, void* ws, size_t ws_bytes, void* stream);
int osb_oscr_ratios(const int64_t* ccr_cnt, const int64_t* fpr_cnt, const int64_t* n_unique,
                    const int64_t* n_known, const int64_t* n_unk, int64_t n,
                    double* ccr, double* fpr, void* stream);
int osb_oscr(const float* scores, int64_t ld, const void* gt, int gt_type, int64_t N, int C,
             int64_t unk_label, float* taus, int64_t* ccr_cnt, int64_t* fpr_cnt,
             double* ccr, double* fpr, int64_t* meta /* [5]: n_unique, n_known, n_correct, n_unk, n_bad */,
             void* ws, size_t ws_bytes, void* stream);

/* OSCR and metrics.confidence of the same evaluation set in ONE pass over the scores (script/plot_all.py:363-372 calls
 * util.calculate_oscr and metrics.confidence on the same `values` with unknown_class = unk_label): osb_oscr plus
 *   conf_sums[2] (double) {sum over known rows of s[i, y_i],
 *                          sum over unk_label rows of float32(1 + offset) - max_c s[i, :valid_cols]}
 * exactly as osb_confidence defines them (fixed summation order, bit-reproducible); the counts of metrics.confidence are
 * meta[1] (kn_count) and meta[3] (neg_count), out-of-range labels meta[4].  unk_label must be negative. */
int osb_oscr_with_confidence(const float* scores, int64_t ld, const void* gt, int gt_type, int64_t N, int C,
                             int64_t unk_label, float one_plus_offset, int valid_cols,
                             float* taus, int64_t* ccr_cnt, int64_t* fpr_cnt, double* ccr, double* fpr,
                             int64_t* meta, double* conf_sums, void* ws, size_t ws_bytes, void* stream);

/* ---- row-sharded OSCR over the GPUs of one box ---------------------------------------------------------------- *
 * The reference evaluates on one device (script/plot_all.py:151-188); this is what a row-sharded evaluation set needs.
 * osb_oscr_exchange (one launch per rank, after osb_oscr_records on the rank's rows): all-gather of the compacted
 * records through NVLink peer memory.  Every rank's records land, in rank order, in EVERY rank's gathered arrays
 * (peer_rec_unk[r] / peer_rec_known[r]: capacity `cap` records each, peer-mapped), and totals_global[4] (local, int64)
 * receives the four stage-A totals summed over the ranks: pass totals_global + 2 / + 0 as n_a / n_b of osb_oscr_curve
 * and totals_global[3] as n_bad.  Every rank then sorts identical inputs, so the curve is replicated and bit-identical
 * to the single-GPU one.  peer_slots[r]: osb_oscr_exchange_slot_bytes(world) bytes per rank, peer-mapped, zeroed once
 * (count words, done stamps and the epoch, which advances per launch: graph capturable).  A rank that waits ~2 s for a
 * peer gives up: totals_global = {0, 0, 0, 2^40} (an impossible n_bad the caller must check).  HOST pointer tables. */
size_t osb_oscr_exchange_slot_bytes(int world);
int osb_oscr_exchange(const uint64_t* rec_unk, const uint64_t* rec_known, const int64_t* totals,
                      void* const* peer_rec_unk, void* const* peer_rec_known, void* const* peer_slots,
                      int rank, int world, int64_t cap, int64_t* totals_global, int max_ctas, void* stream);

/* ---- CCR@FPR (script/plot_all.py:346-352,378-384) ------------------------------------------ *
 * For each of nq <= 16 queries q (host doubles): idx = first argmin_j |fpr[j]-q| over the
 * n_unique[0]-1 thresholds the reference keeps (NaN counts as the minimum, like numpy.argmin).
 * out_idx[nq] int64 (-1 when there are no thresholds), out_fpr[nq], out_ccr[nq] double (device).
 * The 10 % acceptance rule is a host-side round() on out_fpr and stays in the Python wrapper.
 */
int osb_ccr_at_fpr(const double* ccr, const double* fpr, const int64_t* n_unique, int64_t n,
                   const double* queries_host, int nq, int64_t* out_idx, double* out_fpr,
                   double* out_ccr, void* stream);

/* ---- score / feature-norm histograms (util.py:202-228 get_histogram) ------------------------- *
 * osb_hist_metrics: the per-row values the reference histograms.  Pass `scores` [N, ld] fp32 (metric 'score':
 * kn_metric[i] = scores[i, gt[i]] for gt[i] >= 0, unk_metric[i] = max scores[i, :valid_cols] for gt[i] == unk_label;
 * valid_cols = C - 1 reproduces drop_bg) OR `features` [N, ldf] fp32 with K columns (metric 'norm': the row's L2 norm,
 * accumulated in fp64 in a fixed order, for both groups); the other pointer is NULL.  Rows outside a group get NaN.
 * counts[2] int64 = rows in each group; minmax[4] fp32 = kn min, kn max, unk min, unk max (np.histogram takes its bin
 * range from them; untouched sentinels when a group is empty -- check counts).
 * osb_histogram: counts[n_edges-1] int64 of `values` in the bins given by `edges` (device doubles, ascending), with
 * np.histogram's rules: edges[j] <= x < edges[j+1], last bin closed on the right, NaN / out-of-range dropped.
 */
int osb_hist_metrics(const float* scores, int64_t ld, const float* features, int64_t ldf, int K,
                     const int64_t* gt, int64_t N, int C, int valid_cols, int64_t unk_label,
                     float* kn_metric, float* unk_metric, int64_t* counts, float* minmax, void* stream);
int osb_histogram(const float* values, int64_t n, const double* edges, int n_edges, int64_t* counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OSB_H_ */
