"""CPU oracle for the open-set head + evaluation path.  TEST INFRASTRUCTURE ONLY.

This module restates, in numpy / float64, the arithmetic of the reference's hot path
(AIML-IfI/openset-imagenet).  It is imported ONLY by ``tests/``, by
``__graft_entry__.smoke()`` and by ``bench.py``'s CPU-baseline / ``--impl reference`` legs.
Nothing under ``openset_imagenet_b200/`` may import it: the product path is CUDA-only and
fails loudly when the extension is missing.

Parity pinning: the reference ships no tests and no golden vectors, so
this oracle is pinned by *executing the reference's own files* (``oracle/ref_loader.py``)
in the build container on seeded inputs; the resulting vectors are committed under
``tests/golden/`` by ``oracle/make_golden.py`` and re-checked on every CPU test run
(``tests/test_oracle_golden.py``), plus the hand-checkable known-answer vectors of
SURVEY.md §9.6.

Every function cites the reference lines it follows (paths relative to the reference root).
"""
from __future__ import annotations

import math

import numpy as np

ENTROPIC, SOFTMAX, GARBAGE = 0, 1, 2
VARIANTS = {"entropic": ENTROPIC, "softmax": SOFTMAX, "garbage": GARBAGE}


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def _logsumexp_rows(z: np.ndarray) -> np.ndarray:
    m = z.max(axis=1, keepdims=True)
    return (m + np.log(np.exp(z - m).sum(axis=1, keepdims=True)))[:, 0]


def bf16_round(a: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even float32 -> bfloat16 -> float32 (numpy has no bf16)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return rounded.astype(np.uint32).view(np.float32).reshape(a.shape)


# --------------------------------------------------------------------------------------
# losses  (openset_imagenet/losses.py:7-29, openset_imagenet/train.py:338-347)
# --------------------------------------------------------------------------------------
def row_terms(variant: int, y: np.ndarray, C: int, unk_weight: float = 1.0,
              class_weights: np.ndarray | None = None, ignore_index: int | None = None):
    """Per-row (mode, omega) and the loss denominator.

    mode 0: ignored row; 1: one-hot target y with row weight omega; 2: uniform target
    unk_weight/C on every class (entropic negatives, losses.py:12-13,24-28).
    Denominators: entropic -> B (CrossEntropyLoss() mean over all rows of a prob-target,
    losses.py:14,29); softmax -> #non-ignored (train.py:343); garbage -> sum of w[y]
    (train.py:347).
    """
    y = np.asarray(y, dtype=np.int64)
    B = y.shape[0]
    mode = np.zeros(B, dtype=np.int64)
    omega = np.zeros(B, dtype=np.float64)
    if variant == ENTROPIC:
        kn = y >= 0                                   # losses.py:18-19
        if np.any(y[kn] >= C):
            raise IndexError("label out of range")    # eye[target] raises in the reference
        mode[kn] = 1
        omega[kn] = 1.0
        mode[~kn] = 2
        # the reference's `ones` buffer is float32: every class gets float32(w/C) (losses.py:12-13)
        omega[~kn] = float(np.float32(unk_weight / C)) * C
        den = float(B)
    elif variant == SOFTMAX:
        ii = -1 if ignore_index is None else ignore_index
        valid = y != ii
        if np.any((y[valid] < 0) | (y[valid] >= C)):
            raise IndexError("label out of range")
        mode[valid] = 1
        omega[valid] = 1.0
        den = float(valid.sum())
    elif variant == GARBAGE:
        ii = -100 if ignore_index is None else ignore_index
        valid = y != ii
        if np.any((y[valid] < 0) | (y[valid] >= C)):
            raise IndexError("label out of range")
        w = np.asarray(class_weights, dtype=np.float32).astype(np.float64)
        mode[valid] = 1
        omega[valid] = w[y[valid]]
        den = float(omega.sum())
    else:
        raise ValueError(variant)
    return mode, omega, den


def loss_from_logits(logits, y, variant: int, unk_weight: float = 1.0, class_weights=None,
                     ignore_index: int | None = None, den: float | None = None):
    """Closed-form float64 loss and grad-logits for the three variants.

    entropic (losses.py:16-29):  T[i,:] = onehot(y_i) if y_i>=0 else unk_weight/C;
        loss = -(1/B) sum_i sum_c T[i,c] logsoftmax(z)[i,c];  dz = (p*sum_c T - T)/B.
    softmax  (train.py:343):     mean NLL over rows with y != -1 (NaN when none).
    garbage  (train.py:347):     sum_i w[y_i] NLL_i / sum_i w[y_i].
    ``den`` overrides the denominator (used by batch-sharded multi-GPU parity tests).
    Returns (loss, dlogits[B,C], loss_numerator, denominator).
    """
    z = np.asarray(logits, dtype=np.float64)
    B, C = z.shape
    mode, omega, den_local = row_terms(variant, y, C, unk_weight, class_weights, ignore_index)
    den = den_local if den is None else float(den)
    lse = _logsumexp_rows(z)
    logp = z - lse[:, None]
    p = np.exp(logp)
    y = np.asarray(y, dtype=np.int64)
    rowloss = np.zeros(B, dtype=np.float64)
    T = np.zeros_like(z)
    k1 = mode == 1
    k2 = mode == 2
    rows = np.nonzero(k1)[0]
    rowloss[rows] = -omega[rows] * logp[rows, y[rows]]
    T[rows, y[rows]] = omega[rows]
    rowloss[k2] = -(omega[k2] / C) * logp[k2].sum(axis=1)
    T[k2, :] = (omega[k2] / C)[:, None]
    S = omega.copy()
    num = float(rowloss.sum())
    with np.errstate(divide="ignore", invalid="ignore"):
        loss = np.float64(num) / np.float64(den)
        dz = (p * S[:, None] - T) / np.float64(den)
    dz[mode == 0] = 0.0 if den != 0 else dz[mode == 0]
    return float(loss), dz, num, den


def row_stats(logits, y, valid_cols: int | None = None):
    """Per-row statistics the fused head emits instead of a scores tensor:
    p_target = softmax(z)[i, y_i] (0 where y_i is not a column), p_max = max_c softmax(z)[i, :valid_cols],
    pred = first argmax over ALL columns (numpy argmax semantics, util.py:109).
    """
    z = np.asarray(logits, dtype=np.float64)
    B, C = z.shape
    vc = C if valid_cols is None else valid_cols
    lse = _logsumexp_rows(z)
    p = np.exp(z - lse[:, None])
    y = np.asarray(y, dtype=np.int64)
    ok = (y >= 0) & (y < C)
    pt = np.zeros(B)
    pt[ok] = p[ok, y[ok]]
    pm = p[:, :vc].max(axis=1)
    pred = np.argmax(z, axis=1)
    return pt, pm, pred


def head_fwd_bwd(x, W, b, y, variant: int, unk_weight: float = 1.0, class_weights=None,
                 ignore_index: int | None = None, den: float | None = None):
    """float64 oracle of linear -> loss -> backward (model.py:23-26,37-38 + loss above).

    x [B,K], W [C,K] (nn.Linear layout), b [C] or None.  Inputs are used as given (round them
    to bf16 first with ``bf16_round`` to model the bf16-in / fp32-accumulate contract).
    Returns dict(loss, logits, dz, dx, dw, db, num, den).
    """
    x = np.asarray(x, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    z = x @ W.T
    if b is not None:
        z = z + np.asarray(b, dtype=np.float64)[None, :]
    loss, dz, num, den = loss_from_logits(z, y, variant, unk_weight, class_weights, ignore_index, den)
    return dict(loss=loss, logits=z, dz=dz, dx=dz @ W, dw=dz.T @ x, db=dz.sum(axis=0), num=num, den=den)


def class_weights_from_labels(labels: np.ndarray, label_count: int) -> np.ndarray:
    """w[c] = len(train) / (count_c * label_count), float32 (dataset.py:77-86)."""
    labels = np.asarray(labels, dtype=np.int64)
    counts = np.bincount(labels, minlength=label_count).astype(np.float64)
    with np.errstate(divide="ignore"):
        w = len(labels) / (counts * label_count)
    return w.astype(np.float32)


# --------------------------------------------------------------------------------------
# confidence  (openset_imagenet/metrics.py:8-42)
# --------------------------------------------------------------------------------------
def resolve_valid_cols(C: int, last_valid_class) -> int:
    """Number of leading columns selected by the Python slice ``[:last_valid_class]`` (metrics.py:39)."""
    return len(range(C)[slice(None, last_valid_class)])


def confidence(scores, target_labels, offset=0.0, unknown_class=-1, last_valid_class=None):
    """Counts exact; per-sample statistics in float32 exactly as torch computes them
    (gather -> s[i,y_i];  float32(1+offset) - max_c s[i,:last_valid]); sums in float64, row order.

    metrics.py:25-26 group definition, :33 known mean, :36-40 negative mean, 0.0 for empty groups.
    Returns (kn_conf, kn_count, neg_conf, neg_count, kn_sum, neg_sum).
    """
    s = np.asarray(scores, dtype=np.float32)
    y = np.asarray(target_labels, dtype=np.int64)
    N, C = s.shape
    unknown = y == unknown_class
    known = (y >= 0) & ~unknown
    kn_count = int(known.sum())
    neg_count = int(unknown.sum())
    if np.any(y[known] >= C):
        raise IndexError("label out of range")
    kn_sum = float(s[known, y[known]].astype(np.float64).sum()) if kn_count else 0.0
    vc = resolve_valid_cols(C, last_valid_class)
    if neg_count:
        one_off = np.float32(1.0 + offset)
        stat = (one_off - s[unknown, :vc].max(axis=1)).astype(np.float32)
        neg_sum = float(stat.astype(np.float64).sum())
    else:
        neg_sum = 0.0
    kn_conf = kn_sum / kn_count if kn_count else 0.0
    neg_conf = neg_sum / neg_count if neg_count else 0.0
    return kn_conf, kn_count, neg_conf, neg_count, kn_sum, neg_sum


# --------------------------------------------------------------------------------------
# OSCR  (openset_imagenet/util.py:90-122)
# --------------------------------------------------------------------------------------
def oscr_counts_quadratic(gt, scores, unk_label=-1):
    """The reference's own O(U*N) algorithm, returning the integer counts as well.

    util.py:100-102 masks, :109-111 per-sample reductions, :113-118 one pass per unique
    known target score except the largest, strict '>'.
    """
    gt = np.asarray(gt).astype(int)
    scores = np.asarray(scores)
    known = gt >= 0
    unk = gt == unk_label
    n_kn, n_unk = int(known.sum()), int(unk.sum())
    pred = scores.argmax(axis=1)
    top = scores.max(axis=1)
    ts = scores[known, gt[known]]
    hit = pred[known] == gt[known]
    taus = np.unique(ts)[:-1]
    ccr_cnt = np.empty(len(taus), dtype=np.int64)
    fpr_cnt = np.empty(len(taus), dtype=np.int64)
    for j, tau in enumerate(taus):
        ccr_cnt[j] = np.count_nonzero(hit & (ts > tau))
        fpr_cnt[j] = np.count_nonzero(unk & (top > tau))
    return taus, ccr_cnt, fpr_cnt, n_kn, n_unk


def oscr_counts(gt, scores, unk_label=-1):
    """O(N log N) restatement (sort + searchsorted), bit-equal to ``oscr_counts_quadratic``
    (equality is asserted in tests/test_oracle_golden.py)."""
    gt = np.asarray(gt).astype(int)
    scores = np.asarray(scores)
    known = gt >= 0
    unk = gt == unk_label
    n_kn, n_unk = int(known.sum()), int(unk.sum())
    pred = scores.argmax(axis=1)
    top = scores.max(axis=1)
    ts = scores[known, gt[known]]
    hit = pred[known] == gt[known]
    taus = np.unique(ts)[:-1]
    cts = np.sort(ts[hit])
    ums = np.sort(top[unk])
    ccr_cnt = (len(cts) - np.searchsorted(cts, taus, side="right")).astype(np.int64)
    fpr_cnt = (len(ums) - np.searchsorted(ums, taus, side="right")).astype(np.int64)
    return taus, ccr_cnt, fpr_cnt, n_kn, n_unk


def float_keys(v):
    """Order-preserving uint32 image of float32 values (-0.0 == +0.0), as the kernels build it."""
    b = np.ascontiguousarray(v, dtype=np.float32).view(np.uint32).copy()
    b[b == np.uint32(0x80000000)] = 0
    neg = (b & np.uint32(0x80000000)) != 0
    return np.where(neg, ~b, b | np.uint32(0x80000000)).astype(np.uint32)


def oscr_records(gt, scores, unk_label=-1):
    """The two record arrays stage A of the CUDA path emits (include/osb.h): (rec_unk, rec_known) uint64,
    record = key << 32 | type, type 0 known & wrong, 1 known & correct (util.py:114), 2 unk (util.py:102)."""
    gt = np.asarray(gt).astype(int)
    scores = np.asarray(scores)
    known, unk = gt >= 0, gt == unk_label
    pred, top = scores.argmax(axis=1), scores.max(axis=1)
    hit = (pred[known] == gt[known]).astype(np.uint64)
    rec_known = (float_keys(scores[known, gt[known]]).astype(np.uint64) << np.uint64(32)) | hit
    rec_unk = (float_keys(top[unk]).astype(np.uint64) << np.uint64(32)) | np.uint64(2)
    return rec_unk, rec_known


def oscr_curve_from_records(rec_a, rec_b):
    """numpy stand-in for osb_oscr_curve: thresholds (ALL unique known keys, as float32) and the counts of type-1 /
    type-2 records with key > threshold.  Type-3 records are padding."""
    rec = np.concatenate([np.asarray(rec_a, dtype=np.uint64), np.asarray(rec_b, dtype=np.uint64)])
    key, typ = (rec >> np.uint64(32)).astype(np.uint32), (rec & np.uint64(3)).astype(np.int64)
    tau_keys = np.unique(key[typ < 2])
    cts, ums = np.sort(key[typ == 1]), np.sort(key[typ == 2])
    ccr_cnt = (len(cts) - np.searchsorted(cts, tau_keys, side="right")).astype(np.int64)
    fpr_cnt = (len(ums) - np.searchsorted(ums, tau_keys, side="right")).astype(np.int64)
    b = np.where(tau_keys & np.uint32(0x80000000), tau_keys & np.uint32(0x7FFFFFFF), ~tau_keys).astype(np.uint32)
    return b.view(np.float32), ccr_cnt, fpr_cnt


def calculate_oscr(gt, scores, unk_label=-1, quadratic=False):
    """(ccr, fpr) float64 arrays exactly as util.py:90-122 returns them."""
    fn = oscr_counts_quadratic if quadratic else oscr_counts
    _, ccr_cnt, fpr_cnt, n_kn, n_unk = fn(gt, scores, unk_label)
    with np.errstate(divide="ignore", invalid="ignore"):
        ccr = ccr_cnt / np.int64(n_kn)
        fpr = fpr_cnt / np.int64(n_unk)
    return np.asarray(ccr, dtype=np.float64), np.asarray(fpr, dtype=np.float64)


# --------------------------------------------------------------------------------------
# CCR@FPR  (openset_imagenet/script/plot_all.py:346-352,378-384)
# --------------------------------------------------------------------------------------
def ccr_at_fpr(ccr, fpr, queries=(1e-3, 1e-2, 0.1, 1.0)):
    """For each q: idx = first argmin |fpr-q|; value = ccr[idx] unless
    round(100*|fpr[idx]-q|/q, 1) >= 10.0 (then None, printed as '---' by the reference)."""
    ccr = np.asarray(ccr, dtype=np.float64)
    fpr = np.asarray(fpr, dtype=np.float64)
    out = []
    for q in queries:
        if len(fpr) == 0:
            raise ValueError("attempt to get argmin of an empty sequence")
        idx = int(np.abs(fpr - q).argmin())
        err = round(100 * abs(fpr[idx] - q) / q, 1)
        out.append((idx, float(fpr[idx]), None if err >= 10.0 else float(ccr[idx])))
    return out


# --------------------------------------------------------------------------------------
# synthetic workloads of the BASELINE.json shapes (shared by tests and bench; numpy only)
# --------------------------------------------------------------------------------------
def get_histogram(array, unk_label=-1, metric='score', bins=100, drop_bg=False, log_space=False,
                  geomspace_limits=(1, 1e2)):
    """util.py:202-228: histograms of the known rows' target score and the unk_label rows' maximum score (metric
    'score'), or of the feature norms of both groups (metric 'norm').  Returns (kn_hist, kn_edges, unk_hist, unk_edges)."""
    score = np.asarray(array['scores'])
    if drop_bg:                                            # util.py:210-211
        score = score[:, :-1]
    gt = np.asarray(array['gt']).astype(np.int64)          # util.py:212
    norms = np.linalg.norm(np.asarray(array['features']), axis=1)      # util.py:213-214
    kn = gt >= 0
    unk = gt == unk_label
    if metric == 'score':                                  # util.py:217-219
        kn_metric = score[kn, gt[kn]]
        unk_metric = np.amax(score[unk], axis=1)
    elif metric == 'norm':                                 # util.py:220-222
        kn_metric = norms[kn]
        unk_metric = norms[unk]
    if log_space:                                          # util.py:223-225
        bins = np.geomspace(geomspace_limits[0], geomspace_limits[1], num=bins)
    kn_hist, kn_edges = np.histogram(kn_metric, bins=bins)
    unk_hist, unk_edges = np.histogram(unk_metric, bins=bins)
    return kn_hist, kn_edges, unk_hist, unk_edges


def make_labels(B: int, C: int, neg_frac: float = 1.0 / 3.0, seed: int = 0, neg_label: int = -1):
    rng = np.random.default_rng(seed)
    y = rng.integers(0, C, size=B, dtype=np.int64)
    y[rng.random(B) < neg_frac] = neg_label
    return y


def make_oscr_case(N: int, C: int, seed: int = 0):
    """cfg 4 of BASELINE.json as specified in SURVEY.md §8d: 50 % known / 25 % -1 / 25 % -2,
    logits 2*N(0,1) with +4 on the target column of knowns, scores = float32 softmax."""
    rng = np.random.default_rng(seed)
    kind = rng.random(N)
    gt = rng.integers(0, C, size=N).astype(np.int64)
    gt[kind >= 0.5] = -1
    gt[kind >= 0.75] = -2
    logits = (2.0 * rng.standard_normal((N, C), dtype=np.float32)).astype(np.float32)
    kn = np.nonzero(gt >= 0)[0]
    logits[kn, gt[kn]] += np.float32(4.0)
    m = logits.max(axis=1, keepdims=True)
    e = np.exp(logits - m, dtype=np.float32)
    scores = (e / e.sum(axis=1, keepdims=True, dtype=np.float32)).astype(np.float32)
    return gt, scores


def log_c(C: int) -> float:
    return math.log(C)
