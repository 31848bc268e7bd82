"""OSCR evaluation of the reference (openset_imagenet/util.py:90-122) and the CCR@FPR lookup of
``plot_imagenet.py`` (script/plot_all.py:346-352,378-384) on the GPU.

``calculate_oscr(gt, scores, unk_label)`` keeps the reference's signature and return type (two
float64 numpy arrays, one entry per unique known target score except the largest) but runs as a
streaming row reduce + radix sort + scan instead of the reference's O(U*N) Python loop; the integer
counts behind every entry are bit-exact.  As in the reference, dropping a garbage-class column is the
caller's job (util.py:186-187 drops it, plot_all.py:359-363 does not).
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib as L

_LABEL_TYPES = {torch.int64: L.LABEL_I64, torch.float32: L.LABEL_F32, torch.int32: L.LABEL_I32}


@dataclass
class OscrDeviceResult:
    """Everything ``calculate_oscr`` computes, still on the device (no host sync has happened).

    Arrays have capacity N; the first ``meta[0] - 1`` entries are the reference's output (the largest
    unique threshold is dropped, util.py:113).  meta = [n_unique, n_known, n_correct, n_unk, n_bad].
    """
    taus: torch.Tensor
    ccr_cnt: torch.Tensor
    fpr_cnt: torch.Tensor
    ccr: torch.Tensor
    fpr: torch.Tensor
    meta: torch.Tensor


def _to_device(a, device, dtype=None) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.to(device, non_blocking=True)


def default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("openset_imagenet_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def calculate_oscr_device(gt, scores, unk_label=-1, device=None) -> OscrDeviceResult:
    """Enqueue the whole OSCR computation on the current stream; host arrays are copied in first."""
    if isinstance(scores, torch.Tensor) and scores.is_cuda:
        dev = scores.device
    else:
        dev = torch.device(device) if device is not None else default_device()
    s = _to_device(scores, dev, torch.float32)
    if s.dim() != 2:
        raise ValueError(f"scores must be [samples, classes], got {tuple(s.shape)}")
    if s.stride(1) != 1 and s.shape[1] > 1:
        s = s.contiguous()
    g = _to_device(gt, dev)
    if g.dtype not in _LABEL_TYPES:
        # anything else goes through the same truncating cast as gt.astype(int) (util.py:100)
        g = g.to(torch.int64)
    g = g.contiguous()
    N, Cc = s.shape
    if g.dim() != 1 or g.shape[0] != N:
        raise ValueError(f"gt must be [samples]={N}, got {tuple(g.shape)}")
    with torch.cuda.device(dev):
        cap = max(N, 1)
        taus = torch.empty(cap, dtype=torch.float32, device=dev)
        ccr_cnt = torch.empty(cap, dtype=torch.int64, device=dev)
        fpr_cnt = torch.empty(cap, dtype=torch.int64, device=dev)
        ccr = torch.empty(cap, dtype=torch.float64, device=dev)
        fpr = torch.empty(cap, dtype=torch.float64, device=dev)
        meta = torch.zeros(5, dtype=torch.int64, device=dev)
        ws = L.workspace("oscr", L.size("osb_oscr_workspace_bytes", N), dev)
        L.call("osb_oscr", L.ptr(s), s.stride(0) if N else Cc, L.ptr(g), _LABEL_TYPES[g.dtype], N, Cc, int(unk_label),
               L.ptr(taus), L.ptr(ccr_cnt), L.ptr(fpr_cnt), L.ptr(ccr), L.ptr(fpr), L.ptr(meta),
               L.ptr(ws), ws.numel(), L.stream_ptr(dev))
    return OscrDeviceResult(taus, ccr_cnt, fpr_cnt, ccr, fpr, meta)


def calculate_oscr_from_stats_device(gt, p_target, p_max_all, pred, num_classes, unk_label=-1) -> OscrDeviceResult:
    """Compact-input OSCR: every row already reduced (by the fused head, ``HeadStats``) to the softmax
    probability of its label, its row maximum and its arg-max, so no [N, C] score matrix exists.  Same
    result as ``calculate_oscr_device`` on the scores those statistics were taken from."""
    dev = L.require_cuda(p_target, p_max_all, pred, gt)
    g = gt if gt.dtype in _LABEL_TYPES else gt.to(torch.int64)
    g = g.contiguous()
    pt, pm, pr = p_target.float().contiguous(), p_max_all.float().contiguous(), pred.to(torch.int32).contiguous()
    N = g.shape[0]
    if not (pt.shape == pm.shape == pr.shape == g.shape):
        raise ValueError("gt, p_target, p_max_all and pred must all be [samples]")
    with torch.cuda.device(dev):
        cap = max(N, 1)
        rec_unk = torch.empty(cap, dtype=torch.int64, device=dev)
        rec_known = torch.empty(cap, dtype=torch.int64, device=dev)
        totals = torch.empty(4, dtype=torch.int64, device=dev)
        taus = torch.empty(cap, dtype=torch.float32, device=dev)
        ccr_cnt = torch.empty(cap, dtype=torch.int64, device=dev)
        fpr_cnt = torch.empty(cap, dtype=torch.int64, device=dev)
        ccr = torch.empty(cap, dtype=torch.float64, device=dev)
        fpr = torch.empty(cap, dtype=torch.float64, device=dev)
        meta = torch.zeros(5, dtype=torch.int64, device=dev)
        st = L.stream_ptr(dev)
        L.call("osb_oscr_records_from_stats", L.ptr(pt), L.ptr(pm), L.ptr(pr), L.ptr(g), _LABEL_TYPES[g.dtype], N,
               int(num_classes), int(unk_label), L.ptr(rec_unk), L.ptr(rec_known), L.ptr(totals), st)
        ws = L.workspace("oscr_curve", L.size("osb_oscr_curve_workspace_bytes", cap), dev)
        L.call("osb_oscr_curve", L.ptr(rec_unk), L.ptr(totals[2:3]), L.ptr(rec_known), L.ptr(totals[0:1]), cap,
               L.ptr(taus), L.ptr(ccr_cnt), L.ptr(fpr_cnt), L.ptr(ccr), L.ptr(fpr), L.ptr(meta), L.ptr(ws), ws.numel(), st)
        meta[4] = totals[3]
    return OscrDeviceResult(taus, ccr_cnt, fpr_cnt, ccr, fpr, meta)


def calculate_oscr(gt, scores, unk_label=-1):
    """Drop-in for ``openset_imagenet.util.calculate_oscr``.

    Args:
        gt: integer (or float, as stored in the reference's npz files) target labels [N]; numpy array or tensor.
        scores: float array [N, C] (or [N, C+1]); numpy array, CPU tensor or CUDA tensor.
        unk_label: label whose samples count towards the false-positive rate (-1 negatives, -2 unknowns).
    Returns: (ccr, fpr) float64 numpy arrays.
    """
    r = calculate_oscr_device(gt, scores, unk_label)
    n_unique, n_known, _, n_unk, n_bad = r.meta.tolist()
    if n_bad:
        raise IndexError(f"{n_bad} label(s) out of bounds for axis 1 with size {scores.shape[1]}")
    n = max(n_unique - 1, 0)
    ccr = r.ccr[:n].cpu().numpy()
    fpr = r.fpr[:n].cpu().numpy()
    if n and n_unk == 0:
        warnings.warn("invalid value encountered in scalar divide", RuntimeWarning, stacklevel=2)
    return ccr, fpr


def calculate_oscr_and_confidence_device(gt, scores, unk_label=-2, offset=0., last_valid_class=None, device=None):
    """OSCR and ``metrics.confidence(scores, gt, offset, unknown_class=unk_label, last_valid_class)`` of the same
    evaluation set from ONE pass over the score matrix (``osb_oscr_with_confidence``).  Returns
    ``(OscrDeviceResult, conf_sums)`` with ``conf_sums`` float64 [2] on the device; the confidence counts are
    ``meta[1]`` (known) and ``meta[3]`` (rows labelled ``unk_label``).  ``unk_label`` must be negative."""
    if isinstance(scores, torch.Tensor) and scores.is_cuda:
        dev = scores.device
    else:
        dev = torch.device(device) if device is not None else default_device()
    s = _to_device(scores, dev, torch.float32)
    if s.dim() != 2:
        raise ValueError(f"scores must be [samples, classes], got {tuple(s.shape)}")
    if s.stride(1) != 1 and s.shape[1] > 1:
        s = s.contiguous()
    g = _to_device(gt, dev)
    if g.dtype not in _LABEL_TYPES:
        g = g.to(torch.int64)
    g = g.contiguous()
    N, Cc = s.shape
    if g.dim() != 1 or g.shape[0] != N:
        raise ValueError(f"gt must be [samples]={N}, got {tuple(g.shape)}")
    if int(unk_label) >= 0:
        raise ValueError("the one-pass table needs a negative unk_label (a row is then known or unknown, never both)")
    valid_cols = len(range(Cc)[slice(None, last_valid_class)])
    one_plus_offset = float(torch.tensor(1.0 + offset, dtype=torch.float32))
    with torch.cuda.device(dev):
        cap = max(N, 1)
        taus = torch.empty(cap, dtype=torch.float32, device=dev)
        ccr_cnt = torch.empty(cap, dtype=torch.int64, device=dev)
        fpr_cnt = torch.empty(cap, dtype=torch.int64, device=dev)
        ccr = torch.empty(cap, dtype=torch.float64, device=dev)
        fpr = torch.empty(cap, dtype=torch.float64, device=dev)
        meta = torch.zeros(5, dtype=torch.int64, device=dev)
        conf_sums = torch.zeros(2, dtype=torch.float64, device=dev)
        ws = L.workspace("oscr", L.size("osb_oscr_workspace_bytes", N), dev)
        L.call("osb_oscr_with_confidence", L.ptr(s), s.stride(0) if N else Cc, L.ptr(g), _LABEL_TYPES[g.dtype], N, Cc, int(unk_label),
               one_plus_offset, valid_cols, L.ptr(taus), L.ptr(ccr_cnt), L.ptr(fpr_cnt), L.ptr(ccr), L.ptr(fpr), L.ptr(meta),
               L.ptr(conf_sums), L.ptr(ws), ws.numel(), L.stream_ptr(dev))
    return OscrDeviceResult(taus, ccr_cnt, fpr_cnt, ccr, fpr, meta), conf_sums


def conf_and_ccr(gt, scores, loss="entropic", unk_label=-2, queries=(1e-3, 1e-2, 0.1, 1.0)):
    """The numbers of one row of the reference's result table (script/plot_all.py:344-387, ``conf_and_ccr_table``) from
    one pass over the scores: ``(confidence, ccr_at)`` with ``confidence = (kn_conf, kn_count, neg_conf, neg_count)``
    as ``metrics.confidence`` returns it for ``offset = 0 if loss == 'garbage' else 1 / (max(gt) + 1)``,
    ``unknown_class = unk_label``, ``last_valid_class = -1 if loss == 'garbage' else None`` (plot_all.py:366-372), and
    ``ccr_at`` = one entry per query q: ``ccr[argmin |fpr - q|]``, or None where the reference prints '---'
    (``round(100 * |fpr - q| / q, 1) >= 10``, plot_all.py:378-384)."""
    gt_t = gt if isinstance(gt, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(gt))
    n_cls = int(gt_t.max()) + 1                                    # numpy.max(gt) + 1, plot_all.py:366 (one small reduction)
    offset = 0. if loss == "garbage" else 1. / n_cls
    last_valid_class = -1 if loss == "garbage" else None
    r, sums = calculate_oscr_and_confidence_device(gt, scores, unk_label, offset, last_valid_class)
    n_q = len(queries)
    if scores.shape[0] == 0:
        raise AttributeError("'int' object has no attribute 'item'")
    idx, f_at, c_at = ccr_at_fpr_device(r.ccr, r.fpr, r.meta[0:1], queries)
    n_unique, n_known, _, n_unk, n_bad = r.meta.tolist()              # the one host sync
    if n_bad:
        raise IndexError(f"{n_bad} label(s) out of bounds for axis 1 with size {scores.shape[1]}")
    s0, s1 = sums.tolist()
    confidence = (s0 / n_known if n_known else 0.0, n_known, s1 / n_unk if n_unk else 0.0, n_unk)
    if n_unique - 1 <= 0:
        raise ValueError("attempt to get argmin of an empty sequence")
    out = []
    for q, fv, cv in zip(queries, f_at.tolist()[:n_q], c_at.tolist()[:n_q]):
        err = round(100 * abs(fv - q) / q, 1)
        out.append(None if err >= 10.0 else cv)
    return confidence, out


def ccr_at_fpr_device(ccr: torch.Tensor, fpr: torch.Tensor, n_unique: torch.Tensor, queries):
    """First argmin |fpr - q| per query on the device; returns (idx int64, fpr_at, ccr_at) tensors."""
    dev = L.require_cuda(ccr, fpr, n_unique)
    q = (L.C.c_double * len(queries))(*[float(v) for v in queries])
    out_idx = torch.empty(len(queries), dtype=torch.int64, device=dev)
    out_fpr = torch.empty(len(queries), dtype=torch.float64, device=dev)
    out_ccr = torch.empty(len(queries), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        L.call("osb_ccr_at_fpr", L.ptr(ccr), L.ptr(fpr), L.ptr(n_unique), ccr.numel(), q, len(queries),
               L.ptr(out_idx), L.ptr(out_fpr), L.ptr(out_ccr), L.stream_ptr(dev))
    return out_idx, out_fpr, out_ccr


def ccr_at_fpr(ccr, fpr, queries=(1e-3, 1e-2, 0.1, 1.0), device=None):
    """Numeric part of ``conf_and_ccr_table`` (script/plot_all.py:346-352,378-384).

    For each query q returns ``(idx, fpr[idx], value)`` with idx = first argmin |fpr - q| and value =
    ccr[idx], or None when ``round(100*|fpr[idx]-q|/q, 1) >= 10.0`` (printed as '---' by the reference).
    ``ccr``/``fpr`` are the arrays ``calculate_oscr`` returned (numpy) or their device tensors.
    """
    if isinstance(fpr, torch.Tensor) and fpr.is_cuda:
        dev = fpr.device
    else:
        dev = torch.device(device) if device is not None else default_device()
    c = _to_device(ccr, dev, torch.float64).contiguous()
    f = _to_device(fpr, dev, torch.float64).contiguous()
    if f.numel() == 0:
        raise ValueError("attempt to get argmin of an empty sequence")
    n_unique = torch.tensor([f.numel() + 1], dtype=torch.int64, device=dev)
    idx, f_at, c_at = ccr_at_fpr_device(c, f, n_unique, queries)
    out = []
    for q, i, fv, cv in zip(queries, idx.tolist(), f_at.tolist(), c_at.tolist()):
        err = round(100 * abs(fv - q) / q, 1)
        out.append((i, fv, None if err >= 10.0 else cv))
    return out


def get_histogram(array, unk_label=-1, metric='score', bins=100, drop_bg=False, log_space=False,
                  geomspace_limits=(1, 1e2), device=None):
    """Histograms of scores or feature norms, the reference's ``util.get_histogram`` (util.py:202-228) with the
    same signature and return value ``(kn_hist, kn_edges, unk_hist, unk_edges)`` (numpy int64 counts, numpy edges).

    ``array`` is the reference's npz dict (keys ``scores``, ``gt``, ``features``; numpy arrays or tensors).  One
    kernel produces the per-row values (known rows: the target's score, ``unk_label`` rows: the row maximum; or the
    feature norm for both) together with the min / max np.histogram derives its bin range from; the edges are then
    numpy's own (``np.histogram_bin_edges`` on [min, max], or ``np.geomspace``), and a second kernel counts with
    np.histogram's rules, so for metric 'score' the counts are bit-exact.  For metric 'norm' the norm is accumulated
    in fp64 (the reference's float32 pairwise sum is not reproducible), so a value within one float32 ulp of an edge
    may land in the neighbouring bin.
    """
    if metric not in ('score', 'norm'):
        raise ValueError("metric must be 'score' or 'norm'")
    src = array['scores'] if metric == 'score' else array['features']
    if isinstance(src, torch.Tensor) and src.is_cuda:
        dev = src.device
    else:
        dev = torch.device(device) if device is not None else default_device()
    gt = _to_device(array['gt'], dev)
    gt = gt.to(torch.int64).contiguous()            # gt.astype(np.int64), util.py:213
    v = _to_device(src, dev, torch.float32)
    if v.dim() != 2 or v.shape[0] != gt.shape[0]:
        raise ValueError("scores / features must be [N, C] with one row per label")
    if v.stride(1) != 1:
        v = v.contiguous()
    N, C = v.shape
    if metric == 'score':
        valid = C - 1 if drop_bg else C             # score[:, :-1], util.py:210-211
        if valid < 1:
            raise ValueError("no score column left")
        if N and int(gt.max()) >= valid:
            raise IndexError(f"label {int(gt.max())} is out of bounds for {valid} score columns")
    kn = torch.empty(N, dtype=torch.float32, device=dev)
    unk = torch.empty(N, dtype=torch.float32, device=dev)
    counts = torch.empty(2, dtype=torch.int64, device=dev)
    minmax = torch.empty(4, dtype=torch.float32, device=dev)
    st = L.stream_ptr(dev)
    if metric == 'score':
        L.call("osb_hist_metrics", L.ptr(v), v.stride(0), None, 0, 0, L.ptr(gt), N, C, valid, int(unk_label),
               L.ptr(kn), L.ptr(unk), L.ptr(counts), L.ptr(minmax), st)
    else:
        L.call("osb_hist_metrics", None, 0, L.ptr(v), v.stride(0), C, L.ptr(gt), N, 0, 0, int(unk_label),
               L.ptr(kn), L.ptr(unk), L.ptr(counts), L.ptr(minmax), st)
    cnt = counts.cpu().numpy()
    # the kernel keeps min / max as order-preserving integers; decode them on the host
    mm_bits = minmax.view(torch.int32).cpu().numpy().astype(np.uint32)
    mm = np.where(mm_bits & np.uint32(0x80000000), mm_bits & np.uint32(0x7fffffff), ~mm_bits).astype(np.uint32).view(np.float32)

    def one(values, n, lo, hi):
        if log_space:
            edges = np.geomspace(geomspace_limits[0], geomspace_limits[1], num=bins)     # util.py:223-225
        else:
            rng = np.array([lo, hi], dtype=np.float32) if n else np.empty(0, dtype=np.float32)
            edges = np.histogram_bin_edges(rng, bins=bins)
        if not np.all(np.diff(edges) >= 0):
            raise ValueError('`bins` must increase monotonically, when an array')
        e = torch.from_numpy(np.ascontiguousarray(edges, dtype=np.float64)).to(dev)
        out = torch.empty(len(edges) - 1, dtype=torch.int64, device=dev)
        L.call("osb_histogram", L.ptr(values), N, L.ptr(e), len(edges), L.ptr(out), st)
        return out.cpu().numpy(), edges

    kn_hist, kn_edges = one(kn, int(cnt[0]), mm[0], mm[1])
    unk_hist, unk_edges = one(unk, int(cnt[1]), mm[2], mm[3])
    return kn_hist, kn_edges, unk_hist, unk_edges
