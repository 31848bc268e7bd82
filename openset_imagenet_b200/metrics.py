"""``metrics.confidence`` of the reference (openset_imagenet/metrics.py:8-42) on the GPU.

One streaming kernel: known rows gather a single element, only rows labelled ``unknown_class`` are read
in full, counts are exact, sums are accumulated in double in a fixed order.  The reference's other
helpers in that file (``predict_objectosphere``, ``auc_score_*``) are dead code there (imported, never
called) and are out of scope.
"""
from __future__ import annotations

import torch

from . import _lib as L


def _valid_cols(C: int, last_valid_class) -> int:
    # python slice semantics of scores[unknown, :last_valid_class] (metrics.py:39)
    return len(range(C)[slice(None, last_valid_class)])


def confidence_sums(scores: torch.Tensor, target_labels: torch.Tensor, offset=0., unknown_class=-1,
                    last_valid_class=None) -> tuple[torch.Tensor, torch.Tensor]:
    """Device-side result without a host sync: ``sums`` float64 [2] = (sum of known target scores, sum of
    1+offset-max over negatives) and ``counts`` int64 [3] = (kn_count, neg_count, labels >= C).
    Additive across shards (all-reduce both for a sharded validation set)."""
    dev = L.require_cuda(scores, target_labels)
    if scores.dim() != 2:
        raise ValueError(f"scores must be [samples, classes], got {tuple(scores.shape)}")
    N, Cc = scores.shape
    if scores.dtype != torch.float32:
        scores = scores.float()
    if scores.stride(1) != 1 and Cc > 1:
        scores = scores.contiguous()
    y = target_labels
    if y.dtype != torch.int64:
        y = y.long()
    y = y.contiguous()
    if y.dim() != 1 or y.shape[0] != N:
        raise ValueError(f"target_labels must be [samples]={N}, got {tuple(y.shape)}")
    vc = _valid_cols(Cc, last_valid_class)
    sums = torch.zeros(2, dtype=torch.float64, device=dev)
    counts = torch.zeros(3, dtype=torch.int64, device=dev)
    if N == 0:
        return sums, counts
    one_plus_offset = float(torch.tensor(1.0 + offset, dtype=torch.float32))   # the scalar torch would use
    with torch.cuda.device(dev):
        ws = L.workspace("conf", L.size("osb_confidence_workspace_bytes", N), dev)
        L.call("osb_confidence", L.ptr(scores), scores.stride(0), L.ptr(y), N, Cc, one_plus_offset,
               int(unknown_class), vc, L.ptr(sums), L.ptr(counts), L.ptr(ws), ws.numel(), L.stream_ptr(dev))
    return sums, counts


def confidence(scores, target_labels, offset=0., unknown_class=-1, last_valid_class=None):
    """Drop-in for ``openset_imagenet.metrics.confidence``.

    Returns ``(kn_conf, kn_count, neg_conf, neg_count)`` as Python ``(float, int, float, int)``:
    the mean target-class score of known samples (label >= 0 and != unknown_class) and the mean of
    ``1 + offset - max_c scores[:, :last_valid_class]`` over samples labelled ``unknown_class``;
    0.0 for an empty group (metrics.py:29-40).  Like the reference it raises on an empty input.
    """
    if scores.shape[0] == 0:
        raise AttributeError("'int' object has no attribute 'item'")      # what sum([]).item() raises there
    sums, counts = confidence_sums(scores, target_labels, offset, unknown_class, last_valid_class)
    return finish_confidence(sums, counts, scores.shape[1])


def finish_confidence(sums: torch.Tensor, counts: torch.Tensor, num_classes: int):
    """(sums [2] f64, counts [3] i64) -> the reference's ``(kn_conf, kn_count, neg_conf, neg_count)`` with its
    empty-group rule (metrics.py:29-40).  The one host sync of the call."""
    s = sums.tolist()
    kn_count, neg_count, bad = (int(v) for v in counts.tolist())
    if bad:
        raise IndexError(f"{bad} known label(s) out of range for {num_classes} classes")
    kn_conf = s[0] / kn_count if kn_count else 0.0
    neg_conf = s[1] / neg_count if neg_count else 0.0
    return kn_conf, kn_count, neg_conf, neg_count


class StreamingConfidence:
    """Confidence of a whole validation set accumulated batch by batch from the fused head's per-row statistics
    (``model.HeadStats``: ``p_target``, ``p_max``), so the ``all_scores [N_val, C]`` tensor of the reference's
    ``validate`` (train.py:170,185) never exists.  No host synchronisation until ``result()``.

    ``offset`` / ``unknown_class`` as in ``confidence``; the slice ``last_valid_class`` is a property of the head
    (``FusedOpenSetHead(last_valid_class=...)``), which already applied it when it reduced ``p_max``.
    """

    def __init__(self, num_classes: int, offset=0., unknown_class=-1, device=None):
        self.dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.C = int(num_classes)
        self.unknown_class = int(unknown_class)
        self.one_plus_offset = float(torch.tensor(1.0 + offset, dtype=torch.float32))
        self.sums = torch.zeros(2, dtype=torch.float64, device=self.dev)
        self.counts = torch.zeros(3, dtype=torch.int64, device=self.dev)

    def reset(self):
        self.sums.zero_()
        self.counts.zero_()

    def update(self, p_target: torch.Tensor, p_max: torch.Tensor, target_labels: torch.Tensor):
        L.require_cuda(p_target, p_max, target_labels)
        y = target_labels if target_labels.dtype == torch.int64 else target_labels.long()
        y = y.contiguous()
        n = y.shape[0]
        if p_target.shape != (n,) or p_max.shape != (n,):
            raise ValueError("p_target, p_max and target_labels must all be [batch]")
        with torch.cuda.device(self.dev):
            L.call("osb_confidence_from_stats", L.ptr(p_target.float().contiguous()), L.ptr(p_max.float().contiguous()), L.ptr(y),
                   n, self.C, self.one_plus_offset, self.unknown_class, 1, L.ptr(self.sums), L.ptr(self.counts),
                   L.stream_ptr(self.dev))

    def result(self):
        """``(kn_conf, kn_count, neg_conf, neg_count)`` with the reference's types and empty-group rule."""
        return finish_confidence(self.sums, self.counts, self.C)

    def result_sharded(self, group=None):
        """``result()`` of a validation set whose batches were sharded over the ranks (one small all-reduce)."""
        from .dist import allreduce_confidence
        return finish_confidence(*allreduce_confidence(self.sums, self.counts, group), self.C)
