"""CPU port of the reference's hot path, op for op, for TIMING the CPU baseline.  TEST INFRASTRUCTURE.

``oracle/oracle.py`` is the closed-form float64 checker; this file instead reproduces the SEQUENCE OF
TORCH / NUMPY OPERATIONS the reference executes, so that a wall-clock measurement of it on the GPU box's
host cores stands in for the reference itself (which cannot travel: /root/reference is absent there).
Used only by ``bench.py`` (``cpu_baseline`` and ``--impl reference``) and by tests that check it against
the real reference and the oracle.

Steps mirrored (paths relative to the reference root):
  * ``EntropicLossPort``     — losses.py:9-14 (buffers) and :16-29 (dense soft-target matrix, boolean-mask
                               assignments with their host syncs, ``CrossEntropyLoss`` in probability mode);
  * ``confidence_port``      — metrics.py:24-42 including the Python-builtin ``sum`` over tensors (:27-28);
  * ``head_step_port``       — train.py:127-139 (zero_grad, forward, loss, ``.item()``, backward) followed by
                               the validation-side ``softmax`` + ``confidence`` (train.py:177,187-192);
  * ``oscr_port_quadratic``  — util.py:100-122, one pass over the data per unique known target score.
"""
from __future__ import annotations

import numpy as np
import torch


class EntropicLossPort:
    def __init__(self, num_of_classes: int, unk_weight: float = 1.0):
        self.n = num_of_classes
        self.identity = torch.eye(num_of_classes)
        self.neg_row = torch.ones(num_of_classes) * (unk_weight / num_of_classes)
        self.ce = torch.nn.CrossEntropyLoss()

    def __call__(self, logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        soft = torch.zeros(logits.shape)
        negative = target < 0
        positive = ~negative
        if torch.any(positive):                                   # host sync in the reference too
            soft[positive, :] = self.identity[target[positive]]
        soft[negative, :] = self.neg_row.expand(torch.sum(negative).item(), self.n)
        return self.ce(logits, soft)


def confidence_port(scores, target_labels, offset=0.0, unknown_class=-1, last_valid_class=None):
    with torch.no_grad():
        is_unknown = target_labels == unknown_class
        is_known = torch.logical_and(target_labels >= 0, ~is_unknown)
        kn_count = sum(is_known).item()            # builtin sum: one tensor add per sample, as in the reference
        neg_count = sum(is_unknown).item()
        kn_conf = neg_conf = 0.0
        if kn_count:
            kn_conf = torch.sum(scores[is_known, target_labels[is_known]]).item() / kn_count
        if neg_count:
            top = torch.max(scores[is_unknown, :last_valid_class], dim=1)[0]
            neg_conf = torch.sum(1.0 + offset - top).item() / neg_count
    return kn_conf, kn_count, neg_conf, neg_count


def head_step_port(linear: torch.nn.Linear, loss_fn, x: torch.Tensor, y: torch.Tensor, offset: float):
    """One batch of the reference path on the CPU.  Returns (loss value, confidence tuple)."""
    linear.zero_grad(set_to_none=True)
    if x.grad is not None:
        x.grad = None
    logits = linear(x)
    j = loss_fn(logits, y)
    value = j.item()
    j.backward()
    with torch.no_grad():
        scores = torch.nn.functional.softmax(logits, dim=1)
        conf = confidence_port(scores, y, offset=offset, unknown_class=-1, last_valid_class=None)
    return value, conf


def oscr_port_quadratic(gt, scores, unk_label=-1):
    gt = gt.astype(int)
    known = gt >= 0
    unknown = gt == unk_label
    n_known, n_unknown = np.sum(known), np.sum(unknown)
    predicted = np.argmax(scores, axis=1)
    top = np.max(scores, axis=1)
    target_score = scores[known][range(known.sum()), gt[known]]
    ccr, fpr = [], []
    for tau in np.unique(target_score)[:-1]:
        ccr.append(((predicted[known] == gt[known]) & (target_score > tau)).sum() / n_known)
        fpr.append((unknown & (top > tau)).sum() / n_unknown)
    return np.array(ccr), np.array(fpr)
