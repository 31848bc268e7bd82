"""Synthetic inputs of the protocol shapes (BASELINE.json configurations) for benchmarks and
smoke runs: there is no network for ImageNet or the protocol files, so features, labels and scores are seeded
random draws with the class counts and negative fractions of the protocols.

``class_weights_from_labels`` is the balanced class weighting of the garbage-class variant
(openset_imagenet/dataset.py:77-86: ``len(train) / (count_c * label_count)``, float32)."""
from __future__ import annotations

import numpy as np


def make_labels(B: int, C: int, neg_frac: float = 1.0 / 3.0, seed: int = 0, neg_label: int = -1) -> np.ndarray:
    """Labels uniform in [0, C), each row independently replaced by ``neg_label`` with probability ``neg_frac``."""
    rng = np.random.default_rng(seed)
    y = rng.integers(0, C, size=B, dtype=np.int64)
    y[rng.random(B) < neg_frac] = neg_label
    return y


def class_weights_from_labels(labels: np.ndarray, label_count: int) -> np.ndarray:
    labels = np.asarray(labels, dtype=np.int64)
    counts = np.bincount(labels, minlength=label_count).astype(np.float64)
    with np.errstate(divide="ignore"):
        w = len(labels) / (counts * label_count)
    return w.astype(np.float32)


def make_oscr_case(N: int, C: int, seed: int = 0):
    """Evaluation set of BASELINE.json's OSCR configuration: 50 % known / 25 % label -1 / 25 % label -2,
    logits 2*N(0,1) with +4 on the target column of known rows, scores = float32 softmax.  Returns (gt, scores)."""
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
