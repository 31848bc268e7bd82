"""Training / validation / array-extraction loops of the reference (openset_imagenet/train.py:104-234) around the
fused head, with the reference's signatures.

What changes against the reference, and nothing else:
  * ``train``: the logits layer, the loss and their backward run as the fused tcgen05 head
    (``model.forward_loss``) when the model offers it; the per-batch ``j.item()`` host sync (train.py:136) is
    replaced by a device-side running sum that is read ONCE per epoch (same ``trackers["j"].avg``);
  * ``validate``: no ``all_scores [N_val, C]`` tensor (train.py:170,185) — confidence is accumulated per batch from
    the head's per-row statistics (``metrics.StreamingConfidence``); the loss average is read once at the end;
  * ``get_arrays``: unchanged semantics and on-disk format (float32 ``gt``, logits, features, softmax scores).
Models without ``forward_loss`` (the reference's own ``ResNet50``) take the reference's path with the drop-in
loss objects of ``losses.py`` and ``metrics.confidence``.
"""
from __future__ import annotations

from collections import defaultdict

import numpy as np
import torch

from .metrics import StreamingConfidence, confidence


class AverageMeter:
    """Running average (openset_imagenet/losses.py:32-60)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val, self.avg, self.sum, self.count = 0, 0, 0, 0

    def update(self, val, count=1):
        self.val = val
        self.sum += val * count
        self.count += count
        self.avg = self.sum / self.count

    def __repr__(self):
        return f"{self.avg:3.3f}"


def _device_of(model) -> torch.device:
    return next(model.parameters()).device


def _confidence_args(cfg, n_classes):
    """train.py:156-163: offset, unknown_class, last_valid_class per loss type."""
    if cfg.loss.type == "garbage":
        return 0., n_classes - 1, -1
    return 1. / n_classes, -1, None


def train(model, data_loader, optimizer, loss_fn, trackers, cfg):
    """One epoch (train.py:104-139).  ``trackers["j"]`` ends with the sample-weighted mean loss of the epoch."""
    for metric in trackers.values():
        metric.reset()
    dev = _device_of(model)
    fused = hasattr(model, "forward_loss")
    loss_sum = torch.zeros(1, dtype=torch.float64, device=dev)
    seen = 0
    last = None
    for images, labels in data_loader:
        model.train()                                  # to collect batch-norm statistics (train.py:124)
        batch_len = labels.shape[0]
        optimizer.zero_grad()
        images = images.to(dev, non_blocking=True)
        labels = labels.to(dev, non_blocking=True)
        if fused:
            j, _ = model.forward_loss(images, labels)
        else:
            logits, _ = model(images)
            j = loss_fn(logits, labels)
        loss_sum += j.detach().double() * batch_len    # no host sync per batch (the reference: j.item(), train.py:136)
        seen += batch_len
        last = j
        j.backward()
        optimizer.step()
    if seen:
        trackers["j"].update(float(loss_sum.item()) / seen, seen)
        trackers["j"].val = float(last.item())


def validate(model, data_loader, loss_fn, n_classes, trackers, cfg):
    """Validation loop (train.py:142-196): mean loss and the two confidences, no score matrix."""
    for metric in trackers.values():
        metric.reset()
    offset, unknown_class, last_valid_class = _confidence_args(cfg, n_classes)
    dev = _device_of(model)
    fused = hasattr(model, "forward_loss")
    model.eval()
    with torch.no_grad():
        loss_sum = torch.zeros(1, dtype=torch.float64, device=dev)
        seen = 0
        if fused:
            # the head reduces p_max over its own `valid_cols` columns: they must be the slice the reference's
            # confidence() would take for this loss type (metrics.py:39, train.py:156-163)
            want_cols = len(range(n_classes)[slice(None, last_valid_class)])
            head = model.logits
            if head.out_features != n_classes or head.valid_cols != want_cols:
                raise ValueError(f"validate(): the fused head reduces the confidence maximum over {head.valid_cols} of "
                                 f"{head.out_features} columns, but loss type {cfg.loss.type!r} with {n_classes} classes needs "
                                 f"{want_cols} (build the head with last_valid_class={last_valid_class!r})")
            conf = StreamingConfidence(n_classes, offset, unknown_class, device=dev)
        else:
            data_len = len(data_loader.dataset)
            all_targets = torch.empty((data_len,), dtype=torch.int64, device=dev)
            all_scores = torch.empty((data_len, n_classes), device=dev)
        for images, labels in data_loader:
            batch_len = labels.shape[0]
            images = images.to(dev, non_blocking=True)
            labels = labels.to(dev, non_blocking=True)
            if fused:
                j, _ = model.forward_loss(images, labels)
                st = model.logits.last_stats
                conf.update(st.p_target, st.p_max, labels)
            else:
                logits, _ = model(images)
                all_scores[seen:seen + batch_len] = torch.nn.functional.softmax(logits.float(), dim=1)
                all_targets[seen:seen + batch_len] = labels
                j = loss_fn(logits, labels)
            loss_sum += j.detach().double() * batch_len
            seen += batch_len
        if seen:
            trackers["j"].update(float(loss_sum.item()) / seen, seen)
        if fused:
            kn_conf, kn_count, neg_conf, neg_count = conf.result()
        else:
            kn_conf, kn_count, neg_conf, neg_count = confidence(all_scores[:seen], all_targets[:seen], offset, unknown_class,
                                                               last_valid_class)
        if kn_count:
            trackers["conf_kn"].update(kn_conf, kn_count)
        if neg_count:
            trackers["conf_unk"].update(neg_conf, neg_count)


def get_arrays(model, loader):
    """Targets, logits, deep features and softmax scores of a whole dataset as numpy arrays (train.py:200-234);
    ``gt`` is float32 like the reference's (train.py:212), which ``util.calculate_oscr`` accepts as is."""
    dev = _device_of(model)
    model.eval()
    with torch.no_grad():
        data_len = len(loader.dataset)
        logits_dim = model.logits.out_features
        features_dim = model.logits.in_features
        all_targets = torch.empty(data_len, device=dev)
        all_logits = torch.empty((data_len, logits_dim), device=dev)
        all_feat = torch.empty((data_len, features_dim), device=dev)
        all_scores = torch.empty((data_len, logits_dim), device=dev)
        index = 0
        for images, labels in loader:
            n = labels.shape[0]
            logit, feature = model(images.to(dev, non_blocking=True))
            logit = logit.float()
            all_targets[index:index + n] = labels.to(dev)
            all_logits[index:index + n] = logit
            all_feat[index:index + n] = feature.float()
            all_scores[index:index + n] = torch.nn.functional.softmax(logit, dim=1)
            index += n
        return (all_targets.cpu().numpy(), all_logits.cpu().numpy(), all_feat.cpu().numpy(), all_scores.cpu().numpy())


def make_trackers():
    return defaultdict(AverageMeter)


# ---- the on-disk array format either side of the OSCR evaluation (script/evaluate.py:133-149, util.py:69-87) ----
def save_arrays(path, gt, logits, features, scores):
    """``{loss}_{val,test}_arr{_curr,_best}.npz`` with the reference's keys."""
    np.savez(path, gt=gt, logits=logits, features=features, scores=scores)


def read_array_list(file_names):
    """util.read_array_list: dict name -> npz, from a dict of paths or a list of ``<name>_arr.npz``-style paths."""
    arrays = defaultdict(dict)
    if isinstance(file_names, dict):
        for key, file in file_names.items():
            arrays[key] = np.load(file)
    else:
        for file in file_names:
            file = str(file)
            arrays[file.split('/')[-1][:-8]] = np.load(file)
    return arrays


def oscr_from_arrays(array, unk_label=-1, drop_last_column=False):
    """``calculate_oscr`` on one loaded npz (``gt`` float32, ``scores`` [N, C']); ``drop_last_column`` strips the
    garbage-class column as util.plot_oscr does (util.py:186-187) and script/plot_all.py:359-363 does not."""
    from .util import calculate_oscr
    scores = array["scores"]
    if drop_last_column:
        scores = scores[:, :-1]
    return calculate_oscr(array["gt"], scores, unk_label=unk_label)
