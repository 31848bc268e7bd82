"""Open-set losses with the reference's call signatures, computed by libosb's CUDA kernels.

Mirrors ``openset_imagenet/losses.py`` (``EntropicOpensetLoss``, :7-29) and the two
``torch.nn.CrossEntropyLoss`` configurations the reference trains with (``train.py:343`` softmax with
``ignore_index=-1``; ``train.py:346-347`` garbage class with balanced class weights).  Each object is
a plain callable ``loss_fn(logits, target) -> 0-dim tensor`` that supports ``.item()`` and
``.backward()`` exactly like the reference's (``train.py:135-138,179-180``).

Differences that matter to a user switching over:
  * CUDA tensors only (sm_100a); CPU tensors raise instead of falling back;
  * no host synchronisation in the loss itself (the reference syncs three times per call,
    losses.py:21,22,26) unless ``validate_labels=True`` (default), which reads one counter back to raise
    the reference's ``IndexError`` for labels >= C.  Pass ``validate_labels=False`` in training loops:
    an out-of-range label then yields a NaN loss instead of an exception.
"""
from __future__ import annotations

import torch

from . import _lib as L

_VARIANT_NAMES = {"entropic": L.ENTROPIC, "softmax": L.SOFTMAX, "garbage": L.GARBAGE}


def _prep_logits(logits: torch.Tensor) -> tuple[torch.Tensor, int]:
    if logits.dim() != 2:
        raise ValueError(f"logits must be [batch, classes], got {tuple(logits.shape)}")
    if logits.dtype == torch.float16 or logits.dtype == torch.float64:
        logits = logits.float()
    if logits.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError(f"unsupported logits dtype {logits.dtype}")
    if logits.stride(1) != 1 and logits.shape[1] > 1:
        logits = logits.contiguous()
    return logits, (L.F32 if logits.dtype == torch.float32 else L.BF16)


def _prep_target(target: torch.Tensor, B: int) -> torch.Tensor:
    if target.dim() != 1 or target.shape[0] != B:
        raise ValueError(f"target must be [batch]={B}, got {tuple(target.shape)}")
    if target.dtype != torch.int64:
        target = target.long()
    return target.contiguous()


class _LossFn(torch.autograd.Function):
    """loss = f(logits, target); backward re-runs the row kernel with the upstream gradient as a
    device scalar, so neither direction needs a host sync."""

    @staticmethod
    def forward(ctx, logits, target, cfg):
        dev = L.require_cuda(logits, target, cfg.class_w)
        in_dtype = logits.dtype
        z, zt = _prep_logits(logits)
        B, Cc = z.shape
        y = _prep_target(target, B)
        st = L.stream_ptr(dev)
        with torch.cuda.device(dev):
            ws = L.workspace("loss", L.size("osb_loss_workspace_bytes", B), dev)
            den = torch.empty(1, dtype=torch.float64, device=dev)
            counts = torch.empty(4, dtype=torch.int64, device=dev)
            L.call("osb_label_stats", L.ptr(y), L.ptr(cfg.class_w), B, Cc, cfg.variant, cfg.ignore_index,
                   L.ptr(den), L.ptr(counts), L.ptr(ws), ws.numel(), 0, st)
            if cfg.den_hook is not None:          # batch-sharded training: all-reduce the denominator
                den = cfg.den_hook(den)
            out = torch.empty(1, dtype=torch.float32, device=dev)
            num = torch.empty(1, dtype=torch.float64, device=dev)
            L.call("osb_loss_from_logits", L.ptr(z), zt, z.stride(0), L.ptr(y), L.ptr(cfg.class_w), B, Cc,
                   cfg.variant, float(cfg.unk_weight), cfg.ignore_index, L.ptr(den), Cc,
                   L.ptr(num), L.ptr(out), None, None, None, None, None, None, 0, None,
                   L.ptr(ws), ws.numel(), st)
        if cfg.validate_labels:
            bad = int(counts[3].item())
            if bad:
                raise IndexError(f"{bad} target label(s) out of range for {Cc} classes")
        ctx.save_for_backward(z, y, den)
        ctx.cfg, ctx.in_dtype = cfg, in_dtype
        cfg.last_counts, cfg.last_num, cfg.last_den = counts, num, den
        return out[0]

    @staticmethod
    def backward(ctx, grad_out):
        z, y, den = ctx.saved_tensors
        cfg = ctx.cfg
        dev = z.device
        B, Cc = z.shape
        zt = L.F32 if z.dtype == torch.float32 else L.BF16
        g = grad_out.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        st = L.stream_ptr(dev)
        with torch.cuda.device(dev):
            ws = L.workspace("loss", L.size("osb_loss_workspace_bytes", B), dev)
            dz = torch.empty_like(z, memory_format=torch.contiguous_format)
            scratch_f = torch.empty(1, dtype=torch.float32, device=dev)
            scratch_d = torch.empty(1, dtype=torch.float64, device=dev)
            L.call("osb_loss_from_logits", L.ptr(z), zt, z.stride(0), L.ptr(y), L.ptr(cfg.class_w), B, Cc,
                   cfg.variant, float(cfg.unk_weight), cfg.ignore_index, L.ptr(den), Cc,
                   L.ptr(scratch_d), L.ptr(scratch_f), None, None, None, None, None,
                   L.ptr(dz), dz.stride(0), L.ptr(g), L.ptr(ws), ws.numel(), st)
        if dz.dtype != ctx.in_dtype:
            dz = dz.to(ctx.in_dtype)
        return dz, None, None


class _Cfg:
    __slots__ = ("variant", "unk_weight", "class_w", "ignore_index", "validate_labels", "den_hook",
                 "last_counts", "last_num", "last_den")

    def __init__(self, variant, unk_weight=1.0, class_w=None, ignore_index=-100, validate_labels=True, den_hook=None):
        self.variant, self.unk_weight, self.class_w = variant, unk_weight, class_w
        self.ignore_index, self.validate_labels, self.den_hook = ignore_index, validate_labels, den_hook
        self.last_counts = self.last_num = self.last_den = None


class EntropicOpensetLoss:
    """Drop-in for ``openset_imagenet.losses.EntropicOpensetLoss`` (losses.py:7-29).

    Known rows (label >= 0) get a one-hot target, every negative label (-1, -2, ...) the uniform target
    ``unk_weight / num_of_classes``; the loss is the mean over ALL rows of the soft-target cross-entropy.
    """

    def __init__(self, num_of_classes, unk_weight=1, validate_labels=True):
        self.class_count = int(num_of_classes)
        self.unknowns_multiplier = unk_weight / self.class_count
        self._cfg = _Cfg(L.ENTROPIC, float(unk_weight), validate_labels=validate_labels)

    def __call__(self, logits, target):
        if logits.dim() == 2 and logits.shape[1] != self.class_count:
            raise RuntimeError(f"logits have {logits.shape[1]} classes, loss was built for {self.class_count}")
        return _LossFn.apply(logits, target, self._cfg)


class CrossEntropyLoss:
    """The subset of ``torch.nn.CrossEntropyLoss`` the reference uses (train.py:343,347): class-index
    targets, mean reduction, optional per-class ``weight`` and ``ignore_index``."""

    def __init__(self, weight=None, ignore_index=-100, validate_labels=True):
        variant = L.GARBAGE if weight is not None else L.SOFTMAX
        if weight is not None:
            L.require_cuda(weight)
            weight = weight.detach().float().contiguous()
        self.weight = weight
        self.ignore_index = int(ignore_index)
        self._cfg = _Cfg(variant, 1.0, weight, self.ignore_index, validate_labels)

    def __call__(self, logits, target):
        if self.weight is not None and logits.dim() == 2 and logits.shape[1] != self.weight.numel():
            raise RuntimeError(f"weight has {self.weight.numel()} entries, logits have {logits.shape[1]} classes")
        return _LossFn.apply(logits, target, self._cfg)


def make_loss(loss_type: str, n_classes: int, unk_weight: float = 1.0, class_weights: torch.Tensor | None = None,
              validate_labels: bool = True):
    """Loss selection of ``train.worker`` (train.py:338-347): 'entropic' | 'softmax' | 'garbage'."""
    if loss_type == "entropic":
        return EntropicOpensetLoss(n_classes, unk_weight, validate_labels)
    if loss_type == "softmax":
        return CrossEntropyLoss(ignore_index=-1, validate_labels=validate_labels)
    if loss_type == "garbage":
        if class_weights is None:
            raise ValueError("the garbage-class loss needs class_weights (dataset.py:77-86)")
        return CrossEntropyLoss(weight=class_weights, validate_labels=validate_labels)
    raise ValueError(f"unknown loss type {loss_type!r}")


def class_weights_from_labels(labels: torch.Tensor, label_count: int) -> torch.Tensor:
    """Balanced class weights ``len(train) / (count_c * label_count)`` in float32 (dataset.py:77-86).
    Host bookkeeping done once per training run; returned on the labels' device."""
    counts = torch.bincount(labels.reshape(-1).long(), minlength=label_count).double()
    return (labels.numel() / (counts * label_count)).float()
