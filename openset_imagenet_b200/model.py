"""The classification head of the reference model, fused (openset_imagenet/model.py:20-26,37-39).

``FusedOpenSetHead`` IS an ``nn.Linear`` (same parameters, same state-dict keys, same
``in_features``/``out_features`` that ``train.get_arrays`` reads, train.py:210-211), so it drops in as
``model.logits``.  Called like a linear layer it returns logits (cuBLAS, for callers that want the
tensor, e.g. the npz writer).  Its ``loss(features, target)`` method is the B200 path: one tcgen05
kernel computes the logits into TMEM and reduces them to the loss, the per-row confidence / OSCR
statistics and the grad-logits without ever writing logits to HBM; backward is two more tcgen05 GEMMs
(dX, dW).  bf16 operands, fp32 accumulation; master weights and dW stay fp32.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from . import _lib as L

_VARIANTS = {"entropic": L.ENTROPIC, "softmax": L.SOFTMAX, "garbage": L.GARBAGE}
_DEFAULT_IGNORE = {"entropic": -100, "softmax": -1, "garbage": -100}


@dataclass
class HeadStats:
    """Per-row statistics of the last fused forward (device tensors, no sync).

    p_target : softmax probability of the label column (0 where the label is not a column)
    p_max    : max probability over the first ``valid_cols`` columns (metrics.py:39)
    p_max_all: max probability over all columns; pred: first arg-max over all columns (util.py:109-110)
    loss_num / den: un-normalised loss sum and denominator of this shard (float64, device)
    label_counts: int64 [4] = one-hot rows, uniform-target rows, ignored rows, out-of-range labels
    """
    p_target: torch.Tensor
    p_max: torch.Tensor
    p_max_all: torch.Tensor
    pred: torch.Tensor
    loss_num: torch.Tensor
    den: torch.Tensor
    label_counts: torch.Tensor


def _round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def _as_bf16_rows(t: torch.Tensor, k64: int) -> torch.Tensor:
    """bf16, row-major, inner dimension zero-padded to k64 (no copy when already in that form)."""
    if t.shape[1] == k64:
        if t.dtype == torch.bfloat16 and t.stride(1) == 1 and t.stride(0) % 8 == 0 and t.data_ptr() % 16 == 0:
            return t
        return t.to(torch.bfloat16).contiguous()
    out = torch.zeros(t.shape[0], k64, dtype=torch.bfloat16, device=t.device)
    out[:, :t.shape[1]] = t
    return out


class _FusedHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, weight, bias, target, head):
        dev = L.require_cuda(features, weight, bias, target, head.class_weights)
        if features.dim() != 2 or features.shape[1] != weight.shape[1]:
            raise ValueError(f"features must be [batch, {weight.shape[1]}], got {tuple(features.shape)}")
        B, K = features.shape
        Cc = weight.shape[0]
        k64 = _round_up(K, 64)
        x = _as_bf16_rows(features.detach(), k64)
        w = head._weight_bf16(weight, k64)
        b = None if bias is None else bias.detach().float().contiguous()
        y = target if target.dtype == torch.int64 else target.long()
        y = y.contiguous()
        if y.dim() != 1 or y.shape[0] != B:
            raise ValueError(f"target must be [batch]={B}, got {tuple(y.shape)}")
        need_grad = any(ctx.needs_input_grad[:3])
        st = L.stream_ptr(dev)
        with torch.cuda.device(dev):
            ws = L.workspace("head", max(L.size("osb_head_workspace_bytes", B, k64, Cc),
                                         L.size("osb_loss_workspace_bytes", B)), dev)
            den = torch.empty(1, dtype=torch.float64, device=dev)
            counts = torch.empty(4, dtype=torch.int64, device=dev)
            L.call("osb_label_stats", L.ptr(y), L.ptr(head.class_weights), B, Cc, head.variant, head.ignore_index,
                   L.ptr(den), L.ptr(counts), L.ptr(ws), ws.numel(), st)
            if head.den_hook is not None:
                den = head.den_hook(den)
            loss = torch.empty(1, dtype=torch.float32, device=dev)
            num = torch.empty(1, dtype=torch.float64, device=dev)
            stats = torch.empty(3, B, dtype=torch.float32, device=dev)
            pred = torch.empty(B, dtype=torch.int32, device=dev)
            lddz = _round_up(Cc, 8)
            dz = torch.empty(B, lddz, dtype=torch.bfloat16, device=dev) if need_grad else None
            L.call("osb_head_forward", L.ptr(x), x.stride(0), L.ptr(w), L.ptr(b), L.ptr(y), L.ptr(head.class_weights),
                   B, k64, Cc, head.variant, float(head.unk_weight), head.ignore_index, L.ptr(den), head.valid_cols,
                   L.ptr(num), L.ptr(loss), L.ptr(stats[0]), L.ptr(stats[1]), L.ptr(stats[2]), L.ptr(pred),
                   L.ptr(dz), lddz, L.ptr(ws), ws.numel(), st)
        head.last_stats = HeadStats(stats[0], stats[1], stats[2], pred, num, den, counts)
        if head.validate_labels:
            bad = int(counts[3].item())
            if bad:
                raise IndexError(f"{bad} target label(s) out of range for {Cc} classes")
        if need_grad:
            ctx.save_for_backward(x, w, dz)
        ctx.meta = (B, K, k64, Cc, lddz, features.dtype, weight.dtype, bias is not None, None if bias is None else bias.dtype)
        return loss[0]

    @staticmethod
    def backward(ctx, grad_out):
        x, w, dz = ctx.saved_tensors
        B, K, k64, Cc, lddz, x_dtype, w_dtype, has_bias, b_dtype = ctx.meta
        dev = x.device
        need_dx, need_dw, need_db = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2] and has_bias
        g = grad_out.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        st = L.stream_ptr(dev)
        with torch.cuda.device(dev):
            ws = L.workspace("head", L.size("osb_head_workspace_bytes", B, k64, Cc), dev)
            dx = torch.empty(B, k64, dtype=torch.bfloat16, device=dev) if need_dx else None
            dw = torch.empty(Cc, k64, dtype=torch.float32, device=dev) if need_dw else None
            db = torch.empty(Cc, dtype=torch.float32, device=dev) if need_db else None
            L.call("osb_head_backward", L.ptr(dz), lddz, L.ptr(x), x.stride(0), L.ptr(w), B, k64, Cc, L.ptr(g),
                   L.ptr(dx), k64, L.ptr(dw), L.ptr(db), L.ptr(ws), ws.numel(), st)
        if dx is not None:
            dx = dx[:, :K] if k64 != K else dx
            dx = dx.to(x_dtype) if dx.dtype != x_dtype else dx
        if dw is not None:
            dw = dw[:, :K] if k64 != K else dw
            dw = dw.to(w_dtype) if dw.dtype != w_dtype else dw
        if db is not None and db.dtype != b_dtype:
            db = db.to(b_dtype)
        return dx, dw, db, None, None


class FusedOpenSetHead(nn.Linear):
    """``nn.Linear(in_features, out_features, bias)`` plus the fused loss path.

    Args beyond nn.Linear's:
        variant: 'entropic' (losses.py:7-29), 'softmax' (train.py:343) or 'garbage' (train.py:346-347).
        unk_weight: ``cfg.loss.w`` of the entropic loss.
        class_weights: float tensor [out_features] for the garbage variant (dataset.py:77-86).
        ignore_index: label skipped by the softmax / garbage variants (defaults: -1 / -100).
        last_valid_class: slice end for the confidence max, None (all columns) or -1 (drop the garbage
            column), as in ``metrics.confidence`` (train.py:156-163).
        validate_labels: read one counter back per call to raise the reference's IndexError for labels
            >= out_features (a host sync); off by default, bad labels then give a NaN loss.
    """

    def __init__(self, in_features, out_features, bias=False, variant="entropic", unk_weight=1.0,
                 class_weights=None, ignore_index=None, last_valid_class=None, validate_labels=False,
                 device=None, dtype=None):
        super().__init__(in_features, out_features, bias=bias, device=device, dtype=dtype)
        if variant not in _VARIANTS:
            raise ValueError(f"unknown variant {variant!r}")
        if variant == "garbage" and class_weights is None:
            raise ValueError("the garbage variant needs class_weights")
        self.variant_name = variant
        self.variant = _VARIANTS[variant]
        self.unk_weight = float(unk_weight)
        self.ignore_index = int(_DEFAULT_IGNORE[variant] if ignore_index is None else ignore_index)
        self.valid_cols = len(range(out_features)[slice(None, last_valid_class)])
        self.validate_labels = bool(validate_labels)
        self.den_hook = None           # set by dist.ShardedHead: all-reduce of the loss denominator
        self.last_stats: HeadStats | None = None
        if class_weights is not None:
            self.register_buffer("class_weights", class_weights.detach().float().contiguous(), persistent=False)
        else:
            self.class_weights = None
        self._w_cache = None           # (version, data_ptr, k64, bf16 tensor)

    def _weight_bf16(self, weight: torch.Tensor, k64: int) -> torch.Tensor:
        key = (weight._version, weight.data_ptr(), k64)
        if self._w_cache is None or self._w_cache[0] != key:
            self._w_cache = (key, _as_bf16_rows(weight.detach(), k64))
        return self._w_cache[1]

    def loss(self, features: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """Fused linear + loss: 0-dim fp32 tensor with grad_fn; per-row statistics in ``last_stats``."""
        return _FusedHeadFn.apply(features, self.weight, self.bias, target, self)


class ResNet50(nn.Module):
    """Drop-in for ``openset_imagenet.model.ResNet50`` (model.py:5-39): torchvision/cuDNN backbone whose
    ``fc`` maps 2048 -> fc_layer_dim, then the ``logits`` layer.  ``forward`` returns
    ``(logits, features)`` like the reference; ``forward_loss`` routes the logits layer through the fused
    tcgen05 head.  State-dict keys are the reference's (``resnet_base.*``, ``logits.weight``)."""

    def __init__(self, fc_layer_dim=1000, out_features=1000, logit_bias=True, **head_kwargs):
        super().__init__()
        from torchvision import models
        resnet_base = models.resnet50(weights=None)
        fc_in_features = resnet_base.fc.in_features
        resnet_base.fc = nn.Linear(in_features=fc_in_features, out_features=fc_layer_dim)
        self.resnet_base = resnet_base
        self.logits = FusedOpenSetHead(fc_layer_dim, out_features, bias=logit_bias, **head_kwargs)

    def forward(self, image):
        features = self.resnet_base(image)
        logits = self.logits(features)
        return logits, features

    def forward_loss(self, image, target):
        features = self.resnet_base(image)
        return self.logits.loss(features, target), features
