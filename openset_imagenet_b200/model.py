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
from .functional import resolve_planes

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
        # the bf16 copy of the module's own weight is cached per version; any other tensor (the per-step product matrix of
        # the collapsed head) is converted afresh
        w = head._weight_bf16(weight, k64) if head._own_weight else _as_bf16_rows(weight.detach(), k64)
        b = None if bias is None else bias.detach().float().contiguous()
        y = target if target.dtype == torch.int64 else target.long()
        y = y.contiguous()
        if y.dim() != 1 or y.shape[0] != B:
            raise ValueError(f"target must be [batch]={B}, got {tuple(y.shape)}")
        need_grad = any(ctx.needs_input_grad[:3])
        planes = resolve_planes(head.precision, Cc, bias is not None)
        st = L.stream_ptr(dev)
        with torch.cuda.device(dev):
            ws = L.workspace("head", L.size("osb_head_workspace_bytes", B, k64, Cc), dev)
            wsl = L.workspace("labels", L.size("osb_loss_workspace_bytes", B), dev)
            den = torch.empty(1, dtype=torch.float64, device=dev)
            counts = torch.empty(4, dtype=torch.int64, device=dev)
            L.call("osb_label_stats", L.ptr(y), L.ptr(head.class_weights), B, Cc, head.variant, head.ignore_index,
                   L.ptr(den), L.ptr(counts), L.ptr(wsl), wsl.numel(), 0, st)
            flags = L.FWD_DEN_FROM_PREV_KERNEL
            if head.den_hook is not None:
                den = head.den_hook(den)
                flags = 0
            loss = torch.empty(1, dtype=torch.float32, device=dev)
            num = torch.empty(1, dtype=torch.float64, device=dev)
            stats = torch.empty(3, B, dtype=torch.float32, device=dev)
            pred = torch.empty(B, dtype=torch.int32, device=dev)
            lddz = _round_up(Cc, 8)
            dz = torch.empty(planes, B, lddz, dtype=torch.bfloat16, device=dev) if need_grad else None
            L.call("osb_head_forward", L.ptr(x), x.stride(0), L.ptr(w), L.ptr(b), L.ptr(y), L.ptr(head.class_weights),
                   B, k64, Cc, head.variant, float(head.unk_weight), head.ignore_index, L.ptr(den), head.valid_cols,
                   L.ptr(num), L.ptr(loss), L.ptr(stats[0]), L.ptr(stats[1]), L.ptr(stats[2]), L.ptr(pred),
                   L.ptr(dz), lddz, planes, flags, 0, L.ptr(ws), ws.numel(), st)
        head.last_stats = HeadStats(stats[0], stats[1], stats[2], pred, num, den, counts)
        if head.validate_labels:
            bad = int(counts[3].item())
            if bad:
                raise IndexError(f"{bad} target label(s) out of range for {Cc} classes")
        if need_grad:
            ctx.save_for_backward(x, w, dz, den)
        ctx.meta = (B, K, k64, Cc, lddz, planes, features.dtype, weight.dtype, bias is not None, None if bias is None else bias.dtype)
        return loss[0]

    @staticmethod
    def backward(ctx, grad_out):
        x, w, dz, den = ctx.saved_tensors
        B, K, k64, Cc, lddz, planes, x_dtype, w_dtype, has_bias, b_dtype = ctx.meta
        dev = x.device
        need_dx, need_dw, need_db = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2] and has_bias
        g = grad_out.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        st = L.stream_ptr(dev)
        # bf16 features (an autocast backbone) get their gradient in bf16; anything wider gets the fp32 accumulator as is
        dx_t = torch.bfloat16 if x_dtype == torch.bfloat16 else torch.float32
        with torch.cuda.device(dev):
            ws = L.workspace("head", L.size("osb_head_workspace_bytes", B, k64, Cc), dev)
            dx = torch.empty(B, k64, dtype=dx_t, device=dev) if need_dx else None
            dw = torch.empty(Cc, k64, dtype=torch.float32, device=dev) if need_dw else None
            db = torch.empty(Cc, dtype=torch.float32, device=dev) if need_db else None
            L.call("osb_head_backward", L.ptr(dz), lddz, planes, L.ptr(den), L.ptr(x), x.stride(0), L.ptr(w), B, k64, Cc, L.ptr(g),
                   L.ptr(dx), k64, L.F32 if dx_t == torch.float32 else L.BF16, L.ptr(dw), L.ptr(db), 0, 0,
                   L.ptr(ws), ws.numel(), st)
        if dx is not None:
            dx = dx[:, :K] if k64 != K else dx
            dx = dx.to(x_dtype) if dx.dtype != x_dtype else dx
        if dw is not None:
            dw = dw[:, :K] if k64 != K else dw
            dw = dw.to(w_dtype) if dw.dtype != w_dtype else dw
        if db is not None and db.dtype != b_dtype:
            db = db.to(b_dtype)
        return dx, dw, db, None, None


_AUTO = "auto"


class FusedOpenSetHead(nn.Linear):
    """``nn.Linear(in_features, out_features, bias)`` plus the fused loss path.

    Args beyond nn.Linear's:
        variant: 'entropic' (losses.py:7-29), 'softmax' (train.py:343) or 'garbage' (train.py:346-347).
        unk_weight: ``cfg.loss.w`` of the entropic loss.
        class_weights: float tensor [out_features] for the garbage variant (dataset.py:77-86).
        ignore_index: label skipped by the softmax / garbage variants (defaults: -1 / -100).
        last_valid_class: slice end for the confidence max, as in ``metrics.confidence``: None (all columns) or -1
            (drop the garbage column).  Default 'auto' = what ``train.validate`` passes for the variant
            (train.py:156-163): -1 for 'garbage', None otherwise.
        validate_labels: read one counter back per call to raise the reference's IndexError for labels
            >= out_features (a host sync); off by default, bad labels then give a NaN loss.
        precision: 'exact' (default) | 'fast' (``functional.resolve_planes``): how the grad-logits reach the backward
            GEMMs.  'exact' keeps their bf16 rounding residual in a second plane (gradients ~1e-5 from float64, the
            1e-3 contract holds for every shape); 'fast' is one bf16 plane (6e-4 .. 1.5e-3 norm-wise, half the
            backward MMAs: worth it only for wide heads on large batches).

    The bf16 copy of the weights the kernels read is cached per weight VERSION (``Tensor._version``): every
    in-place update through the autograd-visible tensor (optimizers, ``copy_`` under ``no_grad``) refreshes it.
    Writes through ``weight.data`` do not bump the version; call ``invalidate_weight_cache()`` after them.
    """

    def __init__(self, in_features, out_features, bias=False, variant="entropic", unk_weight=1.0,
                 class_weights=None, ignore_index=None, last_valid_class=_AUTO, validate_labels=False,
                 precision="exact", device=None, dtype=None):
        super().__init__(in_features, out_features, bias=bias, device=device, dtype=dtype)
        if variant not in _VARIANTS:
            raise ValueError(f"unknown variant {variant!r}")
        if variant == "garbage" and class_weights is None:
            raise ValueError("the garbage variant needs class_weights")
        resolve_planes(precision, out_features, bias)        # validates the name
        self.variant_name = variant
        self.variant = _VARIANTS[variant]
        self.unk_weight = float(unk_weight)
        self.ignore_index = int(_DEFAULT_IGNORE[variant] if ignore_index is None else ignore_index)
        if isinstance(last_valid_class, str):
            if last_valid_class != _AUTO:
                raise ValueError("last_valid_class must be None, an int or 'auto'")
            last_valid_class = -1 if variant == "garbage" else None
        self.last_valid_class = last_valid_class
        self.valid_cols = len(range(out_features)[slice(None, last_valid_class)])
        self.validate_labels = bool(validate_labels)
        self.precision = precision
        self.den_hook = None           # set by dist.ShardedHead: all-reduce of the loss denominator
        self.last_stats: HeadStats | None = None
        if class_weights is not None:
            self.register_buffer("class_weights", class_weights.detach().float().contiguous(), persistent=False)
        else:
            self.class_weights = None
        self._w_cache = None           # (version, data_ptr, k64, bf16 tensor)
        self._own_weight = True

    def invalidate_weight_cache(self) -> None:
        """Drop the cached bf16 weights (needed only after a write through ``weight.data``, which no version counter sees)."""
        self._w_cache = None

    def _weight_bf16(self, weight: torch.Tensor, k64: int) -> torch.Tensor:
        key = (weight._version, weight.data_ptr(), k64)
        if self._w_cache is None or self._w_cache[0] != key:
            self._w_cache = (key, _as_bf16_rows(weight.detach(), k64))
        return self._w_cache[1]

    def loss(self, features: torch.Tensor, target: torch.Tensor, weight=None, bias=None) -> torch.Tensor:
        """Fused linear + loss: 0-dim fp32 tensor with grad_fn; per-row statistics in ``last_stats``.
        ``weight`` / ``bias`` override the module's own parameters (the collapsed two-layer head of ``ResNet50``)."""
        self._own_weight = weight is None
        if weight is None:
            weight, bias = self.weight, self.bias
        return _FusedHeadFn.apply(features, weight, bias, target, self)


class ResNet50(nn.Module):
    """Drop-in for ``openset_imagenet.model.ResNet50`` (model.py:5-39): torchvision/cuDNN backbone whose
    ``fc`` maps 2048 -> fc_layer_dim, then the ``logits`` layer.  ``forward`` returns
    ``(logits, features)`` like the reference.  State-dict keys are the reference's (``resnet_base.*``,
    ``logits.weight``), so its checkpoints load with ``strict=True`` (train.py:63-101).

    ``forward_loss(image, target)`` is the B200 path.  ``fused`` selects what runs on the tcgen05 head:
      * 'logits'    : only the ``logits`` linear (fc_layer_dim -> out_features) + loss; ``resnet_base.fc`` stays cuBLAS;
      * 'collapsed' : (default) the two stacked linears have no non-linearity between them (model.py:37-38), so
        ``logits(fc(x)) = x (W2 W1)^T + W2 b1 (+ b2)``: the 2048 -> out_features product matrix is formed per step
        (a [C, fc_dim] x [fc_dim, 2048] product of PARAMETERS, independent of the batch) and the whole
        2048 -> out_features GEMM, the loss and both backward GEMMs run as the fused head on the 2048-d pooled
        features; autograd carries dW_eff back into ``logits.weight``, ``resnet_base.fc.weight`` and ``.bias``.
    """

    def __init__(self, fc_layer_dim=1000, out_features=1000, logit_bias=True, fused="collapsed", **head_kwargs):
        super().__init__()
        from torchvision import models
        if fused not in ("logits", "collapsed"):
            raise ValueError("fused must be 'logits' or 'collapsed'")
        resnet_base = models.resnet50(weights=None)
        fc_in_features = resnet_base.fc.in_features
        resnet_base.fc = nn.Linear(in_features=fc_in_features, out_features=fc_layer_dim)
        self.resnet_base = resnet_base
        self.logits = FusedOpenSetHead(fc_layer_dim, out_features, bias=logit_bias, **head_kwargs)
        self.fused = fused

    def forward(self, image):
        features = self.resnet_base(image)
        logits = self.logits(features)
        return logits, features

    def pooled_features(self, image):
        """The 2048-d input of ``resnet_base.fc`` (torchvision's ResNet forward up to the flatten)."""
        rb = self.resnet_base
        x = rb.maxpool(rb.relu(rb.bn1(rb.conv1(image))))
        x = rb.layer4(rb.layer3(rb.layer2(rb.layer1(x))))
        return torch.flatten(rb.avgpool(x), 1)

    def collapsed_weights(self):
        """(W_eff [out, 2048], b_eff [out]) with autograd back to the three parameter tensors."""
        fc, lg = self.resnet_base.fc, self.logits
        w_eff = lg.weight.float() @ fc.weight.float()
        b_eff = lg.weight.float() @ fc.bias.float()
        if lg.bias is not None:
            b_eff = b_eff + lg.bias.float()
        return w_eff, b_eff

    def forward_loss(self, image, target, return_features=False):
        """(loss, features): ``features`` is the reference's second output (fc_layer_dim deep features) when
        ``return_features`` (or in 'logits' mode, where it is computed anyway), else None."""
        if self.fused == "logits":
            features = self.resnet_base(image)
            return self.logits.loss(features, target), features
        pooled = self.pooled_features(image)
        with torch.autocast("cuda", enabled=False):
            w_eff, b_eff = self.collapsed_weights()
        loss = self.logits.loss(pooled, target, weight=w_eff, bias=b_eff)
        features = self.resnet_base.fc(pooled) if return_features else None
        return loss, features
