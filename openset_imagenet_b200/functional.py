"""Allocation-free entry to the fused head: a ``HeadPlan`` owns every buffer one (B, K, C) step needs and
calls the C ABI directly, so a training loop (or a CUDA graph) replays ``forward`` / ``backward`` with no
per-step allocations, no autograd bookkeeping and no host synchronisation.

``model.FusedOpenSetHead.loss`` is the autograd-friendly wrapper over the same three entry points.
"""
from __future__ import annotations

import torch

from . import _lib as L

_VARIANTS = {"entropic": L.ENTROPIC, "softmax": L.SOFTMAX, "garbage": L.GARBAGE}
_DEFAULT_IGNORE = {"entropic": -100, "softmax": -1, "garbage": -100}


def resolve_planes(precision: str, C: int = 0, bias: bool = False) -> int:
    """dZ planes for a precision mode.
    'exact' (default): the bf16 rounding residual of the grad-logits travels in a second plane and the backward GEMMs
        accumulate both in fp32 -- dW / db / dX within ~1e-5 of float64, at twice the backward MMA count (free for the
        launch-latency-bound protocol heads, ~1.6x the step time of the 1000-class stress configuration);
    'fast': one bf16 plane, i.e. plain bf16 tensor-core operands: 6e-4 .. 1.5e-3 norm-wise in dW / dX depending on the
        shape (measured), more in db, whose column sums cancel.  The benchmarked mode of the large-head configuration."""
    if precision == "exact":
        return 2
    if precision == "fast":
        return 1
    raise ValueError(f"precision must be 'exact' or 'fast', got {precision!r}")


class HeadPlan:
    """Buffers + launch arguments for linear(K -> C) + open-set loss + backward on a fixed batch size.

    Outputs (device tensors, overwritten by every call):
      loss [1] f32, loss_num [1] f64, den [1] f64, label_counts [4] i64,
      p_target / p_max / p_max_all [B] f32, pred [B] i32,
      dz [planes, B, round_up(C,8)] bf16 (denominator-free grad-logits; plane 1 = bf16 rounding residual of plane 0),
      dx [B, K] bf16 or f32, dw [C, K] f32, db [C] f32 (when bias).
    """

    def __init__(self, B: int, K: int, C: int, variant: str = "entropic", unk_weight: float = 1.0,
                 class_weights: torch.Tensor | None = None, ignore_index: int | None = None,
                 last_valid_class=None, bias: bool = False, device=None, grad_buffer: torch.Tensor | None = None,
                 workspace: torch.Tensor | None = None, precision: str = "exact", dx_dtype: torch.dtype = torch.bfloat16,
                 n_tail: int = 1):
        if K % 64 != 0:
            raise ValueError("HeadPlan needs in_features % 64 == 0 (FusedOpenSetHead pads other widths)")
        if dx_dtype not in (torch.bfloat16, torch.float32):
            raise ValueError("dx_dtype must be torch.bfloat16 or torch.float32")
        self.dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.B, self.K, self.C = int(B), int(K), int(C)
        self.variant_name = variant
        self.variant = _VARIANTS[variant]
        self.unk_weight = float(unk_weight)
        self.ignore_index = int(_DEFAULT_IGNORE[variant] if ignore_index is None else ignore_index)
        self.valid_cols = len(range(C)[slice(None, last_valid_class)])
        self.class_w = None if class_weights is None else class_weights.detach().to(self.dev).float().contiguous()
        self.planes = resolve_planes(precision, C, bias)
        self.dx_dtype = dx_dtype
        d = self.dev
        self.lddz = (C + 7) // 8 * 8
        ws_bytes = max(L.size("osb_head_workspace_bytes", B, K, C), L.size("osb_loss_workspace_bytes", B)) + 4096
        if workspace is None:
            workspace = torch.zeros(ws_bytes, dtype=torch.uint8, device=d)
        if workspace.dtype != torch.uint8 or workspace.numel() < ws_bytes:
            raise ValueError(f"workspace must be a zero-filled uint8 tensor of at least {ws_bytes} bytes")
        self.ws = workspace
        self.ws_bytes_needed = ws_bytes
        self.den = torch.empty(1, dtype=torch.float64, device=d)
        # entropic loss: the denominator is the row count (losses.py:29), known without reading the labels
        self.den_const = torch.full((1,), float(B), dtype=torch.float64, device=d)
        # the label statistics have their own workspace: they may run beside the forward kernel
        self.ws_label = torch.zeros(L.size("osb_loss_workspace_bytes", B) + 256, dtype=torch.uint8, device=d)
        self.label_counts = torch.empty(4, dtype=torch.int64, device=d)
        # [dW | db | tail] may live in one caller-owned float32 buffer (the packed all-reduce buffer of dist.py), so
        # that the kernels write straight into it and no gather copy is needed; tail[0] is the loss
        n_flat = C * K + (C if bias else 0) + int(n_tail)
        if grad_buffer is None:
            grad_buffer = torch.empty(n_flat, dtype=torch.float32, device=d)
        if grad_buffer.dtype != torch.float32 or grad_buffer.numel() < n_flat or not grad_buffer.is_contiguous():
            raise ValueError(f"grad_buffer must be a contiguous float32 tensor with at least {n_flat} elements")
        self.flat = grad_buffer[:n_flat]
        self.n_tail = int(n_tail)
        self.loss = self.flat[n_flat - self.n_tail:n_flat - self.n_tail + 1]
        self.loss_num = torch.empty(1, dtype=torch.float64, device=d)
        self.stats = torch.empty(3, B, dtype=torch.float32, device=d)
        self.p_target, self.p_max, self.p_max_all = self.stats[0], self.stats[1], self.stats[2]
        self.pred = torch.empty(B, dtype=torch.int32, device=d)
        self.dz = torch.empty(self.planes, B, self.lddz, dtype=torch.bfloat16, device=d)
        self.dx = torch.empty(B, K, dtype=dx_dtype, device=d)
        self.dw = self.flat[:C * K].view(C, K)
        self.db = self.flat[C * K:C * K + C] if bias else None
        self.dw_direct = bool(L.size("osb_head_dw_is_direct", B, K, C))
        self.fused_dx = bool(L.size("osb_head_forward_dx_is_fused", B, K, C))
        # label_stats, fwd, dx, dw (+ the split-K reduce unless the GEMM writes dW directly) (+ db); a captured step runs
        # forward and dX as one kernel where the shape allows
        self.launches_per_step = 4 + (0 if self.dw_direct else 1) + (1 if bias else 0)
        self.launches_per_captured_step = self.launches_per_step        # set by capture()

    @property
    def den_step(self) -> torch.Tensor:
        """The denominator the kernels of a single-GPU step divide by."""
        return self.den_const if self.variant == L.ENTROPIC else self.den

    def _check(self, x, w, y):
        if x.shape != (self.B, self.K) or x.dtype != torch.bfloat16 or not x.is_contiguous():
            raise ValueError("x must be a contiguous bf16 [B, K] tensor")
        if w.shape != (self.C, self.K) or w.dtype != torch.bfloat16 or not w.is_contiguous():
            raise ValueError("w must be a contiguous bf16 [C, K] tensor")
        if y.shape != (self.B,) or y.dtype != torch.int64:
            raise ValueError("y must be int64 [B]")

    def label_stats(self, y, den_scale_hook=None, max_ctas=0):
        L.call("osb_label_stats", L.ptr(y), L.ptr(self.class_w), self.B, self.C, self.variant, self.ignore_index,
               L.ptr(self.den), L.ptr(self.label_counts), L.ptr(self.ws_label), self.ws_label.numel(), int(max_ctas),
               L.stream_ptr(self.dev))
        if den_scale_hook is not None:
            den_scale_hook(self.den)

    def forward(self, x, w, y, bias=None, write_dz=True, want_pred=True, den="step", flags=0, max_ctas=0):
        """``den``: device double the loss is divided by ("step": this plan's own, None: ``loss`` = the numerator)."""
        if isinstance(den, str):
            den = self.den_step
        L.call("osb_head_forward", L.ptr(x), self.K, L.ptr(w), L.ptr(bias), L.ptr(y), L.ptr(self.class_w),
               self.B, self.K, self.C, self.variant, self.unk_weight, self.ignore_index,
               L.ptr(den), self.valid_cols,
               L.ptr(self.loss_num), L.ptr(self.loss), L.ptr(self.p_target), L.ptr(self.p_max),
               L.ptr(self.p_max_all) if want_pred else None, L.ptr(self.pred) if want_pred else None,
               L.ptr(self.dz) if write_dz else None, self.lddz, self.planes, int(flags), int(max_ctas),
               L.ptr(self.ws), self.ws.numel(), L.stream_ptr(self.dev))

    def forward_dx(self, x, w, y, bias=None, gscale=None, den="step", flags=0, max_ctas=0):
        """Forward and dX of a training step from ONE call (``osb_head_forward_dx``): one kernel where ``self.fused_dx``,
        whose CTA pairs go on with dX tiles under the forward's last epilogue; no per-row predictions."""
        if isinstance(den, str):
            den = self.den_step
        L.call("osb_head_forward_dx", L.ptr(x), self.K, L.ptr(w), L.ptr(bias), L.ptr(y), L.ptr(self.class_w),
               self.B, self.K, self.C, self.variant, self.unk_weight, self.ignore_index, L.ptr(den), self.valid_cols,
               L.ptr(self.loss_num), L.ptr(self.loss), L.ptr(self.p_target), L.ptr(self.p_max),
               L.ptr(self.dz), self.lddz, self.planes, L.ptr(gscale), L.ptr(self.dx), self.K, self._dx_code(),
               int(flags), int(max_ctas), L.ptr(self.ws), self.ws.numel(), L.stream_ptr(self.dev))

    def _dx_code(self):
        return L.F32 if self.dx_dtype == torch.float32 else L.BF16

    def backward_dx(self, x, w, gscale=None, den=None, max_ctas=0):
        L.call("osb_head_backward", L.ptr(self.dz), self.lddz, self.planes, L.ptr(self.den_step if den is None else den),
               L.ptr(x), self.K, L.ptr(w), self.B, self.K, self.C, L.ptr(gscale), L.ptr(self.dx), self.K, self._dx_code(),
               None, None, int(max_ctas), 0, L.ptr(self.ws), self.ws.numel(), L.stream_ptr(self.dev))

    def backward(self, x, w, gscale=None, den=None, max_ctas=0):
        """dX, then dW (+ db), from ONE call: with two batch splits the dX kernel clears dW on its way in and both splits of
        the dW GEMM add their tile at the destination without waiting for each other (``osb_head_backward`` with dx and dw)."""
        L.call("osb_head_backward", L.ptr(self.dz), self.lddz, self.planes, L.ptr(self.den_step if den is None else den),
               L.ptr(x), self.K, L.ptr(w), self.B, self.K, self.C, L.ptr(gscale), L.ptr(self.dx), self.K, self._dx_code(),
               L.ptr(self.dw), L.ptr(self.db), int(max_ctas), 0, L.ptr(self.ws), self.ws.numel(), L.stream_ptr(self.dev))

    def backward_dw(self, x, w, gscale=None, den=None, flags=0):
        """``flags=L.BWD_DZ_FROM_PREV_FORWARD``: this launch is directly behind ``forward(..., flags=L.FWD_PUBLISH_ROW_TILES)``
        on the stream; the dW GEMM then starts on the SMs the forward leaves idle and takes dZ row tile by row tile."""
        L.call("osb_head_backward", L.ptr(self.dz), self.lddz, self.planes, L.ptr(self.den_step if den is None else den),
               L.ptr(x), self.K, L.ptr(w), self.B, self.K, self.C, L.ptr(gscale), None, self.K, self._dx_code(),
               L.ptr(self.dw), L.ptr(self.db), 0, int(flags), L.ptr(self.ws), self.ws.numel(), L.stream_ptr(self.dev))

    def backward_dw_partials(self, x, local_reduce=False, flags=0):
        """Only the dW GEMM: un-scaled split-K partials stay in the workspace (input of the peer-memory all-reduce),
        summed into the first one when ``local_reduce``; with a bias the un-scaled column sums of dZ go to the db slots
        of the packed buffer.  ``flags`` as in ``backward_dw``."""
        L.call("osb_head_backward_dw_partials", L.ptr(self.dz), self.lddz, self.planes, L.ptr(x), self.K, self.B, self.K, self.C,
               L.ptr(self.db), 1 if local_reduce else 0, int(flags), L.ptr(self.ws), self.ws.numel(), L.stream_ptr(self.dev))

    def capture(self, x, w, y, bias=None, gscale=None, parallel_backward=None, handover=False, fused_dx=True):
        """Capture one whole step on (x, w, y) into a CUDA graph and return it; ``graph.replay()`` re-runs its kernels
        (``launches_per_captured_step`` of them: three at the benchmarked shape -- label statistics, forward + dX as one
        kernel, dW GEMM) on the CURRENT CONTENTS of those tensors with a single launch call.  The small-batch
        configurations (B <= 1024) are bound by launch latency, not by the kernels, so this is the path to use in a
        training loop.  ``parallel_backward`` / ``handover`` / ``fused_dx`` override the launch form chosen from the
        batch size; every form gives the same bits."""
        self._check(x, w, y)
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        # Small batches leave most SMs idle and every kernel is a few microseconds of latency: dX and dW only depend on
        # dZ, so they become parallel branches of the graph.  Large batches fill the GPU with either kernel and stay serial:
        # under a full load the chip is power-limited (the SM clock falls from 1.85 to 1.58 GHz between 16 and 148 busy
        # SMs within one launch), so kernels that overlap only share the same budget -- measured at 16384 x 2048 -> 1000,
        # interleaved runs on one box: serial 182.9 us, parallel branches with the dW GEMM launched first 182.0 us, the same
        # with the row-tile hand-over (``handover=True``: the first dW CTA pairs run under the forward on the SMs it leaves
        # idle and are done 25 us before the others) 187.1 us; time lines in profiles/r02_probe_power_and_fill.txt.
        fork = self.B <= 4096 if parallel_backward is None else bool(parallel_backward)
        pub = L.FWD_PUBLISH_ROW_TILES if (fork and handover) else 0
        take = L.BWD_DZ_FROM_PREV_FORWARD if (fork and handover) else 0
        branch = torch.cuda.Stream(device=self.dev) if fork else None
        # Entropic loss: the denominator is a constant, so the label statistics (counts for the caller) are a branch
        # beside the backward kernels instead of the first link of the chain.
        side_labels = self.variant == L.ENTROPIC
        lbranch = torch.cuda.Stream(device=self.dev) if side_labels else None

        def labels_beside():
            # at most 8 CTAs: next to the dW GEMM (64 CTA pairs on 74 SM pairs at the large configuration) they take
            # SMs nobody else wants, whichever kernel the hardware places first
            lbranch.wait_stream(side)
            with torch.cuda.stream(lbranch):
                self.label_stats(y, max_ctas=8)

        # Large batches (serial order): forward and dX as ONE kernel where the shape allows.  Neutral to -3.5 % at 8192 rows
        # and more, slower below (tools/fused_shapes.py), where the parallel branches are the default anyway.
        fuse = self.fused_dx and not fork and fused_dx
        self.launches_per_captured_step = self.launches_per_step - (1 if fuse else 0)

        def body():
            if not side_labels:
                # the forward kernel runs under the label statistics it follows and waits for them only before it
                # divides the loss (OSB_FWD_DEN_FROM_PREV_KERNEL); the grad-logits do not depend on the denominator
                self.label_stats(y)
                if fuse:
                    self.forward_dx(x, w, y, bias, gscale, flags=L.FWD_DEN_FROM_PREV_KERNEL)
                else:
                    self.forward(x, w, y, bias, want_pred=False, flags=L.FWD_DEN_FROM_PREV_KERNEL | pub)
            elif fuse:
                self.forward_dx(x, w, y, bias, gscale)
            else:
                self.forward(x, w, y, bias, want_pred=False, flags=pub)
            if fork:
                if side_labels:
                    labels_beside()
                # dX on the branch; the longer chain (dW GEMM + split-K reduce) stays on the stream of the forward kernel,
                # where each kernel is launched programmatically behind the one in front of it
                branch.wait_stream(side)
                if handover:
                    # both kernels are programmatic dependents of the forward and the one launched first takes the SMs
                    # first: the dW GEMM, which consumes dZ as it is published
                    self.backward_dw(x, w, gscale, flags=take)
                    with torch.cuda.stream(branch):
                        self.backward_dx(x, w, gscale)
                else:
                    with torch.cuda.stream(branch):
                        self.backward_dx(x, w, gscale)
                    self.backward_dw(x, w, gscale)
                side.wait_stream(branch)
            else:
                if side_labels:
                    labels_beside()
                if fuse:
                    self.backward_dw(x, w, gscale)          # dX came out of the forward's kernel
                else:
                    self.backward(x, w, gscale)
            if side_labels:
                side.wait_stream(lbranch)

        with torch.cuda.stream(side):
            body()                                           # warm-up outside capture (function attributes, descriptors)
            side.synchronize()
            for b_ in (branch, lbranch):
                if b_ is not None:
                    b_.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                body()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        return graph

    def step(self, x, w, y, bias=None, gscale=None, den_scale_hook=None, check=False, want_pred=False):
        """label stats -> fused forward -> dX -> dW (+db): the whole head fwd+bwd as separate launches (4 to 6).  A training step
        has no use for the per-row arg-max / max probability (the reference's train loop reads the loss only,
        train.py:124-139), so they are off unless ``want_pred``; p_target and p_max are always written."""
        if check:
            self._check(x, w, y)
        self.label_stats(y, den_scale_hook)
        den = self.den if den_scale_hook is not None else None
        self.forward(x, w, y, bias, want_pred=want_pred, den="step" if den is None else den)
        self.backward(x, w, gscale, den=den)


class OscrPlan:
    """Buffers + launch arguments for OSCR over a fixed number of samples: ``run(scores, gt)`` enqueues the two
    kernels (row reduce, cooperative sort + scan) with no allocation and no host synchronisation, so an
    evaluation loop (or a CUDA graph) can replay it.  Outputs (device tensors, overwritten by every call):
      taus [N] f32, ccr_cnt / fpr_cnt [N] i64, ccr / fpr [N] f64, meta [5] i64 =
      (n_unique, n_known, n_correct, n_unk, n_bad); the reference's arrays are the first ``meta[0] - 1`` entries.
    """

    _LABELS = {torch.int64: L.LABEL_I64, torch.float32: L.LABEL_F32, torch.int32: L.LABEL_I32}

    def __init__(self, N: int, C: int, unk_label: int = -1, device=None, counts: bool = True):
        """``counts=False`` computes only what ``util.calculate_oscr`` of the reference returns (ccr, fpr); the
        thresholds and the integer counts behind them (``taus``, ``ccr_cnt``, ``fpr_cnt``) are then ``None``."""
        self.dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.N, self.C, self.unk_label = int(N), int(C), int(unk_label)
        d, cap = self.dev, max(int(N), 1)
        self.taus = torch.empty(cap, dtype=torch.float32, device=d) if counts else None
        self.ccr_cnt = torch.empty(cap, dtype=torch.int64, device=d) if counts else None
        self.fpr_cnt = torch.empty(cap, dtype=torch.int64, device=d) if counts else None
        self.ccr = torch.empty(cap, dtype=torch.float64, device=d)
        self.fpr = torch.empty(cap, dtype=torch.float64, device=d)
        self.meta = torch.zeros(5, dtype=torch.int64, device=d)
        self.ws = torch.zeros(L.size("osb_oscr_workspace_bytes", self.N) + 256, dtype=torch.uint8, device=d)
        self.launches_per_call = 2

    def run(self, scores: torch.Tensor, gt: torch.Tensor):
        if scores.shape != (self.N, self.C) or scores.dtype != torch.float32 or scores.stride(1) != 1:
            raise ValueError("scores must be a float32 [N, C] tensor with unit column stride")
        if gt.shape != (self.N,) or gt.dtype not in self._LABELS or not gt.is_contiguous():
            raise ValueError("gt must be a contiguous int64 / int32 / float32 [N] tensor")
        L.call("osb_oscr", L.ptr(scores), scores.stride(0), L.ptr(gt), self._LABELS[gt.dtype], self.N, self.C, self.unk_label,
               L.ptr(self.taus), L.ptr(self.ccr_cnt), L.ptr(self.fpr_cnt), L.ptr(self.ccr), L.ptr(self.fpr), L.ptr(self.meta),
               L.ptr(self.ws), self.ws.numel(), L.stream_ptr(self.dev))
        return self
