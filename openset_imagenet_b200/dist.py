"""Batch-sharded head and sharded OSCR over the GPUs of one box (one process per GPU, NCCL over NVLink).

The reference has no collectives at all (its multi-GPU mode is one independent training job per GPU,
script/train_all.py:96-118); what is exchanged here follows from the arithmetic:

* head: rows are independent given W, so the batch shards freely.  The loss is a global mean, so the
  denominator must be global: B_global for the entropic loss (known without communication), one
  all-reduced double for the softmax / garbage variants.  After backward ONE all-reduce(sum) of the
  fp32 [dW | db | loss_num] buffer makes every rank hold the full-batch gradient; dX stays local.
* OSCR: stage A (97 % of the bytes) shards by rows.  Thresholds are the global set of unique known
  target scores, so the known records are all-gathered (8 B per sample slot); every rank sorts its own
  unk records together with ALL known records (remote ones demoted to "threshold only"), derives the
  identical threshold list with its own counts, and ONE all-reduce(sum) of the int64 counts merges the
  shards.  Counts at a fixed threshold are additive across shards, so the result is
  bit-identical to the single-GPU one.

All of this works on any backend: the CPU tests run the control flow with gloo (world size 2) on fake
stage results; the GPU path uses nccl.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def bind_host_thread_to_gpu_node(device_index: int | None = None) -> bool:
    """Pin the calling host thread to the CPUs next to this rank's GPU (NVML's ideal CPU affinity) so that pinned host
    buffers allocated afterwards -- and the copies out of them -- stay on the GPU's own NUMA node / PCIe root complex.
    With eight ranks on a two-socket box the default placement sends most host-to-device copies across the socket
    interconnect.  Returns False (and changes nothing) when NVML is not available."""
    try:
        import pynvml
        pynvml.nvmlInit()
        idx = torch.cuda.current_device() if device_index is None else int(device_index)
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(idx))
        return True
    except Exception:
        return False


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


# ------------------------------------------------------------------------------------------------
# head
# ------------------------------------------------------------------------------------------------
class ShardedHead:
    """Wraps a ``FusedOpenSetHead`` whose batch is sharded across the process group.

    ``loss(features, target)`` returns this shard's contribution ``loss_num_local / den_global``; summing
    it over ranks gives the reference's full-batch loss, and ``backward()`` then yields
    ``d(full loss)/d(local features)`` plus this shard's part of dW/db.  ``allreduce_grads()`` sums the
    packed [dW | db | loss] buffer with one NCCL all-reduce.
    """

    def __init__(self, head, group=None):
        self.head = head
        self.group = group
        head.den_hook = self._global_den
        self._flat = None

    def _global_den(self, den: torch.Tensor) -> torch.Tensor:
        # shards may differ in size, so even the entropic row count is summed
        _, ws = world()
        if ws > 1:
            dist.all_reduce(den, op=dist.ReduceOp.SUM, group=self.group)
        return den

    def loss(self, features, target):
        return self.head.loss(features, target)

    def pack_grads(self, local_loss: torch.Tensor) -> torch.Tensor:
        parts = [self.head.weight.grad.reshape(-1)]
        if self.head.bias is not None and self.head.bias.grad is not None:
            parts.append(self.head.bias.grad.reshape(-1))
        parts.append(local_loss.detach().reshape(1).float())
        n = sum(p.numel() for p in parts)
        if self._flat is None or self._flat.numel() != n:
            self._flat = torch.empty(n, dtype=torch.float32, device=parts[0].device)
        torch.cat(parts, out=self._flat)
        return self._flat

    def allreduce_grads(self, local_loss: torch.Tensor, async_op: bool = False):
        """One all-reduce for everything the optimiser and the logger need.  Returns (work, unpack)."""
        flat = self.pack_grads(local_loss)
        _, ws = world()
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op) if ws > 1 else None

        def unpack():
            if work is not None and async_op:
                work.wait()
            off = 0
            w = self.head.weight
            w.grad.copy_(flat[off:off + w.numel()].view_as(w))
            off += w.numel()
            if self.head.bias is not None and self.head.bias.grad is not None:
                b = self.head.bias
                b.grad.copy_(flat[off:off + b.numel()])
                off += b.numel()
            return flat[off]

        return work, unpack


class PeerGradAllReduce:
    """The head's dW all-reduce over NVLink peer memory, fused with the split-K reduction (``osb_head_dw_allreduce``).

    Every pointer table lists, per rank, the DEVICE address of that rank's buffer as mapped into this process:
    the head workspace (split-K partials), the packed ``[dW | db | tail]`` buffer, and a small flag area.
    ``from_symmetric_memory`` builds them with ``torch.distributed._symmetric_memory``; tests on one GPU pass
    plain pointers.  ``launch()`` enqueues ONE kernel on the current stream; it has no per-step argument (the epoch
    lives in the flag area), so it can be captured into a CUDA graph.  The result is bit-identical on every rank.
    ``extra_sum`` = [db summed and scaled like dW | plain sums of the tail (loss numerators)].
    """

    def __init__(self, plan, ws_ptrs, flat_ptrs, flag_ptrs, rank, world, max_ctas=32, flags_local=None):
        from . import _lib as L
        self._L = L
        self.plan, self.rank, self.world, self.max_ctas = plan, int(rank), int(world), int(max_ctas)
        off = L.size("osb_head_dw_partials_offset")
        n = plan.C * plan.K
        self.n_extra = plan.flat.numel() - n
        self.n_plain = plan.n_tail
        P = L.C.c_void_p * world
        self._partials = P(*[int(p) + off for p in ws_ptrs])
        self._dw = P(*[int(p) for p in flat_ptrs])
        self._extra = P(*[int(p) + 4 * n for p in flat_ptrs])
        self._flags = P(*[int(p) for p in flag_ptrs])
        self._flags_local = int(flag_ptrs[rank])
        # many splits (small batches): they are summed on their own GPU first, ONE matrix per rank crosses NVLink; the two
        # splits of a large batch are pulled as they are (the reduce is fused into the all-reduce, one pass less)
        self.gemm_splits = L.size("osb_head_dw_splits", plan.B, plan.K, plan.C)
        self.local_reduce = self.gemm_splits > 2
        self.splits = 1 if self.local_reduce else self.gemm_splits
        self.n = n
        self.extra_sum = torch.zeros(max(self.n_extra, 1), dtype=torch.float32, device=plan.dev)
        self.mc_partials = self.mc_dw = None          # multicast addresses (NVSwitch), set by from_symmetric_memory
        self.stream = torch.cuda.Stream(device=plan.dev)
        self._ready = torch.cuda.Event()
        self._done = torch.cuda.Event()
        self._keep = None

    @classmethod
    def from_symmetric_memory(cls, B, K, C, variant, group=None, max_ctas=32, device=None, multicast=True, **plan_kwargs):
        """Allocates workspace, gradient buffer, flags and denominator slots in symmetric memory and builds the
        ``HeadPlan`` on them."""
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib as L
        from .functional import HeadPlan
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        rank, ws_ = world()
        grp = group if group is not None else dist.group.WORLD
        bias = bool(plan_kwargs.get("bias", False))
        n_tail = int(plan_kwargs.pop("n_tail", 1))
        n_flat = C * K + (C if bias else 0) + n_tail
        ws_bytes = max(L.size("osb_head_workspace_bytes", B, K, C), L.size("osb_loss_workspace_bytes", B)) + 4096
        wsb = symm_mem.empty(ws_bytes, dtype=torch.uint8, device=dev)
        flat = symm_mem.empty(n_flat, dtype=torch.float32, device=dev)
        flags = symm_mem.empty(L.size("osb_head_dw_allreduce_flag_bytes", ws_) // 4, dtype=torch.int32, device=dev)
        slots = symm_mem.empty(L.size("osb_peer_sum_bytes", ws_) // 4, dtype=torch.int32, device=dev)
        wsb.zero_(); flat.zero_(); flags.zero_(); slots.zero_()
        handles = [symm_mem.rendezvous(t, grp) for t in (wsb, flat, flags, slots)]
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)
        plan = HeadPlan(B, K, C, variant, device=dev, grad_buffer=flat, workspace=wsb, n_tail=n_tail, **plan_kwargs)
        obj = cls(plan, handles[0].buffer_ptrs, handles[1].buffer_ptrs, handles[2].buffer_ptrs, rank, ws_, max_ctas=max_ctas)
        obj.slot_ptrs = (L.C.c_void_p * ws_)(*[int(p) for p in handles[3].buffer_ptrs])
        # NVSwitch multicast mappings of the partials and of the gradient buffer, when the fabric offers them: the
        # all-reduce then sums inside the switch (multimem.ld_reduce) and broadcasts with one store (multimem.st)
        mc_ws = int(getattr(handles[0], "multicast_ptr", 0) or 0)
        mc_flat = int(getattr(handles[1], "multicast_ptr", 0) or 0)
        if multicast and mc_ws and mc_flat:
            obj.mc_partials, obj.mc_dw = mc_ws + L.size("osb_head_dw_partials_offset"), mc_flat
        obj._keep = (wsb, flat, flags, slots, handles)
        return obj

    def launch(self, den, gscale=None, stream=None):
        """One all-reduce of the partials currently in the workspaces, on ``stream`` (default: the current one)."""
        L = self._L
        st = L.stream_ptr(self.plan.dev) if stream is None else stream.cuda_stream
        L.call("osb_head_dw_allreduce", self._partials, self._dw, self._extra if self.n_extra else None, self._flags,
               self.rank, self.world, self.splits, self.n, self.n, L.ptr(self.extra_sum), self.n_extra, self.n_plain,
               L.ptr(gscale), L.ptr(den), self.mc_partials, self.mc_dw, self.max_ctas, st)

    def run(self, den, gscale=None):
        """dW GEMM partials must already be enqueued on the current stream.  Runs the all-reduce on the side stream
        and returns its completion event."""
        cur = torch.cuda.current_stream(self.plan.dev)
        self._ready.record(cur)
        self.stream.wait_event(self._ready)
        self.launch(den, gscale, self.stream)
        self._done.record(self.stream)
        return self._done

    def failed_epoch(self) -> int:
        """0, or the epoch at which this rank gave up waiting for a peer (host sync).  After a failure dW was NOT
        written and the summed loss is NaN."""
        L = self._L
        out = L.C.c_uint32(0)
        L.call("osb_head_dw_allreduce_error", self._flags_local, self.world, L.C.byref(out))
        return int(out.value)

    def check(self) -> None:
        ep = self.failed_epoch()
        if ep:
            raise RuntimeError(f"rank {self.rank}: gradient all-reduce gave up waiting for a peer at epoch {ep}; "
                               "gradients of that step were not written")


class ShardedHeadPlan:
    """Allocation-free batch-sharded head step (``functional.HeadPlan`` per rank + ONE all-reduce), capturable as ONE
    CUDA graph per input buffer (``capture``).

    The kernels write dW partials, db and this shard's loss numerator into peer-mapped buffers; the all-reduce kernel
    sums them over NVLink, divides by the GLOBAL denominator and leaves the full-batch dW / db on every rank, while the dX
    GEMM (local, needs no communication) runs on ``SMs - sm_reserve`` CTAs so that the collective finds free SMs.
    Denominator of the global mean: the entropic loss counts every row, so with equal shards it is ``B_local * world``
    and needs no communication; the softmax / garbage variants sum one double over the ranks through peer memory
    (``osb_peer_sum``) on a branch beside the forward kernel, whose grad-logits do not depend on it.
    Without symmetric memory (``peer_allreduce=False``) the same step runs eagerly with NCCL collectives.
    """

    def __init__(self, B, K, C, variant="entropic", group=None, sm_reserve=None, device=None, peer_allreduce=False,
                 multicast=True, **plan_kwargs):
        from . import _lib as L
        from .functional import HeadPlan
        self.group = group
        self.rank, self.world = world()
        self.variant = variant
        bias = bool(plan_kwargs.get("bias", False))
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        self.peer = None
        if peer_allreduce and self.world > 1:
            # dW all-reduce over NVLink peer memory fused with the split-K reduction (no NCCL on the data path)
            self.peer = PeerGradAllReduce.from_symmetric_memory(B, K, C, variant, group=group, max_ctas=sm_reserve or 16, device=dev,
                                                                multicast=multicast, **plan_kwargs)
            if sm_reserve is None:
                # either way the all-reduce is paced by round trips over NVLink (the switch answers a multimem.ld_reduce
                # in a few microseconds): throughput = bytes in flight, 128 KB per CTA
                sm_reserve = 16
                self.peer.max_ctas = sm_reserve
            self.plan = self.peer.plan
            self.flat = self.plan.flat
        else:
            self.flat = torch.zeros(C * K + (C if bias else 0) + 1, dtype=torch.float32, device=dev)
            self.plan = HeadPlan(B, K, C, variant, device=dev, grad_buffer=self.flat, **plan_kwargs)
        self._L = L
        self._sms = L.size("osb_sm_count")
        self.sm_reserve = int(sm_reserve or 16) if self.world > 1 else 0
        p = self.plan
        self.entropic = variant == "entropic"
        # global denominator: a constant for the entropic loss, the peer / NCCL sum of the shard denominators otherwise
        self.den_global = (torch.full((1,), float(p.B * self.world), dtype=torch.float64, device=dev) if self.entropic
                           else torch.zeros(1, dtype=torch.float64, device=dev))
        self._lbranch = torch.cuda.Stream(device=dev)
        self._abranch = torch.cuda.Stream(device=dev)
        self.fused_step = self.world == 1 and self.plan.fused_dx and self.plan.B > 4096      # as HeadPlan.capture decides
        # our kernels per step: label statistics, forward, dW GEMM, dX, and the split-K reduce (NCCL path) or the peer
        # all-reduce that replaces it (+ the denominator's peer sum for softmax / garbage, + db with a bias)
        self.launches_per_step = ((5 if self.peer is not None else self.plan.launches_per_step - (1 if bias else 0)) + (1 if bias else 0)
                                  + (1 if (self.peer is not None and not self.entropic) else 0)
                                  + (1 if (self.peer is not None and self.peer.local_reduce) else 0))

    # ---- one step; every call below only ENQUEUES, so the same body runs eagerly or under stream capture ----
    def _body(self, x, w, y, bias=None, marks=None):
        p, L = self.plan, self._L
        cur = torch.cuda.current_stream(self.dev)
        mark = (lambda i: marks[i].record(cur)) if marks is not None else (lambda i: None)
        dx_ctas = self._sms - self.sm_reserve if self.world > 1 else 0
        if self.peer is not None:
            # branch: label statistics (+ the denominator's sum over the ranks) beside the forward kernel
            self._lbranch.wait_stream(cur)
            with torch.cuda.stream(self._lbranch):
                p.label_stats(y, max_ctas=8)
                if not self.entropic:
                    L.call("osb_peer_sum", L.ptr(p.den), L.ptr(self.den_global), 1, self.peer.slot_ptrs, self.rank, self.world,
                           self._lbranch.cuda_stream)
            mark(0)
            p.forward(x, w, y, bias, want_pred=False, den=None)        # training step: no per-row predictions
            mark(1)
            p.backward_dw_partials(x, local_reduce=self.peer.local_reduce)
            mark(2)
            self._abranch.wait_stream(cur)
            if not self.entropic:
                self._abranch.wait_stream(self._lbranch)
                cur.wait_stream(self._lbranch)
            if marks is not None and len(marks) > 5:
                marks[4].record(self._abranch)
            self.peer.launch(self.den_global, stream=self._abranch)
            if marks is not None and len(marks) > 5:
                marks[5].record(self._abranch)
            p.backward_dx(x, w, den=self.den_global, max_ctas=dx_ctas)
            mark(3)
            cur.wait_stream(self._abranch)
            cur.wait_stream(self._lbranch)
            return None
        # NCCL path (eager only): the denominator first, then one all-reduce of the packed [dW | db | loss] buffer
        p.label_stats(y)
        if not self.entropic:
            self.den_global.copy_(p.den)
            if self.world > 1:
                dist.all_reduce(self.den_global, op=dist.ReduceOp.SUM, group=self.group)
        mark(0)
        if self.world == 1 and self.fused_step:
            # one GPU, large batch: the kernels a captured step runs -- forward and dX as ONE kernel, then the dW GEMM
            p.forward_dx(x, w, y, bias, den=self.den_global)
            mark(1)
            p.backward_dw(x, w, den=self.den_global)
            mark(2)
            mark(3)
            return None
        p.forward(x, w, y, bias, want_pred=False, den=self.den_global)
        mark(1)
        p.backward_dw(x, w, den=self.den_global)
        mark(2)
        work = None
        if self.world > 1:
            work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        p.backward_dx(x, w, den=self.den_global, max_ctas=dx_ctas)
        mark(3)
        return work

    def step(self, x, w, y, bias=None, marks=None):
        """Enqueue one step.  Returns the async NCCL work handle (NCCL path) or None; ``finish(work)`` makes the
        current stream wait for it.  After that ``plan.dw`` / ``plan.db`` hold the full-batch gradients on every rank
        and ``full_loss()`` the full-batch loss; ``plan.dx`` is this shard's d loss / d features."""
        return self._body(x, w, y, bias, marks)

    def finish(self, work):
        if work is not None:
            work.wait()

    def capture(self, x, w, y, bias=None):
        """The whole sharded step on (x, w, y) as ONE CUDA graph (peer-memory path): every rank must replay the same
        number of times."""
        if self.world > 1 and self.peer is None:
            raise RuntimeError("capture needs the peer-memory all-reduce (NCCL collectives of the fallback path are not captured)")
        p = self.plan
        p._check(x, w, y)
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            self._body(x, w, y, bias)            # warm-up outside capture; also one epoch of the collective on every rank
            side.synchronize()
            self._lbranch.synchronize(); self._abranch.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                self._body(x, w, y, bias)
        torch.cuda.current_stream(self.dev).wait_stream(side)
        return graph

    def full_loss(self) -> torch.Tensor:
        """Full-batch loss (0-dim float64 device tensor): summed numerators over the global denominator."""
        if self.peer is not None:
            return self.peer.extra_sum[self.peer.n_extra - 1].double() / self.den_global[0]
        return self.plan.loss[0].double()

    def db_full(self):
        """Full-batch bias gradient (None without a bias)."""
        if self.plan.db is None:
            return None
        return self.peer.extra_sum[:self.plan.C] if self.peer is not None else self.plan.db

    def check(self) -> None:
        if self.peer is not None:
            self.peer.check()


def confidence_sharded(scores, target_labels, offset=0., unknown_class=-1, last_valid_class=None, group=None):
    """``metrics.confidence`` (metrics.py:8-42) over a row-sharded validation set: every rank passes ITS rows and
    receives the tuple of the whole set.  The two float64 sums and the two counts are additive across shards
    (SURVEY 8e): one all-reduce of four numbers."""
    from .metrics import confidence_sums, finish_confidence
    sums, counts = confidence_sums(scores, target_labels, offset, unknown_class, last_valid_class)
    return finish_confidence(*allreduce_confidence(sums, counts, group), scores.shape[1])


def allreduce_confidence(sums: torch.Tensor, counts: torch.Tensor, group=None):
    """Sum the (float64 sums [2], int64 counts [3]) pair of ``metrics.confidence_sums`` / ``StreamingConfidence`` over
    the ranks.  The counts travel as float64 in the same buffer (exact below 2^53)."""
    _, ws = world()
    if ws == 1:
        return sums, counts
    buf = torch.cat([sums.double(), counts.double()])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf[:2], buf[2:].round().to(torch.int64)


# ------------------------------------------------------------------------------------------------
# OSCR
# ------------------------------------------------------------------------------------------------
_TYPE_MASK = 3
_SKIP_RECORD = -1            # int64 view of key 0xFFFFFFFF | type 3: a padding record the scan ignores


def merge_counts(ccr_cnt: torch.Tensor, fpr_cnt: torch.Tensor, group=None):
    """The "single small all-reduce" of the sharded OSCR: int64 [ccr counts | fpr counts]."""
    n = ccr_cnt.numel()
    buf = torch.cat([ccr_cnt, fpr_cnt])
    _, ws = world()
    if ws > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf[:n], buf[n:]


def gather_known_records(rec_known: torch.Tensor, rank: int, ws: int, group=None) -> torch.Tensor:
    """All-gather the fixed-size (skip-padded) known-record array of every shard and demote the remote
    shards' records to "threshold only" (type 0): they define thresholds everywhere but are counted only by
    the shard that owns them.  Works on any backend (int64 view of the 8-byte records)."""
    if ws == 1:
        return rec_known
    n = rec_known.numel()
    out = torch.empty(ws * n, dtype=rec_known.dtype, device=rec_known.device)
    dist.all_gather_into_tensor(out, rec_known.contiguous(), group=group)
    v = out.view(ws, n)
    remote = torch.ones(ws, 1, dtype=torch.bool, device=out.device)
    remote[rank] = False
    keep = ((v & _TYPE_MASK) == _TYPE_MASK) | ~remote          # padding stays padding; local rows keep their type
    return torch.where(keep, v, v & ~_TYPE_MASK).view(-1)


def calculate_oscr_sharded(gt, scores, unk_label=-1, group=None):
    """OSCR over a row-sharded evaluation set: every rank passes ITS rows (same shard size on every rank,
    pad with label -3 rows if needed) and receives the full (ccr, fpr) float64 device tensors plus meta.
    Bit-identical to ``util.calculate_oscr`` on the concatenated data."""
    from . import _lib as L
    from .util import _LABEL_TYPES

    dev = L.require_cuda(scores, gt)
    rank, ws = world()
    s = scores if scores.dtype == torch.float32 else scores.float()
    if s.stride(1) != 1:
        s = s.contiguous()
    g = gt if gt.dtype in _LABEL_TYPES else gt.long()
    g = g.contiguous()
    n, c = s.shape
    st = L.stream_ptr(dev)
    cap = max(ws * n + n, 1)
    with torch.cuda.device(dev):
        rec_unk = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
        rec_known = torch.full((max(n, 1),), _SKIP_RECORD, dtype=torch.int64, device=dev)
        totals = torch.zeros(4, dtype=torch.int64, device=dev)
        L.call("osb_oscr_records", L.ptr(s), s.stride(0) if n else c, L.ptr(g), _LABEL_TYPES[g.dtype], n, c, int(unk_label),
               L.ptr(rec_unk), L.ptr(rec_known), L.ptr(totals), st)
        rec_b = gather_known_records(rec_known, rank, ws, group)
        n_b = torch.tensor([rec_b.numel()], dtype=torch.int64, device=dev)
        taus = torch.empty(cap, dtype=torch.float32, device=dev)
        ccr_cnt = torch.zeros(cap, dtype=torch.int64, device=dev)
        fpr_cnt = torch.zeros(cap, dtype=torch.int64, device=dev)
        meta = torch.zeros(5, dtype=torch.int64, device=dev)
        cws = L.workspace("oscr_curve", L.size("osb_oscr_curve_workspace_bytes", cap), dev)
        L.call("osb_oscr_curve", L.ptr(rec_unk), L.ptr(totals[2:3]), L.ptr(rec_b), L.ptr(n_b), cap,
               L.ptr(taus), L.ptr(ccr_cnt), L.ptr(fpr_cnt), None, None, L.ptr(meta), L.ptr(cws), cws.numel(), st)
        # every rank derived the same thresholds; the counts are per shard: ONE all-reduce merges them
        tot_g = totals.clone()
        if ws > 1:
            dist.all_reduce(tot_g, op=dist.ReduceOp.SUM, group=group)
        ccr_cnt, fpr_cnt = merge_counts(ccr_cnt, fpr_cnt, group)
        ccr = torch.empty(cap, dtype=torch.float64, device=dev)
        fpr = torch.empty(cap, dtype=torch.float64, device=dev)
        L.call("osb_oscr_ratios", L.ptr(ccr_cnt), L.ptr(fpr_cnt), L.ptr(meta[0:1]), L.ptr(tot_g[0:1]), L.ptr(tot_g[2:3]),
               cap, L.ptr(ccr), L.ptr(fpr), st)
    meta_g = torch.cat([meta[0:1], tot_g])
    return ccr, fpr, ccr_cnt, fpr_cnt, taus, meta_g


class OscrShardedPlan:
    """OSCR over a row-sharded evaluation set with persistent buffers and no NCCL on the data path: stage A on this
    rank's rows, ONE peer-memory all-gather of the compacted records (``osb_oscr_exchange``), then the sort + scan on
    every rank (identical inputs -> the curve is replicated and bit-identical to ``util.calculate_oscr`` on the
    concatenated data).  Three launches, no allocation, no host synchronisation: capturable into a CUDA graph
    (``capture``); every rank must run / replay the same number of times.

    ``n_local``: rows of this rank's shard (the same on every rank; pad a ragged shard with rows labelled -3, which
    belong to neither group).  Outputs (device tensors, overwritten by every call): ``ccr`` / ``fpr`` float64 and, with
    ``counts=True``, ``taus`` float32, ``ccr_cnt`` / ``fpr_cnt`` int64, all of capacity ``n_local * world``;
    ``meta`` int64 [5] = (n_unique, n_known, n_correct, n_unk, 0) and ``totals`` int64 [4] = global stage-A totals
    (n_known, n_correct, n_unk, n_bad).  The reference's arrays are the first ``meta[0] - 1`` entries.
    """

    def __init__(self, n_local: int, C: int, unk_label: int = -1, group=None, device=None, counts: bool = True, max_ctas: int = 32):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib as L
        self._L = L
        self.rank, self.world = world()
        self.dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.n_local, self.C, self.unk_label, self.max_ctas = int(n_local), int(C), int(unk_label), int(max_ctas)
        d, W = self.dev, self.world
        self.cap = max(self.n_local * W, 1)
        grp = group if group is not None else dist.group.WORLD
        # gathered record arrays and the slot area live in symmetric (peer-mapped) memory
        self.all_unk = symm_mem.empty(self.cap, dtype=torch.int64, device=d)
        self.all_known = symm_mem.empty(self.cap, dtype=torch.int64, device=d)
        self.slots = symm_mem.empty(L.size("osb_oscr_exchange_slot_bytes", W) // 4 + 4, dtype=torch.int32, device=d)
        self.slots.zero_()
        self._handles = [symm_mem.rendezvous(t, grp) for t in (self.all_unk, self.all_known, self.slots)]
        torch.cuda.synchronize(d)
        dist.barrier(group=group)
        P = L.C.c_void_p * W
        self._p_unk = P(*[int(p) for p in self._handles[0].buffer_ptrs])
        self._p_known = P(*[int(p) for p in self._handles[1].buffer_ptrs])
        self._p_slots = P(*[int(p) for p in self._handles[2].buffer_ptrs])
        n1 = max(self.n_local, 1)
        self.rec_unk = torch.empty(n1, dtype=torch.int64, device=d)
        self.rec_known = torch.empty(n1, dtype=torch.int64, device=d)
        self.totals_local = torch.zeros(4, dtype=torch.int64, device=d)
        self.totals = torch.zeros(4, dtype=torch.int64, device=d)
        self.taus = torch.empty(self.cap, dtype=torch.float32, device=d) if counts else None
        self.ccr_cnt = torch.empty(self.cap, dtype=torch.int64, device=d) if counts else None
        self.fpr_cnt = torch.empty(self.cap, dtype=torch.int64, device=d) if counts else None
        self.ccr = torch.empty(self.cap, dtype=torch.float64, device=d)
        self.fpr = torch.empty(self.cap, dtype=torch.float64, device=d)
        self.meta = torch.zeros(5, dtype=torch.int64, device=d)
        self.ws = torch.zeros(L.size("osb_oscr_curve_workspace_bytes", self.cap) + 256, dtype=torch.uint8, device=d)
        self.launches_per_call = 3

    def run(self, scores: torch.Tensor, gt: torch.Tensor):
        from .util import _LABEL_TYPES
        L = self._L
        if scores.shape != (self.n_local, self.C) or scores.dtype != torch.float32 or scores.stride(1) != 1:
            raise ValueError("scores must be this rank's float32 [n_local, C] shard with unit column stride")
        if gt.shape != (self.n_local,) or gt.dtype not in _LABEL_TYPES or not gt.is_contiguous():
            raise ValueError("gt must be a contiguous int64 / int32 / float32 [n_local] tensor")
        st = L.stream_ptr(self.dev)
        L.call("osb_oscr_records", L.ptr(scores), scores.stride(0) if self.n_local else self.C, L.ptr(gt), _LABEL_TYPES[gt.dtype],
               self.n_local, self.C, self.unk_label, L.ptr(self.rec_unk), L.ptr(self.rec_known), L.ptr(self.totals_local), st)
        L.call("osb_oscr_exchange", L.ptr(self.rec_unk), L.ptr(self.rec_known), L.ptr(self.totals_local), self._p_unk, self._p_known,
               self._p_slots, self.rank, self.world, self.cap, L.ptr(self.totals), self.max_ctas, st)
        L.call("osb_oscr_curve", L.ptr(self.all_unk), L.ptr(self.totals[2:3]), L.ptr(self.all_known), L.ptr(self.totals[0:1]), self.cap,
               L.ptr(self.taus), L.ptr(self.ccr_cnt), L.ptr(self.fpr_cnt), L.ptr(self.ccr), L.ptr(self.fpr), L.ptr(self.meta),
               L.ptr(self.ws), self.ws.numel(), st)
        return self

    def capture(self, scores: torch.Tensor, gt: torch.Tensor):
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            self.run(scores, gt)                     # warm-up outside capture: one epoch of the exchange on every rank
            side.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                self.run(scores, gt)
        torch.cuda.current_stream(self.dev).wait_stream(side)
        return graph

    def result(self):
        """(ccr, fpr) float64 numpy arrays like ``util.calculate_oscr`` (one host sync); raises like it on bad labels."""
        n_unique = int(self.meta[0].item())
        n_bad = int(self.totals[3].item())
        if n_bad >= (1 << 40):
            raise RuntimeError("sharded OSCR: a rank gave up waiting for a peer during the record exchange")
        if n_bad:
            raise IndexError(f"{n_bad} label(s) out of bounds for axis 1 with size {self.C}")
        n = max(n_unique - 1, 0)
        return self.ccr[:n].cpu().numpy(), self.fpr[:n].cpu().numpy()
