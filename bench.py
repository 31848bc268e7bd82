#!/usr/bin/env python
"""Benchmark of the B200-native open-set head + OSCR path (contract: see the task statement / DESIGN.md §6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload head|bg|oscr|resnet]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Default workload ("head") = BASELINE.json's large-head configuration, the largest one that fits one GPU and
the one its tensor-core target is quoted on: entropic open-set loss over a 2048 -> 1000 head on 16384
bf16 feature rows PER GPU (1/3 negatives), forward + backward (loss, confidence statistics, dX, dW).
A "step" is one such pass; with N GPUs the batch is sharded (weak scaling: 16384 rows per GPU) and the fp32
dW (+loss) is all-reduced over NVLink peer memory every step, the whole sharded step replayed from ONE CUDA graph.
The same JSON line carries sub-objects for the other BASELINE configurations: "oscr" (OSCR over 1M x 116 synthetic
scores, the second headline metric, sharded by rows for N > 1), "bg" (garbage-class head, Protocol 3 shape) and, at
N = 1, "resnet" (ResNet-50 + fused head training step).

One JSON line is printed by rank 0.  `value` is device throughput with inputs resident in HBM; `e2e` is
the same step driven through the C ABI from pinned HOST buffers (H2D of features + labels and D2H of the loss inside
the timed region; dX and dW stay on the device for the backbone / optimiser).  `parity.head` compares the benchmarked
kernels with the float64 oracle on a sub-batch (and, for N > 1, the all-reduced dW across ranks and against one GPU).
`--impl reference` times the reference's OWN files (losses.py, metrics.py: vendored verbatim into oracle/_ref by
oracle/vendor_ref.py) on the host cores, on the same configuration.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402
import torch         # noqa: E402

HEAD = dict(B=16384, K=2048, C=1000, variant="entropic", neg_frac=1.0 / 3.0, precision="fast")
BG = dict(B=1024, K=2048, C=152, variant="garbage", neg_frac=1.0 / 3.0, precision="exact")
OSCR = dict(N=1_000_000, C=116)
METRIC = "EOS head fwd+bwd samples/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tc_burst=d["bf16_tflops"], tc_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, src="fallback")


def load_traffic():
    """DRAM bytes per launch from the committed ncu --set full captures (profiles/traffic.json), or {}."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return {k: v for k, v in json.load(open(p)).items() if not k.startswith("_")}
    except Exception:
        return {}


class ClockSampler:
    """Polls SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                 "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4, "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.0005)               # NVML answers in ~0.1 ms: >= 1 kHz sampling

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        return {"sm_mhz": (statistics.median(self.samples) if self.samples else None), "sm_max_mhz": self.max_mhz,
                "samples": len(self.samples), "reasons": sorted(self.reasons)}


def pick_tensor_peak(peaks, clocks):
    """The cuBLAS bf16 peak a tensor-bound measurement is divided by, chosen from the clock record of the SAME region:
    a power-capped region (sw_power_cap seen) is compared with the sustained figure, anything else with the burst one."""
    capped = "sw_power_cap" in (clocks or {}).get("reasons", [])
    kind = "sustained" if capped else "burst"
    return (peaks["tc_sustained"] if capped else peaks["tc_burst"]), \
        f"{peaks['src']} cuBLAS bf16 {kind} ({'sw_power_cap seen' if capped else 'no power cap'} in this region's clock record)"


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def make_head_inputs(cfg, dev, nbuf, seed):
    from openset_imagenet_b200 import synth as O
    g = torch.Generator(device=dev).manual_seed(seed)
    xs = [torch.randn(cfg["B"], cfg["K"], device=dev, generator=g).bfloat16() for _ in range(nbuf)]
    if cfg["variant"] == "garbage":
        ys_np = [O.make_labels(cfg["B"], cfg["C"] - 1, cfg["neg_frac"], seed=seed + i, neg_label=cfg["C"] - 1) for i in range(nbuf)]
        # class weights are a property of the training set (dataset.py:77-86), the same on every rank
        y_set = O.make_labels(50 * cfg["B"], cfg["C"] - 1, cfg["neg_frac"], seed=99, neg_label=cfg["C"] - 1)
        cw = torch.tensor(O.class_weights_from_labels(np.concatenate([y_set, np.arange(cfg["C"])]), cfg["C"]), device=dev)
    else:
        ys_np = [O.make_labels(cfg["B"], cfg["C"], cfg["neg_frac"], seed=seed + i) for i in range(nbuf)]
        cw = None
    ys = [torch.tensor(y, device=dev) for y in ys_np]
    # the same weights on every rank (data parallel): their own generator
    gw = torch.Generator(device=dev).manual_seed(4321)
    w = (torch.randn(cfg["C"], cfg["K"], device=dev, generator=gw) * (1.0 / cfg["K"] ** 0.5)).bfloat16()
    return xs, ys, ys_np, w, cw


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def head_parity(cfg, x, w, y_np, cw, plan_kw, rows):
    """The benchmarked kernels (same precision mode, same dX dtype) on the first `rows` rows of a bench buffer against the
    float64 oracle: relative error of the loss, dW (norm-wise), dX.  A bf16 dX is compared with the oracle ROUNDED to its
    storage format, and once more as the fp32 accumulator (dx_dtype float32).  The one place the oracle is used here."""
    from openset_imagenet_b200.functional import HeadPlan
    from oracle import oracle as O
    K, C = cfg["K"], cfg["C"]
    dev = x.device
    rows = min(rows, x.shape[0])
    xs, ys = x[:rows].contiguous(), torch.tensor(y_np[:rows], device=dev)
    ov = {"entropic": O.ENTROPIC, "softmax": O.SOFTMAX, "garbage": O.GARBAGE}[cfg["variant"]]
    ref = O.head_fwd_bwd(xs.float().cpu().numpy(), w.float().cpu().numpy(), None, y_np[:rows], ov,
                         class_weights=None if cw is None else cw.cpu().numpy())
    out = {"rows": rows, "tolerance": 1e-3, "oracle": "oracle/oracle.py head_fwd_bwd (float64) on the bf16 operands the kernels read"}
    for tag, dt in (("", torch.bfloat16), ("_f32", torch.float32)):
        plan = HeadPlan(rows, K, C, cfg["variant"], dx_dtype=dt, **plan_kw)
        plan.step(xs, w, ys)
        torch.cuda.synchronize()
        if not tag:
            out["precision"] = "exact (grad-logits as a bf16 pair)" if plan.planes == 2 else "fast (one bf16 plane of grad-logits)"
            out["loss_rel"] = abs(float(plan.loss) - ref["loss"]) / abs(ref["loss"])
            out["dw_rel"] = _rel(plan.dw.cpu().numpy(), ref["dw"])
            out["dx_bf16_rel_vs_bf16_oracle"] = _rel(plan.dx.float().cpu().numpy(), O.bf16_round(ref["dx"].astype(np.float32)))
        else:
            out["dx_f32_rel"] = _rel(plan.dx.cpu().numpy(), ref["dx"])
    out["ok"] = bool(out["loss_rel"] < 1e-3 and out["dw_rel"] < 1e-3 and out["dx_f32_rel"] < 1e-3)
    return out


def sharded_parity(cfg, sp, x, w, y, plan_kw, rank, world, dev):
    """N > 1: one more step on buffer 0, then (i) the all-reduced dW / loss must be bit-identical on every rank and (ii)
    equal, within fp32 summation-order noise, to ONE GPU's dW / loss on the concatenated batch."""
    import torch.distributed as dist
    from openset_imagenet_b200.functional import HeadPlan
    sp.finish(sp.step(x, w, y))
    torch.cuda.synchronize()
    sp.check()
    dw = sp.plan.dw.detach().clone()
    loss = sp.full_loss().reshape(1).clone()
    sig = torch.cat([dw.view(torch.int32).sum(dtype=torch.int64).reshape(1), dw.double().abs().sum().view(torch.int64).reshape(1),
                     loss.view(torch.int64)])
    sigs = [torch.empty_like(sig) for _ in range(world)]
    dist.all_gather(sigs, sig)
    identical = all(torch.equal(sigs[0], s_) for s_ in sigs)
    xs = [torch.empty_like(x) for _ in range(world)] if rank == 0 else None
    ys = [torch.empty_like(y) for _ in range(world)] if rank == 0 else None
    dist.gather(x, xs, dst=0)
    dist.gather(y, ys, dst=0)
    out = {"dw_and_loss_bit_identical_across_ranks": bool(identical)}
    if rank == 0:
        big = HeadPlan(cfg["B"] * world, cfg["K"], cfg["C"], cfg["variant"], **plan_kw)
        big.step(torch.cat(xs), w, torch.cat(ys))
        torch.cuda.synchronize()
        out["dw_rel_vs_single_gpu_concatenated_batch"] = _rel(dw.cpu().numpy(), big.dw.cpu().numpy())
        out["loss_rel_vs_single_gpu_concatenated_batch"] = abs(float(loss) - float(big.loss)) / abs(float(big.loss))
        out["tolerance"] = 1e-4       # fp32 sums in another order (different split-K / tile partition of the longer batch)
        out["ok"] = bool(identical and out["dw_rel_vs_single_gpu_concatenated_batch"] < 1e-4
                         and out["loss_rel_vs_single_gpu_concatenated_batch"] < 1e-4)
        del big
    dist.barrier()
    return out


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_head(args, cfg, rank, world, dev, peaks, light=False):
    """light: sub-object mode (no e2e, no per-kernel loop): throughput + parity only."""
    import torch.distributed as dist
    from openset_imagenet_b200 import _lib as L
    from openset_imagenet_b200.dist import ShardedHeadPlan

    B, K, C = cfg["B"], cfg["K"], cfg["C"]
    nbuf = 3                                        # 3 x 64 MB of features + 64 MB dX + 32 MB dZ: far larger than the 126 MB L2
    xs, ys, ys_np, w, cw = make_head_inputs(cfg, dev, nbuf, seed=1234 + rank)
    plan_kw = dict(class_weights=cw, last_valid_class=-1 if cfg["variant"] == "garbage" else None, device=dev,
                   precision=cfg.get("precision", "exact"))
    allreduce = "none"
    if world > 1:
        try:            # dW all-reduce over NVLink peer memory, fused with the split-K reduction (no NCCL on the data path)
            sp = ShardedHeadPlan(B, K, C, cfg["variant"], peer_allreduce=True, sm_reserve=args.sm_reserve or None,
                                 multicast=not args.no_multicast, **plan_kw)
            allreduce = ("peer-memory kernel fused with the split-K reduction (osb_head_dw_allreduce), " +
                         ("summed inside the NVSwitch (multimem.ld_reduce / multimem.st), " if sp.peer.mc_partials is not None else "peer loads, ") +
                         "overlapped with dX")
        except Exception as e:                                  # no symmetric memory on this box: NCCL all-reduce
            if rank == 0:
                print(f"peer-memory all-reduce unavailable ({type(e).__name__}: {e}); using NCCL", file=sys.stderr)
            sp = ShardedHeadPlan(B, K, C, cfg["variant"], sm_reserve=args.sm_reserve or 32, **plan_kw)
            allreduce = "NCCL all-reduce of the packed fp32 [dW | loss] buffer, overlapped with dX"
    else:
        sp = ShardedHeadPlan(B, K, C, cfg["variant"], **plan_kw)
    plan = sp.plan

    def eager_step(i, marks=None):
        sp.finish(sp.step(xs[i % nbuf], w, ys[i % nbuf], marks=marks))

    # The launches of a step are replayed from ONE CUDA graph per rotating input buffer, so the timed loop has no
    # per-launch host work; with N > 1 the graph holds the whole sharded step including the peer-memory all-reduce (its
    # epoch lives in device memory).  Only the NCCL fallback stays eager.
    launch_mode = "eager"
    run_step = eager_step
    n_launch = sp.launches_per_step
    if not args.eager and (world == 1 or sp.peer is not None):
        graphs = [(plan.capture(xs[i], w, ys[i]) if world == 1 else sp.capture(xs[i], w, ys[i])) for i in range(nbuf)]
        n_launch = plan.launches_per_captured_step if world == 1 else sp.launches_per_step
        launch_mode = "CUDA-graph replay of the whole step" + (" incl. the peer-memory all-reduce" if world > 1 else
                                                                f" ({n_launch} kernels" + ("; forward and dX are one kernel)" if plan.fused_dx and n_launch < plan.launches_per_step else ")"))

        def run_step(i):
            graphs[i % nbuf].replay()

    # spin-up: ~30 ms of untimed replays first, so that a short run (the driver's 20 steps are 1-4 ms) is not measured on a
    # GPU that is still leaving its idle clocks / first-touching peer mappings; then the W warm-up steps of the contract
    for i in range(max(20, int(30e3 / (40.0 if B <= 2048 else 200.0)))):
        run_step(i)
    for i in range(args.warmup):
        run_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev.index) as clk:
        if world > 1:
            # the ranks leave the host barrier tens of microseconds apart: one more untimed step, whose all-reduce makes the
            # GPUs meet, so that the K timed steps start aligned on the device
            run_step(args.warmup)
        t0.record()
        for i in range(args.steps):
            run_step(i)
        t1.record()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sp.check()
    ms = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    loss_val = float(sp.full_loss().item()) if world > 1 else float(plan.loss.item())
    ms_per_step = ms / args.steps
    value = world * B / (ms_per_step * 1e-3)
    clocks = clk.summary()
    splitk = bool(L.size("osb_head_forward_is_splitk", B, K, C))
    fwd_name = "head_fwd_splitk_kernel" if splitk else "head_fwd_kernel"
    gemm_flops = 2.0 * B * K * C
    step_us = ms_per_step * 1e3
    peak, peak_kind = pick_tensor_peak(peaks, clocks)
    roof = {"bound": "tensor", "unit": "TFLOP/s", "peak": peak, "peak_kind": peak_kind,
            "step": {"flops": 3 * gemm_flops, "us": round(step_us, 2), "achieved": 3 * gemm_flops / step_us / 1e6,
                     "frac": 3 * gemm_flops / step_us / 1e6 / peak},
            "algorithmic_flops_per_sample": 6.0 * K * C}
    parity = {}
    if rank == 0:
        parity["head"] = head_parity(cfg, xs[0], w, ys_np[0], cw, plan_kw, rows=B if B <= 2048 else 2048)
    if world > 1:
        parity["sharded"] = sharded_parity(cfg, sp, xs[0], w, ys[0], plan_kw, rank, world, dev)
    if light:
        return dict(value=value, ms_per_step=ms_per_step, roofline=roof, e2e=None, clocks=clocks, loss=loss_val,
                    launches=n_launch * args.steps, allreduce=allreduce, launch_mode=launch_mode, parity=parity,
                    precision=parity.get("head", {}).get("precision"))

    # ---- per-kernel times: a second, shorter loop with CUDA events between the launches of an EAGER step, queued behind
    # ~40 ms of replays of the same step, so that the host is ahead of the GPU (the events bracket kernels that run back
    # to back) and the GPU is in the power state of the headline loop.  Medians over the steps.  The dominant kernel's
    # RAW bracket is what `roofline.achieved` is computed from; no rescaling.
    n_ev = max(10, min(args.steps, 50))
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(n_ev)]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    with ClockSampler(dev.index) as clk_k:
        if run_step is not eager_step:
            for i in range(max(50, int(40e3 / max(step_us, 1.0)))):
                run_step(i)
        for i in range(n_ev):
            eager_step(i, evs[i])
        torch.cuda.synchronize()
    sp.check()
    k_fwd = statistics.median(e[0].elapsed_time(e[1]) for e in evs) * 1e3
    k_dw = statistics.median(e[1].elapsed_time(e[2]) for e in evs) * 1e3
    k_dx = statistics.median(e[2].elapsed_time(e[3]) for e in evs) * 1e3
    dw_name = "head_dw2_kernel" + ((" (direct: split 1 adds at the destination)" if plan.dw_direct else "+dw_reduce_kernel")
                                   if (world == 1 or sp.peer is None) else " (partials; the reduce is the all-reduce)")
    fused_name = "head_fwd_kernel<FUSE> (forward and dX GEMMs in one kernel)"
    if sp.fused_step:
        kernels = {fused_name: k_fwd, dw_name: k_dw}
    else:
        kernels = {fwd_name: k_fwd, dw_name: k_dw, "head_dx2_kernel": k_dx}
    k_flops = {k: (2.0 if k == fused_name else 1.0) * gemm_flops for k in kernels}
    dom = max(kernels, key=kernels.get)
    clocks_k = clk_k.summary()
    peak_k, peak_kind_k = pick_tensor_peak(peaks, clocks_k)
    bracket_sum = sum(kernels.values())
    roof.update({
        "kernel": dom, "achieved": k_flops[dom] / kernels[dom] / 1e6, "peak": peak_k, "peak_kind": peak_kind_k,
        "kernel_flops": k_flops[dom],
        "traffic": ((load_traffic().get(dom) or load_traffic().get(dom.split(" ")[0], {})).get("bytes")
                    if (B, K, C) == (16384, 2048, 1000) else None),
        "traffic_note": "dram read+write bytes of one launch from the committed ncu --set full capture (profiles/traffic.json)",
        "kernels_us": {k: round(v, 2) for k, v in kernels.items()},
        "kernels_tflops": {k: round(k_flops[k] / v / 1e6, 1) for k, v in kernels.items()},
        "duration_note": "kernels_us = medians of raw CUDA-event brackets around the launches of an eager step (the same kernels as the "
                         "captured step) queued behind graph replays of that step; achieved = the dominant kernel's algorithmic flops "
                         "(2*B*K*C per GEMM; the fused kernel holds two) / its bracket",
        "kernels_us_share_of_step": {k: round(v * step_us / bracket_sum, 2) for k, v in kernels.items()},
        "clocks_kernel_loop": clocks_k,
    })
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["step"]["peak"], roof["step"]["peak_kind"] = peak, peak_kind
    if world > 1 and sp.peer is not None:
        roof["allreduce_us"] = round(statistics.median(e[4].elapsed_time(e[5]) for e in evs) * 1e3, 2)
        roof["allreduce_note"] = "dw_allreduce_kernel on its own stream, beside dX: launch to completion incl. both cross-GPU barriers"

    # ---- end to end from pinned host buffers through the C ABI: every step copies ITS features and labels host ->
    # device and reads its loss back.  Two device input slots: the copy of step i+1 (copy stream) overlaps the kernels
    # of step i, as a training loop with a prefetching loader does.
    e2e = None
    if not args.no_e2e:
        hx = [x.cpu().pin_memory() for x in xs[:2]]
        hy = [y.cpu().pin_memory() for y in ys[:2]]
        lossh = torch.empty(1, dtype=torch.float32).pin_memory()
        xd = [torch.empty_like(xs[0]) for _ in range(2)]
        yd = [torch.empty_like(ys[0]) for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        ready = [torch.cuda.Event() for _ in range(2)]       # slot filled
        freed = [torch.cuda.Event() for _ in range(2)]       # slot consumed
        n_e2e = max(3, min(args.steps, 40))
        loss_src = sp.peer.extra_sum[sp.peer.n_extra - 1:] if (world > 1 and sp.peer is not None) else plan.loss

        def prefetch(i):
            sl = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[sl])
                xd[sl].copy_(hx[i % 2], non_blocking=True)
                yd[sl].copy_(hy[i % 2], non_blocking=True)
                ready[sl].record(copy_stream)

        def e2e_loop(n):
            for sl in range(2):
                freed[sl].record(main)
            prefetch(0)
            for i in range(n):
                sl = i % 2
                if i + 1 < n:
                    prefetch(i + 1)
                main.wait_event(ready[sl])
                sp.finish(sp.step(xd[sl], w, yd[sl]))
                freed[sl].record(main)
                lossh.copy_(loss_src, non_blocking=True)

        e2e_loop(3)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        e2e_loop(n_e2e)
        b.record()
        torch.cuda.synchronize()
        ems = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ems], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        e2e = {"value": world * B * n_e2e / (ems * 1e-3), "unit": "samples/s", "steps": n_e2e,
               "h2d_bytes_per_step": B * K * 2 + B * 8, "d2h_bytes_per_step": 4,
               "ms_per_step": ems / n_e2e,
               "note": "features+labels copied from pinned host memory every step (prefetched one step ahead on a copy stream), "
                       "loss read back every step; dX and dW stay on the device for the backbone / optimiser",
               "host_thread_cpus": len(os.sched_getaffinity(0))}
    return dict(value=value, ms_per_step=ms_per_step, roofline=roof, e2e=e2e, clocks=clocks, loss=loss_val,
                launches=n_launch * args.steps, allreduce=allreduce, launch_mode=launch_mode, parity=parity,
                precision=parity.get("head", {}).get("precision"))


def run_oscr(args, rank, world, dev, peaks):
    """OSCR ms @ 1M scores.  N > 1: rows sharded, known records all-gathered, counts merged by one all-reduce.
    Two score buffers alternate, each (464 MB at N = 1) larger than the 126 MB L2, so no iteration finds its input
    in L2; nothing is written to flush, so no dirty lines are written back inside the timed region."""
    import torch.distributed as dist
    from openset_imagenet_b200 import synth
    from openset_imagenet_b200.functional import OscrPlan

    N, C = OSCR["N"], OSCR["C"]
    gt_np, s_np = synth.make_oscr_case(N, C, seed=0)
    n_loc = N // world
    sl = slice(rank * n_loc, (rank + 1) * n_loc)
    bufs = [torch.tensor(s_np[sl], device=dev) for _ in range(2)]
    gt = torch.tensor(gt_np[sl], device=dev)
    sharded_mode = None
    if world > 1:
        from openset_imagenet_b200.dist import OscrShardedPlan, calculate_oscr_sharded
        if n_loc * C * 4 < (192 << 20):                  # shards smaller than L2: add a flush buffer to the rotation
            bufs += [torch.tensor(s_np[sl], device=dev) for _ in range(2)]
        try:
            # stage A on the shard, ONE peer-memory all-gather of the records, sort + scan on every rank: 3 launches from a graph
            splan = OscrShardedPlan(n_loc, C, -1, device=dev)
            sgraphs = [splan.capture(b_, gt) for b_ in bufs]
            sharded_mode = "peer-memory all-gather of the stage-A records (osb_oscr_exchange), curve replicated; CUDA-graph replay of 3 kernels"

            def run(i):
                sgraphs[i % len(sgraphs)].replay()
                return (splan.ccr, splan.fpr, splan.ccr_cnt, splan.fpr_cnt, splan.taus, torch.cat([splan.meta[0:1], splan.totals]))
        except Exception as e:
            if rank == 0:
                print(f"peer-memory OSCR exchange unavailable ({type(e).__name__}: {e}); using NCCL collectives", file=sys.stderr)
            sharded_mode = "NCCL all-gather of the known records + all-reduce of the counts (eager)"

            def run(i):
                return calculate_oscr_sharded(gt, bufs[i % len(bufs)], -1)
    else:
        plan = OscrPlan(N, C, -1, device=dev, counts=False)      # the reference's outputs: ccr, fpr (float64)
        # the two launches of one evaluation replayed from a CUDA graph (one graph per input buffer): no per-call
        # host work between the row reduce and the cooperative sort + scan kernel
        graphs = []
        side = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(side):
            for b_ in bufs:
                plan.run(b_, gt)
            side.synchronize()
            for b_ in bufs:
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_, stream=side):
                    plan.run(b_, gt)
                graphs.append(g_)
        torch.cuda.current_stream(dev).wait_stream(side)

        def run(i):
            graphs[i % 2].replay()
            return plan
    for i in range(3):
        r = run(i)
    torch.cuda.synchronize()
    times = []
    iters = 20
    for i in range(iters):
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = run(i)
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    ms = statistics.median(times)
    if world == 1:
        meta = r.meta.tolist()
        n_thr = max(meta[0] - 1, 0)
    else:
        n_thr = max(int(r[5][0].item()) - 1, 0)
    alg_bytes = 4.0 * N * C + 8.0 * N + 16.0 * n_thr
    out = {"metric": "OSCR ms @1M scores", "value": ms, "unit": "ms", "higher_is_better": False, "iters": iters,
           "config": {"workload": f"OSCR CCR/FPR over N={N} x C={C} float32 scores, 50% known / 25% -1 / 25% -2, unk_label=-1",
                      "rows_per_gpu": n_loc, "n_thresholds": n_thr,
                      "outputs": "ccr, fpr float64 [n_thresholds] as util.calculate_oscr returns them" if world == 1 else
                                 "ccr, fpr float64 plus thresholds and int64 counts",
                      "l2": "two alternating score buffers, each larger than L2; no flush writes",
                      "launch": "CUDA-graph replay of the two kernels" if world == 1 else sharded_mode},
           "roofline": {"bound": "hbm", "achieved": alg_bytes / (ms * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                        "peak_kind": f"{peaks['src']} copy bandwidth", "algorithmic_bytes": alg_bytes,
                        "traffic": (sum(load_traffic().get(k, {}).get("bytes", 0) for k in ("oscr_records_kernel", "oscr_curve_kernel")) or None)
                        if (world == 1 and N == 1_000_000 and C == 116) else None,
                        "traffic_note": "dram read+write bytes of both launches from the committed ncu --set full capture (profiles/traffic.json)",
                        "kernels": "oscr_records_kernel (one pass over the scores) + oscr_curve_kernel (sort + scan, L2 resident)"}}
    out["roofline"]["frac"] = out["roofline"]["achieved"] / out["roofline"]["peak"]
    if rank == 0 and world > 1:
        # the sharded result must be bit-identical to the single-GPU one over the whole evaluation set
        ref = OscrPlan(N, C, -1, device=dev).run(torch.tensor(s_np, device=dev), torch.tensor(gt_np, device=dev))
        nu = int(ref.meta[0].item())
        same = (int(r[5][0].item()) == nu and torch.equal(r[2][:nu], ref.ccr_cnt[:nu]) and torch.equal(r[3][:nu], ref.fpr_cnt[:nu])
                and torch.equal(r[4][:nu], ref.taus[:nu]) and torch.equal(r[0][:nu], ref.ccr[:nu])
                and torch.equal(r[1][:nu].nan_to_num(-1.0), ref.fpr[:nu].nan_to_num(-1.0)))
        out["parity"] = {"sharded_bit_identical_to_single_gpu": bool(same)}
    if rank == 0 and world == 1 and not args.no_e2e:
        import openset_imagenet_b200 as osb
        from oracle import oracle as O
        hs, hg = torch.from_numpy(s_np).pin_memory(), torch.from_numpy(gt_np).pin_memory()
        osb.calculate_oscr(hg, hs, -1)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            ccr, fpr = osb.calculate_oscr(hg, hs, -1)
        e_ms = (time.perf_counter() - t0) / reps * 1e3
        out["e2e"] = {"value": e_ms, "unit": "ms", "h2d_bytes_per_step": N * C * 4 + N * 8, "d2h_bytes_per_step": 16 * n_thr + 40}
        # ---- CPU baseline leg (the only place the oracle is used here) ----
        t0 = time.perf_counter()
        o_ccr, o_fpr = O.calculate_oscr(gt_np, s_np, -1)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        n_q = 8000
        from oracle import ref_loader, ref_port
        quad = ref_loader.load("util").calculate_oscr if ref_loader.available() else ref_port.oscr_port_quadratic
        t0 = time.perf_counter()
        with np.errstate(all="ignore"):
            quad(gt_np[:n_q], s_np[:n_q], -1)
        q_ms = (time.perf_counter() - t0) * 1e3
        out["cpu_baseline"] = {"value": cpu_ms, "unit": "ms", "cores": 1, "kind": "port",
                               "sample": f"O(N log N) numpy restatement (oracle.calculate_oscr, bit-equal to the reference) on the full N={N}; "
                                         f"numpy sort is single-threaded",
                               "reference_quadratic": {"n": n_q, "ms": q_ms, "extrapolated_ms_at_1M": q_ms * (N / n_q) ** 2,
                                                       "kind": "reference" if ref_loader.available() else "port",
                                                       "note": "the reference's own util.calculate_oscr (O(U*N) loop, util.py:113-118) timed at n rows; 1M figure is an N^2 extrapolation"}}
        out["parity"] = {"bit_exact_vs_oracle": bool(np.array_equal(ccr, o_ccr) and np.array_equal(fpr, o_fpr, equal_nan=True))}
    return out


def run_resnet(args, rank, world, dev):
    """BASELINE.json cfg 2: ResNet-50 (torchvision / cuDNN, bf16 autocast, channels_last) + fused entropic head,
    Protocol-1 shape (116 classes), B = 256 synthetic 224x224 images per GPU; whole training step and the head's share.
    The model is the reference's (fc 2048 -> 116 with bias, logits 116 -> 116 without, train.py:350-352); the two
    stacked linears run COLLAPSED as one 2048 -> 116 fused tcgen05 head (model.ResNet50, fused='collapsed')."""
    from openset_imagenet_b200 import synth
    from openset_imagenet_b200.model import ResNet50
    B, C = 256, 116
    torch.manual_seed(0)
    model = ResNet50(fc_layer_dim=C, out_features=C, logit_bias=False, variant="entropic").to(dev).to(memory_format=torch.channels_last)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9)
    imgs = [torch.randn(B, 3, 224, 224, device=dev).contiguous(memory_format=torch.channels_last) for _ in range(2)]
    ys = [torch.tensor(synth.make_labels(B, C, 1 / 3, seed=i), device=dev) for i in range(2)]

    def step(i):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            j, _ = model.forward_loss(imgs[i % 2], ys[i % 2])
        j.backward()
        opt.step()
        return j

    steps, warm = min(args.steps, 30), max(3, min(args.warmup, 5))
    for i in range(warm):
        step(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        j = step(i)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    # the head alone on the same shapes (pooled features [B, 2048] bf16 -> 116 classes, collapsed weights), forward + backward
    feats = torch.randn(B, 2048, device=dev, dtype=torch.bfloat16, requires_grad=True)

    def head_only():
        w_eff, b_eff = model.collapsed_weights()
        model.logits.loss(feats, ys[0], weight=w_eff, bias=b_eff).backward()

    for _ in range(3):
        head_only()
    torch.cuda.synchronize()
    a.record()
    for _ in range(20):
        head_only()
    b.record()
    torch.cuda.synchronize()
    head_us = a.elapsed_time(b) / 20 * 1e3
    return {"metric": "ResNet-50 + EOS training step samples/s", "value": world * B / (ms * 1e-3), "unit": "samples/s",
            "ms_per_step": ms, "steps": steps, "warmup": warm, "head_fwd_bwd_us": head_us, "head_share": head_us * 1e-3 / ms,
            "loss": float(j.item()),
            "config": {"workload": "ResNet-50 + fused entropic head, Protocol 1 shape (116 classes), B=256 synthetic 224x224 images, "
                                   "bf16 autocast, channels_last, SGD momentum", "global_batch": B * world,
                       "head": "fc (2048->116, bias) and logits (116->116) collapsed into one 2048->116 tcgen05 head per step; "
                               "autograd through torch.autograd.Function (eager launches)"}}


def reference_head_step(cfg, rows):
    """The reference's own code for one step of the path on the host cores: nn.Linear -> losses.EntropicOpensetLoss (or the
    CrossEntropyLoss configurations of train.py:343,347) -> .item() -> backward (train.py:132-139), then the validation
    side of the same batch: softmax + metrics.confidence (train.py:177,187-192).  losses.py / metrics.py are the
    reference's files, loaded unmodified (oracle/ref_loader.py) from /root/reference or from the verbatim copies under
    oracle/_ref; only when neither exists does the op-for-op port (oracle/ref_port.py) stand in."""
    from openset_imagenet_b200 import synth as O
    from oracle import ref_loader
    torch.manual_seed(0)
    K, C = cfg["K"], cfg["C"]
    lin = torch.nn.Linear(K, C, bias=False)
    x = torch.randn(rows, K, requires_grad=True)
    if ref_loader.available():
        losses, metrics = ref_loader.load("losses"), ref_loader.load("metrics")
        kind, conf_fn = "reference", metrics.confidence
        src = f"the reference's losses.py + metrics.py ({ref_loader.source()})"
    else:
        from oracle import ref_port
        losses = None
        kind, conf_fn = "port", ref_port.confidence_port
        src = "oracle/ref_port.py (reference files not available)"
    if cfg["variant"] == "garbage":
        y_np = O.make_labels(rows, C - 1, cfg["neg_frac"], seed=7, neg_label=C - 1)
        cw = torch.tensor(O.class_weights_from_labels(np.concatenate([y_np, np.arange(C)]), C))
        loss_fn = torch.nn.CrossEntropyLoss(weight=cw)                       # train.py:346-347
        offset, unknown_class, last_valid = 0.0, C - 1, -1                   # train.py:156-159
    else:
        y_np = O.make_labels(rows, C, cfg["neg_frac"], seed=7)
        loss_fn = losses.EntropicOpensetLoss(C, 1.0) if losses is not None else ref_port.EntropicLossPort(C, 1.0)
        offset, unknown_class, last_valid = 1.0 / C, -1, None                # train.py:160-163
    y = torch.tensor(y_np)

    def step():
        lin.zero_grad(set_to_none=True)
        x.grad = None
        logits = lin(x)
        j = loss_fn(logits, y)
        val = j.item()
        j.backward()
        with torch.no_grad():
            scores = torch.nn.functional.softmax(logits, dim=1)
            conf = conf_fn(scores, y, offset=offset, unknown_class=unknown_class, last_valid_class=last_valid)
        return val, conf

    return step, kind, src


def time_reference(cfg, rows, steps, warmup):
    step, kind, src = reference_head_step(cfg, rows)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    total = sum(ts)
    return rows * len(ts) / total, total / len(ts) * 1e3, kind, src


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line of the contract, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def workload_name(cfg):
    return (f"{cfg['variant']} open-set head fwd+bwd, {cfg['K']}->{cfg['C']} classes, {cfg['B']} bf16 feature rows per GPU, "
            f"{cfg['neg_frac']:.2f} negatives (BASELINE.json {'large-head stress config' if cfg is HEAD else 'BG config'})")


def main():
    # Libraries print to stdout (NCCL's version banner does at every debug level above NONE): keep fd 1 for the
    # JSON line alone and send everything else to stderr.
    global _REAL_STDOUT
    # `kill -USR1 <pid>` (or `timeout -s USR1`) dumps every thread's Python stack to stderr: where a stuck run is waiting
    import faulthandler
    import signal
    faulthandler.register(signal.SIGUSR1, all_threads=True, chain=False)
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="head", choices=["head", "bg", "oscr", "resnet"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--eager", action="store_true", help="launch the kernels one by one instead of replaying a CUDA graph")
    ap.add_argument("--no-oscr", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the bg / resnet sub-objects of the default line")
    ap.add_argument("--sm-reserve", type=int, default=0, help="SMs left to the gradient all-reduce while dX runs (N > 1); "
                    "0 = 16")
    ap.add_argument("--no-multicast", action="store_true", help="all-reduce with peer loads instead of NVSwitch multimem (N > 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank, world, local = dist_env()
    cfg = HEAD if args.workload != "bg" else BG
    peaks = load_peaks()

    if args.impl == "reference":
        if rank != 0:
            return
        torch.set_num_threads(os.cpu_count() or 1)
        # the same configuration as our arm: the whole job's rows per step (gpus x rows per GPU), the same step and
        # warm-up counts, bounded only if a probe says the run would exceed ~4 minutes
        rows = cfg["B"] * max(args.gpus, 1)
        steps, warmup = args.steps, args.warmup
        _, probe_ms, _, _ = time_reference(cfg, min(rows, 4096), 1, 1)
        est_s = probe_ms * 1e-3 * rows / min(rows, 4096) * (steps + warmup)
        bounded = None
        if est_s > 240.0:
            steps = max(3, int(steps * 240.0 / est_s))
            warmup = min(warmup, 3)
            bounded = f"steps bounded to {steps} (+{warmup} warm-up): {args.steps} steps were estimated at {est_s:.0f} s"
        value, ms_step, kind, src = time_reference(cfg, rows, steps, warmup)
        line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "impl": "reference",
                "config": {"workload": workload_name(cfg), "global_batch": rows, "in_features": cfg["K"], "classes": cfg["C"],
                           "parallelism": f"host CPU, {torch.get_num_threads()} threads"},
                "cpu_baseline": {"value": value, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": kind,
                                 "sample": f"{rows} rows per step x {steps} steps of the same workload, float32, torch CPU "
                                           f"({torch.get_num_threads()} threads): nn.Linear -> loss -> .item() -> backward -> softmax -> "
                                           f"confidence with {src}" + (f"; {bounded}" if bounded else "")},
                "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA (sm_100a) device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_bound = False
    if world > 1:
        # host buffers and the threads that feed them next to this rank's GPU (the e2e leg copies 67 MB per step and rank)
        from openset_imagenet_b200.dist import bind_host_thread_to_gpu_node
        numa_bound = bind_host_thread_to_gpu_node(local)
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"        # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    import openset_imagenet_b200  # noqa: F401  (loads libosb.so; raises if missing)

    def sub_head(cfg_):
        r_ = run_head(args, cfg_, rank, world, dev, peaks, light=True)
        return {"metric": METRIC.replace("EOS", "BG" if cfg_ is BG else "EOS"), "value": r_["value"], "unit": "samples/s",
                "ms_per_step": r_["ms_per_step"], "us_per_step": r_["ms_per_step"] * 1e3, "steps": args.steps, "warmup": args.warmup,
                "config": {"workload": workload_name(cfg_), "global_batch": cfg_["B"] * world, "launch": r_["launch_mode"],
                           "precision": r_["precision"]},
                "gpu_launches": r_["launches"], "loss": r_["loss"], "parity": r_["parity"], "clocks": r_["clocks"]}

    if args.workload == "resnet":
        o = run_resnet(args, rank, world, dev)
        if rank == 0:
            o.update({"n_gpus": world, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic"})
            emit(o)
    elif args.workload == "oscr":
        o = run_oscr(args, rank, world, dev, peaks)
        if rank == 0:
            o.update({"n_gpus": world, "steps": 15, "warmup": 3, "ms_per_step": o["value"], "scaling": "strong", "vs_baseline": None,
                      "dtype": "f32 keys / int64 counts", "data": "synthetic", "gpu_launches": o["iters"] * 2})
            emit(o)
    else:
        r = run_head(args, cfg, rank, world, dev, peaks)
        oscr = None if args.no_oscr else run_oscr(args, rank, world, dev, peaks)
        bg = sub_head(BG) if (cfg is HEAD and not args.no_sub) else None
        resnet = run_resnet(args, rank, world, dev) if (cfg is HEAD and not args.no_sub and world == 1) else None
        if rank == 0:
            line = {"metric": METRIC, "value": r["value"], "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                    "config": {"workload": workload_name(cfg), "global_batch": cfg["B"] * world, "in_features": cfg["K"],
                               "classes": cfg["C"], "parallelism": f"batch-sharded x{world}; {r['allreduce']}" if world > 1 else "single GPU",
                               "precision": r["precision"],
                               "l2": "3 rotating 64 MB feature buffers + 64 MB dX + 32 MB dZ per step: working set > 126 MB L2",
                               "spin_up": "~30 ms of untimed replays of the same step before the W warm-up steps",
                               "launch": r["launch_mode"]},
                    "roofline": r["roofline"], "e2e": r["e2e"], "clocks": r["clocks"], "gpu_launches": r["launches"],
                    "loss": r["loss"], "parity": r["parity"], "peaks": peaks}
            if world == 1 and not args.no_e2e:
                torch.set_num_threads(os.cpu_count() or 1)
                v, ms_step, kind, src = time_reference(cfg, cfg["B"], 5, 2)
                line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": kind,
                                        "sample": f"{cfg['B']} rows per step x 5 steps (2 warm-up) of the same workload, float32 torch CPU: "
                                                  f"nn.Linear -> loss -> .item() -> backward -> softmax -> confidence with {src}"}
            if oscr is not None:
                line["oscr"] = oscr
                if "parity" in oscr:
                    line["parity"]["oscr"] = oscr["parity"]
            if bg is not None:
                line["bg"] = bg
            if resnet is not None:
                line["resnet"] = resnet
            emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
