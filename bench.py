#!/usr/bin/env python
"""Benchmark of the B200-native open-set head + OSCR path (contract: see the task statement / DESIGN.md §6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload head|bg|oscr]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Default workload ("head") = BASELINE.json's large-head configuration, the largest one that fits one GPU and
the one its tensor-core target is quoted on: entropic open-set loss over a 2048 -> 1000 head on 16384
bf16 feature rows PER GPU (1/3 negatives), forward + backward (loss, confidence statistics, dX, dW).
A "step" is one such pass; with N GPUs the batch is sharded (weak scaling: 16384 rows per GPU) and the fp32
dW (+loss) buffer is all-reduced with NCCL every step.  The same JSON line carries an "oscr" object: OSCR
over 1M x 116 synthetic scores (BASELINE.json's second headline metric), sharded by rows for N > 1.

One JSON line is printed by rank 0.  `value` is device throughput with inputs resident in HBM; `e2e` is
the same step driven through the C ABI from pinned HOST buffers (H2D of features+labels, D2H of loss, dX
and dW inside the timed region).  `--impl reference` times the CPU port of the reference's own path
(oracle/ref_port.py) on the host cores, on a bounded sample of the same workload.
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

HEAD = dict(B=16384, K=2048, C=1000, variant="entropic", neg_frac=1.0 / 3.0)
BG = dict(B=1024, K=2048, C=152, variant="garbage", neg_frac=1.0 / 3.0)
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
            time.sleep(0.004)

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
        cw = torch.tensor(O.class_weights_from_labels(np.concatenate(ys_np + [np.arange(cfg["C"])]), cfg["C"]), device=dev)
    else:
        ys_np = [O.make_labels(cfg["B"], cfg["C"], cfg["neg_frac"], seed=seed + i) for i in range(nbuf)]
        cw = None
    ys = [torch.tensor(y, device=dev) for y in ys_np]
    w = (torch.randn(cfg["C"], cfg["K"], device=dev, generator=g) * (1.0 / cfg["K"] ** 0.5)).bfloat16()
    return xs, ys, ys_np, w, cw


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_head(args, cfg, rank, world, dev, peaks):
    import torch.distributed as dist
    from openset_imagenet_b200.dist import ShardedHeadPlan

    B, K, C = cfg["B"], cfg["K"], cfg["C"]
    nbuf = 3                                        # 3 x 64 MB of features + 64 MB dX + 32 MB dZ: far larger than the 126 MB L2
    xs, ys, ys_np, w, cw = make_head_inputs(cfg, dev, nbuf, seed=1234 + rank)
    plan_kw = dict(class_weights=cw, last_valid_class=-1 if cfg["variant"] == "garbage" else None, device=dev)
    allreduce = "none"
    if world > 1:
        try:            # dW all-reduce over NVLink peer memory, fused with the split-K reduction (no NCCL on the data path)
            sp = ShardedHeadPlan(B, K, C, cfg["variant"], peer_allreduce=True, sm_reserve=args.sm_reserve, **plan_kw)
            allreduce = "peer-memory kernel fused with the split-K reduction (osb_head_dw_allreduce), overlapped with dX"
        except Exception as e:                                  # no symmetric memory on this box: NCCL all-reduce
            if rank == 0:
                print(f"peer-memory all-reduce unavailable ({type(e).__name__}: {e}); using NCCL", file=sys.stderr)
            sp = ShardedHeadPlan(B, K, C, cfg["variant"], sm_reserve=args.sm_reserve, **plan_kw)
            allreduce = "NCCL all-reduce of the packed fp32 [dW | loss] buffer, overlapped with dX"
    else:
        sp = ShardedHeadPlan(B, K, C, cfg["variant"], **plan_kw)
    plan = sp.plan

    def step(i, ev=None):
        """label stats -> fused forward -> dW -> [all-reduce of dW|loss  overlapped with  dX]"""
        x, y = xs[i % nbuf], ys[i % nbuf]
        if ev is None:
            sp.finish(sp.step(x, w, y))
            return
        late_labels = sp.begin(x, w, y, mark=ev[0].record)      # (label statistics ->) forward; see ShardedHeadPlan.begin
        ev[1].record()
        if sp.peer is not None:
            plan.backward_dw_partials(x)
            ev[2].record()
            work = sp.peer.run()
        else:
            plan.backward_dw(x, w)
            ev[2].record()
            work = dist.all_reduce(sp.flat, op=dist.ReduceOp.SUM, async_op=True) if world > 1 else None
        if world > 1:
            sp._L.load().osb_set_cta_limit(sp._sms - sp.sm_reserve)
        plan.backward_dx(x, w)
        if world > 1:
            sp._L.load().osb_set_cta_limit(0)
        ev[3].record()
        sp.end(y, late_labels)
        sp.finish(work)

    # One GPU: the five launches of a step are replayed from a CUDA graph (one graph per rotating input buffer), so
    # the timed loop has no per-launch host work.  N > 1 stays eager: the all-reduce epoch changes every step.
    launch_mode = "eager"
    run_step = step
    if world == 1 and not args.eager:
        graphs = [plan.capture(xs[i], w, ys[i]) for i in range(nbuf)]
        launch_mode = "CUDA-graph replay of the 5 kernels of a step"

        def run_step(i):
            graphs[i % nbuf].replay()

    for i in range(args.warmup):
        run_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev.index) as clk:
        t0.record()
        for i in range(args.steps):
            run_step(i)
        t1.record()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    loss_val = float(sp.full_loss().item())
    ms_per_step = ms / args.steps
    value = world * B / (ms_per_step * 1e-3)

    step_us_hint = ms_per_step * 1e3
    # per-kernel times: a second, shorter loop with events between the launches (kept out of the headline loop)
    # On one GPU the loop is queued behind ~40 ms of graph replays of the same step: the host has every launch and event
    # queued before the first timed kernel starts (the events bracket kernels that run back to back, not the host's launch
    # gaps), and the GPU is in the same power state as in the headline loop (behind an idle gap the first kernels run at
    # boost clocks: 73 vs 90 us for the same forward kernel on one box).
    n_ev = max(10, min(args.steps, 50))
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(n_ev)]
    torch.cuda.synchronize()
    if world == 1 and run_step is not step:
        for i in range(max(50, int(40e3 / max(step_us_hint, 1.0)))):
            run_step(i)
    for i in range(n_ev):
        step(i, evs[i])
    torch.cuda.synchronize()
    # median over the steps: a launch the host issued late shows up as a gap inside one bracket, not in the median
    k_fwd = statistics.median(e[0].elapsed_time(e[1]) for e in evs) * 1e3
    k_dw = statistics.median(e[1].elapsed_time(e[2]) for e in evs) * 1e3
    k_dx = statistics.median(e[2].elapsed_time(e[3]) for e in evs) * 1e3

    gemm_flops = 2.0 * B * K * C
    # batches of up to ~15 row tiles of a head of <= 256 classes take the split-K forward (cluster of 8 CTAs per tile)
    splitk = C <= 256 and K >= 512 and (B + 127) // 128 <= sp._L.load().osb_debug_head_splitk_max_clusters()
    kernels = {"head_fwd_splitk_kernel" if splitk else "head_fwd_kernel": k_fwd,
               "head_dw2_kernel+dw_reduce_kernel": k_dw, "head_dx2_kernel": k_dx}
    dom = max(kernels, key=kernels.get)
    step_us = ms_per_step * 1e3
    # Duration of a kernel INSIDE the timed region: the eager brackets carry an event record on either side and none of
    # the prologue overlap of the graph-replayed step (their sum is 3-10 % above the step), so each kernel is given its
    # share of the measured step time.  The raw brackets stay in `kernels_us_eager`.
    bracket_sum = sum(kernels.values())
    eager = dict(kernels)
    if world == 1 and run_step is not step and bracket_sum > step_us:
        kernels = {k: v * step_us / bracket_sum for k, v in kernels.items()}
        duration_note = ("kernels_us = share of the CUDA-event-timed step (graph replay) each kernel takes; shares from an eager "
                         "loop with events between the launches (kernels_us_eager, medians)")
    else:
        duration_note = "kernels_us = medians of CUDA-event brackets around the launches of an eager loop"
    roof = {
        "bound": "tensor", "kernel": dom, "unit": "TFLOP/s",
        "achieved": gemm_flops / kernels[dom] / 1e6, "peak": peaks["tc_sustained"],
        "peak_kind": f"{peaks['src']} cuBLAS bf16 sustained (kernels timed inside a long step); burst is {peaks['tc_burst']}",
        "traffic": (load_traffic().get(dom, {}).get("bytes") if (B, K, C) == (16384, 2048, 1000) else None),
        "traffic_note": "dram read+write bytes of one launch from the committed ncu --set full capture (profiles/traffic.json)",
        "kernels_us": {k: round(v, 2) for k, v in kernels.items()},
        "kernels_us_eager": {k: round(v, 2) for k, v in eager.items()},
        "duration_note": duration_note,
        "kernels_tflops": {k: round(gemm_flops / v / 1e6, 1) for k, v in kernels.items()},
        "step": {"flops": 3 * gemm_flops, "us": round(step_us, 2), "achieved": 3 * gemm_flops / step_us / 1e6,
                 "frac": 3 * gemm_flops / step_us / 1e6 / peaks["tc_sustained"]},
        "algorithmic_flops_per_sample": 6.0 * K * C,
    }
    roof["frac"] = roof["achieved"] / roof["peak"]

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
                lossh.copy_(plan.loss, non_blocking=True)

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
                       "loss read back every step; dX and dW stay on the device for the backbone / optimiser"}
    return dict(value=value, ms_per_step=ms_per_step, roofline=roof, e2e=e2e, clocks=clk.summary(), loss=loss_val,
                launches=sp.launches_per_step * args.steps, allreduce=allreduce, launch_mode=launch_mode)


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
    if world > 1:
        from openset_imagenet_b200.dist import calculate_oscr_sharded
        if n_loc * C * 4 < (192 << 20):                  # shards smaller than L2: add a flush buffer to the rotation
            bufs += [torch.tensor(s_np[sl], device=dev) for _ in range(2)]

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
                      "launch": "CUDA-graph replay of the two kernels" if world == 1 else "eager (torch.distributed collectives between the stages)"},
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
        from oracle import ref_port
        t0 = time.perf_counter()
        ref_port.oscr_port_quadratic(gt_np[:n_q], s_np[:n_q], -1)
        q_ms = (time.perf_counter() - t0) * 1e3
        out["cpu_baseline"] = {"value": cpu_ms, "unit": "ms", "cores": 1, "kind": "port",
                               "sample": f"O(N log N) numpy restatement (oracle.calculate_oscr, bit-equal to the reference) on the full N={N}; "
                                         f"numpy sort is single-threaded",
                               "reference_quadratic": {"n": n_q, "ms": q_ms, "extrapolated_ms_at_1M": q_ms * (N / n_q) ** 2,
                                                       "note": "reference's own O(U*N) loop (util.py:113-118) timed at n rows; 1M figure is an N^2 extrapolation"}}
        out["parity"] = {"bit_exact_vs_oracle": bool(np.array_equal(ccr, o_ccr) and np.array_equal(fpr, o_fpr, equal_nan=True))}
    return out


def run_resnet(args, rank, world, dev):
    """BASELINE.json cfg 2: ResNet-50 (torchvision / cuDNN, bf16 autocast, channels_last) + fused entropic head,
    Protocol-1 shape (116 classes), B = 256 synthetic 224x224 images per GPU; whole training step and the head's share."""
    import openset_imagenet_b200 as osb
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
    # the head alone on the same shapes (features [B, 116] bf16 -> 116 classes), forward + backward
    feats = torch.randn(B, C, device=dev, dtype=torch.bfloat16, requires_grad=True)
    for _ in range(3):
        model.logits.loss(feats, ys[0]).backward()
    torch.cuda.synchronize()
    a.record()
    for _ in range(20):
        model.logits.loss(feats, ys[0]).backward()
    b.record()
    torch.cuda.synchronize()
    head_us = a.elapsed_time(b) / 20 * 1e3
    return {"metric": "ResNet-50 + EOS training step samples/s", "value": world * B / (ms * 1e-3), "unit": "samples/s",
            "ms_per_step": ms, "steps": steps, "warmup": warm, "head_fwd_bwd_us": head_us, "head_share": head_us * 1e-3 / ms,
            "loss": float(j.item()),
            "config": {"workload": "ResNet-50 + fused entropic head, Protocol 1 shape (116 classes), B=256 synthetic 224x224 images, "
                                   "bf16 autocast, channels_last, SGD momentum", "global_batch": B * world}}


def head_cpu_baseline(cfg, sample_rows, iters):
    """The reference path op for op on the host cores (oracle/ref_port.py), on `sample_rows` rows of the workload."""
    from openset_imagenet_b200 import synth as O
    from oracle import ref_port
    torch.manual_seed(0)
    K, C = cfg["K"], cfg["C"]
    lin = torch.nn.Linear(K, C, bias=False)
    loss_fn = ref_port.EntropicLossPort(C, 1.0)
    x = torch.randn(sample_rows, K, requires_grad=True)
    y = torch.tensor(O.make_labels(sample_rows, C, cfg["neg_frac"], seed=7))
    ref_port.head_step_port(lin, loss_fn, x, y, 1.0 / C)
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        ref_port.head_step_port(lin, loss_fn, x, y, 1.0 / C)
        ts.append(time.perf_counter() - t0)
    return sample_rows / statistics.median(ts), ts


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line of the contract, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries print to stdout (NCCL's version banner does at every debug level above NONE): keep fd 1 for the
    # JSON line alone and send everything else to stderr.
    global _REAL_STDOUT
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
    ap.add_argument("--eager", action="store_true", help="launch the kernels one by one instead of replaying a CUDA graph (N=1)")
    ap.add_argument("--no-oscr", action="store_true")
    ap.add_argument("--sm-reserve", type=int, default=32, help="SMs left to the gradient all-reduce while dX runs (N > 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank, world, local = dist_env()
    cfg = HEAD if args.workload != "bg" else BG
    workload_name = (f"{cfg['variant']} open-set head fwd+bwd, {cfg['K']}->{cfg['C']} classes, {cfg['B']} bf16 feature rows per GPU, "
                     f"{cfg['neg_frac']:.2f} negatives (BASELINE.json {'large-head stress config' if cfg is HEAD else 'BG config'})")
    peaks = load_peaks()

    if args.impl == "reference":
        if rank != 0:
            return
        torch.set_num_threads(os.cpu_count() or 1)
        # bounded sample: size the per-step sample so steps+warmup stay within ~2 minutes
        v0, ts = head_cpu_baseline(cfg, 256, 1)
        per_row = ts[0] / 256
        rows = int(min(cfg["B"], max(64, 90.0 / max(per_row * (args.steps + args.warmup), 1e-9))))
        rows = 1 << (rows.bit_length() - 1)
        rows = max(64, min(rows, 4096))
        value, ts = head_cpu_baseline(cfg, rows, max(1, min(args.steps, 200)))
        line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": len(ts), "warmup": 1,
                "ms_per_step": statistics.median(ts) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "impl": "reference",
                "config": {"workload": workload_name, "sample_rows_per_step": rows},
                "cpu_baseline": {"value": value, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                                 "sample": f"{rows} rows per step of the same workload, float32, torch CPU ({torch.get_num_threads()} threads): "
                                           "nn.Linear -> EntropicOpensetLoss port -> backward -> softmax -> confidence port (oracle/ref_port.py)"},
                "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA (sm_100a) device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"        # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    import openset_imagenet_b200  # noqa: F401  (loads libosb.so; raises if missing)

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
        if rank == 0:
            line = {"metric": METRIC, "value": r["value"], "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                    "config": {"workload": workload_name, "global_batch": cfg["B"] * world, "in_features": cfg["K"],
                               "classes": cfg["C"], "parallelism": f"batch-sharded x{world}; {r['allreduce']}" if world > 1 else "single GPU",
                               "l2": "3 rotating 64 MB feature buffers + 64 MB dX + 32 MB dZ per step: working set > 126 MB L2",
                               "launch": r["launch_mode"]},
                    "roofline": r["roofline"], "e2e": r["e2e"], "clocks": r["clocks"], "gpu_launches": r["launches"],
                    "loss": r["loss"], "peaks": peaks}
            if world == 1 and not args.no_e2e:
                torch.set_num_threads(os.cpu_count() or 1)
                rows = 2048
                v, ts = head_cpu_baseline(cfg, rows, 5)
                line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                                        "sample": f"{rows} rows of the same workload x 5 steps, float32 torch CPU: nn.Linear -> EntropicOpensetLoss "
                                                  "port -> backward -> softmax -> confidence port (oracle/ref_port.py)"}
            if oscr is not None:
                line["oscr"] = oscr
            emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
