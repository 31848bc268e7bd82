"""Where does a sharded small-batch step spend its time?  torchrun -N 2 tools/shard_probe.py [B K C variant]
Times graph replays of the step with parts removed (no all-reduce, no denominator exchange) and the eager step."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from openset_imagenet_b200 import synth, _lib as L
from openset_imagenet_b200.dist import ShardedHeadPlan

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, K, C = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (1024, 2048, 152)
variant = sys.argv[4] if len(sys.argv) > 4 else "garbage"
reserve = int(os.environ.get("RESERVE", "16"))
g = torch.Generator(device=dev).manual_seed(rank)
x = torch.randn(B, K, device=dev, generator=g).bfloat16()
w = (torch.randn(C, K, device=dev, generator=torch.Generator(device=dev).manual_seed(7)) / K ** 0.5).bfloat16()
if variant == "garbage":
    y = torch.tensor(synth.make_labels(B, C - 1, 1 / 3, seed=rank, neg_label=C - 1), device=dev)
    kw = dict(class_weights=torch.ones(C, device=dev), last_valid_class=-1)
else:
    y = torch.tensor(synth.make_labels(B, C, 1 / 3, seed=rank), device=dev); kw = {}
kw["precision"] = os.environ.get("OSB_PRECISION", "fast" if C > 256 else "exact")

import time
host_us = {}
def timeit(fn, n=200, warm=20, name=None):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    h0 = time.perf_counter()
    for _ in range(n): fn()
    h1 = time.perf_counter()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n * 1e3, (h1 - h0) / n * 1e6], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if name: host_us[name] = float(t[1])
    return float(t[0])

sp = ShardedHeadPlan(B, K, C, variant, peer_allreduce=True, sm_reserve=reserve, device=dev, **kw)
res = {}
# NVSwitch multicast (multimem.ld_reduce / multimem.st) against the peer-load path: same gradients?
sp_nomc = ShardedHeadPlan(B, K, C, variant, peer_allreduce=True, sm_reserve=reserve, device=dev, multicast=False, **kw)
res["multicast available"] = float(sp.peer.mc_partials is not None)
for s_ in (sp, sp_nomc):
    s_.finish(s_.step(x, w, y))
torch.cuda.synchronize()
a_, b_ = sp.plan.dw.double(), sp_nomc.plan.dw.double()
res["dW multicast vs peer loads, rel"] = float((a_ - b_).norm() / b_.norm()) * 1e6
res["  (x 1e-6)  loss diff"] = float((sp.full_loss() - sp_nomc.full_loss()).abs()) * 1e6
gn = sp_nomc.capture(x, w, y)
res["graph full, peer loads (no multicast)"] = timeit(gn.replay)
sp_nomc.plan.backward_dw_partials(x, local_reduce=sp_nomc.peer.local_reduce)
res["all-reduce alone, peer loads"] = timeit(lambda: sp_nomc.peer.launch(sp_nomc.den_global))
graph = sp.capture(x, w, y)
res["graph full"] = timeit(graph.replay, name="graph full")
res["eager full"] = timeit(lambda: sp.finish(sp.step(x, w, y)))
# without the all-reduce kernel
orig_launch = sp.peer.launch
sp.peer.launch = lambda *a, **k: None
g2 = sp.capture(x, w, y)
res["graph, no all-reduce"] = timeit(g2.replay, name="graph, no all-reduce")
sp.peer.launch = orig_launch
# all-reduce alone, back to back (both barriers, no GEMMs)
sp.plan.backward_dw_partials(x, local_reduce=sp.peer.local_reduce)
res["all-reduce kernel alone"] = timeit(lambda: sp.peer.launch(sp.den_global))
if variant != "entropic":
    res["peer_sum alone"] = timeit(lambda: L.call("osb_peer_sum", L.ptr(sp.plan.den), L.ptr(sp.den_global), 1, sp.peer.slot_ptrs, rank, world, L.stream_ptr(dev)))
# single-GPU plan graph for reference
g1 = sp.plan.capture(x, w, y)
res["single-GPU graph (no exchange)"] = timeit(g1.replay, name="single-GPU graph")
# the bench harness: 3 rotating input buffers / graphs, 5 warm-up + 20 timed steps behind a barrier, several trials
xs3 = [x] + [torch.randn(B, K, device=dev, generator=g).bfloat16() for _ in range(2)]
graphs = [graph] + [sp.capture(xx, w, y) for xx in xs3[1:]]
def trial(sampler=False):
    import threading, time
    stop = threading.Event(); th = None
    for i in range(5): graphs[i % 3].replay()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    if sampler:
        import pynvml
        pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(local)
        def run():
            while not stop.is_set():
                pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                try: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception: pass
                time.sleep(0.0005)
        th = threading.Thread(target=run, daemon=True); th.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(20): graphs[i % 3].replay()
    b.record(); torch.cuda.synchronize()
    if th is not None: stop.set(); th.join()
    t = torch.tensor([a.elapsed_time(b) / 20 * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t), 1)
res["bench-like trials (20 steps)"] = [trial() for _ in range(6)]
res["bench-like trials + NVML sampler thread"] = [trial(True) for _ in range(6)]
sp.check()
if rank == 0:
    for k, v in res.items(): print(f"{k:34s} {v if isinstance(v, list) else round(v, 1)} us")
    print("host us per call (max over ranks):", {k: round(v, 1) for k, v in host_us.items()}, "cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
dist.barrier(); dist.destroy_process_group()
