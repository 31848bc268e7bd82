"""Do N GPUs of one box slow each other down?  torchrun -N 8 tools/multi_gpu_load_probe.py
Every rank replays the SINGLE-GPU head step graph (no communication at all) for ~1.5 s, first alone (ranks take turns), then
all together, and reports its own step time, SM clock, power draw and throttle reasons."""
import os, sys, time, threading, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist, pynvml
from openset_imagenet_b200 import synth
from openset_imagenet_b200.functional import HeadPlan

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, K, C = 16384, 2048, 1000
g = torch.Generator(device=dev).manual_seed(rank)
xs = [torch.randn(B, K, device=dev, generator=g).bfloat16() for _ in range(3)]
w = (torch.randn(C, K, device=dev, generator=g) / K ** 0.5).bfloat16()
y = torch.tensor(synth.make_labels(B, C, 1 / 3, seed=rank), device=dev)
plan = HeadPlan(B, K, C, "entropic", device=dev, precision="fast")
graphs = [plan.capture(x, w, y) for x in xs]
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(local)

def load(seconds):
    stop = threading.Event(); clk, pw, rs = [], [], set()
    def samp():
        while not stop.is_set():
            clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)); pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)
            r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
            for name, bit in (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal", 0x20), ("hw_thermal", 0x40), ("hw_power_brake", 0x80)):
                if r & bit: rs.add(name)
            time.sleep(0.01)
    th = threading.Thread(target=samp, daemon=True); th.start()
    n = int(seconds / 180e-6)
    for i in range(50): graphs[i % 3].replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): graphs[i % 3].replay()
    b.record(); torch.cuda.synchronize()
    stop.set(); th.join()
    half = clk[len(clk) // 2:] or [0]
    return a.elapsed_time(b) / n * 1e3, statistics.median(half), max(pw or [0]), sorted(rs)

def rate(fn, n=300):
    for i in range(30): fn(i)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): fn(i)
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) / n * 1e3, 1)

if os.environ.get("VARIANTS"):
    from openset_imagenet_b200.dist import ShardedHeadPlan
    res = {}
    res["A plain HeadPlan graph, all ranks at once"] = rate(lambda i: graphs[i % 3].replay())
    sp = ShardedHeadPlan(B, K, C, "entropic", peer_allreduce=True, sm_reserve=16, device=dev, precision="fast")
    res["A2 same, after symmetric memory + peer set-up"] = rate(lambda i: graphs[i % 3].replay())
    gB = [sp.plan.capture(x, w, y) for x in xs]
    res["B HeadPlan graph on symmetric-memory buffers, no exchange"] = rate(lambda i: gB[i % 3].replay())
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    def eager_marks(i):
        p = sp.plan
        evs[0].record(); p.forward(xs[i % 3], w, y, want_pred=False); evs[1].record(); p.backward_dx(xs[i % 3], w); evs[2].record()
        p.backward_dw_partials(xs[i % 3]); evs[3].record(); p.label_stats(y); evs[4].record()
    for i in range(20): eager_marks(i)
    torch.cuda.synchronize()
    res["B eager brackets fwd/dx/dw-partials/labels"] = [round(evs[j].elapsed_time(evs[j + 1]) * 1e3, 1) for j in range(4)]
    gD = [sp.capture(x, w, y) for x in xs]
    res["D sharded graph (peer all-reduce)"] = rate(lambda i: gD[i % 3].replay())
    res["A3 plain HeadPlan graph again"] = rate(lambda i: graphs[i % 3].replay())
    out = [None] * world
    dist.all_gather_object(out, res)
    if rank == 0:
        for k in res:
            print(f"{k:62s}", [o[k] for o in out])
    dist.barrier(); dist.destroy_process_group(); sys.exit(0)

alone = None
for r in range(world):
    dist.barrier()
    if r == rank: alone = load(0.5)
    torch.cuda.synchronize()
dist.barrier()
together = load(1.5)
out = [None] * world
dist.all_gather_object(out, (rank, alone, together, pynvml.nvmlDeviceGetPowerManagementLimit(h) / 1e3, pynvml.nvmlDeviceGetTemperature(h, 0)))
if rank == 0:
    print("rank | alone: us/step  MHz  W  reasons | together: us/step  MHz  W  reasons | limit W, temp C")
    for r, a, t, lim, temp in out:
        print(f"{r:4d} | {a[0]:7.1f} {a[1]:6.0f} {a[2]:6.0f} {a[3]} | {t[0]:7.1f} {t[1]:6.0f} {t[2]:6.0f} {t[3]} | {lim:.0f} {temp}")
dist.barrier(); dist.destroy_process_group()
