"""Developer tool: phase stamps (globaltimer, ns) of head_fwd_splitk_kernel, one row per CTA."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes, numpy as np, torch
from openset_imagenet_b200 import _lib as L
from openset_imagenet_b200.functional import HeadPlan
from oracle import oracle as O
dev = torch.device("cuda:0")
B, K, C = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (1024, 2048, 152)))
x = torch.randn(B, K, device=dev).bfloat16()
w = (torch.randn(C, K, device=dev) / K ** 0.5).bfloat16()
y = torch.tensor(O.make_labels(B, C, 1 / 3, seed=1), device=dev)
plan = HeadPlan(B, K, C, device=dev, precision="fast")
lib = L.load()
lib.osb_debug_set_head_trace.argtypes = [ctypes.c_void_p]
trace = torch.zeros(512, 64, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
plan.label_stats(y)
for _ in range(3): plan.forward(x, w, y)
torch.cuda.synchronize()
names = ["entry", "setup", "mma_done", "sync2", "dsmem_out", "sync3", "reduced", "dz_out", "loss_done", "exit", "lds_j0", "argmax", "exp", "sums"]
order = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9]
big = torch.randn(8192, 8192, device=dev).bfloat16()
for mode in ("warm", "cold-L2", "hot-clocks"):
    if mode == "cold-L2": flush.fill_(1)
    if mode == "hot-clocks":
        for _ in range(40): big @ big          # ~30 ms of tensor work right before the traced launch
        for _ in range(5): plan.forward(x, w, y, want_pred=False)
    trace.zero_()
    lib.osb_debug_set_head_trace(trace.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.forward(x, w, y, want_pred=False); e1.record()
    torch.cuda.synchronize()
    lib.osb_debug_set_head_trace(None)
    full = trace.cpu().numpy()
    t = full[:, :14]
    live = t[:, 0] > 0
    t = t[live]
    fl = full[live]
    mhz = np.median((fl[:, 21] - fl[:, 20]) / np.maximum(fl[:, 9] - fl[:, 0], 1)) * 1e3
    print(f"    SM clock during the kernel: {mhz:.0f} MHz")
    base = t[:, 0].min()
    print(f"--- {mode}: event time {e0.elapsed_time(e1) * 1e3:.1f} us, {live.sum()} CTAs; ns since the first CTA's entry (median / max over CTAs)")
    for i in order:
        n = names[i]
        col = t[:, i] - base
        print(f"  {n:10s} median {int(np.median(col)):7d}  max {int(col.max()):7d}  min {int(col.min()):7d}")
