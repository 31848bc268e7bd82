"""Developer tool: per-tile clock stamps of head_fwd_kernel (MMA thread and one epilogue thread)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes, numpy as np, torch
from openset_imagenet_b200 import _lib as L
from openset_imagenet_b200.functional import HeadPlan
from oracle import oracle as O
dev = torch.device("cuda:0")
B, K, C = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (16384, 2048, 1000)))
x = torch.randn(B, K, device=dev).bfloat16()
w = (torch.randn(C, K, device=dev) / K ** 0.5).bfloat16()
y = torch.tensor(O.make_labels(B, C, 1 / 3, seed=1), device=dev)
plan = HeadPlan(B, K, C, device=dev, precision="fast")
lib = L.load()
lib.osb_debug_set_head_trace.argtypes = [ctypes.c_void_p]
trace = torch.zeros(512, 64, dtype=torch.int64, device=dev)
plan.label_stats(y)
for _ in range(3): plan.forward(x, w, y)
torch.cuda.synchronize()
lib.osb_debug_set_head_trace(trace.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
plan.forward(x, w, y, write_dz=len(sys.argv) <= 4)
e1.record()
torch.cuda.synchronize()
print(f"event time of the traced launch: {e0.elapsed_time(e1) * 1e3:.1f} us")
lib.osb_debug_set_head_trace(None)
t = trace.cpu().numpy()
names = ["mma_start", "first_A_full", "mma_issued", "epi_tfull", "sweepA_done", "xchg_done", "sweepB_done"]
gbase = t[:, 0][t[:, 0] > 0].min()
nodz = len(sys.argv) > 4
for cta in (0, 2, 4, 6, 121):
    row = t[cta]
    if row[0] == 0 and row[3] == 0: continue
    base = gbase
    for it in range(4):
        r = row[it * 8: it * 8 + 7]
        if r.max() == 0: continue
        print(f"cta {cta:3d} tile#{it}: " + "  ".join(f"{n}={int(v - base):7d}" for n, v in zip(names, r)))

r = t[0]
if r[8 * 3 + 7] and r[7]:
    print(f"SM clock during the launch: {(r[8 * 3 + 7] - r[7]) / max(1, (r[8 * 3] - r[0])) * 1e3:.0f} MHz")

kb = t[256:258].reshape(-1)
kb = kb[kb > 0]
if len(kb) > 1:
    print("cycles between k-block issues (CTA 0, tiles 0-3):", " ".join(str(int(v)) for v in np.diff(kb)))

sub = t[258:259].reshape(-1)[:32].reshape(4, 8)
for it in range(4):
    b = t[0][it * 8 + 3]
    if b:
        names2 = ["tfull", "ld_done", "max_done", "expA(4)", "partner", "xchg(5)", "dz_done", "outputs(6)"]
        vals = [b, sub[it][0], sub[it][1], t[0][it * 8 + 4], sub[it][2], t[0][it * 8 + 5], sub[it][3], t[0][it * 8 + 6]]
        print(f"cta 0 epilogue tile#{it} (ns after t_full): " + "  ".join(f"{n}={int(v - b)}" for n, v in zip(names2, vals)))

se = t[384:392].reshape(-1)[:512].reshape(256, 2)
se = se[se[:, 0] > 0]
if len(se):
    print(f"CTA entry: first {int(se[:, 0].min() - gbase)} last {int(se[:, 0].max() - gbase)}; CTA exit: first {int(se[:, 1].min() - gbase)} last {int(se[:, 1].max() - gbase)} (ns relative to the first MMA start); span {int(se[:, 1].max() - se[:, 0].min())} ns over {len(se)} CTAs")

ex = t[259].reshape(-1)
b = t[0][3 * 8 + 6]
if b:
    print("cta 0 exit path (ns after its last tile's epilogue end):", " ".join(f"{n}={int(ex[i] - b)}" for i, n in enumerate(["loop_end", "tma_wait", "ticket", "cluster_sync"]) if ex[i]),
          "| warps reach the final sync:", " ".join(str(int(ex[8 + w] - b)) for w in range(10) if ex[8 + w]),
          "| CTA0 entry->exit:", int(t[384][1] - t[384][0]))
