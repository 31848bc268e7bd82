"""Phase timing of the cooperative OSCR curve kernel from its SM-clock stamps (developer tool)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from openset_imagenet_b200 import _lib as L
from openset_imagenet_b200.util import calculate_oscr_device
from oracle import oracle as O

N, C = 1_000_000, 116
dev = torch.device("cuda:0")
gt, s = O.make_oscr_case(N, C, seed=0)
st, gtt = torch.tensor(s, device=dev), torch.tensor(gt, device=dev)
for _ in range(3):
    calculate_oscr_device(gtt, st, -1)
torch.cuda.synchronize()
lib = L.load()
trace = torch.zeros(148, 32, dtype=torch.int64, device=dev)
lib.osb_debug_set_curve_trace.argtypes = [ctypes.c_void_p]
lib.osb_debug_set_curve_trace(trace.data_ptr())
calculate_oscr_device(gtt, st, -1)
torch.cuda.synchronize()
lib.osb_debug_set_curve_trace(None)
t = trace.cpu().numpy().astype(np.float64)
names = {0: "start", 1: "ranked+counts", 2: "barrier1", 3: "offsets", 4: "scattered", 5: "barrier2"}
prev = t[:, 0]
print("per-phase SM cycles: median / max over CTAs (1 cycle ~ 0.51 ns)")
for ps in range(4):
    for k in range(6):
        cur = t[:, ps * 6 + k]
        if not (ps == 0 and k == 0):
            d = cur - prev
            print(f"pass {ps} {names[k]:14s} med {np.median(d):8.0f} max {d.max():8.0f} min {d.min():8.0f}")
        prev = cur
for k, nm in [(24, "scan start"), (25, "scan tile"), (26, "barrier"), (27, "emit")]:
    cur = t[:, k]
    d = cur - prev
    print(f"scan   {nm:14s} med {np.median(d):8.0f} max {d.max():8.0f} min {d.min():8.0f}")
    prev = cur
b = t[:, 6]
for k, nm in [(28, "p1 zeroed"), (29, "p1 round 0 done"), (30, "p1 rounds done"), (31, "p1 block sync")]:
    print(f"pass 1 detail {nm:16s} med {np.median(t[:, k] - b):8.0f} (since pass start)")
print("total cycles (median CTA):", np.median(t[:, 27] - t[:, 0]))
