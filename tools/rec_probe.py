"""Timing experiments on the OSCR records kernel with parts switched off (developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openset_imagenet_b200 import _lib as L
N, C = 1_000_000, 116
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
s = torch.softmax(2 * torch.randn(N, C, device=dev, generator=g), 1)
u = torch.rand(N, device=dev, generator=g)
lab = torch.randint(0, C, (N,), device=dev, generator=g)
gt = torch.where(u < 0.5, lab, torch.where(u < 0.75, torch.full_like(lab, -1), torch.full_like(lab, -2)))
recu = torch.empty(N, dtype=torch.int64, device=dev); reck = torch.empty(N, dtype=torch.int64, device=dev)
tot = torch.zeros(4, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lib = L.load()
stp = L.stream_ptr(dev)
def run(): L.call("osb_oscr_records", s.data_ptr(), C, gt.data_ptr(), 0, N, C, -1, recu.data_ptr(), reck.data_ptr(), tot.data_ptr(), stp)
for mode in (0, 1, 2, 3):
    lib.osb_debug_set_records_mode(mode)
    ts = []
    for i in range(12):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    print(f"mode {mode}: med {ts[len(ts)//2]:.1f} us min {ts[0]:.1f} us  totals {tot.tolist()}")
lib.osb_debug_set_records_mode(0)
