"""Developer tool: one replay of the whole captured head step on a time line -- when do the CTAs of the forward, the dW
GEMM and the dX GEMM start real work and finish (globaltimer stamps of the kernels' trace hooks)?
    python tools/timeline.py [serial|fork|handover]        (OSB_FWD_ALL_GROUPS=1: forward on 18 groups)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes, numpy as np, torch
from openset_imagenet_b200 import _lib as L
from openset_imagenet_b200.functional import HeadPlan
from oracle import oracle as O

mode = sys.argv[1] if len(sys.argv) > 1 else "serial"
dev = torch.device("cuda:0")
B, K, C = 16384, 2048, 1000
x = torch.randn(B, K, device=dev).bfloat16()
w = (torch.randn(C, K, device=dev) / K ** 0.5).bfloat16()
y = torch.tensor(O.make_labels(B, C, 1 / 3, seed=1), device=dev)
plan = HeadPlan(B, K, C, device=dev, precision="fast", dx_dtype=torch.bfloat16)
lib = L.load()
for f in (lib.osb_debug_set_head_trace, lib.osb_debug_set_dw_trace):
    f.argtypes = [ctypes.c_void_p]
trace = torch.zeros(512, 64, dtype=torch.int64, device=dev)
trace_dw = torch.zeros(160, 8, dtype=torch.int64, device=dev)
lib.osb_debug_set_head_trace(trace.data_ptr())          # the pointers are baked into the captured launches
lib.osb_debug_set_dw_trace(trace_dw.data_ptr())
kw = {"serial": dict(parallel_backward=False), "fork": dict(parallel_backward=True), "handover": dict(parallel_backward=True, handover=True)}[mode]
graph = plan.capture(x, w, y, **kw)
for _ in range(20):
    graph.replay()
torch.cuda.synchronize()
trace.zero_(); trace_dw.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); graph.replay(); e1.record()
torch.cuda.synchronize()
lib.osb_debug_set_head_trace(None); lib.osb_debug_set_dw_trace(None)
t, tw = trace.cpu().numpy(), trace_dw.cpu().numpy()
fw = t[384:389].reshape(-1)[:2 * 160].reshape(-1, 2)
fw = fw[fw[:, 0] > 0]
t0 = fw[:, 0].min()
us = lambda v: (np.asarray(v, dtype=np.float64) - t0) / 1e3
q = lambda v: f"min {np.min(v):6.1f}  med {np.median(v):6.1f}  max {np.max(v):6.1f}"
print(f"{mode}: event time of the replay {e0.elapsed_time(e1) * 1e3:.1f} us; times below in us after the first forward CTA started")
print(f"forward  {len(fw):3d} CTAs   entry {q(us(fw[:, 0]))}   exit {q(us(fw[:, 1]))}")
dw = tw[tw[:, 0] > 0]
print(f"dW GEMM  {len(dw):3d} CTAs   entry {q(us(dw[:, 0]))}   mainloop done {q(us(dw[:, 1]))}   exit {q(us(dw[:, 3]))}")
order = np.argsort(dw[:, 0])
early = dw[order][:48:2]
print("   first CTAs by entry (entry -> mainloop done -> exit): " + "  ".join(f"{a:.0f}>{b:.0f}>{c:.0f}" for a, b, c in zip(us(early[:, 0]), us(early[:, 1]), us(early[:, 3]))))
late = dw[order][-8::2]
print("   last CTAs by entry: " + "  ".join(f"{a:.0f}>{b:.0f}>{c:.0f}" for a, b, c in zip(us(late[:, 0]), us(late[:, 1]), us(late[:, 3]))))
dx = t[:148]
first = dx[:, 0]
ends = np.array([row[[it * 8 + 4 for it in range(8)]].max() for row in dx])
ok = first > 0
print(f"dX GEMM  {int(ok.sum()):3d} CTAs   first MMA {q(us(first[ok]))}   last tile drained {q(us(ends[ok]))}")
print(f"step end (last stamp) {max(us(ends[ok]).max(), us(dw[:, 3]).max()):.1f} us")
