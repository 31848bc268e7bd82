"""Developer tool: two captured head steps replayed back to back -- how long is the GPU between the last CTA of one step and
the first forward CTA of the next (graph launch gap), and how do the three kernels line up inside a step?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes, numpy as np, torch
from openset_imagenet_b200 import _lib as L
from openset_imagenet_b200.functional import HeadPlan
from oracle import oracle as O

dev = torch.device("cuda:0")
B, K, C = 16384, 2048, 1000
xs = [torch.randn(B, K, device=dev).bfloat16() for _ in range(2)]
w = (torch.randn(C, K, device=dev) / K ** 0.5).bfloat16()
y = torch.tensor(O.make_labels(B, C, 1 / 3, seed=1), device=dev)
plan = HeadPlan(B, K, C, device=dev, precision="fast", dx_dtype=torch.bfloat16)
lib = L.load()
for f in (lib.osb_debug_set_head_trace, lib.osb_debug_set_dw_trace):
    f.argtypes = [ctypes.c_void_p]
tr = [torch.zeros(512, 64, dtype=torch.int64, device=dev) for _ in range(2)]
trw = [torch.zeros(160, 8, dtype=torch.int64, device=dev) for _ in range(2)]
graphs = []
for i in range(2):
    lib.osb_debug_set_head_trace(tr[i].data_ptr()); lib.osb_debug_set_dw_trace(trw[i].data_ptr())
    graphs.append(plan.capture(xs[i], w, y))
lib.osb_debug_set_head_trace(None); lib.osb_debug_set_dw_trace(None)
for _ in range(30):
    graphs[0].replay(); graphs[1].replay()
torch.cuda.synchronize()
for t_ in tr + trw:
    t_.zero_()
torch.cuda.synchronize()
for _ in range(6):
    graphs[0].replay(); graphs[1].replay()
torch.cuda.synchronize()


def spans(t, tw):
    t, tw = t.cpu().numpy(), tw.cpu().numpy()
    fw = t[384:389].reshape(-1)[:320].reshape(-1, 2); fw = fw[fw[:, 0] > 0]
    dw = tw[tw[:, 0] > 0]
    dx = t[:148]; first = dx[:, 0]; ok = first > 0
    ends = np.array([row[[it * 8 + 4 for it in range(8)]].max() for row in dx])
    return dict(f0=fw[:, 0].min(), f1=fw[:, 1].max(), x0=first[ok].min(), x1=ends[ok].max(), w0=dw[:, 0].min(), w1=dw[:, 3].max())


a, b = spans(tr[0], trw[0]), spans(tr[1], trw[1])
for nm, s in (("step A", a), ("step B", b)):
    z = s["f0"]
    print(f"{nm}: forward 0 -> {(s['f1'] - z) / 1e3:.1f} | dX {(s['x0'] - z) / 1e3:.1f} -> {(s['x1'] - z) / 1e3:.1f} | dW {(s['w0'] - z) / 1e3:.1f} -> {(s['w1'] - z) / 1e3:.1f} us")
print(f"gap between the last CTA of step A and the first forward CTA of step B: {(b['f0'] - max(a['w1'], a['x1'])) / 1e3:.1f} us")
print(f"period A -> B (forward start to forward start): {(b['f0'] - a['f0']) / 1e3:.1f} us")
