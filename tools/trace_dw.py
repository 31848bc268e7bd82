"""Developer tool: clock stamps of head_dw2_kernel (entry, accumulator complete, drain done, exit)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes, numpy as np, torch
from openset_imagenet_b200 import _lib as L
dev = torch.device("cuda:0")
B, K, C = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (16384, 2048, 1000)))
lddz = (C + 7) // 8 * 8
dz = (torch.randn(B, lddz, device=dev) * 1e-3).bfloat16()
x = torch.randn(B, K, device=dev).bfloat16()
dw = torch.empty(C, K, device=dev)
ws = torch.zeros(L.size("osb_head_workspace_bytes", B, K, C) + (1 << 20), dtype=torch.uint8, device=dev)
den = torch.ones(1, dtype=torch.float64, device=dev)
st = L.stream_ptr(dev)
lib = L.load()
lib.osb_debug_set_head_trace.argtypes = [ctypes.c_void_p]
def bdw(): L.call("osb_head_backward", dz.data_ptr(), lddz, 1, den.data_ptr(), x.data_ptr(), K, x.data_ptr(), B, K, C, None, None, K, L.BF16, dw.data_ptr(), None, 0, 0, ws.data_ptr(), ws.numel(), st)
for _ in range(3): bdw()
torch.cuda.synchronize()
trace = torch.zeros(512, 64, dtype=torch.int64, device=dev)
lib.osb_debug_set_head_trace(trace.data_ptr())
bdw()
torch.cuda.synchronize()
lib.osb_debug_set_head_trace(None)
t = trace.cpu().numpy().reshape(-1)[:148 * 8].reshape(148, 8)
t = t[t[:, 0] > 0]
b = t[:, 0].min()
print(f"{len(t)} CTAs; entry {int(t[:,0].min()-b)}..{int(t[:,0].max()-b)}  accumulator complete {int(t[:,1].min()-b)}..{int(t[:,1].max()-b)}  drain done {int(t[:,2].min()-b)}..{int(t[:,2].max()-b)}  exit {int(t[:,3].min()-b)}..{int(t[:,3].max()-b)} (ns)")
print("per-CTA drain time (ns): min", int((t[:,2]-t[:,1]).min()), "median", int(np.median(t[:,2]-t[:,1])), "max", int((t[:,2]-t[:,1]).max()))
