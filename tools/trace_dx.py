"""Developer tool: per-tile clock stamps of head_dx2_kernel (MMA thread of the leader CTA and one epilogue thread)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes, numpy as np, torch
from openset_imagenet_b200 import _lib as L
dev = torch.device("cuda:0")
B, K, C = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (16384, 2048, 1000)))
_al = int(os.environ.get("OSB_LDDZ_ALIGN", "8")); lddz = (C + _al - 1) // _al * _al
dz = (torch.randn(B, lddz, device=dev) * 1e-3).bfloat16()
x = torch.randn(B, K, device=dev).bfloat16()
w = (torch.randn(C, K, device=dev) / K ** 0.5).bfloat16()
dx = torch.empty(B, K, dtype=torch.bfloat16, device=dev)
ws = torch.zeros(L.size("osb_head_workspace_bytes", B, K, C) + (1 << 20), dtype=torch.uint8, device=dev)
den = torch.ones(1, dtype=torch.float64, device=dev)
st = L.stream_ptr(dev)
lib = L.load()
lib.osb_debug_set_head_trace.argtypes = [ctypes.c_void_p]
def bdx(): L.call("osb_head_backward", dz.data_ptr(), lddz, 1, den.data_ptr(), x.data_ptr(), K, w.data_ptr(), B, K, C, None, dx.data_ptr(), K, L.BF16, None, None, int(os.environ.get("OSB_CTAS", "0")), 0, ws.data_ptr(), ws.numel(), st)
lib.osb_debug_set_head_dbg(int(os.environ.get('OSB_DBG', '0')))
for _ in range(3): bdx()
torch.cuda.synchronize()
trace = torch.zeros(512, 64, dtype=torch.int64, device=dev)
lib.osb_debug_set_head_trace(trace.data_ptr())
bdx()
torch.cuda.synchronize()
lib.osb_debug_set_head_trace(None)
t = trace.cpu().numpy()
names = ["mma_start", "first_full", "mma_issued", "epi_tfull", "epi_done"]
gbase = t[:, 0][t[:, 0] > 0].min()
for cta in (0, 1, 2, 74, 146):
    row = t[cta]
    for it in range(8):
        r = row[it * 8: it * 8 + 5]
        if r.max() == 0: continue
        print(f"cta {cta:3d} tile#{it}: " + "  ".join(f"{n}={int(v - gbase) if v else -1:7d}" for n, v in zip(names, r)))
r = t[0]
if r[8 * 6 + 7] and r[7]:
    print(f"SM clock during the launch: {(r[8 * 6 + 7] - r[7]) / max(1, (r[8 * 6] - r[0])) * 1e3:.0f} MHz")

kb = t[256:258].reshape(-1)[:64]
kb = kb[kb > 0]
if len(kb) > 1:
    print("cycles between k-block issues (CTA 0, tiles 1-2):", " ".join(str(int(v)) for v in np.diff(kb)))
