"""Developer probe: is a GEMM tile paced by its own SM (per-SM L2 port) or by the chip (L2 / crossbar throughput)?
Runs the dX GEMM and the forward on fewer and fewer CTAs and prints the time one CTA pair spends per tile: constant ->
per-SM limit; falling with fewer CTAs -> the SMs slow each other down through the shared L2 path."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openset_imagenet_b200 import _lib as L
from openset_imagenet_b200.functional import HeadPlan

dev = torch.device("cuda:0")
B, K, C = 16384, 2048, 1000
plan = HeadPlan(B, K, C, "entropic", device=dev, precision="fast", dx_dtype=torch.bfloat16)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(B, K, device=dev, generator=g).bfloat16()
w = (torch.randn(C, K, device=dev, generator=g) / K ** 0.5).bfloat16()
y = torch.randint(0, C, (B,), device=dev)
plan.step(x, w, y)
torch.cuda.synchronize()


def timed(fn, n=12):
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)


for ctas in (148, 128, 112, 96, 74, 64, 38, 20, 8):
    t = timed(lambda: plan.backward_dx(x, w, max_ctas=ctas))
    pairs = ctas // 2
    rounds = -(-512 // pairs)
    print(f"dX  ctas {ctas:3d}: {t:7.1f} us  rounds {rounds:3d}  us per tile and pair {t / rounds:5.2f}")
for ctas in (148, 64, 32, 16, 8):
    t = timed(lambda: plan.forward(x, w, y, want_pred=False, max_ctas=ctas))
    groups = max(1, ctas // 8)
    rounds = -(-64 // groups)
    groups = -(-64 // rounds)
    print(f"fwd ctas {ctas:3d}: {t:7.1f} us  groups {groups:2d} rounds {rounds:3d}  us per round {t / rounds:5.2f}")
