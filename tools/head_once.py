"""Run the fused head step a few times at the benchmark size (profiling target for ncu; developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openset_imagenet_b200.functional import HeadPlan
from oracle import oracle as O

B, K, C = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (16384, 2048, 1000)))
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
dev = torch.device("cuda:0")
x = torch.randn(B, K, device=dev).bfloat16()
w = (torch.randn(C, K, device=dev) / K ** 0.5).bfloat16()
y = torch.tensor(O.make_labels(B, C, 1 / 3, seed=1), device=dev)
from openset_imagenet_b200 import _lib as L
plan = HeadPlan(B, K, C, device=dev, precision=os.environ.get("OSB_PRECISION", "fast"))
fused = plan.fused_dx and B > 4096 and os.environ.get("OSB_HEAD_ONCE_UNFUSED") is None
for _ in range(reps):
    if fused:        # the kernels of a captured step at this size: label statistics, forward + dX as one kernel, dW GEMM
        plan.label_stats(y)
        plan.forward_dx(x, w, y)
        plan.backward_dw(x, w)
    else:
        plan.step(x, w, y)
torch.cuda.synchronize()
print("loss", float(plan.loss[0]) if hasattr(plan, "loss") else "ok")
