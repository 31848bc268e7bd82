"""Run the OSCR path a few times at the benchmark size (profiling target for ncu; developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from openset_imagenet_b200.util import calculate_oscr_device

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
C = int(sys.argv[2]) if len(sys.argv) > 2 else 116
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
z = 2 * torch.randn(N, C, device=dev, generator=g)
lab = torch.randint(0, C, (N,), device=dev, generator=g)
u = torch.rand(N, device=dev, generator=g)
gt = torch.where(u < 0.5, lab, torch.where(u < 0.75, torch.full_like(lab, -1), torch.full_like(lab, -2)))
z[torch.arange(N, device=dev)[gt >= 0], gt[gt >= 0]] += 4
s = torch.softmax(z, 1)
del z
for _ in range(reps):
    r = calculate_oscr_device(gt, s, -1)
torch.cuda.synchronize()
print("meta", r.meta.tolist())
