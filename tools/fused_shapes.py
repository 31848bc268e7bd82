"""Developer tool: captured head step, us per replay, for several shapes and launch forms
(serial with forward and dX as one kernel / serial with separate kernels / parallel branches)."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openset_imagenet_b200.functional import HeadPlan
from oracle import oracle as O

dev = torch.device("cuda:0")


def time_graph(graph, n=200):
    for _ in range(30):
        graph.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            graph.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / n * 1e3)
    return statistics.median(ts)


for B, K, C in [(1024, 2048, 1000), (2048, 2048, 1000), (4096, 2048, 1000), (8192, 2048, 1000), (16384, 2048, 1000), (4096, 2048, 300), (16384, 1024, 512)]:
    x = torch.randn(B, K, device=dev).bfloat16()
    w = (torch.randn(C, K, device=dev) / K ** 0.5).bfloat16()
    y = torch.tensor(O.make_labels(B, C, 1 / 3, seed=1), device=dev)
    plan = HeadPlan(B, K, C, device=dev, precision="fast", dx_dtype=torch.bfloat16)
    res = {}
    for name, kw in (("fused", dict(parallel_backward=False, fused_dx=True)), ("serial", dict(parallel_backward=False, fused_dx=False)),
                     ("fork", dict(parallel_backward=True))):
        res[name] = time_graph(plan.capture(x, w, y, **kw))
    print(f"B={B:6d} K={K} C={C:5d}: " + "  ".join(f"{k} {v:7.1f} us" for k, v in res.items()))
