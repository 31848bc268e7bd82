"""All-reduce timing next to the head kernels (developer tool; run under torchrun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
os.environ.setdefault("NCCL_DEBUG", "WARN")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from openset_imagenet_b200.dist import ShardedHeadPlan
from openset_imagenet_b200 import synth
B, K, C = 16384, 2048, 1000
x = torch.randn(B, K, device=dev).bfloat16(); w = (torch.randn(C, K, device=dev) / 45).bfloat16()
y = torch.tensor(synth.make_labels(B, C, 1 / 3, seed=rank), device=dev)

def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n * 1e3], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

flat = torch.zeros(C * K + 1, device=dev)
flat16 = torch.zeros(C * K + 8, device=dev, dtype=torch.bfloat16)
res = {}
res["allreduce fp32 8.2MB alone"] = timeit(lambda: dist.all_reduce(flat))
res["allreduce bf16 4.1MB alone"] = timeit(lambda: dist.all_reduce(flat16))
for reserve in (0, 16, 32, 48):
    sp = ShardedHeadPlan(B, K, C, "entropic", device=dev, sm_reserve=reserve)
    res[f"step reserve={reserve}"] = timeit(lambda: sp.finish(sp.step(x, w, y)))
sp = ShardedHeadPlan(B, K, C, "entropic", device=dev, sm_reserve=16)
def no_ar():
    sp.plan.label_stats(y, sp._den_hook); sp.plan.forward(x, w, y); sp.plan.backward_dw(x, w); sp.plan.backward_dx(x, w)
res["step without all-reduce"] = timeit(no_ar)
def ar_after():
    no_ar(); dist.all_reduce(sp.flat)
res["step + all-reduce after dx (no overlap)"] = timeit(ar_after)
# ---- peer-memory all-reduce fused with the split-K reduction ----
nccl = ShardedHeadPlan(B, K, C, "entropic", device=dev, sm_reserve=32)
nccl.finish(nccl.step(x, w, y)); torch.cuda.synchronize()
want_dw, want_loss = nccl.plan.dw.clone(), nccl.plan.loss.clone()
for reserve in (8, 16, 24, 32):
    pp = ShardedHeadPlan(B, K, C, "entropic", device=dev, sm_reserve=reserve, peer_allreduce=True)
    pp.finish(pp.step(x, w, y)); torch.cuda.synchronize()
    err = (pp.plan.dw - want_dw).abs().max().item() / want_dw.abs().max().item()
    gathered = [torch.empty_like(pp.plan.dw) for _ in range(world)]
    dist.all_gather(gathered, pp.plan.dw.contiguous())
    same = all(torch.equal(g, gathered[0]) for g in gathered)
    lerr = abs(float(pp.full_loss()) - float(want_loss)) / abs(float(want_loss))
    flag_err = int(pp.peer._keep[2][2 * world].item())
    res[f"peer all-reduce step reserve={reserve}"] = timeit(lambda: pp.finish(pp.step(x, w, y)))
    if rank == 0:
        print(f"peer path reserve={reserve}: max rel diff vs NCCL {err:.2e}, loss rel diff {lerr:.2e}, identical on all ranks {same}, error word {flag_err}")
if rank == 0:
    for k, v in res.items(): print(f"{k:45s} {v:8.1f} us")
dist.barrier(); dist.destroy_process_group()
