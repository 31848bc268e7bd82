"""Per-kernel timing of the fused head and the OSCR stages (developer tool; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from openset_imagenet_b200 import _lib as L
from openset_imagenet_b200.model import FusedOpenSetHead
from oracle import oracle as O

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

FLUSH_READ = os.environ.get("OSB_FLUSH") == "read"      # evict by reading (no dirty lines left in L2) instead of writing

def timeit(fn, iters=20, warm=3, flush_l2=True):
    for _ in range(warm): fn()
    ts = []
    for _ in range(iters):
        if flush_l2:
            if FLUSH_READ: flush.view(torch.int32).sum()
            else: flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]

def head_case(B, K, C, variant="entropic"):
    x = torch.randn(B, K, device=dev).bfloat16()
    y = torch.tensor(O.make_labels(B, C, 1 / 3, seed=1), device=dev)
    cw = torch.ones(C, device=dev) if variant == "garbage" else None
    if variant == "garbage": y = y.clamp(min=0)
    head = FusedOpenSetHead(K, C, variant=variant, class_weights=cw, device=dev)
    w = head.weight.detach().bfloat16().contiguous()
    st = L.stream_ptr(dev)
    ws = torch.zeros(L.size("osb_head_workspace_bytes", B, K, C) + (1 << 20), dtype=torch.uint8, device=dev)
    wsl = torch.zeros(L.size("osb_loss_workspace_bytes", B) + 256, dtype=torch.uint8, device=dev)
    PLANES = int(os.environ.get("OSB_PLANES", "1"))
    den = torch.full((1,), float(B), dtype=torch.float64, device=dev)
    counts = torch.empty(4, dtype=torch.int64, device=dev)
    loss = torch.empty(1, device=dev); num = torch.empty(1, dtype=torch.float64, device=dev)
    stats = torch.empty(3, B, device=dev); pred = torch.empty(B, dtype=torch.int32, device=dev)
    _al = int(os.environ.get("OSB_LDDZ_ALIGN", "8")); lddz = (C + _al - 1) // _al * _al
    dz = torch.empty(PLANES, B, lddz, dtype=torch.bfloat16, device=dev)
    dx = torch.empty(B, K, dtype=torch.bfloat16, device=dev); dw = torch.empty(C, K, device=dev)
    vnum = {"entropic": 0, "softmax": 1, "garbage": 2}[variant]
    def lab(): L.call("osb_label_stats", y.data_ptr(), L.ptr(cw), B, C, vnum, -100, den.data_ptr(), counts.data_ptr(), wsl.data_ptr(), wsl.numel(), 0, st)
    def fwd(): L.call("osb_head_forward", x.data_ptr(), K, w.data_ptr(), None, y.data_ptr(), L.ptr(cw), B, K, C, vnum, 1.0, -100, den.data_ptr(), C,
                      num.data_ptr(), loss.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(), stats[2].data_ptr(), pred.data_ptr(), dz.data_ptr(), lddz, PLANES, 0, 0, ws.data_ptr(), ws.numel(), st)
    def fwd_nodz(): L.call("osb_head_forward", x.data_ptr(), K, w.data_ptr(), None, y.data_ptr(), L.ptr(cw), B, K, C, vnum, 1.0, -100, den.data_ptr(), C,
                      num.data_ptr(), loss.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(), stats[2].data_ptr(), pred.data_ptr(), None, lddz, PLANES, 0, 0, ws.data_ptr(), ws.numel(), st)
    def bdx(): L.call("osb_head_backward", dz.data_ptr(), lddz, PLANES, den.data_ptr(), x.data_ptr(), K, w.data_ptr(), B, K, C, None, dx.data_ptr(), K, L.BF16, None, None, 0, 0, ws.data_ptr(), ws.numel(), st)
    def bdw(): L.call("osb_head_backward", dz.data_ptr(), lddz, PLANES, den.data_ptr(), x.data_ptr(), K, w.data_ptr(), B, K, C, None, None, K, L.BF16, dw.data_ptr(), None, 0, 0, ws.data_ptr(), ws.numel(), st)
    def bdwp(): L.call("osb_head_backward_dw_partials", dz.data_ptr(), lddz, PLANES, x.data_ptr(), K, B, K, C, None, 0, 0, ws.data_ptr(), ws.numel(), st)
    def step(): lab(); fwd(); bdx(); bdw()
    fl = 2.0 * B * K * C
    print(f"--- head B={B} K={K} C={C} {variant}: GEMM flops each {fl/1e9:.2f} GF")
    lib = L.load()
    lib.osb_debug_set_head_dbg(int(os.environ.get('OSB_DBG', '0')))
    def fwd_persist():
        lib.osb_debug_set_head_splitk(0); fwd(); lib.osb_debug_set_head_splitk(1)
    for name, fn, f in [("label_stats", lab, 0), ("fwd(K1)", fwd, fl), ("fwd persistent", fwd_persist, fl), ("fwd no dz", fwd_nodz, fl), ("dx(K2)", bdx, fl),
                        ("dw(K3+reduce)", bdw, fl), ("dw GEMM only", bdwp, fl), ("step", step, 3 * fl)]:
        med, mn = timeit(fn)
        med2, mn2 = timeit(fn, flush_l2=False)
        print(f"{name:16s} cold-L2 med {med:9.1f} us min {mn:9.1f} us | warm med {med2:9.1f} us" + (f" | {f/med/1e6:8.1f} TF/s cold, {f/med2/1e6:8.1f} warm" if f else ""))
    # whole step through HeadPlan: eager (5 launches) and replayed from a CUDA graph
    from openset_imagenet_b200.functional import HeadPlan
    plan = HeadPlan(B, K, C, variant, class_weights=cw, last_valid_class=-1 if variant == "garbage" else None, device=dev)
    graph = plan.capture(x, w, y)
    for name, fn in [("plan.step eager", lambda: plan.step(x, w, y)), ("plan graph", graph.replay)]:
        med, mn = timeit(fn); med2, mn2 = timeit(fn, flush_l2=False)
        print(f"{name:16s} cold-L2 med {med:9.1f} us min {mn:9.1f} us | warm med {med2:9.1f} us | {3*fl/med/1e6:8.1f} TF/s cold")
    # cuBLAS reference points for the same three GEMMs
    xf, wf, dzf = x, w, dz[:, :C].contiguous()
    for name, fn in [("cublas X@W^T", lambda: xf @ wf.t()), ("cublas dZ@W", lambda: dzf @ wf), ("cublas dZ^T@X", lambda: dzf.t() @ xf)]:
        med, mn = timeit(fn); print(f"{name:16s} cold-L2 med {med:9.1f} us min {mn:9.1f} us | {fl/med/1e6:8.1f} TF/s")

def oscr_case(N, C):
    """Inputs are larger than L2 and two score buffers alternate, so no iteration finds its input in L2 and no
    dirty flush lines have to be written back inside the timed region."""
    from openset_imagenet_b200.util import calculate_oscr_device
    gt, s = O.make_oscr_case(N, C, seed=0)
    sts = [torch.tensor(s, device=dev), torch.tensor(s, device=dev)]
    gtt = torch.tensor(gt, device=dev)
    stp = L.stream_ptr(dev)
    recu = torch.empty(N, dtype=torch.int64, device=dev); reck = torch.empty(N, dtype=torch.int64, device=dev)
    totals = torch.zeros(4, dtype=torch.int64, device=dev)
    cws = torch.zeros(L.size("osb_oscr_curve_workspace_bytes", N), dtype=torch.uint8, device=dev)
    taus = torch.empty(N, dtype=torch.float32, device=dev)
    cc = torch.empty(N, dtype=torch.int64, device=dev); fc = torch.empty(N, dtype=torch.int64, device=dev)
    ccr = torch.empty(N, dtype=torch.float64, device=dev); fpr = torch.empty(N, dtype=torch.float64, device=dev)
    meta = torch.zeros(5, dtype=torch.int64, device=dev)
    it = [0]
    def nxt():
        it[0] += 1
        return sts[it[0] & 1]
    def rr(): L.call("osb_oscr_records", nxt().data_ptr(), C, gtt.data_ptr(), 0, N, C, -1, recu.data_ptr(), reck.data_ptr(), totals.data_ptr(), stp)
    def so(): L.call("osb_oscr_curve", recu.data_ptr(), totals[2:3].data_ptr(), reck.data_ptr(), totals[0:1].data_ptr(), N,
                     taus.data_ptr(), cc.data_ptr(), fc.data_ptr(), ccr.data_ptr(), fpr.data_ptr(), meta.data_ptr(), cws.data_ptr(), cws.numel(), stp)
    def full(): calculate_oscr_device(gtt, nxt(), -1)
    bytes_alg = N * C * 4 + N * 8
    print(f"--- oscr N={N} C={C}: algorithmic bytes {bytes_alg/1e6:.1f} MB (two alternating score buffers, no flush)")
    for name, fn in [("records (A)", rr), ("curve (B)", so), ("full", full)]:
        med, mn = timeit(fn, flush_l2=False)
        print(f"{name:16s} med {med:9.1f} us min {mn:9.1f} us | {bytes_alg/med/1e3:8.1f} GB/s-equivalent")
    from openset_imagenet_b200.metrics import confidence_sums
    med, mn = timeit(lambda: confidence_sums(nxt(), gtt, 1.0 / C, -1, None), flush_l2=False)
    print(f"{'confidence':16s} med {med:9.1f} us min {mn:9.1f} us")

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "head"):
        head_case(16384, 2048, 1000)
    if which in ("all", "head", "small"):
        head_case(2048, 2048, 152, "garbage")
        head_case(1024, 2048, 152, "garbage")
        head_case(256, 2048, 116)
        head_case(64, 2048, 116)
    if which in ("all", "oscr"):
        oscr_case(1_000_000, 116)
