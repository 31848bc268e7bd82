"""GPU parity: OSCR (util.py:90-122) and CCR@FPR (plot_all.py:346-384).  Everything is bit-exact:
integer counts, the float64 ratios derived from them, threshold values and their order."""
import warnings

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

OSCR_CASES = ["synth_m1", "synth_m2", "ties", "bgcol", "kat_ties", "kat_nounk", "kat_oneknown", "kat_noknown"]


def _calc(gt, s, unk):
    import openset_imagenet_b200 as osb
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return osb.calculate_oscr(gt, s, unk_label=unk)


@pytest.mark.parametrize("name", OSCR_CASES)
def test_oscr_golden_bit_exact(golden, name):
    g = golden["oscr"]
    ccr, fpr = _calc(g[f"{name}_gt"], g[f"{name}_scores"], int(g[f"{name}_unk"]))
    assert isinstance(ccr, np.ndarray) and ccr.dtype == np.float64 and fpr.dtype == np.float64
    np.testing.assert_array_equal(ccr, g[f"{name}_ccr"])
    np.testing.assert_array_equal(fpr, g[f"{name}_fpr"])


@pytest.mark.parametrize("N,C", [(1, 2), (2, 2), (100, 3), (4096, 116), (4097, 116), (20000, 117), (30000, 152),
                                 (5000, 1000), (70000, 10)])
@pytest.mark.parametrize("unk", [-1, -2])
def test_oscr_random_bit_exact(N, C, unk):
    from openset_imagenet_b200.util import calculate_oscr_device
    gt, s = O.make_oscr_case(N, C, seed=N + C)
    taus, ccr_cnt, fpr_cnt, n_kn, n_unk = O.oscr_counts(gt, s, unk)
    r = calculate_oscr_device(gt, s, unk)
    meta = r.meta.tolist()
    n = max(meta[0] - 1, 0)
    assert n == len(taus) and meta[1] == n_kn and meta[3] == n_unk and meta[4] == 0
    np.testing.assert_array_equal(r.taus[:n].cpu().numpy(), taus)
    np.testing.assert_array_equal(r.ccr_cnt[:n].cpu().numpy(), ccr_cnt)
    np.testing.assert_array_equal(r.fpr_cnt[:n].cpu().numpy(), fpr_cnt)
    ccr, fpr = _calc(gt, s, unk)
    o_ccr, o_fpr = O.calculate_oscr(gt, s, unk)
    np.testing.assert_array_equal(ccr, o_ccr)
    np.testing.assert_array_equal(fpr, o_fpr)


def test_oscr_strided_wide_and_unaligned_inputs():
    """Row pitch > C, very wide rows and a misaligned base pointer take the non-bulk-copy path of stage A."""
    from openset_imagenet_b200.util import calculate_oscr_device
    gt, s = O.make_oscr_case(3000, 37, seed=21)
    o = O.oscr_counts(gt, s, -1)
    wide = torch.zeros(3000, 64, device="cuda")
    wide[:, :37] = torch.tensor(s)
    flat = torch.zeros(3000 * 37 + 1, device="cuda")
    flat[1:] = torch.tensor(s).reshape(-1)
    for s_in in (wide[:, :37], flat[1:].view(3000, 37)):
        r = calculate_oscr_device(torch.tensor(gt, device="cuda"), s_in, -1)
        n = int(r.meta[0]) - 1
        np.testing.assert_array_equal(r.taus[:n].cpu().numpy(), o[0])
        np.testing.assert_array_equal(r.ccr_cnt[:n].cpu().numpy(), o[1])
        np.testing.assert_array_equal(r.fpr_cnt[:n].cpu().numpy(), o[2])
    gt, s = O.make_oscr_case(300, 5000, seed=22)            # 20 KB rows: wider than a bulk-copy stage
    ccr, fpr = _calc(gt, s, -2)
    o_ccr, o_fpr = O.calculate_oscr(gt, s, -2)
    np.testing.assert_array_equal(ccr, o_ccr)
    np.testing.assert_array_equal(fpr, o_fpr)


def test_oscr_records_stage_multiset():
    """Stage A alone: the compacted records are exactly the oracle's (key, type) multiset."""
    from openset_imagenet_b200 import _lib as L
    N, C = 50_000, 116
    gt, s = O.make_oscr_case(N, C, seed=5)
    dev = torch.device("cuda")
    st, gtt = torch.tensor(s, device=dev), torch.tensor(gt, device=dev)
    recu = torch.empty(N, dtype=torch.int64, device=dev)
    reck = torch.empty(N, dtype=torch.int64, device=dev)
    tot = torch.empty(4, dtype=torch.int64, device=dev)
    L.call("osb_oscr_records", L.ptr(st), C, L.ptr(gtt), L.LABEL_I64, N, C, -1, L.ptr(recu), L.ptr(reck), L.ptr(tot), L.stream_ptr(dev))
    n_kn, n_ok, n_unk, n_bad = tot.tolist()
    kn = gt >= 0
    pred = s.argmax(1)
    assert (n_kn, n_unk, n_bad) == (int(kn.sum()), int((gt == -1).sum()), 0)
    assert n_ok == int((pred[kn] == gt[kn]).sum())

    def recs(vals, types):
        keys = vals.view(np.uint32).astype(np.uint64) | np.uint64(0x80000000)      # scores are >= 0
        return np.sort((keys << np.uint64(32)) | types.astype(np.uint64))

    want_k = recs(s[kn, gt[kn]], (pred[kn] == gt[kn]))
    want_u = recs(s[gt == -1].max(1), np.full(n_unk, 2))
    np.testing.assert_array_equal(np.sort(reck[:n_kn].cpu().numpy().view(np.uint64)), want_k)
    np.testing.assert_array_equal(np.sort(recu[:n_unk].cpu().numpy().view(np.uint64)), want_u)


@pytest.mark.parametrize("shards", [2, 3])
def test_oscr_shard_merge_is_bit_exact(shards):
    """The multi-GPU algorithm on one GPU: per-shard records, 'all-gathered' known records with remote ones
    demoted and skip padding, per-shard curve, summed counts == single-shard result (through the C ABI)."""
    from openset_imagenet_b200 import _lib as L
    from openset_imagenet_b200.util import calculate_oscr_device
    N, C = 30_000, 31
    n = N // shards
    gt, s = O.make_oscr_case(n * shards, C, seed=9)
    s[::5] = np.round(s[::5], 2)                                  # plenty of ties across shards
    dev = torch.device("cuda")
    stp = L.stream_ptr(dev)
    full = calculate_oscr_device(gt, s, -1)
    nu = int(full.meta[0])
    rec_u, rec_k, tots = [], [], []
    for r in range(shards):
        st = torch.tensor(s[r * n:(r + 1) * n], device=dev)
        gtt = torch.tensor(gt[r * n:(r + 1) * n], device=dev)
        ru = torch.empty(n, dtype=torch.int64, device=dev)
        rk = torch.full((n,), -1, dtype=torch.int64, device=dev)
        tt = torch.empty(4, dtype=torch.int64, device=dev)
        L.call("osb_oscr_records", L.ptr(st), C, L.ptr(gtt), L.LABEL_I64, n, C, -1, L.ptr(ru), L.ptr(rk), L.ptr(tt), stp)
        rec_u.append(ru); rec_k.append(rk); tots.append(tt)
    cap = shards * n + n
    ws = torch.zeros(L.size("osb_oscr_curve_workspace_bytes", cap), dtype=torch.uint8, device=dev)
    sum_c = torch.zeros(cap, dtype=torch.int64, device=dev)
    sum_f = torch.zeros(cap, dtype=torch.int64, device=dev)
    for r in range(shards):
        parts = [rk if q == r else torch.where((rk & 3) == 3, rk, rk & ~3) for q, rk in enumerate(rec_k)]
        rec_b = torch.cat(parts)
        n_b = torch.tensor([rec_b.numel()], dtype=torch.int64, device=dev)
        taus = torch.empty(cap, dtype=torch.float32, device=dev)
        cc = torch.zeros(cap, dtype=torch.int64, device=dev)
        fc = torch.zeros(cap, dtype=torch.int64, device=dev)
        meta = torch.zeros(5, dtype=torch.int64, device=dev)
        L.call("osb_oscr_curve", L.ptr(rec_u[r]), L.ptr(tots[r][2:3]), L.ptr(rec_b), L.ptr(n_b), cap,
               L.ptr(taus), L.ptr(cc), L.ptr(fc), None, None, L.ptr(meta), L.ptr(ws), ws.numel(), stp)
        assert int(meta[0]) == nu
        np.testing.assert_array_equal(taus[:nu].cpu().numpy(), full.taus[:nu].cpu().numpy())
        sum_c += cc; sum_f += fc
    np.testing.assert_array_equal(sum_c[:nu].cpu().numpy(), full.ccr_cnt[:nu].cpu().numpy())
    np.testing.assert_array_equal(sum_f[:nu].cpu().numpy(), full.fpr_cnt[:nu].cpu().numpy())


def test_oscr_from_head_statistics():
    """Compact-input variant: (p_target, p_max_all, pred) per row instead of the score matrix."""
    from openset_imagenet_b200.util import calculate_oscr_device, calculate_oscr_from_stats_device
    N, C = 40_000, 116
    gt, s = O.make_oscr_case(N, C, seed=12)
    dev = torch.device("cuda")
    st, gtt = torch.tensor(s, device=dev), torch.tensor(gt, device=dev)
    full = calculate_oscr_device(gtt, st, -2)
    pmax = st.max(1).values
    pred = torch.tensor(s.argmax(1), device=dev)                   # numpy's first arg-max
    ptgt = st.gather(1, gtt.clamp(min=0).view(-1, 1)).view(-1)
    r = calculate_oscr_from_stats_device(gtt, ptgt, pmax, pred, C, -2)
    assert r.meta.tolist() == full.meta.tolist()
    n = int(r.meta[0])
    for a, b in ((r.taus, full.taus), (r.ccr_cnt, full.ccr_cnt), (r.fpr_cnt, full.fpr_cnt), (r.ccr, full.ccr), (r.fpr, full.fpr)):
        np.testing.assert_array_equal(a[:n].cpu().numpy(), b[:n].cpu().numpy())


def test_oscr_workspace_is_reusable_and_graph_capturable():
    """Back-to-back calls on one workspace (counters must be left zero) and replay from a CUDA graph."""
    from openset_imagenet_b200.util import calculate_oscr_device
    gt, s = O.make_oscr_case(20_000, 116, seed=3)
    o = O.oscr_counts(gt, s, -1)
    dev = torch.device("cuda")
    st, gtt = torch.tensor(s, device=dev), torch.tensor(gt, device=dev)
    for _ in range(3):
        r = calculate_oscr_device(gtt, st, -1)
        n = int(r.meta[0]) - 1
        np.testing.assert_array_equal(r.ccr_cnt[:n].cpu().numpy(), o[1])
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        r = calculate_oscr_device(gtt, st, -1)                # allocate the workspace of this stream before capture
        side.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            r = calculate_oscr_device(gtt, st, -1)
        for _ in range(2):
            r.ccr_cnt.zero_()
            graph.replay()
            side.synchronize()
            n = int(r.meta[0]) - 1
            np.testing.assert_array_equal(r.ccr_cnt[:n].cpu().numpy(), o[1])
            np.testing.assert_array_equal(r.fpr_cnt[:n].cpu().numpy(), o[2])


def test_oscr_heavy_ties_and_special_values():
    rng = np.random.default_rng(0)
    N, C = 9000, 20
    s = (rng.integers(0, 4, size=(N, C)) / 4.0).astype(np.float32)     # only 4 distinct values
    s[::7, 3] = -0.0
    s[::11, 5] = 0.0
    s[::13] *= -1.0                                                     # negative scores sort below positive ones
    gt = O.make_labels(N, C, 0.4, seed=4)
    for unk in (-1,):
        ccr, fpr = _calc(gt, s, unk)
        o_ccr, o_fpr = O.calculate_oscr(gt, s, unk)
        np.testing.assert_array_equal(ccr, o_ccr)
        np.testing.assert_array_equal(fpr, o_fpr)


def test_oscr_inputs_float_gt_tensor_inputs_and_errors():
    import openset_imagenet_b200 as osb
    gt, s = O.make_oscr_case(3000, 31, seed=8)
    o = O.calculate_oscr(gt, s, -2)
    for g_in, s_in in [(gt.astype(np.float32), s), (torch.tensor(gt), torch.tensor(s)),
                       (torch.tensor(gt).cuda(), torch.tensor(s).cuda()), (gt.astype(np.int32), s)]:
        ccr, fpr = osb.calculate_oscr(g_in, s_in, unk_label=-2)
        np.testing.assert_array_equal(ccr, o[0])
        np.testing.assert_array_equal(fpr, o[1])
    bad = gt.copy()
    bad[5] = 31
    with pytest.raises(IndexError):
        osb.calculate_oscr(bad, s)
    with pytest.warns(RuntimeWarning):
        osb.calculate_oscr(np.array([0, 1, 1, 0]), s[:4, :2])           # no unknowns -> NaN fpr


def test_oscr_full_size_properties():
    """cfg 4 size (1M x 116): size-independent properties + bit-exactness against the O(N log N) oracle."""
    from openset_imagenet_b200.util import calculate_oscr_device
    N, C = 1_000_000, 116
    gt, s = O.make_oscr_case(N, C, seed=0)
    r = calculate_oscr_device(gt, s, -1)
    meta = r.meta.tolist()
    n = meta[0] - 1
    taus = r.taus[:n].cpu().numpy()
    cc, fc = r.ccr_cnt[:n].cpu().numpy(), r.fpr_cnt[:n].cpu().numpy()
    assert np.all(np.diff(taus) > 0)                                    # strictly ascending unique thresholds
    assert np.all(np.diff(cc) <= 0) and np.all(np.diff(fc) <= 0)        # monotone counts
    assert cc[0] <= meta[2] and fc[0] <= meta[3] and cc.min() >= 0 and fc.min() >= 0
    assert meta[1] == int((gt >= 0).sum()) and meta[3] == int((gt == -1).sum())
    o_taus, o_cc, o_fc, _, _ = O.oscr_counts(gt, s, -1)
    np.testing.assert_array_equal(taus, o_taus)
    np.testing.assert_array_equal(cc, o_cc)
    np.testing.assert_array_equal(fc, o_fc)


@pytest.mark.parametrize("scale,shift", [(1.0, 0.0), (37.5, -11.0), (1e-30, 0.0), (0.0, 0.75)])
def test_oscr_radix_passes_cover_the_whole_key(scale, shift):
    """The radix passes must cover all 32 bits of the order-preserving key: scores as probabilities, raw logits of both
    signs and large magnitude, denormal-sized values and a constant matrix (every key equal) all give the oracle's curve
    bit for bit."""
    from openset_imagenet_b200.util import calculate_oscr_device
    rng = np.random.default_rng(11)
    N, C = 60000, 24
    s = (rng.standard_normal((N, C)) * scale + shift).astype(np.float32)
    gt = O.make_labels(N, C, 0.4, seed=12)
    gt[::9] = np.where(gt[::9] < 0, -2, gt[::9])
    o_taus, o_cc, o_fc, n_kn, n_unk = O.oscr_counts(gt, s, -1)
    r = calculate_oscr_device(gt, s, -1)
    n = max(int(r.meta[0]) - 1, 0)
    assert n == len(o_taus) and int(r.meta[1]) == n_kn and int(r.meta[3]) == n_unk
    np.testing.assert_array_equal(r.taus[:n].cpu().numpy(), o_taus)
    np.testing.assert_array_equal(r.ccr_cnt[:n].cpu().numpy(), o_cc)
    np.testing.assert_array_equal(r.fpr_cnt[:n].cpu().numpy(), o_fc)


def test_oscr_more_records_than_one_tile_per_cta():
    """More records than 148 CTAs x 8192 hold at once: every CTA sorts several tiles per pass (the records no longer stay
    in registers between the count and the scatter phase).  1.6 M rows x 4 columns, all of them known or unk."""
    from openset_imagenet_b200.util import calculate_oscr_device
    rng = np.random.default_rng(5)
    N, C = 1_600_000, 4
    s = rng.random((N, C), dtype=np.float32)
    s[::3] = np.round(s[::3], 3)                                        # plenty of ties as well
    gt = rng.integers(-1, C, size=N).astype(np.int64)
    o_taus, o_cc, o_fc, n_kn, n_unk = O.oscr_counts(gt, s, -1)
    r = calculate_oscr_device(gt, s, -1)
    n = max(int(r.meta[0]) - 1, 0)
    assert n == len(o_taus) and int(r.meta[1]) == n_kn and int(r.meta[3]) == n_unk and n_kn + n_unk == N
    np.testing.assert_array_equal(r.taus[:n].cpu().numpy(), o_taus)
    np.testing.assert_array_equal(r.ccr_cnt[:n].cpu().numpy(), o_cc)
    np.testing.assert_array_equal(r.fpr_cnt[:n].cpu().numpy(), o_fc)


@pytest.mark.parametrize("name", ["synth_m1", "synth_m2"])
def test_ccr_at_fpr_golden(golden, name):
    import openset_imagenet_b200 as osb
    g = golden["oscr"]
    ref = g[f"{name}_ccr_at_fpr"]
    got = osb.ccr_at_fpr(g[f"{name}_ccr"], g[f"{name}_fpr"], queries=list(ref[:, 0]))
    for (idx, f, c), row in zip(got, ref):
        assert idx == int(row[1]) and f == row[2]
        assert (c is None and np.isnan(row[3])) or c == row[3]


def test_ccr_at_fpr_ties_nan_and_empty():
    import openset_imagenet_b200 as osb
    fpr = np.array([1.0, 0.5, 0.5, 0.25, 0.25, 0.0])
    ccr = np.array([0.9, 0.8, 0.7, 0.6, 0.5, 0.4])
    got = osb.ccr_at_fpr(ccr, fpr, queries=[0.5, 0.26, 1e-3])
    want = O.ccr_at_fpr(ccr, fpr, queries=[0.5, 0.26, 1e-3])
    assert got == want and got[0][0] == 1                                # first index among equals
    nan_fpr = np.array([np.nan, np.nan])
    assert osb.ccr_at_fpr(np.array([0.1, 0.2]), nan_fpr, [0.1])[0][0] == 0
    with pytest.raises(ValueError):
        osb.ccr_at_fpr(np.array([]), np.array([]), [0.1])


def test_oscr_plan_with_and_without_count_outputs():
    """OscrPlan(counts=False) writes only the reference's outputs (ccr, fpr); both modes agree bit for bit."""
    from openset_imagenet_b200.functional import OscrPlan
    gt, s = O.make_oscr_case(60_000, 116, seed=14)
    o_ccr, o_fpr = O.calculate_oscr(gt, s, -1)
    st, gtt = torch.tensor(s, device="cuda"), torch.tensor(gt, device="cuda")
    full = OscrPlan(60_000, 116, -1).run(st, gtt)
    lean = OscrPlan(60_000, 116, -1, counts=False).run(st, gtt)
    assert lean.taus is None and lean.ccr_cnt is None and lean.fpr_cnt is None
    assert lean.meta.tolist() == full.meta.tolist()
    n = int(lean.meta[0]) - 1
    for r in (full, lean):
        np.testing.assert_array_equal(r.ccr[:n].cpu().numpy(), o_ccr)
        np.testing.assert_array_equal(r.fpr[:n].cpu().numpy(), o_fpr)


# ---- util.get_histogram (util.py:202-228) ---------------------------------------------------------
@pytest.mark.parametrize("name", ["score_m1", "score_m2_b37", "score_dropbg", "score_log"])
def test_get_histogram_scores_bit_exact_vs_reference_golden(golden, name):
    from oracle.make_golden import HIST_CASES
    from openset_imagenet_b200.util import get_histogram
    g = golden["hist"]
    kw = HIST_CASES[name]
    gt = g["gt"]
    if kw.get("drop_bg"):
        gt = np.where(gt >= 23, 0, gt).astype(np.float32)
    kh, ke, uh, ue = get_histogram({"scores": g["scores"], "gt": gt, "features": g["features"]}, **kw)
    np.testing.assert_array_equal(kh, g[f"{name}_kn_hist"])
    np.testing.assert_array_equal(uh, g[f"{name}_unk_hist"])
    np.testing.assert_array_equal(ke, g[f"{name}_kn_edges"])
    np.testing.assert_array_equal(ue, g[f"{name}_unk_edges"])
    assert kh.dtype == np.int64 and uh.dtype == np.int64


@pytest.mark.parametrize("name", ["norm_m1", "norm_log"])
def test_get_histogram_norms_vs_reference_golden(golden, name):
    """Feature norms are accumulated in fp64 on the GPU (the reference's float32 pairwise sum is not reproducible), so a
    norm within one float32 ulp of an edge may change bins: totals are exact, at most a few counts move to a neighbour."""
    from oracle.make_golden import HIST_CASES
    from openset_imagenet_b200.util import get_histogram
    g = golden["hist"]
    kw = HIST_CASES[name]
    kh, ke, uh, ue = get_histogram({"scores": g["scores"], "gt": g["gt"], "features": g["features"]}, **kw)
    for got, ref in ((kh, g[f"{name}_kn_hist"]), (uh, g[f"{name}_unk_hist"])):
        assert abs(int(got.sum()) - int(ref.sum())) <= 2
        assert np.abs(np.cumsum(got) - np.cumsum(ref)).max() <= 2
    np.testing.assert_allclose(ke, g[f"{name}_kn_edges"], rtol=1e-6)
    np.testing.assert_allclose(ue, g[f"{name}_unk_edges"], rtol=1e-6)


def test_get_histogram_large_and_empty_groups():
    """1M x 116 scores against the oracle (numpy), and a set without unknowns (np.histogram's (0, 1) default range)."""
    from openset_imagenet_b200.util import get_histogram
    gt, s = O.make_oscr_case(1_000_000, 116, seed=0)
    arr = {"scores": s, "gt": gt, "features": np.zeros((len(gt), 1), dtype=np.float32)}
    for unk in (-1, -2):
        got = get_histogram(arr, unk_label=unk, bins=100)
        ref = O.get_histogram(arr, unk_label=unk, bins=100)
        for a, b in zip(got, ref):
            np.testing.assert_array_equal(a, b)
    small = {"scores": s[:100], "gt": np.abs(gt[:100]), "features": np.zeros((100, 1), dtype=np.float32)}
    got = get_histogram(small, bins=10)
    ref = O.get_histogram(small, bins=10)
    for a, b in zip(got, ref):
        np.testing.assert_array_equal(a, b)
    assert got[2].sum() == 0


def test_oscr_record_exchange_two_emulated_ranks_give_the_single_gpu_curve():
    """osb_oscr_exchange (the peer-memory all-gather of the sharded path), two ranks emulated on ONE GPU: rank 1 runs first
    with rank 0's count words and done stamp pre-published, then rank 0 finds rank 1's in its slot area.  Both gathered
    arrays must hold [rank 0 records | rank 1 records], and the curve sorted from them is bit-identical to the single-GPU
    one on the concatenated rows."""
    from openset_imagenet_b200 import _lib as L
    from openset_imagenet_b200.functional import OscrPlan
    dev = torch.device("cuda")
    n, C = 6000, 37
    gt, s = O.make_oscr_case(2 * n, C, seed=23)
    s[::5] = np.round(s[::5], 2)                                     # ties within and across the shards
    st = L.stream_ptr(dev)
    rec, tot = [], []
    for r in range(2):
        ru = torch.empty(n, dtype=torch.int64, device=dev); rk = torch.empty(n, dtype=torch.int64, device=dev)
        t = torch.zeros(4, dtype=torch.int64, device=dev)
        sc = torch.tensor(s[r * n:(r + 1) * n], device=dev); g = torch.tensor(gt[r * n:(r + 1) * n], device=dev)
        L.call("osb_oscr_records", L.ptr(sc), C, L.ptr(g), L.LABEL_I64, n, C, -1, L.ptr(ru), L.ptr(rk), L.ptr(t), st)
        rec.append((ru, rk)); tot.append(t)
    torch.cuda.synchronize()
    cap = 2 * n
    all_u = [torch.full((cap,), -7, dtype=torch.int64, device=dev) for _ in range(2)]
    all_k = [torch.full((cap,), -7, dtype=torch.int64, device=dev) for _ in range(2)]
    nsl = L.size("osb_oscr_exchange_slot_bytes", 2) // 4 + 4
    slots = [torch.zeros(nsl, dtype=torch.int32, device=dev) for _ in range(2)]
    P = L.C.c_void_p * 2
    pu, pk, ps = P(*[t.data_ptr() for t in all_u]), P(*[t.data_ptr() for t in all_k]), P(*[t.data_ptr() for t in slots])
    # pre-publish rank 0 in rank 1's slot area: count words (value, epoch 1) at [rank 0][j], done stamp at [16 * 4 * 2 + 0]
    t0 = tot[0].tolist()
    for j in range(4):
        slots[1][(0 * 4 + j) * 2] = t0[j]; slots[1][(0 * 4 + j) * 2 + 1] = 1
    slots[1][16 * 4 * 2 + 0] = 1
    tg = [torch.zeros(4, dtype=torch.int64, device=dev) for _ in range(2)]
    for r in (1, 0):
        L.call("osb_oscr_exchange", L.ptr(rec[r][0]), L.ptr(rec[r][1]), L.ptr(tot[r]), pu, pk, ps, r, 2, cap, L.ptr(tg[r]), 8, st)
        torch.cuda.synchronize()
    want_tot = (tot[0] + tot[1]).tolist()
    for r in range(2):
        assert tg[r].tolist() == want_tot
        nk0, nk1, nu0, nu1 = int(tot[0][0]), int(tot[1][0]), int(tot[0][2]), int(tot[1][2])
        assert torch.equal(all_k[r][:nk0].sort().values, rec[0][1][:nk0].sort().values)
        assert torch.equal(all_k[r][nk0:nk0 + nk1].sort().values, rec[1][1][:nk1].sort().values)
        assert torch.equal(all_u[r][:nu0].sort().values, rec[0][0][:nu0].sort().values)
        assert torch.equal(all_u[r][nu0:nu0 + nu1].sort().values, rec[1][0][:nu1].sort().values)
        assert bool((all_k[r][nk0 + nk1:] == -7).all()) and bool((all_u[r][nu0 + nu1:] == -7).all())
    ref = OscrPlan(2 * n, C, -1).run(torch.tensor(s, device=dev), torch.tensor(gt, device=dev))
    nu = int(ref.meta[0])
    for r in range(2):
        taus = torch.empty(cap, dtype=torch.float32, device=dev)
        cc = torch.empty(cap, dtype=torch.int64, device=dev); fc = torch.empty(cap, dtype=torch.int64, device=dev)
        ccr = torch.empty(cap, dtype=torch.float64, device=dev); fpr = torch.empty(cap, dtype=torch.float64, device=dev)
        meta = torch.zeros(5, dtype=torch.int64, device=dev)
        ws = torch.zeros(L.size("osb_oscr_curve_workspace_bytes", cap) + 256, dtype=torch.uint8, device=dev)
        L.call("osb_oscr_curve", L.ptr(all_u[r]), L.ptr(tg[r][2:3]), L.ptr(all_k[r]), L.ptr(tg[r][0:1]), cap, L.ptr(taus), L.ptr(cc), L.ptr(fc),
               L.ptr(ccr), L.ptr(fpr), L.ptr(meta), L.ptr(ws), ws.numel(), st)
        torch.cuda.synchronize()
        assert meta.tolist()[:4] == ref.meta.tolist()[:4]
        assert torch.equal(taus[:nu], ref.taus[:nu]) and torch.equal(cc[:nu], ref.ccr_cnt[:nu]) and torch.equal(fc[:nu], ref.fpr_cnt[:nu])
        assert torch.equal(ccr[:nu], ref.ccr[:nu]) and torch.equal(fpr[:nu], ref.fpr[:nu])
    # a rank whose peer never shows up (epoch 2 words are missing) gives up and poisons n_bad
    L.load().osb_debug_set_exchange_timeout(20_000_000)
    try:
        L.call("osb_oscr_exchange", L.ptr(rec[0][0]), L.ptr(rec[0][1]), L.ptr(tot[0]), pu, pk, ps, 0, 2, cap, L.ptr(tg[0]), 8, st)
        torch.cuda.synchronize()
    finally:
        L.load().osb_debug_set_exchange_timeout(0)
    assert int(tg[0][3]) >= (1 << 40) and int(tg[0][0]) == 0


@pytest.mark.parametrize("N,C,loss", [(20000, 116, "entropic"), (20000, 117, "garbage"), (5000, 10, "garbage"), (3001, 1000, "entropic"),
                                        (70000, 152, "garbage")])
def test_one_pass_table_is_oscr_plus_confidence(N, C, loss):
    """script/plot_all.py:363-372 runs util.calculate_oscr and metrics.confidence over the same scores; the one-pass entry
    must give the same curve bit for bit and the confidence of the oracle (counts exact, float64 sums to 1e-9)."""
    import openset_imagenet_b200 as osb
    from openset_imagenet_b200.util import calculate_oscr_and_confidence_device, calculate_oscr_device, conf_and_ccr
    n_known = C - 1 if loss == "garbage" else C
    gt, s = O.make_oscr_case(N, C, seed=N + C)
    gt = np.where(gt >= n_known, n_known - 1, gt)                    # the garbage column is never a label of a test sample
    gt[0] = n_known - 1                                              # so that numpy.max(gt) + 1 is the class count
    offset = 0. if loss == "garbage" else 1. / n_known
    lvc = -1 if loss == "garbage" else None
    r, sums = calculate_oscr_and_confidence_device(gt, s, -2, offset, lvc)
    ref = calculate_oscr_device(gt, s, -2)
    nu = int(ref.meta[0])
    assert r.meta.tolist() == ref.meta.tolist()
    for a, b in ((r.taus, ref.taus), (r.ccr_cnt, ref.ccr_cnt), (r.fpr_cnt, ref.fpr_cnt), (r.ccr, ref.ccr), (r.fpr, ref.fpr)):
        assert torch.equal(a[:nu], b[:nu])
    kn, n_kn, ng, n_ng, kn_sum, ng_sum = O.confidence(s, gt, offset, -2, lvc)
    assert (int(r.meta[1]), int(r.meta[3])) == (n_kn, n_ng)
    got = sums.tolist()
    assert got[0] == pytest.approx(kn_sum, rel=1e-12, abs=1e-9) and got[1] == pytest.approx(ng_sum, rel=1e-12, abs=1e-9)
    # the table row against the two separate drop-in calls
    conf, ccr_at = conf_and_ccr(gt, s, loss=loss, unk_label=-2)
    want_conf = osb.confidence(torch.tensor(s, device="cuda"), torch.tensor(gt, device="cuda"), offset, -2, lvc)
    assert conf[1] == want_conf[1] and conf[3] == want_conf[3]
    assert conf[0] == pytest.approx(want_conf[0], rel=1e-12) and conf[2] == pytest.approx(want_conf[2], rel=1e-12)
    ccr, fpr = _calc(gt, s, -2)
    want_at = [v for _, _, v in osb.ccr_at_fpr(ccr, fpr)]
    assert ccr_at == want_at


def test_one_pass_table_strided_input_takes_the_fallback_row_reduce():
    from openset_imagenet_b200.util import calculate_oscr_and_confidence_device
    gt, s = O.make_oscr_case(4000, 37, seed=3)
    wide = torch.zeros(4000, 41, device="cuda")
    wide[:, :37] = torch.tensor(s, device="cuda")
    r, sums = calculate_oscr_and_confidence_device(gt, wide[:, :37], -1, 0., -1)     # row pitch 41: no bulk copies
    _, n_kn, _, n_ng, kn_sum, ng_sum = O.confidence(s, gt, 0., -1, -1)
    assert (int(r.meta[1]), int(r.meta[3])) == (n_kn, n_ng)
    got = sums.tolist()
    assert got[0] == pytest.approx(kn_sum, rel=1e-12, abs=1e-9) and got[1] == pytest.approx(ng_sum, rel=1e-12, abs=1e-9)
