"""Pin the oracle: every golden vector made from the reference's own code (oracle/make_golden.py),
the known-answer vectors of SURVEY.md §9.6, and — when /root/reference is mounted — the live
reference on fresh random cases."""
import warnings

import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref_loader

EOS_CASES = ["cfg1", "w05", "allneg", "allknown", "single", "p3"]


@pytest.mark.parametrize("name", EOS_CASES)
def test_eos_matches_reference_fp32(golden, name):
    g = golden["losses"]
    loss, dz, _, _ = O.loss_from_logits(g[f"eos_{name}_logits"], g[f"eos_{name}_y"], O.ENTROPIC,
                                        unk_weight=float(g[f"eos_{name}_w"]))
    # the reference ran in float32: agreement is limited by ITS rounding, not the oracle's
    assert loss == pytest.approx(float(g[f"eos_{name}_loss"]), rel=2e-6, abs=1e-6)
    np.testing.assert_allclose(dz, g[f"eos_{name}_grad"], rtol=0, atol=3e-8)


def test_eos_matches_reference_fp64(golden):
    g = golden["losses"]
    loss, dz, _, _ = O.loss_from_logits(g["eos_f64_logits"], g["eos_f64_y"], O.ENTROPIC, unk_weight=float(g["eos_f64_w"]))
    assert loss == pytest.approx(float(g["eos_f64_loss"]), rel=1e-13)
    np.testing.assert_allclose(dz, g["eos_f64_grad"], rtol=0, atol=1e-15)


def test_eos_minus2_equals_minus1(golden):
    g = golden["losses"]
    z, y = g["eos_w05_logits"], g["eos_w05_y"].copy()
    a = O.loss_from_logits(z, y, O.ENTROPIC, 0.5)[0]
    y[y < 0] = -1
    assert a == O.loss_from_logits(z, y, O.ENTROPIC, 0.5)[0]


def test_eos_allneg_lower_bound(golden):
    g = golden["losses"]
    loss = O.loss_from_logits(g["eos_allneg_logits"], g["eos_allneg_y"], O.ENTROPIC)[0]
    assert np.isfinite(loss) and loss >= O.log_c(116)


def test_eos_label_out_of_range():
    with pytest.raises(IndexError):
        O.loss_from_logits(np.zeros((2, 3)), np.array([0, 3]), O.ENTROPIC)


def test_softmax_variant(golden):
    g = golden["losses"]
    loss, dz, _, den = O.loss_from_logits(g["sm_logits"], g["sm_y"], O.SOFTMAX)
    assert den == int((g["sm_y"] != -1).sum())
    assert loss == pytest.approx(float(g["sm_loss"]), rel=2e-6)
    np.testing.assert_allclose(dz, g["sm_grad"], rtol=0, atol=3e-8)
    assert np.isnan(g["sm_allignored_loss"])
    assert np.isnan(O.loss_from_logits(g["sm_logits"][:8], np.full(8, -1), O.SOFTMAX)[0])


def test_garbage_variant(golden):
    g = golden["losses"]
    cw = O.class_weights_from_labels(g["bg_train_labels"], 152)
    np.testing.assert_array_equal(cw, g["bg_cw"])
    loss, dz, _, _ = O.loss_from_logits(g["bg_logits"], g["bg_y"], O.GARBAGE, class_weights=cw)
    assert loss == pytest.approx(float(g["bg_loss"]), rel=2e-6)
    np.testing.assert_allclose(dz, g["bg_grad"], rtol=0, atol=3e-8)


def test_head_fwd_bwd(golden):
    g = golden["losses"]
    r = O.head_fwd_bwd(g["head_x"], g["head_W"], g["head_b"], g["head_y"], O.ENTROPIC)
    assert r["loss"] == pytest.approx(float(g["head_loss"]), rel=1e-12)
    np.testing.assert_allclose(r["dx"], g["head_dx"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(r["dw"], g["head_dw"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(r["db"], g["head_db"], rtol=0, atol=1e-14)


@pytest.mark.parametrize("name", ["eos", "eos_mixed", "table", "bg", "noneg", "noknown"])
def test_confidence(golden, name):
    g = golden["confidence"]
    off, uc, lvc = g[f"{name}_args"]
    lvc = None if np.isnan(lvc) else int(lvc)
    r = O.confidence(g[f"{name}_scores"], g[f"{name}_y"], float(off), int(uc), lvc)
    ref = g[f"{name}_out"]
    assert (r[1], r[3]) == (int(ref[1]), int(ref[3]))                 # counts bit-exact
    # the reference sums in float32; the oracle in float64: agree to float32 summation error
    assert r[0] == pytest.approx(ref[0], rel=2e-6, abs=1e-7)
    assert r[2] == pytest.approx(ref[2], rel=2e-6, abs=1e-7)


def test_confidence_kat():
    s = np.array([[.5, .5], [.2, .8], [.6, .4], [.7, .3]], dtype=np.float32)
    r = O.confidence(s, np.array([0, 1, -2, -2]), offset=0.5, unknown_class=-1)
    assert r[:4] == (pytest.approx(0.65), 2, 0.0, 0)                   # -2 is in neither group
    r = O.confidence(s, np.array([0, 1, -1, -1]), offset=0.5, unknown_class=-1)
    assert r[3] == 2 and r[2] == pytest.approx((1.5 - .6 + 1.5 - .7) / 2, rel=1e-6)
    assert O.resolve_valid_cols(5, None) == 5 and O.resolve_valid_cols(5, -1) == 4 and O.resolve_valid_cols(5, 3) == 3


OSCR_CASES = ["synth_m1", "synth_m2", "ties", "bgcol", "kat_ties", "kat_nounk", "kat_oneknown", "kat_noknown"]


@pytest.mark.parametrize("name", OSCR_CASES)
@pytest.mark.parametrize("quadratic", [False, True])
def test_oscr_bit_exact(golden, name, quadratic):
    g = golden["oscr"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ccr, fpr = O.calculate_oscr(g[f"{name}_gt"], g[f"{name}_scores"], int(g[f"{name}_unk"]), quadratic=quadratic)
    assert ccr.dtype == np.float64 and fpr.dtype == np.float64
    np.testing.assert_array_equal(ccr, g[f"{name}_ccr"])              # bit-exact incl. NaN positions
    np.testing.assert_array_equal(fpr, g[f"{name}_fpr"])


def test_oscr_kat_values(golden):
    g = golden["oscr"]
    np.testing.assert_array_equal(g["kat_ties_ccr"], [0.75, 0.5])
    np.testing.assert_array_equal(g["kat_ties_fpr"], [1.0, 1.0])
    np.testing.assert_array_equal(g["kat_nounk_ccr"], [0.75, 0.5, 0.25])
    assert np.all(np.isnan(g["kat_nounk_fpr"])) and len(g["kat_nounk_fpr"]) == 3
    assert len(g["kat_oneknown_ccr"]) == 0 and len(g["kat_noknown_fpr"]) == 0


def test_oscr_shards_are_additive():
    gt, s = O.make_oscr_case(4000, 16, seed=21)
    taus, c, f, nk, nu = O.oscr_counts(gt, s, -1)
    c2 = np.zeros_like(c)
    f2 = np.zeros_like(f)
    for sh in range(8):
        sl = slice(sh * 500, (sh + 1) * 500)
        g_, s_ = gt[sl], s[sl]
        kn = g_ >= 0
        hit = s_.argmax(1)[kn] == g_[kn]
        cts = np.sort(s_[kn, g_[kn]][hit])
        ums = np.sort(s_.max(1)[g_ == -1])
        c2 += len(cts) - np.searchsorted(cts, taus, side="right")
        f2 += len(ums) - np.searchsorted(ums, taus, side="right")
    np.testing.assert_array_equal(c, c2)
    np.testing.assert_array_equal(f, f2)


@pytest.mark.parametrize("name", ["synth_m1", "synth_m2"])
def test_ccr_at_fpr(golden, name):
    g = golden["oscr"]
    ref = g[f"{name}_ccr_at_fpr"]
    got = O.ccr_at_fpr(g[f"{name}_ccr"], g[f"{name}_fpr"], queries=list(ref[:, 0]))
    for (idx, f, c), row in zip(got, ref):
        assert idx == int(row[1]) and f == row[2]
        assert (c is None and np.isnan(row[3])) or c == row[3]


@pytest.mark.parametrize("name", ["score_m1", "score_m2_b37", "score_dropbg", "score_log", "norm_m1", "norm_log"])
def test_get_histogram(golden, name):
    """oracle.get_histogram against the reference's util.get_histogram (util.py:202-228) on the committed vectors."""
    from oracle.make_golden import HIST_CASES
    g = golden["hist"]
    kw = HIST_CASES[name]
    gt = g["gt"]
    if kw.get("drop_bg"):
        gt = np.where(gt >= 23, 0, gt).astype(np.float32)
    kh, ke, uh, ue = O.get_histogram({"scores": g["scores"], "gt": gt, "features": g["features"]}, **kw)
    np.testing.assert_array_equal(kh, g[f"{name}_kn_hist"])
    np.testing.assert_array_equal(uh, g[f"{name}_unk_hist"])
    np.testing.assert_array_equal(ke, g[f"{name}_kn_edges"])
    np.testing.assert_array_equal(ue, g[f"{name}_unk_edges"])
    assert kh.sum() > 0 and uh.sum() > 0


def test_bf16_round():
    import torch
    a = np.random.default_rng(0).standard_normal(10000).astype(np.float32) * 37
    np.testing.assert_array_equal(O.bf16_round(a), torch.tensor(a).bfloat16().float().numpy())


# ---- live reference (build container only) -----------------------------------------------------
needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="reference not mounted (GPU box)")


@needs_ref
@pytest.mark.parametrize("seed", range(4))
def test_live_reference_eos(seed):
    import torch
    L = ref_loader.load("losses")
    rng = np.random.default_rng(seed)
    B, C = int(rng.integers(1, 70)), int(rng.integers(2, 160))
    w = float(rng.choice([0.25, 1.0, 2.0]))
    z = (3 * rng.standard_normal((B, C))).astype(np.float32)
    y = O.make_labels(B, C, float(rng.random()), seed=seed, neg_label=int(rng.choice([-1, -2])))
    zt = torch.tensor(z, requires_grad=True)
    loss = L.EntropicOpensetLoss(C, w)(zt, torch.tensor(y))
    loss.backward()
    o_loss, o_dz, _, _ = O.loss_from_logits(z, y, O.ENTROPIC, w)
    assert o_loss == pytest.approx(loss.item(), rel=3e-6, abs=1e-6)
    np.testing.assert_allclose(o_dz, zt.grad.numpy(), rtol=0, atol=5e-8)


@needs_ref
@pytest.mark.parametrize("seed", range(3))
def test_live_reference_oscr_and_confidence(seed):
    import torch
    U, M = ref_loader.load("util"), ref_loader.load("metrics")
    gt, s = O.make_oscr_case(700 + 300 * seed, 12 + seed, seed=40 + seed)
    for unk in (-1, -2):
        ccr, fpr = U.calculate_oscr(gt, s, unk_label=unk)
        occr, ofpr = O.calculate_oscr(gt, s, unk)
        np.testing.assert_array_equal(occr, ccr)
        np.testing.assert_array_equal(ofpr, fpr)
    r = M.confidence(torch.tensor(s), torch.tensor(gt), offset=0.1, unknown_class=-1)
    o = O.confidence(s, gt, 0.1, -1)
    assert (o[1], o[3]) == (r[1], r[3])
    assert o[0] == pytest.approx(r[0], rel=2e-6) and o[2] == pytest.approx(r[2], rel=2e-6)


def test_synthetic_generators_of_the_product_match_the_oracle_copies():
    """bench.py draws its inputs from openset_imagenet_b200.synth (the product may not import the oracle);
    both generators must stay identical so that parity checks see the same data."""
    from openset_imagenet_b200 import synth
    np.testing.assert_array_equal(synth.make_labels(500, 116, 1 / 3, seed=3), O.make_labels(500, 116, 1 / 3, seed=3))
    g1, s1 = synth.make_oscr_case(2000, 31, seed=4)
    g2, s2 = O.make_oscr_case(2000, 31, seed=4)
    np.testing.assert_array_equal(g1, g2)
    np.testing.assert_array_equal(s1, s2)
    y = synth.make_labels(4000, 20, 0.3, seed=5, neg_label=19)
    np.testing.assert_array_equal(synth.class_weights_from_labels(y, 20), O.class_weights_from_labels(y, 20))
