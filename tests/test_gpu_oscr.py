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
