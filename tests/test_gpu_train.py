"""GPU tests of the loops around the path: train / validate / get_arrays with the reference's
signatures, streaming confidence from the fused head's per-row statistics, and the npz array format feeding OSCR."""
import types

import numpy as np
import pytest
import torch
from torch import nn
from torch.utils.data import DataLoader, TensorDataset

from oracle import oracle as O

pytestmark = pytest.mark.gpu

K, C = 128, 10


class TinyModel(nn.Module):
    """Same interface as model.ResNet50 (forward -> (logits, features), forward_loss, .logits) on a toy backbone."""

    def __init__(self, variant, **kw):
        super().__init__()
        from openset_imagenet_b200.model import FusedOpenSetHead
        self.resnet_base = nn.Sequential(nn.Linear(32, K), nn.ReLU())
        self.logits = FusedOpenSetHead(K, C + (1 if variant == "garbage" else 0), bias=False, variant=variant, **kw)

    def forward(self, image):
        features = self.resnet_base(image)
        return self.logits(features), features

    def forward_loss(self, image, target):
        features = self.resnet_base(image)
        return self.logits.loss(features, target), features


def _data(n, variant, seed=0):
    rng = np.random.default_rng(seed)
    x = torch.tensor(rng.standard_normal((n, 32)), dtype=torch.float32)
    neg = C if variant == "garbage" else -1
    y = torch.tensor(O.make_labels(n, C, 0.3, seed=seed + 1, neg_label=neg))
    return TensorDataset(x, y)


def _cfg(loss_type, batch_size):
    return types.SimpleNamespace(loss=types.SimpleNamespace(type=loss_type, w=1.0), batch_size=batch_size, parallel=True)


@pytest.mark.parametrize("variant", ["entropic", "garbage"])
def test_validate_streams_confidence_without_a_score_matrix(variant):
    import openset_imagenet_b200 as osb
    torch.manual_seed(0)
    n_classes = C + (1 if variant == "garbage" else 0)
    kw = {}
    if variant == "garbage":
        kw = dict(class_weights=torch.ones(n_classes), last_valid_class=-1)
    model = TinyModel(variant, **kw).cuda()
    ds = _data(1000, variant)
    loader = DataLoader(ds, batch_size=96)                        # the last batch is ragged (1000 = 10*96 + 40)
    trackers = osb.train.make_trackers()
    osb.train.validate(model, loader, None, n_classes, trackers, _cfg(variant, 96))
    # reference-style evaluation of the same model: softmax of its logits, materialised, through the oracle
    with torch.no_grad():
        x, y = ds.tensors
        feats = model.resnet_base(x.cuda())
        fb = O.bf16_round(feats.float().cpu().numpy())
        wb = O.bf16_round(model.logits.weight.detach().float().cpu().numpy())
        z = fb @ wb.T
    s = np.exp(z - z.max(1, keepdims=True)); s = (s / s.sum(1, keepdims=True)).astype(np.float32)
    off, unk, lvc = (0., n_classes - 1, -1) if variant == "garbage" else (1. / n_classes, -1, None)
    kn, n_kn, ng, n_ng = O.confidence(s, y.numpy(), off, unk, lvc)[:4]
    assert trackers["conf_kn"].count == n_kn and trackers["conf_unk"].count == n_ng        # counts exact
    assert trackers["conf_kn"].avg == pytest.approx(kn, rel=2e-3)
    assert trackers["conf_unk"].avg == pytest.approx(ng, rel=2e-3)
    vname = O.GARBAGE if variant == "garbage" else O.ENTROPIC
    cw = np.ones(n_classes, dtype=np.float32) if variant == "garbage" else None
    # sample-weighted mean of the per-batch losses, as AverageMeter accumulates them (train.py:179-180)
    tot = 0.0
    for i in range(0, 1000, 96):
        zb, yb = z[i:i + 96], y.numpy()[i:i + 96]
        tot += O.loss_from_logits(zb, yb, vname, class_weights=cw)[0] * len(yb)
    assert trackers["j"].avg == pytest.approx(tot / 1000, rel=2e-3) and trackers["j"].count == 1000


def test_train_epoch_with_the_fused_head_learns_and_tracks_the_mean_loss():
    import openset_imagenet_b200 as osb
    torch.manual_seed(1)
    model = TinyModel("entropic").cuda()
    loader = DataLoader(_data(768, "entropic", seed=3), batch_size=128)
    opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9)
    trackers = osb.train.make_trackers()
    losses = []
    for _ in range(4):
        osb.train.train(model, loader, opt, None, trackers, _cfg("entropic", 128))
        losses.append(trackers["j"].avg)
        assert trackers["j"].count == 768
    assert losses[-1] < losses[0] - 0.05 and all(np.isfinite(losses))


def test_get_arrays_npz_round_trip_feeds_oscr(tmp_path):
    import openset_imagenet_b200 as osb
    torch.manual_seed(2)
    model = TinyModel("entropic").cuda()
    ds = _data(500, "entropic", seed=5)
    y = ds.tensors[1].clone()
    y[::7] = -2                                                  # unknown unknowns of a test set
    loader = DataLoader(TensorDataset(ds.tensors[0], y), batch_size=64)
    gt, logits, features, scores = osb.train.get_arrays(model, loader)
    assert gt.dtype == np.float32 and gt.shape == (500,)          # train.py:212 stores targets as float
    assert logits.shape == (500, C) and features.shape == (500, K) and scores.shape == (500, C)
    np.testing.assert_allclose(scores.sum(1), 1.0, rtol=1e-5)
    path = tmp_path / "entropic_test_arr_best.npz"
    osb.train.save_arrays(path, gt, logits, features, scores)
    arrays = osb.train.read_array_list([path])
    (name, arr), = arrays.items()
    assert name == "entropic_test_arr_best.npz"[:-8]              # util.py:84 strips the last 8 characters
    for unk in (-1, -2):
        ccr, fpr = osb.train.oscr_from_arrays(arr, unk_label=unk)
        o_ccr, o_fpr = O.calculate_oscr(arr["gt"], arr["scores"], unk)
        np.testing.assert_array_equal(ccr, o_ccr)
        np.testing.assert_array_equal(fpr, o_fpr)


def test_streaming_confidence_matches_materialised_confidence_exactly_on_the_same_statistics():
    import openset_imagenet_b200 as osb
    gt, s = O.make_oscr_case(7000, 31, seed=3)
    st = torch.tensor(s, device="cuda"); y = torch.tensor(gt, device="cuda")
    ptgt = st.gather(1, y.clamp(min=0).view(-1, 1)).view(-1)
    pmax = st.max(1).values
    sc = osb.StreamingConfidence(31, offset=1 / 31, unknown_class=-1)
    for i in range(0, 7000, 1234):                               # ragged batches
        sc.update(ptgt[i:i + 1234], pmax[i:i + 1234], y[i:i + 1234])
    got = sc.result()
    want = osb.confidence(st, y, 1 / 31, -1, None)
    assert got[1] == want[1] and got[3] == want[3]
    assert got[0] == pytest.approx(want[0], rel=1e-12) and got[2] == pytest.approx(want[2], rel=1e-12)
