"""GPU parity: row-wise losses, label statistics and confidence against the oracle and the golden vectors
made from the reference.  Tolerances: loss / grad-logits 1e-3 relative (north star; fp32 logits are in
practice ~1e-6), counts bit-exact, confidence sums 1e-6 relative against the float64 oracle."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
REL = 1e-3


def _dev():
    return torch.device("cuda:0")


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def _run_loss(loss_fn, z, y, dtype=torch.float32):
    zt = torch.tensor(z, device=_dev(), dtype=dtype, requires_grad=True)
    loss = loss_fn(zt, torch.tensor(y, device=_dev()))
    loss.backward()
    return float(loss.item()), zt.grad.float().cpu().numpy()


@pytest.mark.parametrize("name", ["cfg1", "w05", "allneg", "allknown", "single", "p3"])
def test_eos_golden(golden, name):
    import openset_imagenet_b200 as osb
    g = golden["losses"]
    z, y, w = g[f"eos_{name}_logits"], g[f"eos_{name}_y"], float(g[f"eos_{name}_w"])
    loss, grad = _run_loss(osb.EntropicOpensetLoss(z.shape[1], w), z, y)
    assert loss == pytest.approx(float(g[f"eos_{name}_loss"]), rel=1e-5, abs=1e-6)
    assert _rel(grad, g[f"eos_{name}_grad"]) < 1e-5
    o_loss, o_dz, _, _ = O.loss_from_logits(z, y, O.ENTROPIC, w)
    assert loss == pytest.approx(o_loss, rel=1e-5, abs=1e-6)
    assert np.abs(grad - o_dz).max() < 1e-6


def test_softmax_and_garbage_golden(golden):
    import openset_imagenet_b200 as osb
    g = golden["losses"]
    loss, grad = _run_loss(osb.CrossEntropyLoss(ignore_index=-1), g["sm_logits"], g["sm_y"])
    assert loss == pytest.approx(float(g["sm_loss"]), rel=1e-5)
    assert _rel(grad, g["sm_grad"]) < 1e-5
    # all rows ignored -> NaN like torch
    loss, _ = _run_loss(osb.CrossEntropyLoss(ignore_index=-1), g["sm_logits"][:8], np.full(8, -1))
    assert np.isnan(loss)
    cw = torch.tensor(g["bg_cw"], device=_dev())
    loss, grad = _run_loss(osb.CrossEntropyLoss(weight=cw), g["bg_logits"], g["bg_y"])
    assert loss == pytest.approx(float(g["bg_loss"]), rel=1e-5)
    assert _rel(grad, g["bg_grad"]) < 1e-5
    assert torch.equal(osb.losses.class_weights_from_labels(torch.tensor(g["bg_train_labels"], device=_dev()), 152).cpu(),
                       torch.tensor(g["bg_cw"]))


@pytest.mark.parametrize("variant", ["entropic", "softmax", "garbage"])
@pytest.mark.parametrize("B,C", [(1, 2), (7, 33), (64, 116), (256, 152), (300, 1000), (33, 1500)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_loss_random_vs_oracle(variant, B, C, dtype):
    import openset_imagenet_b200 as osb
    rng = np.random.default_rng(B * 1000 + C)
    z = (3 * rng.standard_normal((B, C))).astype(np.float32)
    if dtype == torch.bfloat16:
        z = O.bf16_round(z)
    if variant == "entropic":
        y = O.make_labels(B, C, 0.4, seed=B + C, neg_label=-1)
        y[::5] = np.where(y[::5] < 0, -2, y[::5])
        fn, ov, kw = osb.EntropicOpensetLoss(C, 0.7), O.ENTROPIC, dict(unk_weight=0.7)
    elif variant == "softmax":
        y = O.make_labels(B, C, 0.3, seed=B + C, neg_label=-1)
        if np.all(y < 0):
            y[0] = 0
        fn, ov, kw = osb.CrossEntropyLoss(ignore_index=-1), O.SOFTMAX, {}
    else:
        y = O.make_labels(B, C - 1, 0.3, seed=B + C, neg_label=C - 1)
        cw = (0.5 + rng.random(C)).astype(np.float32)
        fn, ov, kw = osb.CrossEntropyLoss(weight=torch.tensor(cw, device=_dev())), O.GARBAGE, dict(class_weights=cw)
    loss, grad = _run_loss(fn, z, y, dtype)
    o_loss, o_dz, _, _ = O.loss_from_logits(z, y, ov, **kw)
    assert loss == pytest.approx(o_loss, rel=REL)
    tol = REL if dtype == torch.float32 else 8e-3          # bf16 OUTPUT rounding of dlogits (2^-9 per element)
    assert _rel(grad, o_dz) < tol


def test_upstream_gradient_and_nonunit_scale():
    import openset_imagenet_b200 as osb
    rng = np.random.default_rng(3)
    z = rng.standard_normal((20, 17)).astype(np.float32)
    y = O.make_labels(20, 17, 0.3, seed=1)
    zt = torch.tensor(z, device=_dev(), requires_grad=True)
    (osb.EntropicOpensetLoss(17)(zt, torch.tensor(y, device=_dev())) * 2.5).backward()
    _, o_dz, _, _ = O.loss_from_logits(z, y, O.ENTROPIC)
    assert np.abs(zt.grad.cpu().numpy() - 2.5 * o_dz).max() < 1e-6


def test_label_errors_match_reference():
    import openset_imagenet_b200 as osb
    z = torch.zeros(4, 5, device=_dev())
    with pytest.raises(IndexError):
        osb.EntropicOpensetLoss(5)(z, torch.tensor([0, 5, 1, -1], device=_dev()))
    loss = osb.EntropicOpensetLoss(5, validate_labels=False)(z, torch.tensor([0, 5, 1, -1], device=_dev()))
    assert torch.isnan(loss)
    with pytest.raises(RuntimeError):
        osb.EntropicOpensetLoss(6)(z, torch.tensor([0, 1, 1, -1], device=_dev()))


def test_noncontiguous_and_fp16_inputs():
    import openset_imagenet_b200 as osb
    rng = np.random.default_rng(5)
    big = torch.tensor(rng.standard_normal((12, 64)).astype(np.float32), device=_dev())
    view = big[:, :40]                                   # row pitch 64, 40 columns
    y = O.make_labels(12, 40, 0.3, seed=2)
    loss = osb.EntropicOpensetLoss(40)(view, torch.tensor(y, device=_dev()))
    assert float(loss) == pytest.approx(O.loss_from_logits(view.cpu().numpy(), y, O.ENTROPIC)[0], rel=1e-5)
    loss16 = osb.EntropicOpensetLoss(40)(view.half(), torch.tensor(y, device=_dev()))
    assert float(loss16) == pytest.approx(float(loss), rel=2e-3)


@pytest.mark.parametrize("name", ["eos", "eos_mixed", "table", "bg", "noneg", "noknown"])
def test_confidence_golden(golden, name):
    import openset_imagenet_b200 as osb
    g = golden["confidence"]
    off, uc, lvc = g[f"{name}_args"]
    lvc = None if np.isnan(lvc) else int(lvc)
    s, y = g[f"{name}_scores"], g[f"{name}_y"]
    r = osb.confidence(torch.tensor(s, device=_dev()), torch.tensor(y, device=_dev()), float(off), int(uc), lvc)
    assert [type(v) for v in r] == [float, int, float, int]
    ref = g[f"{name}_out"]
    assert (r[1], r[3]) == (int(ref[1]), int(ref[3]))
    assert r[0] == pytest.approx(ref[0], rel=2e-6, abs=1e-7) and r[2] == pytest.approx(ref[2], rel=2e-6, abs=1e-7)
    o = O.confidence(s, y, float(off), int(uc), lvc)
    assert r[0] == pytest.approx(o[0], rel=1e-12, abs=1e-15) and r[2] == pytest.approx(o[2], rel=1e-12, abs=1e-15)


@pytest.mark.parametrize("N,C,lvc", [(1, 3, None), (1000, 116, None), (4097, 152, -1), (50000, 117, None),
                                     (3000, 1000, None), (513, 30, 7)])
def test_confidence_random(N, C, lvc):
    import openset_imagenet_b200 as osb
    rng = np.random.default_rng(N + C)
    s = torch.softmax(torch.tensor(rng.standard_normal((N, C)).astype(np.float32) * 2), 1).numpy()
    y = O.make_labels(N, C, 0.3, seed=N)
    y[rng.random(N) < 0.1] = -2
    for uc in (-1, -2, C - 1):
        r = osb.confidence(torch.tensor(s, device=_dev()), torch.tensor(y, device=_dev()), 1.0 / C, uc, lvc)
        o = O.confidence(s, y, 1.0 / C, uc, lvc)
        assert (r[1], r[3]) == (o[1], o[3])
        assert r[0] == pytest.approx(o[0], rel=1e-9, abs=1e-12) and r[2] == pytest.approx(o[2], rel=1e-9, abs=1e-12)
    # run-to-run bit reproducibility of the sums
    from openset_imagenet_b200.metrics import confidence_sums
    a = confidence_sums(torch.tensor(s, device=_dev()), torch.tensor(y, device=_dev()), 0.1)[0].cpu()
    b = confidence_sums(torch.tensor(s, device=_dev()), torch.tensor(y, device=_dev()), 0.1)[0].cpu()
    assert torch.equal(a, b)


def test_confidence_empty_and_bad_labels():
    import openset_imagenet_b200 as osb
    with pytest.raises(AttributeError):
        osb.confidence(torch.zeros(0, 4, device=_dev()), torch.zeros(0, dtype=torch.long, device=_dev()))
    with pytest.raises(IndexError):
        osb.confidence(torch.rand(3, 4, device=_dev()), torch.tensor([0, 4, -1], device=_dev()))
