"""SURVEY 8f row 3 on the real model: ``model.ResNet50`` (torchvision backbone, reference parameter layout) with the
fused head against the reference's own ``ResNet50`` + ``EntropicOpensetLoss`` (model.py:8-39, losses.py:7-29, loaded
unmodified) on the GPU: a checkpoint of one loads into the other, and one training step gives the same loss and the
same gradients on ``logits.weight`` and ``resnet_base.fc.{weight,bias}``.

Both models see the same bf16-representable pooled features and parameters (a forward pre-hook on ``resnet_base.fc``
and a rounding of the two head matrices), so what is compared is the head path itself.  Tolerances: the loss 1e-3; the
parameter gradients 5e-3 norm-wise -- one bf16 rounding of an INTERMEDIATE operand is part of either fused mode (the
collapsed 2048 -> C product matrix, or the C-dim features between the two stacked linears), on top of the 1e-3 kernel
contract that tests/test_gpu_head.py holds on given operands."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from oracle import ref_loader

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.available(), reason="reference model.py / losses.py not available")]
DEV = "cuda:0"
C = 116


def _round_bf16_(t):
    with torch.no_grad():
        t.copy_(t.bfloat16().float())


def _models(fused, variant="entropic", **kw):
    from openset_imagenet_b200.model import ResNet50
    torch.manual_seed(0)
    ref = ref_loader.load("model").ResNet50(fc_layer_dim=C, out_features=C, logit_bias=False).to(DEV)
    for p in (ref.resnet_base.fc.weight, ref.resnet_base.fc.bias, ref.logits.weight):
        _round_bf16_(p)
    with torch.no_grad():
        ref.logits.weight.mul_(4.0)                       # default init gives near-uniform softmax outputs; sharpen them a bit
    ours = ResNet50(fc_layer_dim=C, out_features=C, logit_bias=False, fused=fused, variant=variant, **kw).to(DEV)
    ours.load_state_dict(ref.state_dict(), strict=True)                   # the reference checkpoint, strict
    for m in (ref, ours):
        m.resnet_base.fc.register_forward_pre_hook(lambda mod, inp: (inp[0].bfloat16().float(),))
        m.train()
    return ref, ours


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize("fused", ["collapsed", "logits"])
def test_resnet50_training_step_matches_the_reference_model(fused):
    ref, ours = _models(fused)
    losses = ref_loader.load("losses")
    import vast.tools as vt                                              # the device-placement stub of oracle/ref_loader.py
    vt.set_device_gpu(0)
    try:
        loss_fn = losses.EntropicOpensetLoss(C, 1.0)
        g = torch.Generator(device=DEV).manual_seed(1)
        images = torch.randn(48, 3, 96, 96, device=DEV, generator=g)
        y = torch.tensor(O.make_labels(48, C, 1 / 3, seed=4), device=DEV)
        logits, feats = ref(images)                                       # train.py:132
        j_ref = loss_fn(logits, y)                                        # train.py:135
        j_ref.backward()
    finally:
        vt.set_device_cpu()
    if fused == "collapsed":
        # pooled features reach the fused head directly; the hook only acts on resnet_base.fc, so round them here
        orig = ours.pooled_features
        ours.pooled_features = lambda im: orig(im).bfloat16().float()
    j, f = ours.forward_loss(images, y, return_features=True)
    j.backward()
    assert float(j) == pytest.approx(float(j_ref), rel=1e-3)
    torch.testing.assert_close(f, feats, rtol=1e-4, atol=1e-4)
    e_w2 = _rel(ours.logits.weight.grad, ref.logits.weight.grad)
    e_w1 = _rel(ours.resnet_base.fc.weight.grad, ref.resnet_base.fc.weight.grad)
    e_b1 = _rel(ours.resnet_base.fc.bias.grad, ref.resnet_base.fc.bias.grad)
    e_conv = _rel(ours.resnet_base.conv1.weight.grad, ref.resnet_base.conv1.weight.grad)
    print(f"{fused}: loss {abs(float(j) - float(j_ref)) / float(j_ref):.2e}  d logits.weight {e_w2:.2e}  d fc.weight {e_w1:.2e}  "
          f"d fc.bias {e_b1:.2e}  d conv1.weight {e_conv:.2e}")
    assert max(e_w2, e_w1, e_b1) < 5e-3
    assert e_conv < 2e-2                                  # the whole backbone backward behind dX (cuDNN, atomics)
    st = ours.logits.last_stats                           # per-row statistics of the real model's head
    p = torch.softmax(logits.detach().double(), 1)
    # bf16 rounding of an intermediate operand moves a logit by ~1e-2 absolute: a sanity bound, not a parity gate
    torch.testing.assert_close(st.p_max_all.double(), p.max(1).values, rtol=3e-2, atol=1e-6)


def test_resnet50_validate_uses_the_reference_slice_for_the_garbage_head():
    """validate() with the garbage variant: the head built by model.ResNet50 defaults to last_valid_class=-1 (the slice the
    reference's confidence() takes, train.py:156-159); a head built for another slice is refused instead of giving a wrong
    negative confidence."""
    import types
    import openset_imagenet_b200 as osb
    from openset_imagenet_b200.model import ResNet50
    from torch.utils.data import DataLoader, TensorDataset
    cw = torch.ones(C + 1, device=DEV)
    model = ResNet50(fc_layer_dim=C + 1, out_features=C + 1, logit_bias=False, variant="garbage", class_weights=cw).to(DEV)
    assert model.logits.valid_cols == C
    ds = TensorDataset(torch.randn(40, 3, 64, 64), torch.tensor(O.make_labels(40, C, 0.3, seed=1, neg_label=C)))
    cfg = types.SimpleNamespace(loss=types.SimpleNamespace(type="garbage", w=1.0), batch_size=16, parallel=True)
    tr = osb.train.make_trackers()
    osb.train.validate(model, DataLoader(ds, batch_size=16), None, C + 1, tr, cfg)
    assert tr["conf_kn"].count + tr["conf_unk"].count == 40 and np.isfinite(tr["j"].avg)
    bad = ResNet50(fc_layer_dim=C + 1, out_features=C + 1, logit_bias=False, variant="garbage", class_weights=cw,
                   last_valid_class=None).to(DEV)
    with pytest.raises(ValueError, match="last_valid_class"):
        osb.train.validate(bad, DataLoader(ds, batch_size=16), None, C + 1, tr, cfg)


def test_weight_cache_follows_optimizer_steps_and_explicit_invalidation():
    from openset_imagenet_b200.model import FusedOpenSetHead
    head = FusedOpenSetHead(256, 40, device=DEV)
    x = torch.randn(64, 256, device=DEV)
    y = torch.tensor(O.make_labels(64, 40, 0.3, seed=2), device=DEV)
    opt = torch.optim.SGD(head.parameters(), lr=0.5)
    l0 = float(head.loss(x, y))
    head.loss(x, y).backward()
    opt.step()                                            # in-place update through the parameter: the version changes
    l1 = float(head.loss(x, y))
    assert l1 < l0
    head.weight.data.zero_()                              # a write the version counter does not see ...
    assert float(head.loss(x, y)) == l1                   # ... keeps the cached bf16 copy (documented restriction)
    head.invalidate_weight_cache()
    assert float(head.loss(x, y)) == pytest.approx(np.log(40) * (2 / 3) + np.log(40) * (1 / 3), rel=1e-3)   # zero weights: log C per row
