"""The drop-in model against the reference's own model.py on the CPU: same parameters, same state-dict keys (the
checkpoint contract of train.py:37-101), strict loading in both directions, same plain forward.  The fused loss path
needs the GPU (tests/test_gpu_model.py)."""
import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference model.py not available")


def _pair(C=116, logit_bias=False):
    from openset_imagenet_b200.model import ResNet50
    torch.manual_seed(0)
    ref = ref_loader.load("model").ResNet50(fc_layer_dim=C, out_features=C, logit_bias=logit_bias)     # train.py:350-352
    ours = ResNet50(fc_layer_dim=C, out_features=C, logit_bias=logit_bias, variant="entropic")
    return ref, ours


@pytest.mark.parametrize("logit_bias", [False, True])
def test_state_dict_contract_and_strict_load_both_ways(logit_bias):
    ref, ours = _pair(logit_bias=logit_bias)
    sd_ref, sd_ours = ref.state_dict(), ours.state_dict()
    assert list(sd_ref) == list(sd_ours)                                   # same keys in the same order
    assert {"resnet_base.fc.weight", "resnet_base.fc.bias", "logits.weight"} <= set(sd_ours)
    assert ("logits.bias" in sd_ours) == logit_bias
    for k in sd_ref:
        assert sd_ref[k].shape == sd_ours[k].shape and sd_ref[k].dtype == sd_ours[k].dtype, k
    ours.load_state_dict(sd_ref, strict=True)                              # a reference checkpoint loads into the drop-in
    ref.load_state_dict(ours.state_dict(), strict=True)                    # ... and the other way round
    assert (ours.logits.in_features, ours.logits.out_features) == (116, 116)       # what train.get_arrays reads, train.py:210-211


def test_plain_forward_is_the_reference_forward():
    ref, ours = _pair(C=12)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ref.eval(); ours.eval()
    x = torch.randn(2, 3, 64, 64)
    with torch.no_grad():
        lr, fr = ref(x)
        lo, fo = ours(x)
        pooled = ours.pooled_features(x)
        w_eff, b_eff = ours.collapsed_weights()
    torch.testing.assert_close(lo, lr)
    torch.testing.assert_close(fo, fr)
    # the collapsed head is the same function: logits(fc(pooled)) = pooled W_eff^T + b_eff (model.py:37-38)
    torch.testing.assert_close(pooled @ w_eff.T + b_eff, lr, rtol=1e-4, atol=1e-5)
