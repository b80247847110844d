"""GPU parity of the encoder modules against golden tensors produced by the REFERENCE modules
(tests/golden/make_encoder_golden.py), with weights from the shared seeded recipe.

Stated bounds for the bf16 tensor-core path against the fp32 reference:
  features: cosine >= 0.999 and max|a-b| / max|b| <= 5e-2
  input gradient: cosine >= 0.99;  per-parameter gradient norms within 10 % (median within 3 %)
"""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from param_recipe import fill_params, make_input, zero_dropout  # noqa: E402

from imagined_speech_translation_b200.brain_encoder import BrainRegionEncoder  # noqa: E402
from imagined_speech_translation_b200.layers import Conv1DWithAttention  # noqa: E402

pytestmark = pytest.mark.gpu


def _cos(a, b):
    return torch.nn.functional.cosine_similarity(a.double().flatten(), b.double().flatten(), dim=0).item()


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


@pytest.fixture(scope="module")
def region():
    return torch.load(os.path.join(HERE, "golden", "encoder_region.pt"))


@pytest.fixture(scope="module")
def brain():
    return torch.load(os.path.join(HERE, "golden", "encoder_brain.pt"))


def _check_grad_norms(module, golden_norms):
    # A bias that feeds a train-mode BatchNorm has an exactly-zero true gradient (the batch mean
    # is subtracted): the reference reports fp32 round-off there, so only parameters whose golden
    # gradient norm is non-negligible are compared.
    floor = 1e-4 * max(golden_norms.values())
    ratios = []
    for n, p in module.named_parameters():
        if n in golden_norms and golden_norms[n] > floor:
            assert p.grad is not None, n
            ratios.append((p.grad.float().norm().item() / golden_norms[n], n))
    r = torch.tensor([x[0] for x in ratios])
    outliers = sorted(ratios, key=lambda t: -abs(t[0] - 1))[:5]
    assert (r - 1).abs().median().item() <= 0.03, f"median grad-norm ratio off: {r.median()}"
    assert abs(outliers[0][0] - 1) <= 0.10, f"grad norms off: {outliers}"
    return len(ratios)


@pytest.mark.parametrize("key", ["stft_train", "stft_eval", "raw_train", "cnn_only"])
def test_region_encoder_matches_reference(region, key):
    rec = region[key]
    cfg = rec["cfg"]
    m = Conv1DWithAttention(cfg["n_channels"], cfg["T"], hidden_dim=768, cnn_only=cfg["cnn_only"])
    fill_params(m, seed=11)
    zero_dropout(m)
    m = m.cuda().train(cfg["train"])
    x = make_input((cfg["B"], cfg["n_channels"], cfg["T"]), seed=21).cuda().requires_grad_(True)
    gout = make_input((cfg["B"], 768), seed=22).cuda()
    out = m(x)
    assert out.shape == rec["out"].shape and out.dtype == torch.float32
    assert _cos(out.cpu(), rec["out"]) >= 0.999
    assert _rel(out.cpu(), rec["out"]) <= 5e-2
    (out * gout).sum().backward()
    assert _cos(x.grad[:, :32].cpu(), rec["dx"]) >= 0.99
    assert _check_grad_norms(m, rec["grad_norm"]) > 50
    if cfg["train"]:   # BatchNorm running statistics follow torch semantics
        assert _rel(m.bn1.running_mean.cpu(), rec["bn1_running_mean"]) <= 2e-2
        assert _rel(m.bn4.running_var.cpu(), rec["bn4_running_var"]) <= 2e-2


@pytest.mark.parametrize("key", ["raw_train", "raw_eval"])
def test_brain_encoder_matches_reference(brain, key):
    rec = brain[key]
    cfg = rec["cfg"]
    m = BrainRegionEncoder(cfg["T"], cfg["counts"], hidden_dim=768)
    fill_params(m, seed=12)
    zero_dropout(m)
    m = m.cuda().train(cfg["train"])
    names = ["frontal", "temporal", "central", "parietal"]
    xs = [make_input((cfg["B"], cfg["counts"][n], cfg["T"]), seed=30 + i).cuda().requires_grad_(True)
          for i, n in enumerate(names)]
    gout = make_input((cfg["B"], 768), seed=40).cuda()
    out = m(xs)
    assert _cos(out.cpu(), rec["out"]) >= 0.999
    assert _rel(out.cpu(), rec["out"]) <= 5e-2
    (out * gout).sum().backward()
    assert _cos(xs[0].grad.cpu(), rec["dx0"]) >= 0.99
    assert _check_grad_norms(m, rec["grad_norm"]) > 200


def test_state_dict_is_interchangeable_with_reference_names(region):
    m = Conv1DWithAttention(16, 33, hidden_dim=768)
    keys = set(m.state_dict().keys())
    for n in region["raw_train"]["grad_norm"]:
        assert n in keys
    for must in ("conv1.weight", "residual1.0.weight", "residual1.1.running_mean", "attn_layers.0.attn.in_proj_weight",
                 "attn_layers.2.ffn.gate.bias", "cross_scale_attn.out_proj.weight", "pos_emb", "cls_token",
                 "multi_scale_proj.2.1.weight", "projection.4.bias", "diversity_head.weight",
                 "se_block.excitation.2.weight", "cnn_to_attn.8.bias"):
        assert must in keys, must


def test_rejects_cpu_tensors():
    m = Conv1DWithAttention(16, 33, hidden_dim=768)
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 16, 33))


@pytest.mark.parametrize("n_channels", [9, 11, 12])
def test_region_sizes_of_the_reference_montage(n_channels):
    """The reference montage gives regions of 16 / 9 / 11 / 12 channels (dataset.py region map): channel counts
    that are not a multiple of 8 run through zero-padded channels.  Checked against the CPU oracle on the same
    weights: features, input gradient, and the gradient of the padded conv1 / residual1 weights."""
    from oracle import encoder_oracle as eo      # checker only
    T, B = 33, 4
    m = Conv1DWithAttention(n_channels, T, hidden_dim=768)
    fill_params(m, seed=5)
    zero_dropout(m)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = make_input((B, n_channels, T), seed=6)
    gout = make_input((B, 768), seed=7)
    xr = x.clone().requires_grad_(True)
    params = {k: sd[k].clone().requires_grad_(True) for k, _ in m.named_parameters()}
    ref = eo.region_encoder({**sd, **params}, xr, train=True)
    (ref * gout).sum().backward()
    m = m.cuda().train()
    xg = x.cuda().requires_grad_(True)
    out = m(xg)
    (out * gout.cuda()).sum().backward()
    assert out.shape == ref.shape
    assert _cos(out.detach().cpu(), ref.detach()) >= 0.999
    assert _cos(xg.grad.cpu(), xr.grad) >= 0.99
    for name in ("conv1.weight", "residual1.0.weight"):
        g = dict(m.named_parameters())[name].grad
        assert g is not None and g.shape == sd[name].shape
        assert _cos(g.cpu(), params[name].grad) >= 0.99, name
