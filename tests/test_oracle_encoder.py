"""Pins the encoder oracle (oracle/encoder_oracle.py) against golden tensors produced by the
REFERENCE modules (tests/golden/make_encoder_golden.py).  CPU only.  The weights come from the
shared seeded recipe applied to a parameter container with the reference's names."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from param_recipe import fill_params, make_input  # noqa: E402

from imagined_speech_translation_b200.brain_encoder import BrainRegionEncoder  # parameter container only
from imagined_speech_translation_b200.layers import Conv1DWithAttention
from oracle import encoder_oracle as eo


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


@pytest.fixture(scope="module")
def region():
    return torch.load(os.path.join(HERE, "golden", "encoder_region.pt"))


@pytest.mark.parametrize("key", ["stft_train", "stft_eval", "raw_train", "cnn_only"])
def test_region_oracle_matches_reference_golden(region, key):
    rec = region[key]
    cfg = rec["cfg"]
    holder = fill_params(Conv1DWithAttention(cfg["n_channels"], cfg["T"], hidden_dim=768, cnn_only=cfg["cnn_only"]), seed=11)
    sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and "running" not in k)
          for k, v in holder.state_dict().items()}
    x = make_input((cfg["B"], cfg["n_channels"], cfg["T"]), seed=21).requires_grad_(True)
    out = eo.region_encoder(sd, x, train=cfg["train"], cnn_only=cfg["cnn_only"])
    assert _rel(out, rec["out"]) <= 2e-5            # fp32 vs fp32, same library kernels
    (out * make_input((cfg["B"], 768), seed=22)).sum().backward()
    assert _rel(x.grad[:, :32], rec["dx"]) <= 1e-4
    floor = 1e-4 * max(rec["grad_norm"].values())
    for n, g in rec["grad_norm"].items():
        if g > floor:
            assert sd[n].grad.norm().item() == pytest.approx(g, rel=1e-3), n


@pytest.mark.parametrize("key", ["raw_train", "raw_eval"])
def test_brain_oracle_matches_reference_golden(key):
    rec = torch.load(os.path.join(HERE, "golden", "encoder_brain.pt"))[key]
    cfg = rec["cfg"]
    holder = fill_params(BrainRegionEncoder(cfg["T"], cfg["counts"], hidden_dim=768), seed=12)
    sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and "running" not in k)
          for k, v in holder.state_dict().items()}
    xs = [make_input((cfg["B"], cfg["counts"][n], cfg["T"]), seed=30 + i).requires_grad_(True)
          for i, n in enumerate(eo.REGIONS)]
    out = eo.brain_encoder(sd, xs, train=cfg["train"])
    assert _rel(out, rec["out"]) <= 5e-5
    (out * make_input((cfg["B"], 768), seed=40)).sum().backward()
    assert _rel(xs[0].grad, rec["dx0"]) <= 2e-4
