"""Golden tensors for the encoder modules, produced by the REFERENCE modules imported from
/root/reference (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_encoder_golden.py

Writes tests/golden/encoder_region.pt and tests/golden/encoder_brain.pt (small: outputs,
sub-sampled gradients and per-parameter gradient norms; the weights come from the shared
seeded recipe in param_recipe.py, so they are not stored)."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/main_model")
from param_recipe import fill_params, make_input, zero_dropout  # noqa: E402

from src.models.layers import Conv1DWithAttention  # noqa: E402  (the reference itself)
from src.models.brain_encoder import BrainRegionEncoder  # noqa: E402

torch.set_num_threads(8)


def run_region(tag, n_channels, T, B, cnn_only, train):
    torch.manual_seed(0)
    m = Conv1DWithAttention(n_channels, T, hidden_dim=768, cnn_only=cnn_only)
    fill_params(m, seed=11)
    zero_dropout(m)
    m.train(train)
    x = make_input((B, n_channels, T), seed=21).requires_grad_(True)
    gout = make_input((B, 768), seed=22)
    out = m(x)
    (out * gout).sum().backward()
    rec = {"cfg": dict(n_channels=n_channels, T=T, B=B, cnn_only=cnn_only, train=train),
           "out": out.detach().clone(),
           "dx": x.grad[:, :32].clone(),
           "grad_norm": {n: p.grad.norm().item() for n, p in m.named_parameters() if p.grad is not None},
           "bn1_running_mean": m.bn1.running_mean.clone(), "bn4_running_var": m.bn4.running_var.clone()}
    print(tag, "out", out.abs().mean().item(), "params", sum(p.numel() for p in m.parameters()))
    return rec


def run_brain(tag, counts, T, B, train):
    torch.manual_seed(0)
    m = BrainRegionEncoder(T, counts, hidden_dim=768)
    fill_params(m, seed=12)
    zero_dropout(m)
    m.train(train)
    xs = [make_input((B, counts[n], T), seed=30 + i).requires_grad_(True)
          for i, n in enumerate(["frontal", "temporal", "central", "parietal"])]
    gout = make_input((B, 768), seed=40)
    out = m(xs)
    (out * gout).sum().backward()
    gn = {n: p.grad.norm().item() for n, p in m.named_parameters() if p.grad is not None}
    print(tag, "out", out.abs().mean().item(), "params", sum(p.numel() for p in m.parameters()))
    return {"cfg": dict(counts=counts, T=T, B=B, train=train), "out": out.detach().clone(),
            "dx0": xs[0].grad.clone(), "grad_norm": gn}


if __name__ == "__main__":
    region = {
        "stft_train": run_region("stft_train", 2064, 33, 4, False, True),
        "stft_eval": run_region("stft_eval", 2064, 33, 4, False, False),
        "raw_train": run_region("raw_train", 16, 125, 3, False, True),
        "cnn_only": run_region("cnn_only", 16, 64, 3, True, True),
    }
    torch.save(region, os.path.join(HERE, "encoder_region.pt"))
    counts = {"frontal": 16, "temporal": 16, "central": 16, "parietal": 16}
    brain = {"raw_train": run_brain("brain_train", counts, 40, 3, True),
             "raw_eval": run_brain("brain_eval", counts, 40, 3, False)}
    torch.save(brain, os.path.join(HERE, "encoder_brain.pt"))
    for f in ("encoder_region.pt", "encoder_brain.pt"):
        print(f, os.path.getsize(os.path.join(HERE, f)))
