"""CPU oracle for the encoder half of the hot path (TEST INFRASTRUCTURE, stock PyTorch, fp32).

A functional restatement of what the reference modules compute, written over a plain
``state_dict`` (the reference's parameter names), so the same function checks the reference's
golden tensors *and* consumes the weights of our own modules:

  region_encoder(sd, x, ...)   <->  Conv1DWithAttention.forward   main_model/src/models/layers.py:129-272
                                    (SqueezeExciteBlock :288-298, FeedForwardNetwork :311-317)
  brain_encoder(sd, xs, ...)   <->  BrainRegionEncoder.forward    main_model/src/models/brain_encoder.py:136-193
                                    (apply_multi_scale_processing :94-113,
                                     compute_dynamic_region_weights :115-134)

Pinned against golden tensors produced by the reference modules themselves
(tests/golden/encoder_region.pt, encoder_brain.pt; tests/test_oracle_encoder.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this; the product path never does.

Op semantics follow SURVEY.md section 3.2: exact (erf) GELU, BatchNorm1d with biased batch
variance in train mode, nn.MultiheadAttention's packed in-projection with q scaled by
1/sqrt(head_dim), pre-LN transformer layers.  Dropout is NOT modelled (parity runs use p = 0).
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def _sub(sd: SD, prefix: str) -> SD:
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def _lin(sd: SD, name: str, x):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _ln(sd: SD, name: str, x):
    w = sd[name + ".weight"]
    return F.layer_norm(x, (w.numel(),), w, sd[name + ".bias"], 1e-5)


def _bn(sd: SD, name: str, x, train: bool):
    """BatchNorm1d on (B, C, T); train=True normalises with the batch statistics (the running
    buffers in ``sd`` are left untouched: the oracle is stateless)."""
    if train:
        return F.batch_norm(x, None, None, sd[name + ".weight"], sd[name + ".bias"], True, 0.0, 1e-5)
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                        sd[name + ".bias"], False, 0.0, 1e-5)


def _mha(sd: SD, name: str, q_in, kv_in, heads: int):
    """nn.MultiheadAttention(batch_first=True), no mask, no dropout: rows [Wq; Wk; Wv] of
    in_proj_weight, heads split along the feature dim, softmax(q k^T / sqrt(hd)) v, out_proj."""
    d = q_in.shape[-1]
    W, b = sd[name + ".in_proj_weight"], sd[name + ".in_proj_bias"]
    q = F.linear(q_in, W[:d], b[:d])
    k = F.linear(kv_in, W[d:2 * d], b[d:2 * d])
    v = F.linear(kv_in, W[2 * d:], b[2 * d:])
    B, Sq, _ = q.shape
    Sk, hd = k.shape[1], d // heads
    q = q.view(B, Sq, heads, hd).transpose(1, 2)
    k = k.view(B, Sk, heads, hd).transpose(1, 2)
    v = v.view(B, Sk, heads, hd).transpose(1, 2)
    p = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, Sq, d)
    return F.linear(o, sd[name + ".out_proj.weight"], sd[name + ".out_proj.bias"])


def _res_block(sd: SD, x, conv: str, bn: str, res: str, pad: int, train: bool):
    y = _bn(sd, bn, F.conv1d(x, sd[conv + ".weight"], sd[conv + ".bias"], padding=pad), train)
    r = _bn(sd, res + ".1", F.conv1d(x, sd[res + ".0.weight"]), train) if (res + ".0.weight") in sd else x
    return F.gelu(y + r)


def region_encoder(sd: SD, x: torch.Tensor, train: bool = True, cnn_only: bool = False,
                   n_heads: int = 8) -> torch.Tensor:
    """(B, C, T) -> (B, hidden).  sd: state_dict of one Conv1DWithAttention."""
    h = _res_block(sd, x, "conv1", "bn1", "residual1", 4, train)
    h = _res_block(sd, h, "conv2", "bn2", "residual2", 3, train)
    h = F.conv1d(h, sd["depthwise_conv.weight"], sd["depthwise_conv.bias"], padding=2, groups=h.shape[1])
    h = F.gelu(_bn(sd, "bn_depth", F.conv1d(h, sd["pointwise_conv.weight"], sd["pointwise_conv.bias"]), train))
    h = _res_block(sd, h, "conv3", "bn3", "residual3", 2, train)
    h = _res_block(sd, h, "conv4", "bn4", "residual4", 1, train)
    # squeeze-excite: mean over time -> 768 -> 48 -> 768 -> sigmoid -> scale
    s = h.mean(dim=2)
    e = torch.sigmoid(_lin(sd, "se_block.excitation.2", torch.relu(_lin(sd, "se_block.excitation.0", s))))
    h = (h * e.unsqueeze(2)).transpose(1, 2)                     # (B, T, 768)

    def finish(feats: List[torch.Tensor]):
        parts = [F.gelu(_ln(sd, f"multi_scale_proj.{i}.1", _lin(sd, f"multi_scale_proj.{i}.0", f)))
                 for i, f in enumerate(feats)]
        z = F.gelu(_ln(sd, "projection.1", _lin(sd, "projection.0", torch.cat(parts, dim=1))))
        final = _ln(sd, "projection.5", _lin(sd, "projection.4", z))
        return final + 0.1 * F.normalize(_lin(sd, "diversity_head", final), dim=-1)

    if cnn_only:
        mean_pool, max_pool = h.mean(dim=1), h.max(dim=1)[0]
        w = torch.softmax((h * mean_pool.unsqueeze(1)).sum(dim=2), dim=1)
        return finish([mean_pool, max_pool, (h * w.unsqueeze(2)).sum(dim=1)])

    h = F.gelu(_ln(sd, "cnn_to_attn.1", _lin(sd, "cnn_to_attn.0", h)))
    h = F.gelu(_ln(sd, "cnn_to_attn.5", _lin(sd, "cnn_to_attn.4", h)))
    h = _lin(sd, "cnn_to_attn.8", h)
    B = h.shape[0]
    h = torch.cat([sd["cls_token"].expand(B, -1, -1), sd["temporal_tokens"].expand(B, -1, -1), h], dim=1)
    S, pos = h.shape[1], sd["pos_emb"]
    if S > pos.shape[1]:
        pos = pos.repeat(1, S // pos.shape[1] + 1, 1)
    h = h + pos[:, :S]
    heads = [n_heads, max(4, n_heads // 2), max(4, n_heads // 2)]
    prev = None
    for i in range(3):
        n = _ln(sd, f"attn_layers.{i}.attn_norm", h)
        h = h + _mha(sd, f"attn_layers.{i}.attn", n, n, heads[i])
        saved = h
        n = _ln(sd, f"attn_layers.{i}.ffn_norm", h)
        gated = F.gelu(_lin(sd, f"attn_layers.{i}.ffn.linear1", n)) * torch.sigmoid(_lin(sd, f"attn_layers.{i}.ffn.gate", n))
        h = h + _lin(sd, f"attn_layers.{i}.ffn.linear2", gated)
        if i > 0:
            h = h + 0.1 * _mha(sd, "cross_scale_attn", h, prev, n_heads // 2)
        prev = saved
    feat = h[:, 0] + 0.3 * h[:, 1:4].mean(dim=1)
    return finish([feat, feat, feat])


REGIONS = ("frontal", "temporal", "central", "parietal")


def brain_encoder(sd: SD, xs: List[torch.Tensor], train: bool = True) -> torch.Tensor:
    """list[4] of (B, C_r, T) -> (B, hidden).  sd: state_dict of one BrainRegionEncoder
    (default flags: cross-region attention on, learned region weights)."""
    x = torch.stack([region_encoder(_sub(sd, f"region_encoders.{n}."), xs[i], train)
                     for i, n in enumerate(REGIONS)], dim=1)                  # (B, 4, d)
    B, R, d = x.shape
    xe = x.transpose(1, 2)
    scales = []
    for i, k in enumerate((3, 7, 15, 31)):
        y = F.gelu(F.conv1d(xe, sd[f"temporal_scales.{i}.weight"], sd[f"temporal_scales.{i}.bias"], padding=k // 2))
        scales.append(y.mean(dim=2))
    ms = torch.stack(scales, dim=1).reshape(B, -1)
    ms = _ln(sd, "diversity_projection.4", _lin(sd, "diversity_projection.3", F.gelu(_lin(sd, "diversity_projection.0", ms))))
    x = x + 0.3 * ms.unsqueeze(1) + 0.4 * sd["region_embeddings.weight"].unsqueeze(0)

    def enhancer(v):
        return _ln(sd, "feature_enhancer.4", _lin(sd, "feature_enhancer.3", F.gelu(_lin(sd, "feature_enhancer.0", v))))

    for i in range(2):                                           # pre-LN TransformerEncoderLayer x2
        p = f"fusion_transformer.layers.{i}"
        n = _ln(sd, p + ".norm1", x)
        x = x + _mha(sd, p + ".self_attn", n, n, 12)
        x = x + _lin(sd, p + ".linear2", F.gelu(_lin(sd, p + ".linear1", _ln(sd, p + ".norm2", x))))
    xc = _mha(sd, "cross_region_attention", x, x, 8)
    x = x + torch.sigmoid(enhancer(x.mean(dim=1))).unsqueeze(1) * xc

    dyn = torch.sigmoid(_lin(sd, "region_gate.3", F.gelu(_lin(sd, "region_gate.0", x.mean(dim=1)))))
    w = F.softmax(0.7 * F.softmax(sd["region_importance"], dim=0).unsqueeze(0) + 0.3 * dyn, dim=1)
    fused = (x * w.unsqueeze(-1)).sum(dim=1)
    return fused + 0.3 * enhancer(fused)
