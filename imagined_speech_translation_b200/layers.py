"""Region feature extractor on the B200 path.

Drop-in for the reference ``Conv1DWithAttention`` / ``SqueezeExciteBlock`` /
``FeedForwardNetwork`` (``main_model/src/models/layers.py:9-317``): identical constructor
arguments, parameter names and shapes (so ``state_dict``s interchange and
``get_optimizer_groups`` / ``initialize_custom_weights`` keep working), identical forward
semantics (SURVEY.md section 3.2 and its op-semantics table).  The torch.nn sub-modules are kept
only as parameter containers; the arithmetic runs channels-last in bf16 with every Conv1d /
Linear contraction on the tcgen05 GEMM (``nn_ops``) in forward, dgrad and wgrad.

No CPU path: inputs must be CUDA tensors.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import nn_ops
from .nn_ops import PAD


def _pointwise_residual(c_in: int, c_out: int) -> nn.Module:
    # reference layers.py:20-27: identity when widths match, else 1x1 conv (no bias) + BatchNorm
    if c_in == c_out:
        return nn.Identity()
    return nn.Sequential(nn.Conv1d(c_in, c_out, kernel_size=1, bias=False), nn.BatchNorm1d(c_out))


class SqueezeExciteBlock(nn.Module):
    """Parameter-compatible with reference layers.py:275-298; forward on (B, T, C) channels-last."""

    def __init__(self, channels, reduction=16):
        super().__init__()
        self.squeeze = nn.AdaptiveAvgPool1d(1)
        self.excitation = nn.Sequential(
            nn.Linear(channels, channels // reduction), nn.ReLU(inplace=True),
            nn.Linear(channels // reduction, channels), nn.Sigmoid())

    def forward(self, x_cl):
        # s = mean_T(x); e = sigmoid(W2 relu(W1 s + b1) + b2); x * e   (tiny: M = B rows)
        s = x_cl.float().mean(dim=1)
        fc1, fc2 = self.excitation[0], self.excitation[2]
        e = torch.sigmoid(F.linear(torch.relu(F.linear(s, fc1.weight, fc1.bias)), fc2.weight, fc2.bias))
        return x_cl * e.to(x_cl.dtype).unsqueeze(1)


class FeedForwardNetwork(nn.Module):
    """Gated FFN, reference layers.py:301-317: W2 drop(gelu(W1 x) * sigmoid(Wg x))."""

    def __init__(self, input_dim, hidden_dim):
        super().__init__()
        self.linear1 = nn.Linear(input_dim, hidden_dim)
        self.linear2 = nn.Linear(hidden_dim, input_dim)
        self.gate = nn.Linear(input_dim, hidden_dim)
        self.dropout = nn.Dropout(0.1)

    def forward(self, x):
        act = F.gelu(nn_ops.linear(x, self.linear1.weight, self.linear1.bias).float())
        gate = torch.sigmoid(nn_ops.linear(x, self.gate.weight, self.gate.bias).float())
        h = self.dropout(act * gate).to(torch.bfloat16)
        return nn_ops.linear(h, self.linear2.weight, self.linear2.bias)


def _mha(mod: nn.MultiheadAttention, q_in, kv_in, self_attn: bool):
    """nn.MultiheadAttention(batch_first=True) semantics (packed in_proj, q scaled by
    1/sqrt(head_dim), softmax, dropout on the probabilities in train mode, out_proj); the
    head-averaged attention weights the reference discards are not produced."""
    B, Sq, d = q_in.shape
    H = mod.num_heads
    hd = d // H
    W, b = mod.in_proj_weight, mod.in_proj_bias
    if self_attn:
        qkv = nn_ops.linear(q_in, W, b)                              # (B, S, 3d)
        q, k, v = qkv.split(d, dim=-1)
    else:
        q = nn_ops.linear(q_in, W[:d], b[:d])
        kv = nn_ops.linear(kv_in, W[d:], b[d:])
        k, v = kv.split(d, dim=-1)
    Sk = k.shape[1]
    q = q.reshape(B, Sq, H, hd).transpose(1, 2)
    k = k.reshape(B, Sk, H, hd).transpose(1, 2)
    v = v.reshape(B, Sk, H, hd).transpose(1, 2)
    scores = torch.matmul(q, k.transpose(-1, -2)).float() * (1.0 / math.sqrt(hd))
    p = torch.softmax(scores, dim=-1)
    if mod.training and mod.dropout > 0.0:
        p = F.dropout(p, mod.dropout)
    o = torch.matmul(p.to(v.dtype), v).transpose(1, 2).reshape(B, Sq, d)
    return nn_ops.linear(o, mod.out_proj.weight, mod.out_proj.bias)


def _layer_norm(x, ln: nn.LayerNorm):
    return F.layer_norm(x.float(), ln.normalized_shape, ln.weight, ln.bias, ln.eps).to(torch.bfloat16)


def run_sequential(seq: nn.Sequential, x):
    """A Sequential of Linear / LayerNorm / GELU / Sigmoid / Dropout applied to bf16 rows with
    our contractions (the containers only hold the parameters)."""
    for m in seq:
        if isinstance(m, nn.Linear):
            x = nn_ops.linear(x, m.weight, m.bias)
        elif isinstance(m, nn.LayerNorm):
            x = _layer_norm(x, m)
        elif isinstance(m, nn.GELU):
            x = F.gelu(x.float()).to(torch.bfloat16)
        elif isinstance(m, nn.Sigmoid):
            x = torch.sigmoid(x.float()).to(torch.bfloat16)
        elif isinstance(m, nn.Dropout):
            x = m(x)
        else:
            raise TypeError(type(m))
    return x


class Conv1DWithAttention(nn.Module):
    """(B, n_channels, n_timepoints) -> (B, hidden_dim).  Reference: layers.py:9-272."""

    def __init__(self, n_channels, n_timepoints, hidden_dim=128, n_heads=8, cnn_only=False):
        super().__init__()
        self.cnn_only = cnn_only
        self.hidden_dim = hidden_dim
        self.n_timepoints = n_timepoints
        # parameter containers, created in the reference's order and under its names
        self.conv1 = nn.Conv1d(n_channels, 128, kernel_size=9, padding=4)
        self.bn1 = nn.BatchNorm1d(128)
        self.residual1 = _pointwise_residual(n_channels, 128)
        self.conv2 = nn.Conv1d(128, 256, kernel_size=7, padding=3)
        self.bn2 = nn.BatchNorm1d(256)
        self.residual2 = _pointwise_residual(128, 256)
        self.depthwise_conv = nn.Conv1d(256, 256, kernel_size=5, padding=2, groups=256)
        self.pointwise_conv = nn.Conv1d(256, 384, kernel_size=1)
        self.bn_depth = nn.BatchNorm1d(384)
        self.conv3 = nn.Conv1d(384, 512, kernel_size=5, padding=2)
        self.bn3 = nn.BatchNorm1d(512)
        self.residual3 = _pointwise_residual(384, 512)
        self.conv4 = nn.Conv1d(512, 768, kernel_size=3, padding=1)
        self.bn4 = nn.BatchNorm1d(768)
        self.residual4 = _pointwise_residual(512, 768)
        self.se_block = SqueezeExciteBlock(768)
        self.dropout_light = nn.Dropout(0.05)
        self.dropout_medium = nn.Dropout(0.1)
        self.dropout_heavy = nn.Dropout(0.15)

        if not cnn_only:
            self.cnn_to_attn = nn.Sequential(
                nn.Linear(768, hidden_dim * 2), nn.LayerNorm(hidden_dim * 2), nn.GELU(), nn.Dropout(0.1),
                nn.Linear(hidden_dim * 2, hidden_dim), nn.LayerNorm(hidden_dim), nn.GELU(), nn.Dropout(0.05),
                nn.Linear(hidden_dim, hidden_dim))
            self.cls_token = nn.Parameter(torch.randn(1, 1, hidden_dim) * 0.02)
            self.temporal_tokens = nn.Parameter(torch.randn(1, 3, hidden_dim) * 0.02)
            self.pos_emb = nn.Parameter(torch.randn(1, n_timepoints + 4, hidden_dim) * 0.02)
            self.attn_layers = nn.ModuleList([
                nn.ModuleDict({
                    'attn_norm': nn.LayerNorm(hidden_dim),
                    'attn': nn.MultiheadAttention(embed_dim=hidden_dim,
                                                  num_heads=n_heads if i == 0 else max(4, n_heads // 2),
                                                  dropout=0.1, batch_first=True),
                    'ffn_norm': nn.LayerNorm(hidden_dim),
                    'ffn': FeedForwardNetwork(hidden_dim, hidden_dim * (4 if i == 0 else 2)),
                }) for i in range(3)])
            self.cross_scale_attn = nn.MultiheadAttention(embed_dim=hidden_dim, num_heads=n_heads // 2,
                                                          dropout=0.1, batch_first=True)
        proj_in = 768 if cnn_only else hidden_dim
        self.multi_scale_proj = nn.ModuleList([
            nn.Sequential(nn.Linear(proj_in, hidden_dim), nn.LayerNorm(hidden_dim), nn.GELU(), nn.Dropout(0.05))
            for _ in range(3)])
        self.projection = nn.Sequential(
            nn.Linear(hidden_dim * 3, hidden_dim * 2), nn.LayerNorm(hidden_dim * 2), nn.GELU(), nn.Dropout(0.1),
            nn.Linear(hidden_dim * 2, hidden_dim), nn.LayerNorm(hidden_dim))
        self.diversity_head = nn.Linear(hidden_dim, hidden_dim)

    # ------------------------------------------------------------------ CNN stack
    def _batch_norm(self, y, bn: nn.BatchNorm1d, mask, n_valid):
        """BatchNorm1d over the valid rows of a (M, C) tensor (train: biased batch variance for
        normalisation, running_var updated with the unbiased one; eval: running statistics)."""
        y = y.float()
        if bn.training:
            ym = y * mask
            mean = ym.sum(0) / n_valid
            var = ((y - mean) * mask).pow(2).sum(0) / n_valid
            with torch.no_grad():
                m = bn.momentum
                bn.running_mean.mul_(1 - m).add_(mean, alpha=m)
                bn.running_var.mul_(1 - m).add_(var * (n_valid / max(n_valid - 1, 1)), alpha=m)
                bn.num_batches_tracked += 1
        else:
            mean, var = bn.running_mean, bn.running_var
        return (y - mean) * torch.rsqrt(var + bn.eps) * bn.weight + bn.bias

    def _res_block(self, h, conv, bn, res, drop, B, T, mask):
        """gelu(bn(conv(h)) + res(h)) on channels-last rows; h: (B, T, C_in) bf16."""
        M = B * (T + 2 * PAD)
        n_valid = B * T
        buf = nn_ops.guard_pad(h)
        y = self._batch_norm(nn_ops.conv1d_cl(buf, conv.weight, conv.bias, M), bn, mask, n_valid)
        if isinstance(res, nn.Identity):
            r = buf[PAD:PAD + M].float()
        else:
            r = self._batch_norm(nn_ops.conv1d_cl(buf, res[0].weight, None, M), res[1], mask, n_valid)
        out = F.gelu(y + r)
        if drop is not None:
            out = drop(out)
        out = (out * mask).to(torch.bfloat16)
        return out.view(B, T + 2 * PAD, -1)[:, PAD:PAD + T]

    def _depthwise_block(self, h, B, T, mask):
        # depthwise k5 (groups = channels) -> pointwise 1x1 -> BN -> GELU   (layers.py:157-161)
        C = h.shape[-1]
        w = self.depthwise_conv.weight                         # (C, 1, 5)
        xp = F.pad(h.float(), (0, 0, 2, 2))                    # (B, T+4, C) zero padded in time
        d = self.depthwise_conv.bias.view(1, 1, C).expand(B, T, C)
        for tap in range(5):
            d = d + xp[:, tap:tap + T] * w[:, 0, tap].view(1, 1, C)
        d = d.to(torch.bfloat16)
        M = B * (T + 2 * PAD)
        buf = nn_ops.guard_pad(d)
        pw = self.pointwise_conv
        y = nn_ops.conv1d_cl(buf, pw.weight, pw.bias, M)
        y = self._batch_norm(y, self.bn_depth, mask, B * T)
        out = self.dropout_medium(F.gelu(y))
        out = (out * mask).to(torch.bfloat16)
        return out.view(B, T + 2 * PAD, -1)[:, PAD:PAD + T]

    def _cnn(self, x):
        B, C, T = x.shape
        if C % 8 != 0:
            raise ValueError("n_channels must be a multiple of 8 on this path (TMA row pitch)")
        h = x.transpose(1, 2).to(torch.bfloat16)               # channels-last
        Tp = T + 2 * PAD
        t = torch.arange(Tp, device=x.device)
        mask = ((t >= PAD) & (t < PAD + T)).float().repeat(B).view(B * Tp, 1)
        h = self._res_block(h, self.conv1, self.bn1, self.residual1, self.dropout_light, B, T, mask)
        h = self._res_block(h, self.conv2, self.bn2, self.residual2, self.dropout_light, B, T, mask)
        h = self._depthwise_block(h, B, T, mask)
        h = self._res_block(h, self.conv3, self.bn3, self.residual3, self.dropout_medium, B, T, mask)
        h = self._res_block(h, self.conv4, self.bn4, self.residual4, None, B, T, mask)
        h = self.se_block(h)
        return self.dropout_heavy(h)                            # (B, T, 768) bf16

    # ------------------------------------------------------------------ heads
    def _mlp(self, seq: nn.Sequential, x):
        return run_sequential(seq, x)

    def _finish(self, feats):
        comb = torch.cat([self._mlp(p, f) for p, f in zip(self.multi_scale_proj, feats)], dim=1)
        final = self._mlp(self.projection, comb).float()
        div = nn_ops.linear(final.to(torch.bfloat16), self.diversity_head.weight, self.diversity_head.bias).float()
        return final + 0.1 * F.normalize(div, dim=-1)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("Conv1DWithAttention (B200 path) needs CUDA tensors: there is no CPU fallback")
        B = x.size(0)
        h = self._cnn(x)                                        # (B, T, 768)
        if self.cnn_only:
            hf = h.float()
            mean_pool = hf.mean(dim=1)
            max_pool = hf.max(dim=1)[0]
            w = torch.softmax((hf * mean_pool.unsqueeze(1)).sum(dim=2), dim=1)
            attn_pool = (hf * w.unsqueeze(2)).sum(dim=1)
            return self._finish([p.to(torch.bfloat16) for p in (mean_pool, max_pool, attn_pool)])

        h = self._mlp(self.cnn_to_attn, h)
        tok = torch.cat([self.cls_token, self.temporal_tokens], dim=1).to(torch.bfloat16).expand(B, -1, -1)
        h = torch.cat([tok, h], dim=1)                          # (B, S = T + 4, d)
        S = h.size(1)
        pos = self.pos_emb
        if S > pos.size(1):                                     # longer than built for: tile (layers.py:222-225)
            pos = pos.repeat(1, S // pos.size(1) + 1, 1)
        h = (h.float() + pos[:, :S]).to(torch.bfloat16)

        prev = None
        for i, layer in enumerate(self.attn_layers):
            a = _mha(layer['attn'], _layer_norm(h, layer['attn_norm']), None, True)
            h = (h.float() + self.dropout_light(a.float())).to(torch.bfloat16)
            saved = h
            f = layer['ffn'](_layer_norm(h, layer['ffn_norm']))
            h = (h.float() + self.dropout_medium(f.float())).to(torch.bfloat16)
            if i > 0:
                c = _mha(self.cross_scale_attn, h, prev, False)
                h = (h.float() + 0.1 * c.float()).to(torch.bfloat16)
            prev = saved

        hf = h.float()
        feat = (hf[:, 0] + 0.3 * hf[:, 1:4].mean(dim=1)).to(torch.bfloat16)
        return self._finish([feat, feat, feat])
