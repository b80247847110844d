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

from . import _lib, fused, nn_ops
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

    def forward(self, xg, B, T, p=0.0, training=True):
        """xg: guarded channels-last rows of (B, T, C); returns dropout_p(x * e) as compact (B*T, C) rows.
        s = mean_T(x); e = sigmoid(W2 relu(W1 s + b1) + b2)   (the two tiny Linears: M = B rows)"""
        s = fused.group_mean(xg, B, T)
        fc1, fc2 = self.excitation[0], self.excitation[2]
        z = torch.relu(nn_ops.linear(s.to(torch.bfloat16), fc1.weight, fc1.bias))
        e = torch.sigmoid(nn_ops.linear(z, fc2.weight, fc2.bias).float())
        return fused.se_scale(xg, e, B, T, p=p, training=training)


class FeedForwardNetwork(nn.Module):
    """Gated FFN, reference layers.py:301-317: W2 drop(gelu(W1 x) * sigmoid(Wg x))."""

    def __init__(self, input_dim, hidden_dim):
        super().__init__()
        self.linear1 = nn.Linear(input_dim, hidden_dim)
        self.linear2 = nn.Linear(hidden_dim, input_dim)
        self.gate = nn.Linear(input_dim, hidden_dim)
        self.dropout = nn.Dropout(0.1)

    def forward(self, x):
        ag = nn_ops.linear_cat(x, self.linear1.weight, self.linear1.bias, self.gate.weight, self.gate.bias)
        h = fused.glu(ag, p=self.dropout.p, training=self.dropout.training)
        return nn_ops.linear(h, self.linear2.weight, self.linear2.bias)


def _mha(mod: nn.MultiheadAttention, q_in, kv_in, self_attn: bool):
    """nn.MultiheadAttention(batch_first=True) semantics (packed in_proj, q scaled by
    1/sqrt(head_dim), softmax, dropout on the probabilities in train mode, out_proj); the
    head-averaged attention weights the reference discards are not produced.  Sequences up to 64
    tokens run the one-CTA-per-(batch, head) attention core, longer ones (the raw front end: S = T + 4 =
    1655 / 2052 / 4100) the flash kernel; either way the scores and probabilities stay on the SM."""
    B, Sq, d = q_in.shape
    H = mod.num_heads
    hd = d // H
    W, b = mod.in_proj_weight, mod.in_proj_bias
    Sk = Sq if self_attn else kv_in.shape[1]
    if fused.attn_supported(Sq, Sk, hd):
        if self_attn:
            qkv = nn_ops.linear(q_in.reshape(B * Sq, d), W, b)                   # (B*S, 3d)
            o = fused.attn_self(qkv, B, Sq, H, p=mod.dropout, training=mod.training)
        else:
            q = nn_ops.linear(q_in.reshape(B * Sq, d), W[:d], b[:d])
            kv = nn_ops.linear(kv_in.reshape(B * Sk, d), W[d:], b[d:])
            o = fused.attn_cross(q, kv, B, Sq, Sk, H, p=mod.dropout, training=mod.training)
        return nn_ops.linear(o, mod.out_proj.weight, mod.out_proj.bias).view(B, Sq, d)
    raise _lib.EegxError(f"attention head_dim {hd} is not one the fused attention kernels serve "
                         f"{fused.ATTN_HEAD_DIMS}; there is no library fallback")


def _layer_norm(x, ln: nn.LayerNorm):
    return fused.layer_norm(x, ln.weight, ln.bias, ln.eps)


def run_sequential(seq: nn.Sequential, x):
    """A Sequential of Linear / LayerNorm / GELU / Sigmoid / Dropout applied to bf16 rows: the
    contractions run on the tcgen05 GEMM and the runs LayerNorm -> GELU -> Dropout / GELU -> Dropout
    collapse into one fused kernel each (the containers only hold the parameters)."""
    mods = list(seq)
    i = 0
    while i < len(mods):
        m = mods[i]
        nxt = mods[i + 1] if i + 1 < len(mods) else None
        nxt2 = mods[i + 2] if i + 2 < len(mods) else None
        if isinstance(m, nn.Linear):
            x = nn_ops.linear(x, m.weight, m.bias)
            i += 1
        elif isinstance(m, nn.LayerNorm):
            gelu = isinstance(nxt, nn.GELU)
            drop = nxt2 if gelu and isinstance(nxt2, nn.Dropout) else (nxt if isinstance(nxt, nn.Dropout) else None)
            x = fused.layer_norm(x, m.weight, m.bias, m.eps, gelu=gelu, p=drop.p if drop is not None else 0.0,
                                 training=drop.training if drop is not None else False)
            i += 1 + int(gelu) + int(drop is not None)
        elif isinstance(m, nn.GELU):
            drop = nxt if isinstance(nxt, nn.Dropout) else None
            x = fused.gelu_dropout(x, p=drop.p if drop is not None else 0.0,
                                   training=drop.training if drop is not None else False)
            i += 1 + int(drop is not None)
        elif isinstance(m, nn.Sigmoid):
            x = torch.sigmoid(x.float()).to(torch.bfloat16)
            i += 1
        elif isinstance(m, nn.Dropout):
            x = m(x)
            i += 1
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
    def _res_block(self, hg, conv, bn, res, drop, B, T, cin_pad=0):
        """gelu(bn(conv(h)) + res(h)) -> dropout -> zeroed padding rows, on guarded channels-last rows
        (reference layers.py:142-174): two GEMMs, two statistics reductions, one fused apply.
        cin_pad: zero input channels appended to the block's input (first block only, see _cnn); the weights get
        matching zero columns through F.pad, so their gradients flow back to the unpadded parameters."""
        M = B * (T + 2 * PAD)
        w = F.pad(conv.weight, (0, 0, 0, cin_pad)) if cin_pad else conv.weight
        ya = nn_ops.conv_g(hg, w, conv.bias, M, bias_grad=not bn.training)
        p = drop.p if drop is not None else 0.0
        if isinstance(res, nn.Identity):
            return fused.bn_act(ya, bn, hg, None, B, T, p=p, training=bn.training, drop_training=self.training)
        wr = F.pad(res[0].weight, (0, 0, 0, cin_pad)) if cin_pad else res[0].weight
        yr = nn_ops.conv_g(hg, wr, None, M)
        return fused.bn_act(ya, bn, yr, res[1], B, T, p=p, training=bn.training, drop_training=self.training)

    def _depthwise_block(self, hg, B, T):
        # depthwise k5 (groups = channels) -> pointwise 1x1 -> BN -> GELU   (layers.py:157-161)
        M = B * (T + 2 * PAD)
        d = fused.dwconv5(hg, self.depthwise_conv.weight, self.depthwise_conv.bias, B, T)
        pw, bn = self.pointwise_conv, self.bn_depth
        y = nn_ops.conv_g(d, pw.weight, pw.bias, M, bias_grad=not bn.training)
        return fused.bn_act(y, bn, None, None, B, T, p=self.dropout_medium.p, training=bn.training,
                            drop_training=self.training)

    def _cnn(self, x):
        B, C, T = x.shape
        # the TMA row pitch of the channels-last rows is 16 bytes = 8 bf16 channels: a region whose channel count is
        # not a multiple of 8 (the reference montage gives 16 / 9 / 11 / 12) gets zero channels appended, with zero
        # weight columns against them -- the products are exact zeros, the result is unchanged
        cin_pad = (-C) % 8
        if cin_pad:
            x = F.pad(x, (0, 0, 0, cin_pad))
        hg = fused.to_rows(x.float())                           # guarded channels-last bf16 rows
        hg = self._res_block(hg, self.conv1, self.bn1, self.residual1, self.dropout_light, B, T, cin_pad=cin_pad)
        hg = self._res_block(hg, self.conv2, self.bn2, self.residual2, self.dropout_light, B, T)
        hg = self._depthwise_block(hg, B, T)
        hg = self._res_block(hg, self.conv3, self.bn3, self.residual3, self.dropout_medium, B, T)
        hg = self._res_block(hg, self.conv4, self.bn4, self.residual4, None, B, T)
        # SE re-scaling + dropout_heavy, compact (B*T, 768) rows out
        h = self.se_block(hg, B, T, p=self.dropout_heavy.p, training=self.training)
        return h.view(B, T, -1)

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
        h = (h + pos[:, :S]).to(torch.bfloat16)                 # bf16 + fp32 -> fp32 add, one rounding
        h = fused.grad_boundary(h, ('region', id(self)))        # attention stack + heads: gradients final here

        prev = None
        mid = len(self.attn_layers) // 2
        for i, layer in enumerate(self.attn_layers):
            if i == mid and i > 0:                              # layers mid.. and the heads: gradients final here
                h = fused.grad_boundary(h, ('region_mid', id(self)))
            a = _mha(layer['attn'], _layer_norm(h, layer['attn_norm']), None, True)
            h = fused.add_dropout(h, a, p=self.dropout_light.p, training=self.training)
            saved = h
            f = layer['ffn'](_layer_norm(h, layer['ffn_norm']))
            h = fused.add_dropout(h, f, p=self.dropout_medium.p, training=self.training)
            if i > 0:
                c = _mha(self.cross_scale_attn, h, prev, False)
                h = fused.add_dropout(h, c, scale=0.1)
            prev = saved

        h4 = h[:, :4].float()
        feat = (h4[:, 0] + 0.3 * h4[:, 1:4].mean(dim=1)).to(torch.bfloat16)
        return self._finish([feat, feat, feat])
