"""Multi-region EEG encoder on the B200 path.

Drop-in for the reference ``BrainRegionEncoder`` (``main_model/src/models/brain_encoder.py:11-193``):
same constructor, parameter names / shapes and forward semantics; the four region encoders are
``layers.Conv1DWithAttention`` and every Linear / Conv1d contraction of the fusion stage runs on
the tcgen05 GEMM.  ``forward(list[4] of (B, C_r, T) CUDA tensors) -> (B, hidden_dim)`` float32.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import fused, grouped, nn_ops
from .layers import Conv1DWithAttention, _layer_norm, _mha, run_sequential
from .nn_ops import PAD

REGION_NAMES = ['frontal', 'temporal', 'central', 'parietal']
_REGION_STREAMS = int(os.environ.get("EEGX_REGION_STREAMS", "4"))   # concurrent region encoders (1..4)


class BrainRegionEncoder(nn.Module):
    def __init__(self, n_timepoints, region_channel_counts, hidden_dim=768,
                 disable_cross_region_attn=False, uniform_region_weight=False, cnn_only=False):
        super().__init__()
        self.region_names = list(REGION_NAMES)
        self.region_channel_counts = region_channel_counts
        self.disable_cross_region_attn = disable_cross_region_attn
        self.uniform_region_weight = uniform_region_weight
        self.n_regions = len(self.region_names)
        self.hidden_dim = hidden_dim
        d = hidden_dim

        self.region_embeddings = nn.Embedding(self.n_regions, d)
        nn.init.normal_(self.region_embeddings.weight, std=0.02)
        self.temporal_scales = nn.ModuleList(
            [nn.Conv1d(d, d, kernel_size=k, padding=k // 2) for k in (3, 7, 15, 31)])
        self.diversity_projection = nn.Sequential(
            nn.Linear(d * 4, d * 2), nn.GELU(), nn.Dropout(0.1), nn.Linear(d * 2, d), nn.LayerNorm(d))
        if not uniform_region_weight:
            self.region_importance = nn.Parameter(torch.randn(self.n_regions) * 0.5)
            self.region_gate = nn.Sequential(
                nn.Linear(d, d // 2), nn.GELU(), nn.Dropout(0.1), nn.Linear(d // 2, self.n_regions), nn.Sigmoid())
        self.region_encoders = nn.ModuleDict({
            name: Conv1DWithAttention(region_channel_counts[name], n_timepoints, d, cnn_only=cnn_only)
            for name in self.region_names})
        if not disable_cross_region_attn:
            layer = nn.TransformerEncoderLayer(d_model=d, nhead=12, dim_feedforward=d * 4, dropout=0.1,
                                               activation='gelu', batch_first=True, norm_first=True)
            self.fusion_transformer = nn.TransformerEncoder(layer, num_layers=2, enable_nested_tensor=False)
            self.cross_region_attention = nn.MultiheadAttention(embed_dim=d, num_heads=8, dropout=0.1,
                                                                batch_first=True)
        self.feature_enhancer = nn.Sequential(
            nn.Linear(d, d * 2), nn.GELU(), nn.Dropout(0.1), nn.Linear(d * 2, d), nn.LayerNorm(d))

    # -- brain_encoder.py:94-113: four Conv1d over the length-4 region axis, GELU, mean, project
    def apply_multi_scale_processing(self, x):
        B, R, d = x.shape                                       # (B, 4, d) bf16, channels-last already
        M = B * (R + 2 * PAD)
        buf = nn_ops.guard_pad(x)
        valid = torch.zeros(R + 2 * PAD, device=x.device)
        valid[PAD:PAD + R] = 1.0
        feats = []
        for conv in self.temporal_scales:
            k = conv.kernel_size[0]
            w = conv.weight
            if k > 2 * R - 1:       # only offsets -(R-1)..(R-1) can ever meet data: drop the dead taps
                c = k // 2
                w = w[:, :, c - (R - 1): c + R]
            y = nn_ops.conv1d_cl(buf, w, conv.bias, M).float().view(B, R + 2 * PAD, d)
            y = F.gelu(y) * valid.view(1, -1, 1)
            feats.append(y.sum(dim=1) / R)
        ms = torch.stack(feats, dim=1).reshape(B, 4 * d).to(torch.bfloat16)
        ms = run_sequential(self.diversity_projection, ms)
        return ms.unsqueeze(1).expand(-1, R, -1)

    def compute_dynamic_region_weights(self, x):
        pooled = x.float().mean(dim=1).to(torch.bfloat16)
        dyn = run_sequential(self.region_gate, pooled).float()
        if hasattr(self, 'region_importance'):
            static = F.softmax(self.region_importance, dim=0)
            return F.softmax(0.7 * static.unsqueeze(0) + 0.3 * dyn, dim=1)
        return F.softmax(dyn, dim=1)

    def _fusion_layer(self, layer: nn.TransformerEncoderLayer, x):
        # norm_first: x + drop(MHA(LN1 x)); x + drop(W2 drop(gelu(W1 LN2 x)))
        tr = layer.training
        a = _mha(layer.self_attn, _layer_norm(x, layer.norm1), None, True)
        x = fused.add_dropout(x, a, p=layer.dropout1.p, training=tr)
        h = nn_ops.linear(_layer_norm(x, layer.norm2), layer.linear1.weight, layer.linear1.bias)
        h = fused.gelu_dropout(h, p=layer.dropout.p, training=tr)
        h = nn_ops.linear(h, layer.linear2.weight, layer.linear2.bias)
        return fused.add_dropout(x, h, p=layer.dropout2.p, training=tr)

    def parameter_stacks(self):
        """Lists of same-named, same-shaped parameters of the region encoders (one per region): laid out back to back
        by ``FlatAdamW(stacks=...)`` they give the lock-step path (``grouped.py``) its stacked views."""
        mods = [self.region_encoders[n] for n in self.region_names]
        tables = [dict(m.named_parameters()) for m in mods]
        out = []
        for name, p in tables[0].items():
            group = [t.get(name) for t in tables]
            if all(q is not None and q.shape == p.shape for q in group):
                out.append(group)
        return out

    def _lock_step(self, eeg_data):
        """Can the regions run as ONE stacked pass (grouped.py)?  Needs the ParamStack views of FlatAdamW, equal
        input shapes and a sequence the fused attention core serves."""
        mods = [self.region_encoders[n] for n in self.region_names]
        if not getattr(self, "lock_step_regions", True) or not grouped.available(mods):
            return None
        shp = eeg_data[0].shape
        if any(x.shape != shp or not x.is_cuda for x in eeg_data) or shp[1] % 8 != 0:
            return None
        S = shp[2] + 4
        for layer in list(mods[0].attn_layers) + [{'attn': mods[0].cross_scale_attn}]:
            if not fused.attn_supported(S, S, self.hidden_dim // layer['attn'].num_heads):
                return None
        return mods

    def _region_features(self, eeg_data):
        """Lock step when possible: every op of the four region encoders is one launch (grouped.py).  Otherwise
        the four encoders are independent until the stack: run them on four streams
        (fork / join by events; also legal under CUDA-graph capture), so their many small kernels
        overlap instead of queueing behind each other (reference: sequential loop,
        brain_encoder.py:148-150)."""
        mods = self._lock_step(eeg_data)
        if mods is not None and grouped.MODE >= 2:
            return grouped.forward(mods, list(eeg_data))
        if not getattr(self, "parallel_regions", True):
            return [self.region_encoders[n](eeg_data[i]) for i, n in enumerate(self.region_names)]
        cur = torch.cuda.current_stream()
        n_streams = max(1, min(len(self.region_names), int(getattr(self, "region_streams", _REGION_STREAMS))))
        if getattr(self, "_streams", None) is None or len(self._streams) != n_streams or \
                self._streams[0].device != eeg_data[0].device:
            self._streams = [torch.cuda.Stream(device=eeg_data[0].device) for _ in range(n_streams)]
        start = cur.record_event()
        feats = []
        for s in self._streams:
            s.wait_event(start)
        for i, name in enumerate(self.region_names):
            s = self._streams[i % n_streams]
            with torch.cuda.stream(s):
                eeg_data[i].record_stream(s)
                f = self.region_encoders[name]._cnn(eeg_data[i]) if mods is not None \
                    else self.region_encoders[name](eeg_data[i])
            f.record_stream(cur)
            feats.append(f)
        for s in self._streams:
            cur.wait_event(s.record_event())
        if mods is not None:        # grouped.MODE == 1: CNN stacks per region (above), attention stacks + heads in lock step
            B, T = eeg_data[0].shape[0], eeg_data[0].shape[2]
            return grouped.attention_and_heads(mods, torch.cat([f.reshape(B * T, -1) for f in feats], dim=0), B, T)
        return feats

    def forward(self, eeg_data):
        feats = self._region_features(eeg_data)
        # (B, 4, d) fp32
        x = feats.transpose(0, 1).contiguous() if torch.is_tensor(feats) else torch.stack(feats, dim=1)
        x = fused.grad_boundary(x, ('fusion', id(self)))         # everything after the region encoders
        ms = self.apply_multi_scale_processing(x.to(torch.bfloat16))
        x = x + 0.3 * ms.float()
        x = x + 0.4 * self.region_embeddings.weight.unsqueeze(0)
        if not self.disable_cross_region_attn:
            xt = x.to(torch.bfloat16)
            for layer in self.fusion_transformer.layers:
                xt = self._fusion_layer(layer, xt)
            xc = _mha(self.cross_region_attention, xt, None, True)
            gate = torch.sigmoid(run_sequential(self.feature_enhancer,
                                                xt.float().mean(dim=1).to(torch.bfloat16)).float()).unsqueeze(1)
            x = xt.float() + gate * xc.float()
        if self.uniform_region_weight or not hasattr(self, 'region_importance'):
            pooled = x.mean(dim=1)
        else:
            w = self.compute_dynamic_region_weights(x)
            pooled = (x * w.unsqueeze(-1)).sum(dim=1)
        enhanced = run_sequential(self.feature_enhancer, pooled.to(torch.bfloat16)).float()
        return pooled + 0.3 * enhanced

    def get_region_weights(self):
        """Same report as the reference (brain_encoder.py:195-214)."""
        if hasattr(self, 'region_importance') and not self.uniform_region_weight:
            return {'names': self.region_names,
                    'softmax': F.softmax(self.region_importance, dim=0).data.cpu().numpy(),
                    'has_dynamic': hasattr(self, 'region_gate')}
        return {'names': self.region_names, 'softmax': [1.0 / self.n_regions] * self.n_regions,
                'has_dynamic': False}
