"""Deterministic parameter / input recipe shared by the golden generator (which fills the
REFERENCE modules) and the tests (which fill ours): same names, same shapes, same values,
independent of construction order.  CPU generator, float32."""
import math

import torch


def fill_params(module, seed=0):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if p.dim() >= 2:
                fan_in = max(p[0].numel(), 1)
                val = torch.randn(p.shape, generator=g) / math.sqrt(fan_in)
            elif name.endswith("weight"):            # BatchNorm / LayerNorm gains
                val = 1.0 + 0.1 * torch.randn(p.shape, generator=g)
            elif name.endswith("bias"):
                val = 0.1 * torch.randn(p.shape, generator=g)
            else:                                    # region_importance
                val = 0.5 * torch.randn(p.shape, generator=g)
            p.copy_(val.to(p.dtype))
    return module


def make_input(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g)


def zero_dropout(module):
    """Parity runs keep .train() (BatchNorm batch statistics) with every dropout disabled
    ('Dropout/RNG parity')."""
    for m in module.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
        if isinstance(m, torch.nn.TransformerEncoderLayer):
            m.dropout.p = m.dropout1.p = m.dropout2.p = 0.0
            m.self_attn.dropout = 0.0
    return module


# ---------------------------------------------------------------------------- composed model / train step
# BART architecture of fnlp/bart-base-chinese (SURVEY.md 8(c) shim 1: the checkpoint is not cached, the config
# reproduces the author's parameter count exactly), with every dropout at 0 for parity runs.
BART_SHAPE = dict(vocab_size=51271, d_model=768, encoder_layers=6, decoder_layers=6, encoder_attention_heads=12,
                  decoder_attention_heads=12, encoder_ffn_dim=3072, decoder_ffn_dim=3072, max_position_embeddings=1024,
                  pad_token_id=0, bos_token_id=101, eos_token_id=102, decoder_start_token_id=101)


def zero_bart_dropout(bart):
    """transformers' BART copies its dropout rates into module attributes at construction: zero those."""
    for name in ("dropout", "attention_dropout", "activation_dropout", "classifier_dropout"):
        if hasattr(bart.config, name):
            setattr(bart.config, name, 0.0)
    for m in bart.modules():
        for attr in ("dropout", "activation_dropout"):
            if isinstance(getattr(m, attr, None), float):
                setattr(m, attr, 0.0)
    return bart


def train_batches(n, B, counts, T, seed):
    """Seeded synthetic micro-batches in the reference loader's format: unit-scale region
    features, labels in [1, V) with a ragged -100 tail, decoder_input_ids = [bos] + labels[:-1]."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        labels = torch.randint(1, 51271, (B, 16), generator=g)
        tail = torch.randint(0, 8, (B,), generator=g)
        for b in range(B):
            if tail[b] > 0:
                labels[b, 16 - int(tail[b]):] = -100
        ids = torch.cat([torch.full((B, 1), 101), labels[:, :-1].clamp_min(0)], dim=1)
        out.append({"eeg": [torch.randn(B, counts[k], T, generator=g) for k in ("frontal", "temporal", "central", "parietal")],
                    "decoder_input_ids": ids, "labels": labels})
    return out


SENTINEL = 7.0


def fill_sentinel(module):
    with torch.no_grad():
        for p in module.parameters():
            p.fill_(SENTINEL)
    return module


def classify_init(p):
    """What an init routine did to a sentinel-filled parameter: untouched / ones / zeros /
    xavier_uniform(gain 0.02) / normal(std 0.02)."""
    p = p.detach().float()
    kind = "other"
    if bool((p == SENTINEL).all()):
        kind = "untouched"
    elif bool((p == 1).all()):
        kind = "ones"
    elif bool((p == 0).all()):
        kind = "zeros"
    else:
        if p.dim() >= 2:
            rf = p[0][0].numel() if p.dim() > 2 else 1
            bound = 0.02 * math.sqrt(6.0 / (p.shape[1] * rf + p.shape[0] * rf))
            if float(p.abs().max()) <= bound * (1 + 1e-5) and float(p.abs().max()) >= 0.5 * bound:
                kind = "xavier_uniform_gain0.02"
        if kind == "other" and 0.012 < float(p.std()) < 0.028 and abs(float(p.mean())) < 0.01:
            kind = "normal_std0.02"
    return {"kind": kind, "shape": list(p.shape)}
