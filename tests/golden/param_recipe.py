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
