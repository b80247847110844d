"""GPU parity of the fused decoder path: the BART decoder + LM head +
cross-entropy on our kernels against the stock transformers forward in fp32 on the same weights
(third-party arithmetic: parity is pinned on the installed transformers).

Stated bounds (bf16 path vs fp32): |loss difference| <= 3e-2 at loss ~ 10.8, logits
max|a-b| / max|b| <= 5e-2, gradient cosine >= 0.99 per compared parameter."""
import pytest
import torch
import torch.nn.functional as F

from imagined_speech_translation_b200 import nn_ops
from imagined_speech_translation_b200.model import BARTDecoder

pytestmark = pytest.mark.gpu


def _cos(a, b):
    return F.cosine_similarity(a.double().flatten(), b.double().flatten(), dim=0).item()


@pytest.mark.parametrize("rows,V,K", [(64, 51271, 768), (37, 1000, 64), (5, 8, 16)])
def test_lm_head_cross_entropy(rows, V, K):
    g = torch.Generator(device="cuda").manual_seed(V)
    h = (torch.randn(rows, K, device="cuda", generator=g)).to(torch.bfloat16).requires_grad_(True)
    w = (torch.randn(V, K, device="cuda", generator=g) * 0.05).requires_grad_(True)
    b = torch.randn(V, device="cuda", generator=g) * 0.1
    labels = torch.randint(0, V, (rows,), device="cuda", generator=g)
    labels[::5] = -100
    loss, logits = nn_ops.lm_head_cross_entropy(h, w, b, labels)
    loss.backward()
    hr = h.detach().float().requires_grad_(True)
    wr = w.detach().to(torch.bfloat16).float().requires_grad_(True)
    ref_logits = hr @ wr.t() + b
    ref = F.cross_entropy(ref_logits, labels, ignore_index=-100)
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 5e-3 * max(1.0, abs(ref.item()))
    assert ((logits.float() - ref_logits).abs().max() / ref_logits.abs().max()).item() <= 1e-2
    assert _cos(h.grad, hr.grad) >= 0.999 and _cos(w.grad, wr.grad) >= 0.999
    assert abs(h.grad.float().norm().item() / hr.grad.norm().item() - 1) <= 2e-2


def test_fused_decoder_matches_transformers():
    torch.manual_seed(0)
    dec = BARTDecoder(768).cuda().train()
    for m in dec.modules():                      # parity runs: dropout off on both sides
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    for layer in dec.bart.model.decoder.layers:
        layer.dropout = 0.0
    dec.bart.model.decoder.dropout = 0.0
    B, L = 6, 16
    g = torch.Generator(device="cuda").manual_seed(1)
    feat = torch.randn(B, 768, device="cuda", generator=g)
    labels = torch.randint(1, 51271, (B, L), device="cuda", generator=g)
    labels[:, 12:] = -100
    ids = torch.cat([torch.full((B, 1), 101, device="cuda"), labels[:, :-1].clamp_min(0)], 1)

    def run(fused_path):
        dec.zero_grad(set_to_none=True)
        dec.fused_decoder = fused_path
        dec.autocast_dtype = torch.bfloat16 if fused_path else None        # reference arm: pure fp32
        f = feat.clone().requires_grad_(True)
        out = dec(f, decoder_input_ids=ids, labels=labels)
        out.loss.backward()
        names = ["eeg_to_bart.0.weight", "bart.model.shared.weight", "bart.model.decoder.layers.0.self_attn.q_proj.weight",
                 "bart.model.decoder.layers.5.fc1.weight", "bart.model.decoder.layers.2.encoder_attn.v_proj.weight",
                 "bart.model.decoder.layers.3.final_layer_norm.weight", "bart.model.decoder.layernorm_embedding.bias",
                 "bart.model.decoder.embed_positions.weight"]
        params = dict(dec.named_parameters())
        return out.loss.item(), out.logits.detach().float().clone(), f.grad.clone(), \
            {n: params[n].grad.detach().float().clone() for n in names}

    loss_f, logits_f, dfeat_f, grads_f = run(True)
    loss_r, logits_r, dfeat_r, grads_r = run(False)
    assert abs(loss_f - loss_r) <= 3e-2
    assert ((logits_f - logits_r).abs().max() / logits_r.abs().max()).item() <= 5e-2
    assert _cos(dfeat_f, dfeat_r) >= 0.99
    for n in grads_r:
        assert _cos(grads_f[n], grads_r[n]) >= 0.99, n
        assert abs(grads_f[n].norm().item() / grads_r[n].norm().item() - 1) <= 0.05, n
