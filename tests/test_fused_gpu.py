"""GPU parity of the fused glue kernels (csrc/fused_rowwise.cu, fused_bn.cu, attn_small.cu) against
plain PyTorch fp32 references of the same ops, forward and backward, through the C ABI.

Tolerances: outputs are bf16 (one ulp = 2^-8 relative), so element-wise results are compared with
max|a-b| / max|b| <= 1e-2 and parameter gradients (fp32 reductions of bf16 data) with <= 2e-2."""
import math

import pytest
import torch
import torch.nn.functional as F

from imagined_speech_translation_b200 import fused

pytestmark = pytest.mark.gpu
PAD = fused.PAD


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12)).item()


def bf(x):
    return x.to(torch.bfloat16)


@pytest.mark.parametrize("rows,C,gelu", [(37 * 5, 768, False), (64, 1536, True), (3, 768, True), (1030, 256, False),
                                         (100, 2048, True), (17, 64, False)])
def test_layer_norm(rows, C, gelu):
    g = torch.Generator(device="cuda").manual_seed(rows + C)
    x = bf(torch.randn(rows, C, device="cuda", generator=g) * 2 + 0.5)
    w = (1 + 0.1 * torch.randn(C, device="cuda", generator=g)).requires_grad_(True)
    b = (0.1 * torch.randn(C, device="cuda", generator=g)).requires_grad_(True)
    dy = bf(torch.randn(rows, C, device="cuda", generator=g))
    xo = x.clone().requires_grad_(True)
    y = fused.layer_norm(xo, w, b, 1e-5, gelu=gelu)
    y.backward(dy)
    xr = x.float().requires_grad_(True)
    wr, br = w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = F.layer_norm(xr, (C,), wr, br, 1e-5)
    if gelu:
        yr = F.gelu(yr)
    yr.backward(dy.float())
    assert rel(y, yr) <= 1e-2
    assert rel(xo.grad, xr.grad) <= 1e-2
    assert rel(w.grad, wr.grad) <= 2e-2
    assert rel(b.grad, br.grad) <= 2e-2


def test_elementwise_ops():
    g = torch.Generator(device="cuda").manual_seed(1)
    a = bf(torch.randn(1000, 768, device="cuda", generator=g)).requires_grad_(True)
    b = bf(torch.randn(1000, 768, device="cuda", generator=g)).requires_grad_(True)
    d = bf(torch.randn(1000, 768, device="cuda", generator=g))
    out = fused.add_dropout(a, b, scale=0.1)
    out.backward(d)
    assert rel(out, a.float() + 0.1 * b.float()) <= 1e-2
    assert rel(a.grad, d) == 0 and rel(b.grad, 0.1 * d.float()) <= 1e-2
    x = bf(torch.randn(999, 3072, device="cuda", generator=g) * 2).requires_grad_(True)
    d = bf(torch.randn(999, 3072, device="cuda", generator=g))
    y = fused.gelu_dropout(x)
    y.backward(d)
    xr = x.detach().float().requires_grad_(True)
    yr = F.gelu(xr)
    yr.backward(d.float())
    assert rel(y, yr) <= 1e-2 and rel(x.grad, xr.grad) <= 1e-2
    ag = bf(torch.randn(500, 2 * 1536, device="cuda", generator=g) * 2).requires_grad_(True)
    d = bf(torch.randn(500, 1536, device="cuda", generator=g))
    y = fused.glu(ag)
    y.backward(d)
    agr = ag.detach().float().requires_grad_(True)
    yr = F.gelu(agr[:, :1536]) * torch.sigmoid(agr[:, 1536:])
    yr.backward(d.float())
    assert rel(y, yr) <= 1e-2 and rel(ag.grad, agr.grad) <= 1e-2


@pytest.mark.parametrize("p", [0.1, 0.3])
def test_dropout_statistics_and_backward_mask(p):
    fused.set_seed(1234)
    n = 1 << 20
    a = torch.zeros(n, device="cuda", dtype=torch.bfloat16).requires_grad_(True)
    b = torch.ones(n, device="cuda", dtype=torch.bfloat16).requires_grad_(True)
    fused.begin_step()
    out = fused.add_dropout(a, b, p=p, training=True)
    keep = (out != 0).float().mean().item()
    assert abs(keep - (1 - p)) < 5e-3
    assert rel(out[out != 0], torch.full((1,), 1 / (1 - p), device="cuda")) <= 1e-2
    out.backward(torch.ones_like(out))
    assert torch.equal(b.grad != 0, out != 0)                 # backward regenerates the forward mask
    fused.begin_step()
    again = fused.add_dropout(a, b, p=p, training=True)
    assert torch.equal(again, out)                            # same (seed, step, site) -> same mask
    fused.advance_rng(a.device)
    fused.begin_step()
    other = fused.add_dropout(a, b, p=p, training=True)
    assert not torch.equal(other, out)                        # new step -> new mask
    eval_out = fused.add_dropout(a, b, p=p, training=False)
    assert torch.equal(eval_out, torch.ones_like(eval_out))


def _guard(x_btc):
    """(B, T, C) -> guarded rows tensor."""
    B, T, C = x_btc.shape
    xp = F.pad(x_btc, (0, 0, PAD, PAD)).reshape(B * (T + 2 * PAD), C)
    return F.pad(xp, (0, 0, PAD, PAD)).contiguous()


def _unguard(g, B, T):
    C = g.shape[1]
    return g[PAD:PAD + B * (T + 2 * PAD)].view(B, T + 2 * PAD, C)[:, PAD:PAD + T]


@pytest.mark.parametrize("B,T,C,res,train", [(4, 33, 128, 2, True), (3, 33, 768, 2, True), (5, 125, 256, 0, True),
                                             (4, 33, 384, 1, True), (4, 33, 256, 2, False), (64, 33, 512, 2, True)])
def test_bn_act(B, T, C, res, train):
    g = torch.Generator(device="cuda").manual_seed(B * 100 + C)
    ya = bf(torch.randn(B, T, C, device="cuda", generator=g) * 1.5 + 0.3)
    yr = bf(torch.randn(B, T, C, device="cuda", generator=g) * 0.7 - 0.2)
    dout = bf(torch.randn(B, T, C, device="cuda", generator=g))
    bn_a, bn_r = torch.nn.BatchNorm1d(C).cuda(), torch.nn.BatchNorm1d(C).cuda()
    ra, rr = torch.nn.BatchNorm1d(C).cuda(), torch.nn.BatchNorm1d(C).cuda()
    with torch.no_grad():
        for m in (bn_a, bn_r):
            m.weight.copy_(1 + 0.2 * torch.randn(C, device="cuda", generator=g))
            m.bias.copy_(0.2 * torch.randn(C, device="cuda", generator=g))
            m.running_mean.copy_(0.1 * torch.randn(C, device="cuda", generator=g))
            m.running_var.copy_(1 + 0.2 * torch.rand(C, device="cuda", generator=g))
        ra.load_state_dict(bn_a.state_dict()); rr.load_state_dict(bn_r.state_dict())
    for m in (bn_a, bn_r, ra, rr):
        m.train(train)
    # garbage in the padding rows of the raw conv outputs must not matter
    yag = _guard(ya); yag[yag.abs().sum(1) == 0] = 7.0
    yrg = _guard(yr); yrg[yrg.abs().sum(1) == 0] = -3.0
    yag.requires_grad_(True); yrg.requires_grad_(True)
    out = fused.bn_act(yag, bn_a, yrg if res else None, bn_r if res == 2 else None, B, T, training=train)
    dg = _guard(dout); dg[dg.abs().sum(1) == 0] = 5.0
    out.backward(dg)
    # reference on (B, C, T) fp32
    xa = ya.float().transpose(1, 2).requires_grad_(True)
    xr = yr.float().transpose(1, 2).requires_grad_(True)
    pre = ra(xa)
    if res == 2:
        pre = pre + rr(xr)
    elif res == 1:
        pre = pre + xr
    ref = F.gelu(pre)
    ref.backward(dout.float().transpose(1, 2))
    assert rel(_unguard(out, B, T), ref.transpose(1, 2)) <= 1e-2
    mask = torch.ones(out.shape[0], dtype=torch.bool, device="cuda")
    mask[PAD:PAD + B * (T + 2 * PAD)].view(B, -1)[:, PAD:PAD + T] = False
    assert float(out[mask].abs().max()) == 0.0                       # padding + guard rows are zero
    assert rel(_unguard(yag.grad, B, T), xa.grad.transpose(1, 2)) <= 1.5e-2
    assert float(yag.grad[mask].abs().max()) == 0.0
    assert rel(bn_a.weight.grad, ra.weight.grad) <= 2e-2 and rel(bn_a.bias.grad, ra.bias.grad) <= 2e-2
    if res == 2:
        assert rel(_unguard(yrg.grad, B, T), xr.grad.transpose(1, 2)) <= 1.5e-2
        assert rel(bn_r.weight.grad, rr.weight.grad) <= 2e-2
    if res == 1:
        assert rel(_unguard(yrg.grad, B, T), xr.grad.transpose(1, 2)) <= 1.5e-2
    if train:
        assert rel(bn_a.running_mean, ra.running_mean) <= 1e-3 and rel(bn_a.running_var, ra.running_var) <= 1e-3
        assert int(bn_a.num_batches_tracked) == 1


@pytest.mark.parametrize("B,T,C", [(4, 33, 256), (3, 125, 64)])
def test_dwconv5(B, T, C):
    g = torch.Generator(device="cuda").manual_seed(T)
    x = bf(torch.randn(B, T, C, device="cuda", generator=g))
    conv = torch.nn.Conv1d(C, C, 5, padding=2, groups=C).cuda()
    dout = bf(torch.randn(B, T, C, device="cuda", generator=g))
    xg = _guard(x).requires_grad_(True)
    w = conv.weight.detach().clone().requires_grad_(True)
    b = conv.bias.detach().clone().requires_grad_(True)
    out = fused.dwconv5(xg, w, b, B, T)
    out.backward(_guard(dout))
    xr = x.float().transpose(1, 2).requires_grad_(True)
    ref = conv(xr)
    ref.backward(dout.float().transpose(1, 2))
    assert rel(_unguard(out, B, T), ref.transpose(1, 2)) <= 1e-2
    assert rel(_unguard(xg.grad, B, T), xr.grad.transpose(1, 2)) <= 1e-2
    assert rel(w.grad, conv.weight.grad) <= 2e-2 and rel(b.grad, conv.bias.grad) <= 2e-2


def test_se_pieces_and_to_rows():
    B, T, C = 6, 33, 768
    g = torch.Generator(device="cuda").manual_seed(5)
    x = bf(torch.randn(B, T, C, device="cuda", generator=g))
    e = torch.rand(B, C, device="cuda", generator=g).requires_grad_(True)
    xg = _guard(x).requires_grad_(True)
    s = fused.group_mean(xg, B, T)
    out = fused.se_scale(xg, e, B, T)
    ds = torch.randn(B, C, device="cuda", generator=g)
    dout = bf(torch.randn(B * T, C, device="cuda", generator=g))
    (s * ds).sum().backward(retain_graph=True)
    g_mean = xg.grad.clone(); xg.grad = None
    out.backward(dout)
    xr = x.float().requires_grad_(True)
    er = e.detach().clone().requires_grad_(True)
    sr = xr.mean(1)
    outr = xr * er.unsqueeze(1)
    assert rel(s, sr) <= 1e-3 and rel(out.view(B, T, C), outr) <= 1e-2
    (sr * ds).sum().backward()
    assert rel(_unguard(g_mean, B, T), xr.grad) <= 1e-2
    xr.grad = None
    outr.backward(dout.float().view(B, T, C))
    assert rel(_unguard(xg.grad, B, T), xr.grad) <= 1e-2 and rel(e.grad, er.grad) <= 2e-2
    # (B, C, T) fp32 -> guarded rows, also from a batch-strided view
    big = torch.randn(B, 3 * C, T, device="cuda", generator=g)
    view = big[:, C:2 * C]
    rows = fused.to_rows(view)
    assert torch.equal(rows, _guard(bf(view.transpose(1, 2))))
    v2 = view.clone().requires_grad_(True)
    fused.to_rows(v2).backward(_guard(bf(torch.ones(B, T, C, device="cuda"))))
    assert torch.equal(v2.grad, torch.ones_like(v2))


def _ref_attn(q, k, v, H, causal):
    B, Sq, d = q.shape
    Sk, hd = k.shape[1], d // H
    qh = q.view(B, Sq, H, hd).transpose(1, 2)
    kh = k.view(B, Sk, H, hd).transpose(1, 2)
    vh = v.view(B, Sk, H, hd).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(hd)
    if causal:
        s = s.masked_fill(torch.ones(Sq, Sk, device=q.device).triu(1).bool(), float("-inf"))
    return (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B, Sq, d)


@pytest.mark.parametrize("B,S,H,hd,causal", [(3, 37, 8, 96, False), (2, 37, 4, 192, False), (5, 4, 12, 64, False),
                                             (2, 16, 12, 64, True), (2, 64, 4, 128, False), (1, 17, 2, 96, True)])
def test_attention_self(B, S, H, hd, causal):
    d = H * hd
    g = torch.Generator(device="cuda").manual_seed(S * H)
    qkv = bf(torch.randn(B * S, 3 * d, device="cuda", generator=g)).requires_grad_(True)
    do = bf(torch.randn(B * S, d, device="cuda", generator=g))
    o = fused.attn_self(qkv, B, S, H, causal=causal)
    o.backward(do)
    r = qkv.detach().float().view(B, S, 3 * d).requires_grad_(True)
    ref = _ref_attn(r[..., :d], r[..., d:2 * d], r[..., 2 * d:], H, causal)
    ref.backward(do.float().view(B, S, d))
    assert rel(o.view(B, S, d), ref) <= 1.5e-2
    assert rel(qkv.grad.view(B, S, 3 * d), r.grad) <= 2e-2


@pytest.mark.parametrize("B,Sq,Sk,H,hd", [(3, 37, 37, 4, 192), (2, 16, 6, 12, 64), (2, 5, 40, 8, 96)])
def test_attention_cross(B, Sq, Sk, H, hd):
    d = H * hd
    g = torch.Generator(device="cuda").manual_seed(Sq * Sk)
    q = bf(torch.randn(B * Sq, d, device="cuda", generator=g)).requires_grad_(True)
    kv = bf(torch.randn(B * Sk, 2 * d, device="cuda", generator=g)).requires_grad_(True)
    do = bf(torch.randn(B * Sq, d, device="cuda", generator=g))
    o = fused.attn_cross(q, kv, B, Sq, Sk, H)
    o.backward(do)
    qr = q.detach().float().view(B, Sq, d).requires_grad_(True)
    kvr = kv.detach().float().view(B, Sk, 2 * d).requires_grad_(True)
    ref = _ref_attn(qr, kvr[..., :d], kvr[..., d:], H, False)
    ref.backward(do.float().view(B, Sq, d))
    assert rel(o.view(B, Sq, d), ref) <= 1.5e-2
    assert rel(q.grad.view(B, Sq, d), qr.grad) <= 2e-2
    assert rel(kv.grad.view(B, Sk, 2 * d), kvr.grad) <= 2e-2


@pytest.mark.parametrize("B,S,H,hd", [(2, 65, 4, 192), (1, 1655, 8, 96), (2, 200, 4, 192), (1, 129, 12, 64),
                                      (1, 2052, 4, 192), (3, 100, 6, 128)])
def test_flash_attention_self(B, S, H, hd):
    """S > 64: the flash kernels (csrc/attn_flash.cu), ragged last tiles included (1655 = T + 4 of the real data)."""
    d = H * hd
    g = torch.Generator(device="cuda").manual_seed(S * H)
    qkv = bf(torch.randn(B * S, 3 * d, device="cuda", generator=g)).requires_grad_(True)
    do = bf(torch.randn(B * S, d, device="cuda", generator=g))
    o = fused.attn_self(qkv, B, S, H)
    o.backward(do)
    r = qkv.detach().float().view(B, S, 3 * d).requires_grad_(True)
    ref = _ref_attn(r[..., :d], r[..., d:2 * d], r[..., 2 * d:], H, False)
    ref.backward(do.float().view(B, S, d))
    assert rel(o.view(B, S, d), ref) <= 1.5e-2
    gq, gr = qkv.grad.view(B, S, 3 * d), r.grad
    for lo in (0, d, 2 * d):                     # dq, dk, dv separately: their scales differ by orders of magnitude
        assert rel(gq[..., lo:lo + d], gr[..., lo:lo + d]) <= 2e-2
    # bit-reproducible (one owner CTA per gradient element, fixed summation order)
    q2 = qkv.detach().clone().requires_grad_(True)
    o2 = fused.attn_self(q2, B, S, H)
    o2.backward(do)
    assert torch.equal(o, o2) and torch.equal(qkv.grad, q2.grad)


@pytest.mark.parametrize("B,Sq,Sk,H,hd", [(2, 300, 70, 4, 192), (1, 40, 1000, 8, 96), (2, 130, 131, 4, 64)])
def test_flash_attention_cross(B, Sq, Sk, H, hd):
    d = H * hd
    g = torch.Generator(device="cuda").manual_seed(Sq * Sk)
    q = bf(torch.randn(B * Sq, d, device="cuda", generator=g)).requires_grad_(True)
    kv = bf(torch.randn(B * Sk, 2 * d, device="cuda", generator=g)).requires_grad_(True)
    do = bf(torch.randn(B * Sq, d, device="cuda", generator=g))
    o = fused.attn_cross(q, kv, B, Sq, Sk, H)
    o.backward(do)
    qr = q.detach().float().view(B, Sq, d).requires_grad_(True)
    kvr = kv.detach().float().view(B, Sk, 2 * d).requires_grad_(True)
    ref = _ref_attn(qr, kvr[..., :d], kvr[..., d:], H, False)
    ref.backward(do.float().view(B, Sq, d))
    assert rel(o.view(B, Sq, d), ref) <= 1.5e-2
    assert rel(q.grad.view(B, Sq, d), qr.grad) <= 2e-2
    assert rel(kv.grad.view(B, Sk, 2 * d)[..., :d], kvr.grad[..., :d]) <= 2e-2
    assert rel(kv.grad.view(B, Sk, 2 * d)[..., d:], kvr.grad[..., d:]) <= 2e-2


@pytest.mark.parametrize("S", [37, 150])
def test_attention_dropout_is_consistent(S):
    """With dropout on, backward must use the forward mask: check dV against the dropped probabilities
    recovered from the forward output (v = identity-like trick).  S = 150 exercises the flash kernels."""
    B, H, hd = 2, 4, 64
    d = H * hd
    fused.set_seed(7)
    g = torch.Generator(device="cuda").manual_seed(3)
    qkv = bf(torch.randn(B * S, 3 * d, device="cuda", generator=g) * 0.3)
    outs = []
    for _ in range(2):
        fused.begin_step()
        x = qkv.clone().requires_grad_(True)
        o = fused.attn_self(x, B, S, H, p=0.25, training=True)
        o.float().square().sum().backward()
        outs.append((o.detach().clone(), x.grad.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    # finite-difference style check of the whole op along a random direction (mask fixed by the seed)
    fused.begin_step()
    x = qkv.clone().requires_grad_(True)
    o = fused.attn_self(x, B, S, H, p=0.25, training=True)
    w = torch.randn(B * S, d, device="cuda", generator=g)
    (o.float() * w).sum().backward()
    dirn = bf(torch.randn(B * S, 3 * d, device="cuda", generator=g))
    eps = 2.0 ** -6
    fs = []
    for sgn in (1.0, -1.0):
        fused.begin_step()
        o2 = fused.attn_self(bf(qkv.float() + sgn * eps * dirn.float()), B, S, H, p=0.25, training=True)
        fs.append((o2.float() * w).sum().item())
    fd = (fs[0] - fs[1]) / (2 * eps)
    an = (x.grad.float() * dirn.float()).sum().item()
    assert abs(fd - an) <= 0.08 * max(abs(fd), abs(an), 1.0)
    # keep rate of the probabilities: with v = 1 the output is sum_j p_ij m_ij, whose mean is 1
    fused.begin_step()
    ones = qkv.clone()
    ones[:, 2 * d:] = 1.0
    o3 = fused.attn_self(ones, B, S, H, p=0.25, training=True).float()
    assert abs(o3.mean().item() - 1.0) <= 0.02
