"""Lock-step execution of the region encoders.

The reference runs its four ``Conv1DWithAttention`` modules one after the other
(``main_model/src/models/brain_encoder.py:148-150``).  Their layers have identical shapes and their own
parameters, so here every op of the forward AND backward pass is ONE launch for all G modules: the
activations of the G regions are stacked along the batch (G * B trials in one buffer, block g = region g),
the parameters are read -- and their gradients accumulated -- through ``optim.ParamStack`` views of the flat
fp32 / bf16-shadow / gradient buffers ``FlatAdamW`` lays out back to back, the GEMMs run as grouped launches
(``ops.gemm(..., grouped=True)``: 4 weight sets, 4x the rows per launch), and the row-wise kernels take the
parameter group from the row index.  Against the per-region path (4 CUDA streams of ~800 small kernels each)
this quarters the launch count and quadruples the work per launch.

Semantics are those of ``layers.Conv1DWithAttention.forward`` (reference ``layers.py:129-272``); the per-region
path stays the fallback whenever the stacked views do not exist (before the first optimizer step, plain
autograd parameters, unequal channel counts, ``cnn_only``) -- ``available()`` decides.

Autograd: activations flow through ``torch.autograd.Function``s as usual; parameters do NOT -- each function
gets one real leaf parameter as an ``anchor`` input (so the node runs in backward even when the activation
input needs no gradient) and adds its parameter gradients straight into the stacked gradient views.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, fused, nn_ops, ops
from .fused import PAD

MODE = int(os.environ.get("EEGX_GROUPED", "2"))      # 0: off, 1: attention stack + heads only, 2: CNN stack too


# ------------------------------------------------------------------------------------------ parameter access
def _stack(p):
    ent = getattr(p, "_eegx_stack", None)
    return None if ent is None else ent[0]


def available(mods: List[nn.Module]) -> bool:
    """True when every parameter of the G modules lives in a ParamStack (same position in every module, still bound
    to the flat buffers) and the shape is one the fused kernels serve."""
    if MODE == 0 or len(mods) < 2:
        return False
    m0 = mods[0]
    if getattr(m0, "cnn_only", False):
        return False
    cache = getattr(m0, "_eegx_group_ok", None)
    names = [n for n, _ in m0.named_parameters()]
    stacks = []
    for n, p in m0.named_parameters():
        st = _stack(p)
        if st is None or len(st) != len(mods) or p._eegx_stack[1] != 0:
            return False
        stacks.append(st)
    key = tuple(id(s) for s in stacks)
    if cache is not None and cache[0] == key and cache[1] == tuple(id(m) for m in mods):
        ok = True
    else:
        ok = True
        for i, m in enumerate(mods):
            params = dict(m.named_parameters())
            if list(params) != names:
                return False
            for n, st in zip(names, stacks):
                if st.params[i] is not params[n]:
                    return False
        m0._eegx_group_ok = (key, tuple(id(m) for m in mods))
    return ok and all(st.bound() for st in stacks)


class LinW:
    """Grouped Linear weights: bf16 (G, N, K) operand view, its fp32 gradient view, bias (G, N) + gradient view;
    optionally a row range of the stacked parameter (the q / kv halves of ``in_proj_weight``)."""
    __slots__ = ("wst", "bst", "rows", "anchor")

    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, rows: Optional[slice] = None):
        self.wst, self.bst, self.rows, self.anchor = _stack(weight), (_stack(bias) if bias is not None else None), rows, weight

    def w16(self):
        w = self.wst.w16()
        return w if self.rows is None else w[:, self.rows]

    def wg(self):
        g = self.wst.g
        return g if self.rows is None else g[:, self.rows]

    def b(self):
        if self.bst is None:
            return None
        b = self.bst.p
        return b if self.rows is None else b[:, self.rows]

    def bg(self):
        g = self.bst.g
        return g if self.rows is None else g[:, self.rows]


# ------------------------------------------------------------------------------------------ reductions
def gcolsum(y3: torch.Tensor, into: torch.Tensor, accumulate: bool = True) -> None:
    """into[g, c] (+)= sum_r y3[g, r, c]; y3 (G, R, C) bf16 with unit column stride, into (G, C) fp32 view."""
    G, R, C_ = y3.shape
    lib = _lib.lib()
    ws = fused._workspace(lib.eegx_colreduce_workspace_bytes(C_), y3.device)
    _lib.check(lib.eegx_colsum_bf16(_lib.ptr(y3), y3.stride(1), G, y3.stride(0), R, C_, _lib.ptr(into), into.stride(0),
                                    int(accumulate), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "eegx_colsum_bf16")


def _split_for(tiles: int, R: int) -> int:
    s = 1
    while tiles * s * 2 <= 160 and R % (s * 2) == 0 and (R // (s * 2)) % 8 == 0 and R // (s * 2) >= 512:
        s *= 2
    return s


def _wgrad_partials(dy3: torch.Tensor, x3: torch.Tensor):
    """dy3 (G, R, N), x3 (G, R, K) bf16 -> fp32 partials (s, G, N, K) of dW[g] = dy[g]^T x[g], the reduction over R
    split into s chunks so that s * G * tiles fills the SMs (fixed order: bit-stable)."""
    G, R, N = dy3.shape
    K = x3.shape[2]
    bn = 64 if K <= 64 else (256 if K > 128 and (K % 256 == 0 or K >= 1024) else 128)
    s = _split_for(G * -(-N // 128) * -(-K // bn), R)
    part = torch.empty(s, G, N, K, dtype=torch.float32, device=dy3.device)
    ch = R // s
    dy4 = dy3.as_strided((G, s, ch, N), (dy3.stride(0), ch * dy3.stride(1), dy3.stride(1), 1), dy3.storage_offset())
    x4 = x3.as_strided((G, s, ch, K), (x3.stride(0), ch * x3.stride(1), x3.stride(1), 1), x3.storage_offset())
    ops.gemm(dy4, x4, a_mn_major=True, b_mn_major=True, out=part.permute(1, 0, 2, 3))
    return part


def gwgrad(dy3: torch.Tensor, x3: torch.Tensor, into: torch.Tensor) -> None:
    """into[g] += dy3[g]^T x3[g]; into: (G, N, K) fp32 view (rows contiguous, uniform group stride)."""
    G, R, N = dy3.shape
    K = x3.shape[2]
    bn = 64 if K <= 64 else (256 if K > 128 and (K % 256 == 0 or K >= 1024) else 128)
    if _split_for(G * -(-N // 128) * -(-K // bn), R) == 1:
        ops.gemm(dy3, x3, a_mn_major=True, b_mn_major=True, grouped=True, out=into, accumulate=True)
        return
    part = _wgrad_partials(dy3, x3)
    if into.stride(1) != K or into.stride(2) != 1:
        raise ValueError("gwgrad: gradient rows must be contiguous")
    _lib.check(_lib.lib().eegx_accumulate_partials_f32(_lib.ptr(part), part.shape[0], G, N * K, _lib.ptr(into),
                                                       into.stride(0), 1, _lib.stream_ptr()),
               "eegx_accumulate_partials_f32")


# ------------------------------------------------------------------------------------------ Linear
class _GLinear(torch.autograd.Function):
    """y[g] = x[g] W[g]^T + b[g] on (G*M, K) bf16 rows (block g = module g)."""

    @staticmethod
    def forward(ctx, x, anchor, lw: LinW, G):
        M, K = x.shape[0] // G, x.shape[1]
        w16 = lw.w16()
        y = ops.gemm(x.view(G, M, K), w16, lw.b(), grouped=True)
        ctx.save_for_backward(x)
        ctx.cfg = (lw, G)
        return y.view(G * M, w16.shape[1])

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        lw, G = ctx.cfg
        M, K = x.shape[0] // G, x.shape[1]
        dy = dy.contiguous()
        N = dy.shape[1]
        dy3 = dy.view(G, M, N)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.gemm(dy3, lw.w16(), b_mn_major=True, grouped=True).view(G * M, K)

        def grads():
            gwgrad(dy3, x.view(G, M, K), lw.wg())
            if lw.bst is not None:
                gcolsum(dy3, lw.bg())
        fused.deferred(grads, dy, x)
        return dx, None, None, None


def glinear(x: torch.Tensor, lw: LinW, G: int) -> torch.Tensor:
    return _GLinear.apply(x.contiguous(), lw.anchor, lw, G)


class _GLinearCat(torch.autograd.Function):
    """[y1 | y2] = x [W1; W2]^T + [b1 | b2] per group in one GEMM (the gated FFN's two up-projections)."""

    @staticmethod
    def forward(ctx, x, anchor, lws, G):
        M, K = x.shape[0] // G, x.shape[1]
        w = _cat_pack(lws)
        b = torch.cat([lw.b() for lw in lws], dim=1)
        y = ops.gemm(x.view(G, M, K), w, b, grouped=True)
        ctx.save_for_backward(x)
        ctx.cfg = (lws, G)
        return y.view(G * M, w.shape[1])

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        lws, G = ctx.cfg
        M, K = x.shape[0] // G, x.shape[1]
        dy = dy.contiguous()
        dy3 = dy.view(G, M, dy.shape[1])
        dx = ops.gemm(dy3, _cat_pack(lws), b_mn_major=True, grouped=True).view(G * M, K) if ctx.needs_input_grad[0] else None
        def grads():
            x3, off = x.view(G, M, K), 0
            for lw in lws:
                n = lw.wst.p.shape[1]
                sl = dy3[:, :, off:off + n]
                gwgrad(sl, x3, lw.wg())
                gcolsum(sl, lw.bg())
                off += n
        fused.deferred(grads, dy, x)
        return dx, None, None, None


def _stack_pack(st, tag, make):
    """Derived bf16 pack of a ParamStack (rebuilt after every optimizer step: nn_ops.clear_pack_cache)."""
    key = ("stack", id(st), tag)
    ver = tuple(q._version for q in st.params)
    hit = nn_ops._pack_cache.get(key)
    if hit is not None and hit[0] == ver and hit[2] is st:
        return hit[1]
    with torch.no_grad():
        val = make(st.w16())
    nn_ops._pack_cache[key] = (ver, val, st)
    return val


def _cat_pack(lws):
    tag = "cat:" + ",".join(str(id(lw.wst)) for lw in lws[1:])
    return _stack_pack(lws[0].wst, tag, lambda _: torch.cat([lw.w16() for lw in lws], dim=1).contiguous())


def glinear_cat(x: torch.Tensor, lws, G: int) -> torch.Tensor:
    return _GLinearCat.apply(x.contiguous(), lws[0].anchor, tuple(lws), G)


# ------------------------------------------------------------------------------------------ LayerNorm
class _GLayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, anchor, wst, bst, G, eps, act, rng, site, p):
        C_ = x.shape[-1]
        rows = x.numel() // C_
        y = torch.empty_like(x)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        _lib.check(_lib.lib().eegx_layernorm_fwd_bf16(
            _lib.ptr(x), _lib.ptr(wst.p), _lib.ptr(bst.p), _lib.ptr(y), _lib.ptr(mean), _lib.ptr(rstd), rows, C_, G,
            wst.p.stride(0), float(eps), int(act), _lib.ptr(rng), site, p, _lib.stream_ptr()), "eegx_layernorm_fwd_bf16")
        ctx.save_for_backward(x, mean, rstd)
        ctx.cfg = (wst, bst, G, int(act), rng, site, p)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd = ctx.saved_tensors
        wst, bst, G, act, rng, site, p = ctx.cfg
        C_ = x.shape[-1]
        rows = x.numel() // C_
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        lib = _lib.lib()
        ws = fused._workspace(lib.eegx_layernorm_bwd_workspace_bytes(C_), x.device)
        _lib.check(lib.eegx_layernorm_bwd_bf16(
            _lib.ptr(dy), _lib.ptr(x), _lib.ptr(wst.p), _lib.ptr(bst.p), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(dx),
            _lib.ptr(wst.g), _lib.ptr(bst.g), 1, _lib.ptr(ws), ws.numel(), rows, C_, G, wst.p.stride(0), act,
            _lib.ptr(rng), site, p, _lib.stream_ptr()), "eegx_layernorm_bwd_bf16")
        return dx, None, None, None, None, None, None, None, None, None


def glayer_norm(x, ln: nn.LayerNorm, G: int, gelu: bool = False, p: float = 0.0, training: bool = True):
    wst, bst = _stack(ln.weight), _stack(ln.bias)
    if wst.p.stride(0) != bst.p.stride(0) or wst.g.stride(0) != wst.p.stride(0) or bst.g.stride(0) != wst.p.stride(0):
        raise _lib.EegxError("grouped LayerNorm: gamma / beta stacks must share one group stride")
    rng, site, p = fused._drop(p, training, x.device)
    return _GLayerNorm.apply(x.contiguous(), ln.weight, wst, bst, G, ln.eps, gelu, rng, site, p)


def run_sequential(seq: nn.Sequential, x, G: int):
    """layers.run_sequential for G stacked modules (Linear / LayerNorm / GELU / Dropout containers)."""
    mods = list(seq)
    i = 0
    while i < len(mods):
        m = mods[i]
        nxt = mods[i + 1] if i + 1 < len(mods) else None
        nxt2 = mods[i + 2] if i + 2 < len(mods) else None
        if isinstance(m, nn.Linear):
            x = glinear(x, LinW(m.weight, m.bias), G)
            i += 1
        elif isinstance(m, nn.LayerNorm):
            gelu = isinstance(nxt, nn.GELU)
            drop = nxt2 if gelu and isinstance(nxt2, nn.Dropout) else (nxt if isinstance(nxt, nn.Dropout) else None)
            x = glayer_norm(x, m, G, gelu=gelu, p=drop.p if drop is not None else 0.0,
                            training=drop.training if drop is not None else False)
            i += 1 + int(gelu) + int(drop is not None)
        elif isinstance(m, nn.GELU):
            drop = nxt if isinstance(nxt, nn.Dropout) else None
            x = fused.gelu_dropout(x, p=drop.p if drop is not None else 0.0,
                                   training=drop.training if drop is not None else False)
            i += 1 + int(drop is not None)
        else:                                       # (no bare Dropout / Sigmoid in the region encoder's containers)
            raise TypeError(type(m))
    return x


# ------------------------------------------------------------------------------------------ token assembly
class _GAssemble(torch.autograd.Function):
    """[cls ; temporal ; h] + pos_emb per group (layers.py:214-225): (G*B*T, d) -> (G*B*(T+4), d)."""

    @staticmethod
    def forward(ctx, h, anchor, cls_st, tmp_st, pos_st, G, B, T):
        d = h.shape[1]
        out = torch.empty(G * B * (T + 4), d, dtype=torch.bfloat16, device=h.device)
        _lib.check(_lib.lib().eegx_assemble_tokens_fwd_bf16(
            _lib.ptr(h), _lib.ptr(cls_st.p), cls_st.p.stride(0), _lib.ptr(tmp_st.p), tmp_st.p.stride(0),
            _lib.ptr(pos_st.p), pos_st.p.stride(0), _lib.ptr(out), G, B, T, d, _lib.stream_ptr()),
            "eegx_assemble_tokens_fwd_bf16")
        ctx.cfg = (cls_st, tmp_st, pos_st, G, B, T, d)
        return out

    @staticmethod
    def backward(ctx, dout):
        cls_st, tmp_st, pos_st, G, B, T, d = ctx.cfg
        S = T + 4
        dout = dout.contiguous()
        dh = None
        if ctx.needs_input_grad[0]:
            dh = torch.empty(G * B * T, d, dtype=torch.bfloat16, device=dout.device)
            _lib.check(_lib.lib().eegx_assemble_tokens_bwd_bf16(_lib.ptr(dout), _lib.ptr(dh), G * B, T, d,
                                                                _lib.stream_ptr()), "eegx_assemble_tokens_bwd_bf16")
        dpos = torch.empty(G, S * d, dtype=torch.float32, device=dout.device)
        gcolsum(dout.view(G, B, S * d), dpos, accumulate=False)
        pos_st.g.view(G, S * d).add_(dpos)
        cls_st.g.view(G, d).add_(dpos[:, :d])
        tmp_st.g.view(G, 3 * d).add_(dpos[:, d:4 * d])
        return dh, None, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------ attention
def _gmha(mod: nn.MultiheadAttention, q_in, kv_in, G: int, GB: int, Sq: int, Sk: int):
    """nn.MultiheadAttention(batch_first=True) for G stacked modules on (G*B*S, d) rows; see layers._mha."""
    d = q_in.shape[1]
    H = mod.num_heads
    W, b = mod.in_proj_weight, mod.in_proj_bias
    if kv_in is None:
        qkv = glinear(q_in, LinW(W, b), G)
        o = fused.attn_self(qkv, GB, Sq, H, p=mod.dropout, training=mod.training)
    else:
        q = glinear(q_in, LinW(W, b, slice(0, d)), G)
        kv = glinear(kv_in, LinW(W, b, slice(d, 3 * d)), G)
        o = fused.attn_cross(q, kv, GB, Sq, Sk, H, p=mod.dropout, training=mod.training)
    return glinear(o, LinW(mod.out_proj.weight, mod.out_proj.bias), G)


# ------------------------------------------------------------------------------------------ CNN stack
def _stacked_buffers(mods, name):
    """The G modules' buffer `name` (running_mean / running_var / num_batches_tracked) as ONE (G, ...) tensor that
    each module's buffer is a view of (re-pointed on first use and whenever .to() / load_state_dict broke it)."""
    bufs = [getattr(m, name) for m in mods]
    cached = mods[0].__dict__.get("_eegx_bufstack_" + name)
    if cached is not None:
        step = cached[0].numel() * cached.element_size()
        if all(b.data_ptr() == cached.data_ptr() + i * step for i, b in enumerate(bufs)):
            return cached
    with torch.no_grad():
        st = torch.stack([b.detach() for b in bufs]).contiguous()
        for i, b in enumerate(bufs):
            b.data = st[i]
    mods[0].__dict__["_eegx_bufstack_" + name] = st
    return st


def _gbn_stats(yg, bns, G, B, T, training):
    C_ = yg.shape[1]
    bn = bns[0]
    track = bn.track_running_stats and bn.running_mean is not None
    if not training:
        rm, rv = _stacked_buffers(bns, "running_mean"), _stacked_buffers(bns, "running_var")
        return rm.float().contiguous(), torch.rsqrt(rv.float() + bn.eps).contiguous()
    mean = torch.empty(G, C_, dtype=torch.float32, device=yg.device)
    rstd = torch.empty_like(mean)
    rm = _stacked_buffers(bns, "running_mean") if track else None
    rv = _stacked_buffers(bns, "running_var") if track else None
    lib = _lib.lib()
    ws = fused._workspace(lib.eegx_colreduce_workspace_bytes(C_), yg.device)
    _lib.check(lib.eegx_bn_stats_bf16(
        _lib.ptr(yg[PAD:]), G, B, T, PAD, C_, float(bn.eps), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(rm), _lib.ptr(rv),
        C_, float(bn.momentum if bn.momentum is not None else 0.1), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
        "eegx_bn_stats_bf16")
    if track:
        _stacked_buffers(bns, "num_batches_tracked").add_(1)
    return mean, rstd


class _GBnAct(torch.autograd.Function):
    """zero_pad(dropout(gelu(bn_a(ya) + residual))) on ONE guarded buffer of G*B trials, per-group statistics and
    affine parameters (fused._BnAct for G stacked modules)."""

    @staticmethod
    def forward(ctx, ya, yr, anchor, bns_a, bns_r, res_mode, G, B, T, training, rng, site, p):
        C_ = ya.shape[1]
        mean_a, rstd_a = _gbn_stats(ya, bns_a, G, B, T, training)
        mean_r = rstd_r = None
        if res_mode == 2:
            mean_r, rstd_r = _gbn_stats(yr, bns_r, G, B, T, training)
        out = torch.empty_like(ya)
        ga, ba = _stack(bns_a[0].weight), _stack(bns_a[0].bias)
        gr = _stack(bns_r[0].weight) if res_mode == 2 else None
        br = _stack(bns_r[0].bias) if res_mode == 2 else None
        _lib.check(_lib.lib().eegx_bn_act_fwd_bf16(
            _lib.ptr(ya[PAD:]), _lib.ptr(mean_a), _lib.ptr(rstd_a), _lib.ptr(ga.p), _lib.ptr(ba.p),
            _lib.ptr(yr[PAD:]) if res_mode else None, _lib.ptr(mean_r), _lib.ptr(rstd_r),
            _lib.ptr(gr.p) if gr is not None else None, _lib.ptr(br.p) if br is not None else None,
            res_mode, _lib.ptr(out[PAD:]), G, B, T, PAD, C_, _lib.ptr(rng), site, p, _lib.stream_ptr()),
            "eegx_bn_act_fwd_bf16")
        ctx.save_for_backward(ya, yr if res_mode else None, mean_a, rstd_a, mean_r, rstd_r)
        ctx.cfg = (ga, ba, gr, br, res_mode, G, B, T, bool(training), rng, site, p)
        return out

    @staticmethod
    def backward(ctx, dout):
        ya, yr, mean_a, rstd_a, mean_r, rstd_r = ctx.saved_tensors
        ga, ba, gr, br, res_mode, G, B, T, training, rng, site, p = ctx.cfg
        C_ = ya.shape[1]
        dout = dout.contiguous()
        da = torch.empty_like(ya)
        dr = torch.empty_like(ya) if res_mode else None
        sums = torch.empty(G, 3, C_, dtype=torch.float32, device=ya.device)
        lib = _lib.lib()
        ws = fused._workspace(lib.eegx_colreduce_workspace_bytes(C_), ya.device)
        _lib.check(lib.eegx_bn_act_bwd_bf16(
            _lib.ptr(dout[PAD:]), _lib.ptr(ya[PAD:]), _lib.ptr(mean_a), _lib.ptr(rstd_a), _lib.ptr(ga.p), _lib.ptr(ba.p),
            _lib.ptr(yr[PAD:]) if res_mode else None, _lib.ptr(mean_r), _lib.ptr(rstd_r),
            _lib.ptr(gr.p) if gr is not None else None, _lib.ptr(br.p) if br is not None else None, res_mode,
            int(training), _lib.ptr(da[PAD:]), _lib.ptr(dr[PAD:]) if res_mode else None, _lib.ptr(sums), _lib.ptr(ws),
            ws.numel(), G, B, T, PAD, C_, _lib.ptr(rng), site, p, _lib.stream_ptr()), "eegx_bn_act_bwd_bf16")
        bufs, vals = [ga.g, ba.g], [sums[:, 1], sums[:, 0]]
        if res_mode == 2:
            bufs += [gr.g, br.g]
            vals += [sums[:, 2], sums[:, 0]]
        torch._foreach_add_(bufs, vals)
        return da, dr, None, None, None, None, None, None, None, None, None, None, None


def gbn_act(ya, bns_a, yr, bns_r, G, B, T, p, training, drop_training):
    res_mode = 0 if yr is None else (2 if bns_r is not None else 1)
    rng, site, p = fused._drop(p, drop_training, ya.device)
    return _GBnAct.apply(ya, yr, bns_a[0].weight, bns_a, bns_r, res_mode, G, B, T, training, rng, site, p)


def _conv_fwd_pack(st):       # (G, Cout, Cin, k) -> (G, Cout, k*Cin) bf16, K index = tap*Cin + ci
    return _stack_pack(st, "gconvf", lambda w: w.permute(0, 1, 3, 2).reshape(w.shape[0], w.shape[1], -1).contiguous())


def _conv_dgrad_pack(st):     # (G, Cout, Cin, k) -> (G, k*Cout, Cin) bf16 with the taps flipped
    return _stack_pack(st, "gconvd",
                       lambda w: w.flip(3).permute(0, 3, 1, 2).reshape(w.shape[0], -1, w.shape[2]).contiguous())


class _GConv(torch.autograd.Function):
    """Conv1d (stride 1, 'same' padding) of G stacked modules on ONE guarded channels-last buffer of G*B trials
    (nn_ops._ConvG for G modules): forward, data gradient and weight gradient are grouped implicit-im2col GEMMs."""

    @staticmethod
    def forward(ctx, xg, anchor, wst, bst, G, Mg, bias_grad):
        _, Cout, Cin, k = wst.p.shape
        p = k // 2
        a = xg.as_strided((G, Mg, k * Cin), (Mg * Cin, Cin, 1), xg.storage_offset() + (PAD - p) * Cin)
        yg = torch.empty(G * Mg + 2 * PAD, Cout, dtype=torch.bfloat16, device=xg.device)
        ops.gemm(a, _conv_fwd_pack(wst), bst.p if bst is not None else None, grouped=True,
                 out=yg[PAD:PAD + G * Mg].view(G, Mg, Cout))
        ctx.save_for_backward(xg)
        ctx.cfg = (wst, bst, G, Mg, bool(bias_grad))
        return yg

    @staticmethod
    def backward(ctx, dyg):
        (xg,) = ctx.saved_tensors
        wst, bst, G, Mg, bias_grad = ctx.cfg
        _, Cout, Cin, k = wst.p.shape
        p = k // 2
        dyg = dyg.contiguous()                       # clean guarded tensor: zero outside the valid rows
        dy3 = dyg[PAD:PAD + G * Mg].view(G, Mg, Cout)
        a_view = xg.as_strided((G, Mg, k * Cin), (Mg * Cin, Cin, 1), xg.storage_offset() + (PAD - p) * Cin)

        def grads():
            part = _wgrad_partials(dy3, a_view)                                              # (s, G, Cout, k*Cin)
            _lib.check(_lib.lib().eegx_accumulate_conv_wgrad_f32(_lib.ptr(part), part.shape[0], G * Cout, Cin, k,
                                                                 _lib.ptr(wst.g), 1, _lib.stream_ptr()),
                       "eegx_accumulate_conv_wgrad_f32")
            if bst is not None and bias_grad:        # a bias in front of a train-mode BatchNorm has a zero gradient
                gcolsum(dy3, bst.g)
        fused.deferred(grads, dyg, xg)
        dxg = None
        if ctx.needs_input_grad[0]:
            a = dyg.as_strided((G, Mg, k * Cout), (Mg * Cout, Cout, 1), dyg.storage_offset() + (PAD - p) * Cout)
            dxg = torch.empty(G * Mg + 2 * PAD, Cin, dtype=torch.bfloat16, device=dyg.device)
            # guard rows stay unwritten: every consumer reads rows [0, G*Mg) only and masks the padding rows
            ops.gemm(a, _conv_dgrad_pack(wst), b_mn_major=True, grouped=True, out=dxg[PAD:PAD + G * Mg].view(G, Mg, Cin))
        return dxg, None, None, None, None, None, None


def gconv(xg, convs, G, Mg, bias_grad=True):
    w = convs[0].weight
    if w.shape[1] % 8 != 0:
        raise ValueError("Conv1d in_channels must be a multiple of 8 on this path")
    bst = _stack(convs[0].bias) if convs[0].bias is not None else None
    return _GConv.apply(xg, w, _stack(w), bst, G, Mg, bias_grad)


class _GDwConv5(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xg, anchor, wst, bst, G, B, T):
        C_ = xg.shape[1]
        out = torch.empty_like(xg)
        _lib.check(_lib.lib().eegx_dwconv5_fwd_bf16(_lib.ptr(xg[PAD:]), _lib.ptr(wst.p), _lib.ptr(bst.p), _lib.ptr(out[PAD:]),
                                                    G, B, T, PAD, C_, _lib.stream_ptr()), "eegx_dwconv5_fwd_bf16")
        ctx.save_for_backward(xg)
        ctx.cfg = (wst, bst, G, B, T)
        return out

    @staticmethod
    def backward(ctx, dout):
        (xg,) = ctx.saved_tensors
        wst, bst, G, B, T = ctx.cfg
        C_ = xg.shape[1]
        dout = dout.contiguous()
        dx = torch.empty_like(xg)
        dwdb = torch.empty(G, 6, C_, dtype=torch.float32, device=xg.device)
        lib = _lib.lib()
        ws = fused._workspace(lib.eegx_colreduce_workspace_bytes(C_), xg.device)
        _lib.check(lib.eegx_dwconv5_bwd_bf16(_lib.ptr(dout[PAD:]), _lib.ptr(xg[PAD:]), _lib.ptr(wst.p), _lib.ptr(dx[PAD:]),
                                             _lib.ptr(dwdb), _lib.ptr(ws), ws.numel(), G, B, T, PAD, C_,
                                             _lib.stream_ptr()), "eegx_dwconv5_bwd_bf16")
        wst.g.view(G, C_, 5).add_(dwdb[:, :5].transpose(1, 2))
        bst.g.add_(dwdb[:, 5])
        return dx, None, None, None, None, None, None


class _GToRows(torch.autograd.Function):
    """G tensors (B, C, T) fp32 -> ONE guarded channels-last bf16 buffer of G*B trials (block g = input g)."""

    @staticmethod
    def forward(ctx, *xs):
        B, C_, T = xs[0].shape
        G = len(xs)
        Tp = T + 2 * PAD
        out = torch.empty(G * B * Tp + 2 * PAD, C_, dtype=torch.bfloat16, device=xs[0].device)
        for g, x in enumerate(xs):
            if not x.is_cuda or x.dtype != torch.float32:
                raise _lib.EegxError("region inputs must be CUDA float32 tensors (no CPU fallback)")
            if x.stride(2) != 1 or x.stride(1) != T or x.stride(0) < C_ * T:
                x = x.contiguous()
            _lib.check(_lib.lib().eegx_nct_to_rows_bf16(_lib.ptr(x), x.stride(0), _lib.ptr(out[PAD + g * B * Tp:]), B, T, PAD,
                                                        C_, _lib.stream_ptr()), "eegx_nct_to_rows_bf16")
        ctx.shape = (G, B, C_, T)
        return out

    @staticmethod
    def backward(ctx, dg):
        G, B, C_, T = ctx.shape
        Tp = T + 2 * PAD
        d = dg[PAD:PAD + G * B * Tp].view(G, B, Tp, C_)[:, :, PAD:PAD + T].transpose(2, 3).float()
        return tuple(d[g] if ctx.needs_input_grad[g] else None for g in range(G))


def _cnn(mods, xs):
    """The CNN stack of layers.Conv1DWithAttention._cnn for G modules -> compact (G*B*T, 768) bf16 rows."""
    G = len(mods)
    B, C_, T = xs[0].shape
    if C_ % 8 != 0:
        raise ValueError("n_channels must be a multiple of 8 on this path (TMA row pitch)")
    m0 = mods[0]
    Mg = B * (T + 2 * PAD)
    tr = m0.training

    def res_block(hg, cname, bname, rname, drop):
        convs, bns = [getattr(m, cname) for m in mods], [getattr(m, bname) for m in mods]
        res = [getattr(m, rname) for m in mods]
        ya = gconv(hg, convs, G, Mg, bias_grad=not bns[0].training)
        p = drop.p if drop is not None else 0.0
        if isinstance(res[0], nn.Identity):
            return gbn_act(ya, bns, hg, None, G, B, T, p, bns[0].training, tr)
        yr = gconv(hg, [r[0] for r in res], G, Mg)
        return gbn_act(ya, bns, yr, [r[1] for r in res], G, B, T, p, bns[0].training, tr)

    hg = _GToRows.apply(*[x.float() for x in xs])
    hg = res_block(hg, "conv1", "bn1", "residual1", m0.dropout_light)
    hg = res_block(hg, "conv2", "bn2", "residual2", m0.dropout_light)
    dw = m0.depthwise_conv
    d = _GDwConv5.apply(hg, dw.weight, _stack(dw.weight), _stack(dw.bias), G, B, T)
    bns = [m.bn_depth for m in mods]
    y = gconv(d, [m.pointwise_conv for m in mods], G, Mg, bias_grad=not bns[0].training)
    hg = gbn_act(y, bns, None, None, G, B, T, m0.dropout_medium.p, bns[0].training, tr)
    hg = res_block(hg, "conv3", "bn3", "residual3", m0.dropout_medium)
    hg = res_block(hg, "conv4", "bn4", "residual4", None)
    # SqueezeExcite: s = mean_T(x); e = sigmoid(W2 relu(W1 s + b1) + b2); dropout_heavy(x * e) as compact rows
    s = fused.group_mean(hg, G * B, T)
    fc1, fc2 = m0.se_block.excitation[0], m0.se_block.excitation[2]
    z = torch.relu(glinear(s.to(torch.bfloat16), LinW(fc1.weight, fc1.bias), G))
    e = torch.sigmoid(glinear(z, LinW(fc2.weight, fc2.bias), G).float())
    return fused.se_scale(hg, e, G * B, T, p=m0.dropout_heavy.p, training=tr)


# ------------------------------------------------------------------------------------------ the stacked forward
def attention_and_heads(mods, h, B: int, T: int):
    """Everything after the CNN stack (layers.py:210-272) for G stacked modules.
    h: (G*B*T, 768) bf16 compact rows (block g = module g).  Returns (G, B, hidden) fp32 features."""
    G = len(mods)
    m0 = mods[0]
    GB, S = G * B, T + 4
    tr = m0.training
    if S != m0.pos_emb.shape[1]:
        raise _lib.EegxError("grouped path: sequence length differs from n_timepoints + 4")
    h = run_sequential(m0.cnn_to_attn, h, G)
    h = _GAssemble.apply(h, m0.cls_token, _stack(m0.cls_token), _stack(m0.temporal_tokens), _stack(m0.pos_emb), G, B, T)
    h = fused.grad_boundary(h, ('region', 'group'))            # attention stack + heads: gradients final here

    prev = None
    mid = len(m0.attn_layers) // 2
    for i, layer in enumerate(m0.attn_layers):
        if i == mid and i > 0:
            h = fused.grad_boundary(h, ('region_mid', 'group'))
        a = _gmha(layer['attn'], glayer_norm(h, layer['attn_norm'], G), None, G, GB, S, S)
        h = fused.add_dropout(h, a, p=m0.dropout_light.p, training=tr)
        saved = h
        ffn = layer['ffn']
        ag = glinear_cat(glayer_norm(h, layer['ffn_norm'], G),
                         (LinW(ffn.linear1.weight, ffn.linear1.bias), LinW(ffn.gate.weight, ffn.gate.bias)), G)
        f = glinear(fused.glu(ag, p=ffn.dropout.p, training=ffn.dropout.training),
                    LinW(ffn.linear2.weight, ffn.linear2.bias), G)
        h = fused.add_dropout(h, f, p=m0.dropout_medium.p, training=tr)
        if i > 0:
            c = _gmha(m0.cross_scale_attn, h, prev, G, GB, S, S)
            h = fused.add_dropout(h, c, scale=0.1)
        prev = saved

    d = h.shape[1]
    h4 = h.view(GB, S, d)[:, :4].float()
    feat = (h4[:, 0] + 0.3 * h4[:, 1:4].mean(dim=1)).to(torch.bfloat16)                    # (G*B, d)
    comb = torch.cat([run_sequential(p, feat, G) for p in m0.multi_scale_proj], dim=1)
    final = run_sequential(m0.projection, comb, G).float()
    dh = m0.diversity_head
    div = glinear(final.to(torch.bfloat16), LinW(dh.weight, dh.bias), G).float()
    return (final + 0.1 * F.normalize(div, dim=-1)).view(G, B, -1)


def forward(mods, xs):
    """G region encoders in lock step: xs = G tensors (B, C, T) -> (G, B, hidden) fp32."""
    B, _, T = xs[0].shape
    h = _cnn(mods, xs)
    return attention_and_heads(mods, h, B, T)
