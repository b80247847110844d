"""Autograd glue between torch tensors and the libeegx contractions.

Every dense contraction of the encoder (nn.Linear and nn.Conv1d call sites of reference
``main_model/src/models/layers.py`` / ``brain_encoder.py``) goes through ``ops.gemm`` -- the
tcgen05 kernel -- in all three directions (forward, data gradient, weight gradient).
Activations are bf16, channels-last; parameters stay fp32 "master" tensors and are packed to
bf16 once per optimizer step (cache keyed on the parameter's version counter).
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import fused, ops

PAD = 4   # largest Conv1d padding in the encoder (kernel 9, layers.py:30)

_pack_cache: dict = {}
_USE_SHADOW = os.environ.get("EEGX_BF16_SHADOW", "1") != "0"     # A/B switch for the AdamW-written bf16 weight shadows


def _cached(param: torch.Tensor, tag: str, make):
    """bf16 pack of a parameter (or of a row-slice view of one), rebuilt when the parameter's
    version counter moves (i.e. after an optimizer step).  The entry keeps the base tensor alive,
    so its id() cannot be recycled while the entry exists."""
    base = param._base if param._base is not None else param
    if not isinstance(base, torch.nn.Parameter):       # a computed tensor (e.g. a zero-padded weight): nothing to key on
        with torch.no_grad():
            return make(param.detach())
    key = (id(base), param.storage_offset(), tuple(param.shape), tag)
    hit = _pack_cache.get(key)
    ver = base._version
    if hit is not None and hit[0] == ver and hit[2] is base:
        return hit[1]
    with torch.no_grad():
        val = make(param.detach())
    _pack_cache[key] = (ver, val, base)
    return val


def clear_pack_cache():
    _pack_cache.clear()


def _bf16_src(p: torch.Tensor) -> torch.Tensor:
    """bf16 values of a parameter (or of a row-slice view of one): the shadow the fused AdamW kernel keeps
    beside the fp32 master when it is current (no cast kernel), else a cast."""
    base = p._base if p._base is not None else p
    sh = getattr(base, "_eegx_w16", None) if _USE_SHADOW else None
    if sh is not None and getattr(base, "_eegx_w16_ver", -1) == base._version and sh.shape == base.shape:
        if p is base:
            return sh
        return sh.as_strided(p.shape, p.stride(), sh.storage_offset() + p.storage_offset() - base.storage_offset())
    return p.detach().to(torch.bfloat16)


def _w_linear(w):            # (N, K) fp32 -> bf16, rows zero-padded to a multiple of 8 (TMA row pitch of dy)
    def make(_):
        p16 = _bf16_src(w)
        n_pad = (-p16.shape[0]) % 8
        return (torch.nn.functional.pad(p16, (0, 0, 0, n_pad)) if n_pad else p16).contiguous()
    return _cached(w, "lin", make)


def _w_conv_fwd(w):          # (Cout, Cin, k) -> (Cout, k*Cin) bf16, K index = tap*Cin + ci
    return _cached(w, "convf", lambda _: _bf16_src(w).permute(0, 2, 1).reshape(w.shape[0], -1).contiguous())


def _w_conv_dgrad(w):        # (Cout, Cin, k) -> (k*Cout, Cin) bf16 with taps flipped (K x N, N contiguous)
    return _cached(w, "convd", lambda _: _bf16_src(w).flip(2).permute(2, 0, 1).reshape(-1, w.shape[1]).contiguous())


def _wgrad(dy: torch.Tensor, x: torch.Tensor, into: Optional[torch.Tensor] = None, raw: bool = False):
    """dW (N, K) fp32 = dy^T x, reduction over the R rows.  The output is small (N x K) while R is
    ~10^4, so a plain launch fills only N*K / (128*256) of the 148 SMs: split the reduction into s
    row-chunks run as GEMM batches (s * tiles ~ one wave) and add the partials in a fixed order.
    into: an (N, K) fp32 gradient buffer to accumulate into (GEMM epilogue `D += ...` / one
    partial-folding pass); then None is returned and autograd has nothing left to add."""
    R, N = dy.shape
    K = x.shape[1]
    bn = 64 if K <= 64 else (256 if K > 128 and (K % 256 == 0 or K >= 1024) else 128)
    tiles = -(-N // 128) * -(-K // bn)
    s = 1
    while tiles * s * 2 <= 160 and R % (s * 2) == 0 and (R // (s * 2)) % 8 == 0 and R // (s * 2) >= 512:
        s *= 2
    if raw:                                           # (s, N, K) partials, the caller folds them
        if s == 1:
            return ops.gemm(dy, x, a_mn_major=True, b_mn_major=True, out_dtype=torch.float32).unsqueeze(0)
        chunk = R // s
        dy3 = dy.as_strided((s, chunk, N), (chunk * dy.stride(0), dy.stride(0), 1), dy.storage_offset())
        x3 = x.as_strided((s, chunk, K), (chunk * x.stride(0), x.stride(0), 1), x.storage_offset())
        return ops.gemm(dy3, x3, a_mn_major=True, b_mn_major=True, out_dtype=torch.float32)
    if s == 1:
        if into is not None:
            ops.gemm(dy, x, a_mn_major=True, b_mn_major=True, out=into, accumulate=True)
            return None
        return ops.gemm(dy, x, a_mn_major=True, b_mn_major=True, out_dtype=torch.float32)
    chunk = R // s
    dy3 = dy.as_strided((s, chunk, N), (chunk * dy.stride(0), dy.stride(0), 1), dy.storage_offset())
    x3 = x.as_strided((s, chunk, K), (chunk * x.stride(0), x.stride(0), 1), x.storage_offset())
    part = ops.gemm(dy3, x3, a_mn_major=True, b_mn_major=True, out_dtype=torch.float32)      # (s, N, K)
    if into is not None and (N * K) % 4 == 0:
        fused.accumulate_partials(part, into)
        return None
    return part.sum(0)


def _grad2d(w):
    """(N, K) view of a 2-D weight's in-place gradient buffer, or None."""
    g = fused.grad_buffer(w)
    return g if g is not None and g.dim() == 2 else None


class _Linear(torch.autograd.Function):
    """y = x W^T + b (optionally GELU) on (M, K) bf16 rows; dW returned in fp32 for the master weight."""

    @staticmethod
    def forward(ctx, x, weight, bias, gelu):
        w16 = _w_linear(weight)
        N, n_pad = weight.shape[0], w16.shape[0] - weight.shape[0]
        b32 = bias.detach().float().contiguous() if bias is not None else None
        if gelu and b32 is None:
            b32 = torch.zeros(N, device=x.device)
        if b32 is not None and n_pad:
            b32 = torch.nn.functional.pad(b32, (0, n_pad))
        if gelu:
            pre = ops.gemm(x, w16, b32)                 # keep the pre-activation for backward
            y = torch.nn.functional.gelu(pre.float()).to(torch.bfloat16)
            ctx.save_for_backward(x, weight, pre)
        else:
            y = ops.gemm(x, w16, b32)
            ctx.save_for_backward(x, weight, None)
        ctx.has_bias = bias is not None
        ctx.bias_ref = bias
        ctx.gelu = gelu
        ctx.n_pad = n_pad
        return y[:, :N] if n_pad else y

    @staticmethod
    def backward(ctx, dy):
        x, weight, pre = ctx.saved_tensors
        N = weight.shape[0]
        dy = torch.nn.functional.pad(dy, (0, ctx.n_pad)) if ctx.n_pad else dy.contiguous()
        if ctx.gelu:
            p = pre.float()
            cdf = 0.5 * (1.0 + torch.erf(p * 0.7071067811865476))
            pdf = torch.exp(-0.5 * p * p) * 0.3989422804014327
            dy = (dy.float() * (cdf + p * pdf)).to(torch.bfloat16).contiguous()
        w16 = _w_linear(weight)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = ops.gemm(dy, w16, b_mn_major=True)                                   # (M,N) @ (N,K)
        into = _grad2d(weight) if ctx.n_pad == 0 and ctx.needs_input_grad[1] else None
        gb = fused.grad_buffer(ctx.bias_ref) if ctx.n_pad == 0 and ctx.has_bias and ctx.needs_input_grad[2] else None
        want_b = ctx.has_bias and ctx.needs_input_grad[2]
        if into is not None and (gb is not None or not want_b):
            # both parameter gradients accumulate in place: off the critical path (fused.deferred)
            def grads():
                _wgrad(dy, x, into)
                if gb is not None:
                    fused.colsum(dy, gb)
            fused.deferred(grads, dy, x)
            return dx, None, None, None
        if ctx.needs_input_grad[1]:
            dw = _wgrad(dy, x, into)                                                     # dy^T @ x
            dw = dw[:N] if dw is not None else None
        if want_b:
            db = fused.colsum(dy, gb)
            db = db[:N] if db is not None else None
        return dx, dw, db, None


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
           gelu: bool = False) -> torch.Tensor:
    """nn.Linear on the last dim of a bf16 tensor, through the tcgen05 GEMM."""
    shp = x.shape
    x2 = x.reshape(-1, shp[-1])
    if x2.stride(-1) != 1 or (x2.stride(0) % 8) != 0:
        x2 = x2.contiguous()
    K = x2.shape[1]
    if K % 8 != 0:      # TMA needs 16-byte row pitch: pad the contraction dim with zeros
        raise ValueError(f"in_features={K} must be a multiple of 8 on this path")
    y = _Linear.apply(x2, weight, bias, gelu)
    return y.reshape(*shp[:-1], weight.shape[0])


def guard_pad(x_cl: torch.Tensor) -> torch.Tensor:
    """(B, T, C) channels-last -> guarded flat buffer ((B*(T+2*PAD) + 2*PAD), C) with zero rows
    around every trial, so a Conv1d over time is a GEMM on overlapping rows (implicit im2col)."""
    B, T, C = x_cl.shape
    xp = torch.nn.functional.pad(x_cl, (0, 0, PAD, PAD))                 # per-trial zero padding
    flat = xp.reshape(B * (T + 2 * PAD), C)
    return torch.nn.functional.pad(flat, (0, 0, PAD, PAD))               # global guard rows


class _ConvCL(torch.autograd.Function):
    """Conv1d (stride 1, 'same' zero padding k//2) on a guarded channels-last buffer.

    buf: ((M + 2*PAD), Cin) bf16 with M = B*(T+2*PAD); returns (M, Cout) bf16 -- rows that fall in
    a trial's padding zone hold don't-care values and must be masked by the caller.
    """

    @staticmethod
    def forward(ctx, buf, weight, bias, M):
        Cout, Cin, k = weight.shape
        p = k // 2
        a = buf.as_strided((M, k * Cin), (Cin, 1), buf.storage_offset() + (PAD - p) * Cin)
        b32 = bias.detach().float().contiguous() if bias is not None else None
        y = ops.gemm(a, _w_conv_fwd(weight), b32)
        ctx.save_for_backward(buf, weight)
        ctx.M = M
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        buf, weight = ctx.saved_tensors
        Cout, Cin, k = weight.shape
        p = k // 2
        M = ctx.M
        dy = dy.contiguous()                              # (M, Cout); zero on padding rows
        dbuf = dw = db = None
        if ctx.needs_input_grad[1]:
            a_view = buf.as_strided((M, k * Cin), (Cin, 1), buf.storage_offset() + (PAD - p) * Cin)
            dwf = _wgrad(dy, a_view)                                                     # (Cout, k*Cin)
            dw = dwf.view(Cout, k, Cin).permute(0, 2, 1).contiguous()
        if ctx.needs_input_grad[0]:
            # dx = conv of the (guarded) dy with the flipped taps: rows overlap again
            dyg = torch.nn.functional.pad(dy, (0, 0, PAD, PAD))
            a = dyg.as_strided((M, k * Cout), (Cout, 1), (PAD - p) * Cout)
            dx = ops.gemm(a, _w_conv_dgrad(weight), b_mn_major=True)            # (M, Cin)
            dbuf = torch.nn.functional.pad(dx, (0, 0, PAD, PAD))
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = fused.colsum(dy)
        return dbuf, dw, db, None


def conv1d_cl(buf: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], M: int) -> torch.Tensor:
    if weight.shape[1] % 8 != 0:
        raise ValueError("Conv1d in_channels must be a multiple of 8 on this path")
    return _ConvCL.apply(buf, weight, bias, M)


def _w_cat(ws):
    """bf16 pack of several (N_i, K) weights stacked along N (cached on the first weight's version)."""
    tag = "cat:" + ",".join(str(id(w)) for w in ws[1:])
    return _cached(ws[0], tag, lambda _: torch.cat([_bf16_src(w) for w in ws], 0).contiguous())


class _LinearCat(torch.autograd.Function):
    """[y1 | y2 | ...] = x [W1; W2; ...]^T + [b1 | b2 | ...] in one GEMM: projections that share their
    input (the two up-projections of the gated FFN, reference layers.py:311-317; q/k/v projections) cost
    one launch forward, one dgrad and one wgrad."""

    @staticmethod
    def forward(ctx, x, *wb):
        n = len(wb) // 2
        ws, bs = wb[:n], wb[n:]
        bcat = torch.cat([b.detach().float() for b in bs])
        y = ops.gemm(x, _w_cat(ws), bcat)
        ctx.save_for_backward(x, *ws)
        ctx.biases = bs
        return y

    @staticmethod
    def backward(ctx, dy):
        x, *ws = ctx.saved_tensors
        bs = ctx.biases
        dy = dy.contiguous()
        dx = ops.gemm(dy, _w_cat(ws), b_mn_major=True) if ctx.needs_input_grad[0] else None
        if all(_grad2d(w) is not None for w in ws) and all(fused.grad_buffer(b) is not None for b in bs):
            def grads():
                off = 0
                for w, b in zip(ws, bs):
                    n = w.shape[0]
                    sl = dy[:, off:off + n]
                    _wgrad(sl, x, _grad2d(w))
                    fused.colsum(sl, fused.grad_buffer(b))
                    off += n
            fused.deferred(grads, dy, x)
            return (dx, *([None] * (2 * len(ws))))
        dws, dbs, off = [], [], 0
        for w, b in zip(ws, bs):
            n = w.shape[0]
            sl = dy[:, off:off + n]                       # column block of dy: row pitch = total N
            dws.append(_wgrad(sl, x, _grad2d(w)))
            dbs.append(fused.colsum(sl, fused.grad_buffer(b)))
            off += n
        return (dx, *dws, *dbs)


def linear_cat(x: torch.Tensor, *wb) -> torch.Tensor:
    """linear_cat(x, w1, b1, w2, b2, ...): (..., K) -> (..., N1 + N2 + ...); all N_i multiples of 8."""
    ws, bs = wb[0::2], wb[1::2]
    shp = x.shape
    x2 = x.reshape(-1, shp[-1])
    if x2.stride(-1) != 1 or (x2.stride(0) % 8) != 0:
        x2 = x2.contiguous()
    if any(w.shape[0] % 8 for w in ws) or shp[-1] % 8:
        raise ValueError("linear_cat: feature sizes must be multiples of 8")
    return _LinearCat.apply(x2, *ws, *bs).reshape(*shp[:-1], sum(w.shape[0] for w in ws))


class _LMHeadCE(torch.autograd.Function):
    """mean cross-entropy of (h W^T + bias) against labels (ignore_index = -100) without fp32 logits:
    bf16 logits from the GEMM, one-pass CE forward, in-kernel (softmax - onehot) backward, then the
    dgrad / wgrad GEMMs.  Returns (loss, logits (rows, V) bf16 view)."""

    @staticmethod
    def forward(ctx, h, weight, bias, labels):
        import ctypes as C
        from . import _lib
        rows, V = h.shape[0], weight.shape[0]
        w16 = _w_linear(weight)                                  # rows padded to a multiple of 8
        ld = w16.shape[0]
        b32 = torch.zeros(ld, dtype=torch.float32, device=h.device)
        if bias is not None:
            b32[:V] = bias.detach().float().reshape(-1)
        logits = ops.gemm(h, w16, b32)                           # (rows, ld) bf16
        labels = labels.reshape(-1).contiguous()
        loss_rows = torch.empty(rows, dtype=torch.float32, device=h.device)
        lse = torch.empty_like(loss_rows)
        _lib.check(_lib.lib().eegx_ce_fwd_bf16(_lib.ptr(logits), ld, _lib.ptr(labels), rows, V, -100,
                                               _lib.ptr(loss_rows), _lib.ptr(lse), _lib.stream_ptr()),
                   "eegx_ce_fwd_bf16")
        n_valid = (labels != -100).sum().float()
        loss = loss_rows.sum() / n_valid
        ctx.save_for_backward(h, weight, logits, labels, lse, n_valid)
        ctx.mark_non_differentiable(logits)
        return loss, logits

    @staticmethod
    def backward(ctx, dloss, _dlogits):
        from . import _lib
        h, weight, logits, labels, lse, n_valid = ctx.saved_tensors
        rows, V, ld = h.shape[0], weight.shape[0], logits.shape[1]
        coef = (dloss.float() / n_valid).reshape(1).contiguous()
        dlogits = torch.empty_like(logits)
        _lib.check(_lib.lib().eegx_ce_bwd_bf16(_lib.ptr(logits), ld, _lib.ptr(labels), _lib.ptr(lse), _lib.ptr(coef),
                                               _lib.ptr(dlogits), rows, V, -100, _lib.stream_ptr()),
                   "eegx_ce_bwd_bf16")
        w16 = _w_linear(weight)
        dh = ops.gemm(dlogits, w16, b_mn_major=True) if ctx.needs_input_grad[0] else None
        dw = _wgrad(dlogits, h)[:V] if ctx.needs_input_grad[1] else None
        return dh, dw, None, None


def lm_head_cross_entropy(h: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                          labels: torch.Tensor):
    """h: (rows, K) bf16; weight: (V, K); bias: (V,) or None; labels: (rows,) int64 with -100 = ignore.
    Returns (mean loss over the non-ignored rows, logits (rows, V) bf16)."""
    loss, logits = _LMHeadCE.apply(h.contiguous(), weight, bias, labels)
    return loss, logits[:, :weight.shape[0]]


class _ConvG(torch.autograd.Function):
    """Conv1d (stride 1, 'same' zero padding k//2) on a guarded channels-last tensor
    xg ((M + 2*PAD), Cin) -> raw output ((M + 2*PAD), Cout): rows PAD .. PAD+M are written, the rows
    of each trial's padding zone hold don't-care values (the BatchNorm kernels skip them)."""

    @staticmethod
    def forward(ctx, xg, weight, bias, M, bias_grad):
        Cout, Cin, k = weight.shape
        p = k // 2
        a = xg.as_strided((M, k * Cin), (Cin, 1), xg.storage_offset() + (PAD - p) * Cin)
        b32 = bias.detach().float().contiguous() if bias is not None else None
        yg = torch.empty(M + 2 * PAD, Cout, dtype=torch.bfloat16, device=xg.device)
        ops.gemm(a, _w_conv_fwd(weight), b32, out=yg[PAD:PAD + M])
        ctx.save_for_backward(xg, weight)
        ctx.cfg = (M, bias is not None, bool(bias_grad))
        return yg

    @staticmethod
    def backward(ctx, dyg):
        xg, weight = ctx.saved_tensors
        M, has_bias, bias_grad = ctx.cfg
        Cout, Cin, k = weight.shape
        p = k // 2
        dyg = dyg.contiguous()                       # clean guarded tensor: zero outside the valid rows
        dy = dyg[PAD:PAD + M]
        dxg = dw = db = None
        if ctx.needs_input_grad[1]:
            a_view = xg.as_strided((M, k * Cin), (Cin, 1), xg.storage_offset() + (PAD - p) * Cin)
            part = _wgrad(dy, a_view, raw=True)                                          # (s, Cout, k*Cin)
            gbuf = fused.grad_buffer(weight)
            dw = None if gbuf is not None else torch.empty_like(weight, dtype=torch.float32)
            fused.accumulate_conv_wgrad(part, gbuf if gbuf is not None else dw, Cin, k, accumulate=gbuf is not None)
        if ctx.needs_input_grad[0]:
            a = dyg.as_strided((M, k * Cout), (Cout, 1), dyg.storage_offset() + (PAD - p) * Cout)
            dxg = torch.empty(M + 2 * PAD, Cin, dtype=torch.bfloat16, device=dyg.device)
            ops.gemm(a, _w_conv_dgrad(weight), b_mn_major=True, out=dxg[PAD:PAD + M])
        if has_bias and ctx.needs_input_grad[2]:
            # a bias in front of a train-mode BatchNorm has an exactly zero gradient
            db = fused.colsum(dy) if bias_grad else torch.zeros(Cout, dtype=torch.float32, device=dyg.device)
        return dxg, dw, db, None, None


def conv_g(xg: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], M: int,
           bias_grad: bool = True) -> torch.Tensor:
    if weight.shape[1] % 8 != 0:
        raise ValueError("Conv1d in_channels must be a multiple of 8 on this path")
    return _ConvG.apply(xg, weight, bias, M, bias_grad)
