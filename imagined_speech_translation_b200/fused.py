"""torch.autograd wrappers over the fused glue kernels of libeegx (csrc/fused_rowwise.cu,
fused_bn.cu, attn_small.cu): every op is one HBM pass forward and one (or two) backward, bf16
activations, fp32 parameters.  These replace the ATen element-wise chains behind the reference
modules' LayerNorm / BatchNorm / GELU / Dropout / residual adds / attention core
(``main_model/src/models/layers.py:129-272``, ``brain_encoder.py:136-193``).

Dropout is counter based: a device-resident ``[seed, step]`` pair plus a per-call-site id; the
backward kernels regenerate the forward mask, so no mask tensors exist and a CUDA graph replays
with fresh masks once ``advance_rng()`` (an in-graph increment) has run.

No CPU path: every function raises on non-CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import _lib

PAD = 4   # guard / per-trial padding rows of the CNN stack (largest Conv1d padding, layers.py:30)

# ------------------------------------------------------------------------------------------ RNG
_rng_state: dict = {}
_site = [0]
_seed = [0x5EED5EED]


def set_seed(seed: int) -> None:
    _seed[0] = int(seed)
    for dev, t in _rng_state.items():
        t.copy_(torch.tensor([_seed[0], 0], dtype=torch.int64))


def rng_state(device) -> torch.Tensor:
    """Device int64[2] = [seed, step] read by every dropout kernel."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    t = _rng_state.get(key)
    if t is None:
        t = torch.tensor([_seed[0], 0], dtype=torch.int64, device=device)
        _rng_state[key] = t
    return t


def advance_rng(device) -> None:
    """step += 1 on the device (captured by CUDA graphs, so every replay draws new masks)."""
    rng_state(device)[1:2].add_(1)


def begin_step() -> None:
    """Restart the call-site numbering (same site ids every step; the step counter changes the masks)."""
    _site[0] = 0


def _next_site() -> int:
    _site[0] = (_site[0] + 1) & 0x7FFFFFFF
    return _site[0]


def _drop(p: float, training: bool, device):
    """(rng pointer, site, p) for a call site; dropout off -> (None, 0, 0.0)."""
    if training and p > 0.0:
        return rng_state(device), _next_site(), float(p)
    return None, 0, 0.0


def _need(t: torch.Tensor, name: str, dtype=torch.bfloat16):
    if not t.is_cuda:
        raise _lib.EegxError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


_ws_cache: dict = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    """Per-(device, stream) scratch for the two-stage reductions (consumed inside the same call)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    t = _ws_cache.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(nbytes, 1 << 22), dtype=torch.uint8, device=device)
        _ws_cache[key] = t
    return t


# ------------------------------------------------------------------------------------------ deferred weight gradients
# Nothing on the critical path of backward waits for a weight / bias gradient: the data gradient of a layer feeds the
# next backward node, its parameter gradients only have to exist when backward ends.  The latency-bound stretches of
# backward (decoder: M = 4096 rows, fusion stage and heads: M <= 1024) leave most SMs idle, so the in-place
# weight-gradient work is issued on a side stream where it fills them; the stream is joined when backward ends (a
# final callback of the autograd engine) and at every gradient-ready boundary of the all-reduce overlap.
import os as _os

DEFER_WGRAD = _os.environ.get("EEGX_DEFER_WGRAD", "1") != "0"
_side_streams: dict = {}
_deferred_pending = [False]


def deferred(fn, *tensors) -> None:
    """Run fn() -- work that only ACCUMULATES parameter gradients in place -- on the side stream.  `tensors`: every
    tensor of the current stream that fn reads (kept alive for the side stream by the caching allocator)."""
    t0 = next((t for t in tensors if t is not None), None)
    if not DEFER_WGRAD or t0 is None or not t0.is_cuda:
        fn()
        return
    dev = t0.device
    cur = torch.cuda.current_stream(dev)
    side = _side_streams.get(dev.index)
    if side is None:
        side = _side_streams[dev.index] = torch.cuda.Stream(device=dev)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        fn()
    for t in tensors:
        if t is not None:
            t.record_stream(side)
    if not _deferred_pending[0]:
        _deferred_pending[0] = True
        torch.autograd.Variable._execution_engine.queue_callback(_join_at_end_of_backward)


def join_side() -> None:
    """The current stream waits for everything deferred so far (gradients complete from its point of view)."""
    for idx, side in _side_streams.items():
        torch.cuda.current_stream(idx).wait_stream(side)


def _join_at_end_of_backward() -> None:
    _deferred_pending[0] = False
    join_side()


def _f32(p: torch.Tensor) -> torch.Tensor:
    p = p.detach()
    return p if p.dtype == torch.float32 and p.is_contiguous() else p.float().contiguous()


def grad_buffer(p) -> Optional[torch.Tensor]:
    """The parameter's existing gradient buffer if a kernel may accumulate into it in place (fp32,
    contiguous, CUDA) -- the flat buffers of FlatAdamW qualify -- else None (autograd accumulates)."""
    if p is None or not p.is_leaf:
        return None
    g = p.grad
    if g is None or g.dtype != torch.float32 or not g.is_cuda or not g.is_contiguous():
        return None
    return g


def colsum(y: torch.Tensor, into: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """Column sums (fp32) of a (rows, C) bf16 matrix (unit column stride): bias gradients, fixed
    summation order.  into: gradient buffer to accumulate into (then None is returned)."""
    if not y.is_cuda or y.dtype != torch.bfloat16 or y.stride(1) != 1:
        raise ValueError("colsum needs a CUDA bf16 matrix with unit column stride")
    rows, C_ = y.shape
    out = into if into is not None else torch.empty(C_, dtype=torch.float32, device=y.device)
    if rows == 0:
        return None if into is not None else out.zero_()
    lib = _lib.lib()
    ws = _workspace(lib.eegx_colreduce_workspace_bytes(C_), y.device)
    _lib.check(lib.eegx_colsum_bf16(_lib.ptr(y), y.stride(0), 1, 0, rows, C_, _lib.ptr(out), 0, int(into is not None),
                                    _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "eegx_colsum_bf16")
    return None if into is not None else out


def accumulate_partials(part: torch.Tensor, into: torch.Tensor) -> None:
    """into += part.sum(0) for fp32 split-K partials (s, ...) in one pass."""
    s = part.shape[0]
    _lib.check(_lib.lib().eegx_accumulate_partials_f32(_lib.ptr(part), s, 1, part.numel() // s, _lib.ptr(into), 0, 1,
                                                       _lib.stream_ptr()), "eegx_accumulate_partials_f32")


def accumulate_conv_wgrad(part: torch.Tensor, dst: torch.Tensor, Cin: int, k: int, accumulate: bool) -> None:
    """dst (Cout, Cin, k) (+)= sum_s part[s] with part (s, Cout, k*Cin) in the conv GEMM's K order."""
    s, Cout = part.shape[0], part.shape[1]
    _lib.check(_lib.lib().eegx_accumulate_conv_wgrad_f32(_lib.ptr(part), s, Cout, Cin, k, _lib.ptr(dst), int(accumulate),
                                                         _lib.stream_ptr()), "eegx_accumulate_conv_wgrad_f32")


# ------------------------------------------------------------------------------------------ LayerNorm
class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, act, rng, site, p):
        _need(x, "x")
        C_ = x.shape[-1]
        rows = x.numel() // C_
        y = torch.empty_like(x)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        w, b = _f32(weight), _f32(bias)
        _lib.check(_lib.lib().eegx_layernorm_fwd_bf16(
            _lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), _lib.ptr(mean), _lib.ptr(rstd), rows, C_, 1, 0,
            float(eps), int(act), _lib.ptr(rng), site, p, _lib.stream_ptr()), "eegx_layernorm_fwd_bf16")
        ctx.save_for_backward(x, weight, bias, mean, rstd)
        ctx.cfg = (int(act), rng, site, p)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias, mean, rstd = ctx.saved_tensors
        act, rng, site, p = ctx.cfg
        C_ = x.shape[-1]
        rows = x.numel() // C_
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        gw, gb = grad_buffer(weight), grad_buffer(bias)
        fused_acc = gw is not None and gb is not None      # write d(gamma), d(beta) straight into the gradients
        dg = gw if fused_acc else torch.empty(C_, dtype=torch.float32, device=x.device)
        db = gb if fused_acc else torch.empty_like(dg)
        lib = _lib.lib()
        ws = _workspace(lib.eegx_layernorm_bwd_workspace_bytes(C_), x.device)
        _lib.check(lib.eegx_layernorm_bwd_bf16(
            _lib.ptr(dy), _lib.ptr(x), _lib.ptr(_f32(weight)), _lib.ptr(_f32(bias)), _lib.ptr(mean), _lib.ptr(rstd),
            _lib.ptr(dx), _lib.ptr(dg), _lib.ptr(db), int(fused_acc), _lib.ptr(ws), ws.numel(), rows, C_, 1, 0, act,
            _lib.ptr(rng), site, p, _lib.stream_ptr()), "eegx_layernorm_bwd_bf16")
        if fused_acc:
            return dx, None, None, None, None, None, None, None
        return dx, dg.to(weight.dtype), db.to(bias.dtype), None, None, None, None, None


def layer_norm(x: torch.Tensor, weight, bias, eps: float = 1e-5, gelu: bool = False, p: float = 0.0,
               training: bool = True) -> torch.Tensor:
    """dropout(gelu?(LayerNorm(x))) over the last dim of a bf16 tensor."""
    rng, site, p = _drop(p, training, x.device)
    return _LayerNorm.apply(x.contiguous(), weight, bias, eps, gelu, rng, site, p)


# ------------------------------------------------------------------------------------------ element-wise
class _AddDropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, scale, rng, site, p):
        _need(a, "a"); _need(b, "b")
        out = torch.empty_like(a)
        _lib.check(_lib.lib().eegx_add_dropout_fwd_bf16(_lib.ptr(a), _lib.ptr(b), _lib.ptr(out), a.numel(),
                                                        float(scale), _lib.ptr(rng), site, p, _lib.stream_ptr()),
                   "eegx_add_dropout_fwd_bf16")
        ctx.cfg = (float(scale), rng, site, p)
        return out

    @staticmethod
    def backward(ctx, dout):
        scale, rng, site, p = ctx.cfg
        dout = dout.contiguous()
        if p == 0.0 and scale == 1.0:
            return dout, dout, None, None, None, None
        db = torch.empty_like(dout)
        _lib.check(_lib.lib().eegx_dropout_scale_bf16(_lib.ptr(dout), _lib.ptr(db), dout.numel(), scale,
                                                      _lib.ptr(rng), site, p, _lib.stream_ptr()),
                   "eegx_dropout_scale_bf16")
        return dout, db, None, None, None, None


def add_dropout(a: torch.Tensor, b: torch.Tensor, scale: float = 1.0, p: float = 0.0,
                training: bool = True) -> torch.Tensor:
    """a + scale * dropout(b) (bf16, same shape)."""
    rng, site, p = _drop(p, training, a.device)
    return _AddDropout.apply(a.contiguous(), b.contiguous(), scale, rng, site, p)


class _GeluDropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rng, site, p):
        _need(x, "x")
        out = torch.empty_like(x)
        _lib.check(_lib.lib().eegx_gelu_dropout_fwd_bf16(_lib.ptr(x), _lib.ptr(out), x.numel(), _lib.ptr(rng), site,
                                                         p, _lib.stream_ptr()), "eegx_gelu_dropout_fwd_bf16")
        ctx.save_for_backward(x)
        ctx.cfg = (rng, site, p)
        return out

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        rng, site, p = ctx.cfg
        dout = dout.contiguous()
        dx = torch.empty_like(x)
        _lib.check(_lib.lib().eegx_gelu_dropout_bwd_bf16(_lib.ptr(dout), _lib.ptr(x), _lib.ptr(dx), x.numel(),
                                                         _lib.ptr(rng), site, p, _lib.stream_ptr()),
                   "eegx_gelu_dropout_bwd_bf16")
        return dx, None, None, None


def gelu_dropout(x: torch.Tensor, p: float = 0.0, training: bool = True) -> torch.Tensor:
    rng, site, p = _drop(p, training, x.device)
    return _GeluDropout.apply(x.contiguous(), rng, site, p)


class _Glu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ag, rng, site, p):
        _need(ag, "ag")
        H2 = ag.shape[-1]
        rows = ag.numel() // H2
        out = torch.empty(*ag.shape[:-1], H2 // 2, dtype=ag.dtype, device=ag.device)
        _lib.check(_lib.lib().eegx_glu_fwd_bf16(_lib.ptr(ag), _lib.ptr(out), rows, H2 // 2, _lib.ptr(rng), site, p,
                                                _lib.stream_ptr()), "eegx_glu_fwd_bf16")
        ctx.save_for_backward(ag)
        ctx.cfg = (rng, site, p)
        return out

    @staticmethod
    def backward(ctx, dout):
        (ag,) = ctx.saved_tensors
        rng, site, p = ctx.cfg
        H2 = ag.shape[-1]
        rows = ag.numel() // H2
        dout = dout.contiguous()
        dag = torch.empty_like(ag)
        _lib.check(_lib.lib().eegx_glu_bwd_bf16(_lib.ptr(dout), _lib.ptr(ag), _lib.ptr(dag), rows, H2 // 2,
                                                _lib.ptr(rng), site, p, _lib.stream_ptr()), "eegx_glu_bwd_bf16")
        return dag, None, None, None


def glu(ag: torch.Tensor, p: float = 0.0, training: bool = True) -> torch.Tensor:
    """(..., 2H) = [a | g] -> dropout(gelu(a) * sigmoid(g)) (..., H)."""
    rng, site, p = _drop(p, training, ag.device)
    return _Glu.apply(ag.contiguous(), rng, site, p)


# ------------------------------------------------------------------------------------------ CNN stack
def guarded_rows(B: int, T: int) -> int:
    return B * (T + 2 * PAD) + 2 * PAD


class _ToRows(torch.autograd.Function):
    """(B, C, T) fp32 -> guarded channels-last bf16 rows ((B*(T+2*PAD) + 2*PAD), C)."""

    @staticmethod
    def forward(ctx, x):
        if not x.is_cuda or x.dtype != torch.float32:
            raise _lib.EegxError("to_rows needs a CUDA float32 tensor (no CPU fallback)")
        B, C_, T = x.shape
        if x.stride(2) != 1 or x.stride(1) != T or x.stride(0) < C_ * T:
            x = x.contiguous()
        out = torch.empty(guarded_rows(B, T), C_, dtype=torch.bfloat16, device=x.device)
        _lib.check(_lib.lib().eegx_nct_to_rows_bf16(_lib.ptr(x), x.stride(0),
                                                    _lib.ptr(out[PAD:]), B, T, PAD, C_, _lib.stream_ptr()),
                   "eegx_nct_to_rows_bf16")
        ctx.shape = (B, C_, T)
        return out

    @staticmethod
    def backward(ctx, dg):
        B, C_, T = ctx.shape
        Tp = T + 2 * PAD
        return dg[PAD:PAD + B * Tp].view(B, Tp, C_)[:, PAD:PAD + T].transpose(1, 2).float()


def to_rows(x: torch.Tensor) -> torch.Tensor:
    return _ToRows.apply(x)


def _bn_stats(yg, bn, B, T, training):
    """(mean, rstd) of a BatchNorm1d over the valid rows of a raw conv output (guarded tensor)."""
    C_ = yg.shape[1]
    if not training:
        return _f32(bn.running_mean), torch.rsqrt(bn.running_var.detach().float() + bn.eps).contiguous()
    mean = torch.empty(C_, dtype=torch.float32, device=yg.device)
    rstd = torch.empty_like(mean)
    lib = _lib.lib()
    ws = _workspace(lib.eegx_colreduce_workspace_bytes(C_), yg.device)
    track = bn.track_running_stats and bn.running_mean is not None
    _lib.check(lib.eegx_bn_stats_bf16(
        _lib.ptr(yg[PAD:]), 1, B, T, PAD, C_, float(bn.eps), _lib.ptr(mean), _lib.ptr(rstd),
        _lib.ptr(bn.running_mean) if track else None, _lib.ptr(bn.running_var) if track else None, 0,
        float(bn.momentum if bn.momentum is not None else 0.1), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
        "eegx_bn_stats_bf16")
    if track:
        bn.num_batches_tracked += 1
    return mean, rstd


class _BnAct(torch.autograd.Function):
    """zero_pad(dropout(gelu(bn_a(ya) + residual))) on guarded tensors."""

    @staticmethod
    def forward(ctx, ya, gamma_a, beta_a, yr, gamma_r, beta_r, bn_a, bn_r, res_mode, B, T, training, rng, site, p):
        _need(ya, "ya")
        C_ = ya.shape[1]
        mean_a, rstd_a = _bn_stats(ya, bn_a, B, T, training)
        mean_r = rstd_r = None
        if res_mode == 2:
            _need(yr, "yr")
            mean_r, rstd_r = _bn_stats(yr, bn_r, B, T, training)
        elif res_mode == 1:
            _need(yr, "yr")
        out = torch.empty_like(ya)
        ga, ba = _f32(gamma_a), _f32(beta_a)
        gr = _f32(gamma_r) if res_mode == 2 else None
        br = _f32(beta_r) if res_mode == 2 else None
        _lib.check(_lib.lib().eegx_bn_act_fwd_bf16(
            _lib.ptr(ya[PAD:]), _lib.ptr(mean_a), _lib.ptr(rstd_a), _lib.ptr(ga), _lib.ptr(ba),
            _lib.ptr(yr[PAD:]) if res_mode else None, _lib.ptr(mean_r), _lib.ptr(rstd_r), _lib.ptr(gr), _lib.ptr(br),
            res_mode, _lib.ptr(out[PAD:]), 1, B, T, PAD, C_, _lib.ptr(rng), site, p, _lib.stream_ptr()),
            "eegx_bn_act_fwd_bf16")
        ctx.save_for_backward(ya, gamma_a, beta_a, yr if res_mode else None, gamma_r if res_mode == 2 else None,
                              beta_r if res_mode == 2 else None, mean_a, rstd_a, mean_r, rstd_r)
        ctx.cfg = (res_mode, B, T, bool(training), rng, site, p)
        return out

    @staticmethod
    def backward(ctx, dout):
        ya, gamma_a, beta_a, yr, gamma_r, beta_r, mean_a, rstd_a, mean_r, rstd_r = ctx.saved_tensors
        res_mode, B, T, training, rng, site, p = ctx.cfg
        C_ = ya.shape[1]
        dout = dout.contiguous()
        da = torch.empty_like(ya)
        dr = torch.empty_like(ya) if res_mode else None
        sums = torch.empty(3, C_, dtype=torch.float32, device=ya.device)
        lib = _lib.lib()
        ws = _workspace(lib.eegx_colreduce_workspace_bytes(C_), ya.device)
        gr = _f32(gamma_r) if res_mode == 2 else None
        br = _f32(beta_r) if res_mode == 2 else None
        _lib.check(lib.eegx_bn_act_bwd_bf16(
            _lib.ptr(dout[PAD:]), _lib.ptr(ya[PAD:]), _lib.ptr(mean_a), _lib.ptr(rstd_a), _lib.ptr(_f32(gamma_a)),
            _lib.ptr(_f32(beta_a)), _lib.ptr(yr[PAD:]) if res_mode else None, _lib.ptr(mean_r), _lib.ptr(rstd_r),
            _lib.ptr(gr), _lib.ptr(br), res_mode, int(training), _lib.ptr(da[PAD:]),
            _lib.ptr(dr[PAD:]) if res_mode else None, _lib.ptr(sums), _lib.ptr(ws), ws.numel(), 1, B, T, PAD, C_,
            _lib.ptr(rng), site, p, _lib.stream_ptr()), "eegx_bn_act_bwd_bf16")
        targets = [(gamma_a, sums[1]), (beta_a, sums[0])] + ([(gamma_r, sums[2]), (beta_r, sums[0])] if res_mode == 2 else [])
        bufs = [grad_buffer(p_) for p_, _ in targets]
        if all(b is not None for b in bufs):       # one multi-tensor add straight into the gradient buffers
            torch._foreach_add_(bufs, [v for _, v in targets])
            return da, None, None, dr, None, None, None, None, None, None, None, None, None, None, None
        dga, dba = sums[1].to(gamma_a.dtype), sums[0].to(beta_a.dtype)
        dgr = sums[2].to(gamma_r.dtype) if res_mode == 2 else None
        dbr = sums[0].to(beta_r.dtype) if res_mode == 2 else None
        return da, dga, dba, dr, dgr, dbr, None, None, None, None, None, None, None, None, None


def bn_act(ya, bn_a, yr, bn_r, B: int, T: int, p: float = 0.0, training: bool = True,
           drop_training: Optional[bool] = None) -> torch.Tensor:
    """gelu(bn_a(ya) + res) with res = bn_r(yr) (bn_r given), yr (identity) or nothing; then dropout(p)
    and zeroing of the padding rows.  ya / yr: raw conv outputs as guarded tensors."""
    res_mode = 0 if yr is None else (2 if bn_r is not None else 1)
    rng, site, p = _drop(p, training if drop_training is None else drop_training, ya.device)
    return _BnAct.apply(ya, bn_a.weight, bn_a.bias, yr, bn_r.weight if bn_r is not None else None,
                        bn_r.bias if bn_r is not None else None, bn_a, bn_r, res_mode, B, T, training, rng, site, p)


class _DwConv5(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xg, weight, bias, B, T):
        _need(xg, "xg")
        C_ = xg.shape[1]
        out = torch.empty_like(xg)
        w = _f32(weight).reshape(C_, 5)
        _lib.check(_lib.lib().eegx_dwconv5_fwd_bf16(_lib.ptr(xg[PAD:]), _lib.ptr(w), _lib.ptr(_f32(bias)),
                                                    _lib.ptr(out[PAD:]), 1, B, T, PAD, C_, _lib.stream_ptr()),
                   "eegx_dwconv5_fwd_bf16")
        ctx.save_for_backward(xg, weight, bias)
        ctx.cfg = (B, T)
        return out

    @staticmethod
    def backward(ctx, dout):
        xg, weight, bias = ctx.saved_tensors
        B, T = ctx.cfg
        C_ = xg.shape[1]
        dout = dout.contiguous()
        dx = torch.empty_like(xg)
        dwdb = torch.empty(6, C_, dtype=torch.float32, device=xg.device)
        lib = _lib.lib()
        ws = _workspace(lib.eegx_colreduce_workspace_bytes(C_), xg.device)
        w = _f32(weight).reshape(C_, 5)
        _lib.check(lib.eegx_dwconv5_bwd_bf16(_lib.ptr(dout[PAD:]), _lib.ptr(xg[PAD:]), _lib.ptr(w), _lib.ptr(dx[PAD:]),
                                             _lib.ptr(dwdb), _lib.ptr(ws), ws.numel(), 1, B, T, PAD, C_,
                                             _lib.stream_ptr()), "eegx_dwconv5_bwd_bf16")
        dw = dwdb[:5].t().reshape(weight.shape).to(weight.dtype)
        return dx, dw, dwdb[5].to(bias.dtype), None, None


def dwconv5(xg, weight, bias, B: int, T: int) -> torch.Tensor:
    """Depthwise Conv1d(k=5, groups=C) on a guarded tensor; weight (C, 1, 5)."""
    return _DwConv5.apply(xg, weight, bias, B, T)


class _GroupMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xg, B, T):
        _need(xg, "xg")
        C_ = xg.shape[1]
        s = torch.empty(B, C_, dtype=torch.float32, device=xg.device)
        _lib.check(_lib.lib().eegx_group_mean_bf16(_lib.ptr(xg[PAD:]), _lib.ptr(s), B, T, PAD, C_, _lib.stream_ptr()),
                   "eegx_group_mean_bf16")
        ctx.cfg = (B, T, tuple(xg.shape))
        return s

    @staticmethod
    def backward(ctx, ds):
        B, T, shape = ctx.cfg
        ds = ds.float().contiguous()
        dx = torch.zeros(shape, dtype=torch.bfloat16, device=ds.device)
        _lib.check(_lib.lib().eegx_group_mean_bwd_bf16(_lib.ptr(ds), _lib.ptr(dx[PAD:]), B, T, PAD, shape[1], 0,
                                                       _lib.stream_ptr()), "eegx_group_mean_bwd_bf16")
        return dx, None, None


def group_mean(xg, B: int, T: int) -> torch.Tensor:
    """mean over the T valid rows of every trial: guarded (rows, C) bf16 -> (B, C) fp32."""
    return _GroupMean.apply(xg, B, T)


class _SEScale(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xg, e, B, T, rng, site, p):
        _need(xg, "xg"); _need(e, "e", torch.float32)
        C_ = xg.shape[1]
        out = torch.empty(B * T, C_, dtype=torch.bfloat16, device=xg.device)
        _lib.check(_lib.lib().eegx_se_scale_fwd_bf16(_lib.ptr(xg[PAD:]), _lib.ptr(e), _lib.ptr(out), B, T, PAD, C_,
                                                     _lib.ptr(rng), site, p, _lib.stream_ptr()),
                   "eegx_se_scale_fwd_bf16")
        ctx.save_for_backward(xg, e)
        ctx.cfg = (B, T, rng, site, p)
        return out

    @staticmethod
    def backward(ctx, dout):
        xg, e = ctx.saved_tensors
        B, T, rng, site, p = ctx.cfg
        C_ = xg.shape[1]
        dout = dout.contiguous()
        dx = torch.zeros_like(xg)
        de = torch.empty_like(e)
        _lib.check(_lib.lib().eegx_se_scale_bwd_bf16(_lib.ptr(dout), _lib.ptr(xg[PAD:]), _lib.ptr(e), _lib.ptr(dx[PAD:]),
                                                     _lib.ptr(de), B, T, PAD, C_, _lib.ptr(rng), site, p,
                                                     _lib.stream_ptr()), "eegx_se_scale_bwd_bf16")
        return dx, de, None, None, None, None, None


def se_scale(xg, e, B: int, T: int, p: float = 0.0, training: bool = True) -> torch.Tensor:
    """dropout(x * e[b, :]) -> compact (B*T, C) bf16 rows; e: (B, C) fp32."""
    rng, site, p = _drop(p, training, xg.device)
    return _SEScale.apply(xg, e.float().contiguous(), B, T, rng, site, p)


# ------------------------------------------------------------------------------------------ attention core
ATTN_MAX_S = 64
ATTN_HEAD_DIMS = (64, 96, 128, 192)


def attn_supported(Sq: int, Sk: int, hd: int) -> bool:
    """Sequences up to ATTN_MAX_S tokens run the one-CTA-per-(batch, head) kernel, longer ones the flash kernel
    (keys / values streamed in tiles, scores never in HBM); both need a head_dim in ATTN_HEAD_DIMS."""
    return Sq >= 1 and Sk >= 1 and hd in ATTN_HEAD_DIMS


def _attn_is_flash(Sq: int, Sk: int) -> bool:
    return Sq > ATTN_MAX_S or Sk > ATTN_MAX_S


def _attn_desc(B, H, Sq, Sk, hd, q, k, v, o, causal):
    d = _lib.AttnDesc()
    d.B, d.H, d.Sq, d.Sk, d.hd = B, H, Sq, Sk, hd
    d.q_rs, d.k_rs, d.v_rs, d.o_rs = q.stride(0), k.stride(0), v.stride(0), o.stride(0)
    d.causal = int(causal)
    d.scale = 1.0 / math.sqrt(hd)
    return d


class _AttnCore(torch.autograd.Function):
    """softmax(q k^T / sqrt(hd)) -> dropout -> (.) v.  `packed` = (B*S, 3d) self-attention input, or
    q = (B*Sq, d) and kv = (B*Sk, 2d) for cross attention."""

    @staticmethod
    def forward(ctx, q_or_qkv, kv, B, Sq, Sk, H, causal, rng, site, p):
        _need(q_or_qkv, "q")
        if kv is None:
            d_model = q_or_qkv.shape[1] // 3
            q, k, v = q_or_qkv[:, :d_model], q_or_qkv[:, d_model:2 * d_model], q_or_qkv[:, 2 * d_model:]
        else:
            _need(kv, "kv")
            d_model = q_or_qkv.shape[1]
            q, k, v = q_or_qkv, kv[:, :d_model], kv[:, d_model:]
        hd = d_model // H
        o = torch.empty(B * Sq, d_model, dtype=torch.bfloat16, device=q.device)
        lse = torch.empty(B * H * Sq, dtype=torch.float32, device=q.device)
        desc = _attn_desc(B, H, Sq, Sk, hd, q, k, v, o, causal)
        fwd = "eegx_attn_flash_fwd_bf16" if _attn_is_flash(Sq, Sk) else "eegx_attn_fwd_bf16"
        _lib.check(getattr(_lib.lib(), fwd)(C.byref(desc), _lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(o),
                                            _lib.ptr(lse), _lib.ptr(rng), site, p, _lib.stream_ptr()), fwd)
        ctx.save_for_backward(q_or_qkv, kv, o, lse)
        ctx.cfg = (B, Sq, Sk, H, causal, rng, site, p)
        return o

    @staticmethod
    def backward(ctx, do):
        q_or_qkv, kv, o, lse = ctx.saved_tensors
        B, Sq, Sk, H, causal, rng, site, p = ctx.cfg
        do = do.contiguous()
        if kv is None:
            d_model = q_or_qkv.shape[1] // 3
            q, k, v = q_or_qkv[:, :d_model], q_or_qkv[:, d_model:2 * d_model], q_or_qkv[:, 2 * d_model:]
            dpacked = torch.empty_like(q_or_qkv)
            dq, dk, dv = dpacked[:, :d_model], dpacked[:, d_model:2 * d_model], dpacked[:, 2 * d_model:]
            dkv = None
        else:
            d_model = q_or_qkv.shape[1]
            q, k, v = q_or_qkv, kv[:, :d_model], kv[:, d_model:]
            dpacked = torch.empty_like(q_or_qkv)
            dkv = torch.empty_like(kv)
            dq, dk, dv = dpacked, dkv[:, :d_model], dkv[:, d_model:]
        hd = d_model // H
        desc = _attn_desc(B, H, Sq, Sk, hd, q, k, v, o, causal)
        if _attn_is_flash(Sq, Sk):
            dsum = torch.empty_like(lse)                 # D_i = dO_i . O_i, written by the dQ kernel
            _lib.check(_lib.lib().eegx_attn_flash_bwd_bf16(
                C.byref(desc), _lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(o), _lib.ptr(do), _lib.ptr(lse),
                _lib.ptr(dsum), _lib.ptr(dq), _lib.ptr(dk), _lib.ptr(dv), dq.stride(0), dk.stride(0), dv.stride(0),
                _lib.ptr(rng), site, p, _lib.stream_ptr()), "eegx_attn_flash_bwd_bf16")
            return dpacked, dkv, None, None, None, None, None, None, None, None
        _lib.check(_lib.lib().eegx_attn_bwd_bf16(
            C.byref(desc), _lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(o), _lib.ptr(do), _lib.ptr(lse),
            _lib.ptr(dq), _lib.ptr(dk), _lib.ptr(dv), dq.stride(0), dk.stride(0), dv.stride(0), _lib.ptr(rng), site, p,
            _lib.stream_ptr()), "eegx_attn_bwd_bf16")
        return dpacked, dkv, None, None, None, None, None, None, None, None


def attn_self(qkv: torch.Tensor, B: int, S: int, H: int, p: float = 0.0, training: bool = True,
              causal: bool = False) -> torch.Tensor:
    """qkv: (B*S, 3*d) bf16 packed in-projection output -> (B*S, d)."""
    rng, site, p = _drop(p, training, qkv.device)
    return _AttnCore.apply(qkv.contiguous(), None, B, S, S, H, causal, rng, site, p)


def attn_cross(q: torch.Tensor, kv: torch.Tensor, B: int, Sq: int, Sk: int, H: int, p: float = 0.0,
               training: bool = True) -> torch.Tensor:
    """q: (B*Sq, d), kv: (B*Sk, 2*d) -> (B*Sq, d)."""
    rng, site, p = _drop(p, training, q.device)
    return _AttnCore.apply(q.contiguous(), kv.contiguous(), B, Sq, Sk, H, False, rng, site, p)


# ------------------------------------------------------------------------------------------------
# Gradient-ready boundaries (data-parallel overlap).  An identity in forward; in backward it tells
# the registered callback that every parameter used DOWNSTREAM of this point has its gradient
# complete (all those nodes were created later, so autograd ran them first, and our backward
# functions accumulate weight gradients inside the same node).  trainer.EEGTrainer uses it to
# start the all-reduce of that slice of the flat gradient buffer while backward continues.
_BOUNDARY_CB = None


def set_grad_boundary_callback(cb):
    global _BOUNDARY_CB
    _BOUNDARY_CB = cb


class _GradBoundary(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, key):
        ctx.key = key
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        cb = _BOUNDARY_CB
        if cb is not None:
            cb(ctx.key)
        return g, None


def grad_boundary(x, key):
    if _BOUNDARY_CB is None or not torch.is_grad_enabled() or not x.requires_grad:
        return x
    return _GradBoundary.apply(x, key)
