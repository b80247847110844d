"""Thin Python entry points over the C ABI for the encoder's dense contractions."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib

EPI_NONE, EPI_BIAS, EPI_BIAS_GELU = 0, 1, 2

# Measurement hook (bench.py): when a list is installed here, every gemm() call is bracketed by
# CUDA events on the launching stream and (start, end, flops) is appended.
GEMM_TIMING = None
LAUNCHES = {"gemm": 0}


def _check_bf16(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise _lib.EegxError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != torch.bfloat16:
        raise ValueError(f"{name} must be bfloat16, got {t.dtype}")
    if t.stride(-1) != 1:
        raise ValueError(f"{name} must have a unit innermost stride")


def gemm(a: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor] = None, *,
         a_mn_major: bool = False, b_mn_major: bool = False, out: Optional[torch.Tensor] = None,
         out_dtype: torch.dtype = torch.bfloat16, gelu: bool = False, accumulate: bool = False,
         alpha: float = 1.0, force_block_n: int = 0, grouped: bool = False) -> torch.Tensor:
    """D = alpha * A @ B^T (+ bias) (+ GELU) (+ D) on tcgen05 tensor cores.

    a: (M, K) [or (K, M) if a_mn_major], b: (N, K) [or (K, N) if b_mn_major].  Leading dims on both:
    3-D = a batch of problems (split-K chunks), or -- with ``grouped`` -- G problem sets with their own weights
    and bias (G, N); 4-D = (G, S, ., .), S problems in each of G groups.  bf16 in, fp32 accumulate, bf16 or fp32
    out, row-major (.., M, N) (``out`` may be any view with unit innermost stride and uniform leading strides).
    """
    _check_bf16(a, "a")
    _check_bf16(b, "b")
    if a.dim() != b.dim() or a.dim() not in (2, 3, 4):
        raise ValueError("a and b must both be 2-D, 3-D or 4-D")
    if grouped and a.dim() != 3:
        raise ValueError("grouped=True takes 3-D operands (G, rows, cols); 4-D operands are grouped by definition")
    lead = tuple(a.shape[:-2])
    if tuple(b.shape[:-2]) != lead:
        raise ValueError("leading (batch / group) sizes differ")
    if a.dim() == 4:
        groups, batch = lead
        sa, sb = (a.stride(0), a.stride(1)), (b.stride(0), b.stride(1))
    elif a.dim() == 3 and grouped:
        groups, batch = lead[0], 1
        sa, sb = (a.stride(0), 0), (b.stride(0), 0)
    elif a.dim() == 3:
        groups, batch = 1, lead[0]
        sa, sb = (0, a.stride(0)), (0, b.stride(0))
    else:
        groups, batch, sa, sb = 1, 1, (0, 0), (0, 0)
    ar, ac = a.shape[-2], a.shape[-1]
    br, bc = b.shape[-2], b.shape[-1]
    M, K = (ac, ar) if a_mn_major else (ar, ac)
    N, Kb = (bc, br) if b_mn_major else (br, bc)
    if K != Kb:
        raise ValueError(f"contraction sizes differ: {K} vs {Kb}")
    if out is None:
        out = torch.empty(lead + (M, N), dtype=out_dtype, device=a.device)
    else:
        if out.dtype not in (torch.bfloat16, torch.float32) or out.stride(-1) != 1:
            raise ValueError("out must be bf16 or fp32 with unit innermost stride")
        if tuple(out.shape) != lead + (M, N):
            raise ValueError(f"out must have shape {lead + (M, N)}, got {tuple(out.shape)}")
        out_dtype = out.dtype
    if a.dim() == 4:
        sd = (out.stride(0), out.stride(1))
    elif a.dim() == 3:
        sd = (out.stride(0), 0) if grouped else (0, out.stride(0))
    else:
        sd = (0, 0)
    d = _lib.GemmDesc()
    d.M, d.N, d.K, d.batch = M, N, K, batch
    d.lda, d.ldb, d.ldd = a.stride(-2), b.stride(-2), out.stride(-2)
    d.stride_a, d.stride_b, d.stride_d = sa[1], sb[1], sd[1]
    d.groups, d.stride_a_g, d.stride_b_g, d.stride_d_g = groups, sa[0], sb[0], sd[0]
    d.a_mn_major, d.b_mn_major = int(a_mn_major), int(b_mn_major)
    d.out_f32 = int(out_dtype == torch.float32)
    d.epilogue = EPI_NONE if bias is None else (EPI_BIAS_GELU if gelu else EPI_BIAS)
    if gelu and bias is None:
        raise ValueError("gelu epilogue needs a bias (pass zeros)")
    d.stride_bias_g = 0
    if bias is not None:
        if bias.dtype != torch.float32 or bias.stride(-1) != 1 or bias.shape[-1] != N:
            raise ValueError("bias must be fp32 with N contiguous entries per problem set")
        if bias.dim() == 2:
            if bias.shape[0] != groups:
                raise ValueError("a (G, N) bias needs G groups")
            d.stride_bias_g = bias.stride(0)
        elif bias.dim() != 1:
            raise ValueError("bias must be (N,) or (G, N)")
    d.accumulate = int(accumulate)
    d.force_block_n = int(force_block_n)
    d.alpha = float(alpha)
    LAUNCHES["gemm"] += 1
    if GEMM_TIMING is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _lib.check(_lib.lib().eegx_gemm_bf16(C.byref(d), _lib.ptr(a), _lib.ptr(b), _lib.ptr(bias),
                                         _lib.ptr(out), _lib.stream_ptr()), "eegx_gemm_bf16")
    if GEMM_TIMING is not None:
        e1.record()
        GEMM_TIMING.append((e0, e1, 2.0 * M * N * K * batch * groups,
                            (batch * groups, M, N, K, int(a_mn_major), int(b_mn_major), out.element_size())))
    return out
