"""GPU parity of the tcgen05 bf16 GEMM (through the C ABI) against a plain PyTorch fp32
reference of the same contraction on the same bf16-rounded inputs.

Tolerances (stated): fp32 output  max|d - ref| / max|ref| <= 2e-5 * sqrt(K/64)  (fp32 accumulation
order only); bf16 output <= 2^-8 (one bf16 rounding of the result)."""
import math

import pytest
import torch

from imagined_speech_translation_b200 import ops

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def _mk(shape, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(shape, generator=g, device="cuda").to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K", [
    (128, 128, 64), (128, 256, 128), (256, 64, 64), (384, 128, 512),
    (9472, 768, 768), (1000, 200, 136), (37, 48, 768), (130, 72, 8), (4096, 3072, 768),
])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_gemm_k_major(M, N, K, out_dtype):
    a, b = _mk((M, K), 1), _mk((N, K), 2)
    d = ops.gemm(a, b, out_dtype=out_dtype)
    ref = a.float() @ b.float().t()
    tol = 2e-5 * math.sqrt(max(K, 64) / 64) if out_dtype == torch.float32 else 2 ** -8
    assert d.shape == (M, N) and d.dtype == out_dtype
    assert _rel(d, ref) <= tol


@pytest.mark.parametrize("block_n", [64, 128, 192, 256])
def test_gemm_tile_widths(block_n):
    a, b = _mk((512, 320), 3), _mk((512, 320), 4)
    d = ops.gemm(a, b, out_dtype=torch.float32, force_block_n=block_n)
    assert _rel(d, a.float() @ b.float().t()) <= 5e-5


@pytest.mark.parametrize("a_mn,b_mn", [(False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 384, 192), (768, 3072, 9472), (200, 136, 1000)])
def test_gemm_mn_major_operands(a_mn, b_mn, M, N, K):
    a = _mk((K, M) if a_mn else (M, K), 5)
    b = _mk((K, N) if b_mn else (N, K), 6)
    d = ops.gemm(a, b, a_mn_major=a_mn, b_mn_major=b_mn, out_dtype=torch.float32)
    A = a.float().t() if a_mn else a.float()
    Bt = b.float() if b_mn else b.float().t()
    assert _rel(d, A @ Bt) <= 2e-5 * math.sqrt(max(K, 64) / 64)


def test_gemm_linear_layer_triple():
    """forward / dgrad / wgrad of y = x W^T + b with no transposes in HBM."""
    M, K, N = 1184, 768, 1536
    x, w, dy = _mk((M, K), 7), _mk((N, K), 8), _mk((M, N), 9)
    bias = torch.randn(N, device="cuda")
    y = ops.gemm(x, w, bias, out_dtype=torch.float32)
    assert _rel(y, x.float() @ w.float().t() + bias) <= 1e-4
    dx = ops.gemm(dy, w, b_mn_major=True, out_dtype=torch.float32)          # dy (M,N) @ W (N,K)
    assert _rel(dx, dy.float() @ w.float()) <= 1e-4
    dw = ops.gemm(dy, x, a_mn_major=True, b_mn_major=True, out_dtype=torch.float32)   # dy^T @ x
    assert _rel(dw, dy.float().t() @ x.float()) <= 1e-4
    dw2 = ops.gemm(dy, x, a_mn_major=True, b_mn_major=True, out=dw.clone(), accumulate=True)
    assert _rel(dw2, 2 * (dy.float().t() @ x.float())) <= 1e-4


def test_gemm_bias_gelu_alpha_epilogue():
    a, b = _mk((300, 256), 10), _mk((520, 256), 11)
    bias = torch.randn(520, device="cuda")
    d = ops.gemm(a, b, bias, gelu=True, alpha=0.125, out_dtype=torch.float32)
    ref = torch.nn.functional.gelu(0.125 * (a.float() @ b.float().t()) + bias)
    assert _rel(d, ref) <= 1e-5


def test_gemm_batched_and_strided():
    a, b = _mk((4, 200, 128), 12), _mk((4, 96, 128), 13)
    d = ops.gemm(a, b, out_dtype=torch.float32)
    assert _rel(d, torch.bmm(a.float(), b.float().transpose(1, 2))) <= 3e-5
    # leading dimension larger than the row (a view into a wider buffer)
    wide = _mk((256, 512), 14)
    a2 = wide[:, 64:64 + 256]
    d2 = ops.gemm(a2, b[0, :, :].repeat(1, 2).contiguous(), out_dtype=torch.float32)
    assert _rel(d2, a2.float() @ b[0].repeat(1, 2).float().t()) <= 3e-5


def test_gemm_rejects_bad_arguments():
    a, b = _mk((128, 64), 1), _mk((128, 64), 2)
    with pytest.raises(ValueError):
        ops.gemm(a.float(), b)
    with pytest.raises(ValueError):
        ops.gemm(a, _mk((128, 72), 3))
    from imagined_speech_translation_b200 import EegxError
    with pytest.raises(EegxError):
        ops.gemm(_mk((128, 70), 1)[:, :64], b)      # lda = 70: not a multiple of 8


def test_gemm_grouped_forward_with_per_group_bias():
    """G problem sets in one launch (the four region encoders' identical layers): own weights, own bias."""
    G, M, K, N = 4, 592, 768, 384
    x, w = _mk((G, M, K), 20), _mk((G, N, K), 21)
    bias = torch.randn(G, N, device="cuda")
    y = ops.gemm(x, w, bias, grouped=True, out_dtype=torch.float32)
    ref = torch.bmm(x.float(), w.float().transpose(1, 2)) + bias[:, None, :]
    assert y.shape == (G, M, N) and _rel(y, ref) <= 1e-4
    yg = ops.gemm(x, w, bias, grouped=True, gelu=True)                       # bf16 out, GELU epilogue
    assert _rel(yg.float(), torch.nn.functional.gelu(ref)) <= 2 ** -7
    dx = ops.gemm(y.to(torch.bfloat16), w, b_mn_major=True, grouped=True, out_dtype=torch.float32)
    assert _rel(dx, torch.bmm(y.to(torch.bfloat16).float(), w.float())) <= 1e-4


@pytest.mark.parametrize("S", [1, 2, 4])
def test_gemm_grouped_split_k_weight_gradient(S):
    """(G, S) problems: the weight gradient of G layers with the reduction split into S chunks; partials land in
    an (S, G, N, K) buffer (strided `out` view) and accumulate into a (G, N, K) gradient buffer when S = 1."""
    G, R, N, K = 4, 1024, 256, 192
    dy, x = _mk((G, R, N), 22), _mk((G, R, K), 23)
    ref = torch.bmm(dy.float().transpose(1, 2), x.float())                  # (G, N, K)
    dy4 = dy.view(G, S, R // S, N)
    x4 = x.view(G, S, R // S, K)
    part = torch.empty(S, G, N, K, device="cuda")
    out = ops.gemm(dy4, x4, a_mn_major=True, b_mn_major=True, out=part.permute(1, 0, 2, 3))
    assert out.data_ptr() == part.data_ptr()
    assert _rel(part.sum(0), ref) <= 1e-4
    if S == 1:
        acc = ref.clone()
        ops.gemm(dy, x, a_mn_major=True, b_mn_major=True, grouped=True, out=acc, accumulate=True)
        assert _rel(acc, 2 * ref) <= 1e-4


def test_gemm_grouped_overlapping_rows_conv():
    """Grouped implicit-im2col: group stride = rows * C_in of ONE guarded buffer holding the G region activations."""
    G, M, Cin, k, Cout = 4, 328, 128, 7, 256
    buf = _mk((G * M + 8, Cin), 24)
    w = _mk((G, Cout, k * Cin), 25)
    a = buf.as_strided((G, M, k * Cin), (M * Cin, Cin, 1), (4 - k // 2) * Cin)
    y = ops.gemm(a, w, grouped=True, out_dtype=torch.float32)
    for g in range(G):
        rows = torch.stack([buf[4 - k // 2 + g * M + m: 4 - k // 2 + g * M + m + k].reshape(-1) for m in range(M)])
        assert _rel(y[g], rows.float() @ w[g].float().t()) <= 1e-4
