"""GPU parity of the reference-actual normalisation (through the C ABI) against
the golden vectors produced by the reference's EEGDataset and against the oracle."""
import os

import numpy as np
import pytest
import torch

import imagined_speech_translation_b200 as pkg
from oracle import preprocess_oracle as po

pytestmark = pytest.mark.gpu
NAMES = po.REGION_ORDER


@pytest.fixture(scope="module")
def norm(golden_dir):
    return np.load(os.path.join(golden_dir, "normalize_ref.npz"))


def _batch(norm):
    return torch.from_numpy(norm["trials"][:, 0]).cuda().contiguous()   # (3, 125, T)


def test_robust_matches_reference_golden(norm):
    idx = {n: norm[f"idx_{n}"] for n in NAMES}
    cen = {n: norm[f"center_{n}"] for n in NAMES}
    sca = {n: norm[f"scale_{n}"] for n in NAMES}
    rn = pkg.RegionNormalizer(idx, cen, sca)
    regs = rn(_batch(norm))
    assert [tuple(r.shape[1:]) for r in regs] == [(16, 96), (9, 96), (11, 96), (12, 96)]
    for r, n in enumerate(NAMES):
        got = regs[r].cpu().numpy()
        assert regs[r].is_contiguous()
        for i in range(got.shape[0]):
            # fp32, 3 flops per element: <= 1e-6 inf-norm relative
            assert po.rel_max_err(got[i], norm[f"robust_{i}_{n}"]) <= 1e-6


def test_fallback_matches_reference_golden(norm):
    idx = {n: norm[f"idx_{n}"] for n in NAMES}
    rn = pkg.RegionNormalizer(idx)
    regs = rn(_batch(norm))
    for r, n in enumerate(NAMES):
        got = regs[r].cpu().numpy()
        for i in range(got.shape[0]):
            assert po.rel_max_err(got[i], norm[f"fallback_{i}_{n}"]) <= 2e-6


def test_mixed_scalers(norm):
    """A region without a scaler takes the fallback while the others are robust-scaled
    (dataset.py:210-216 decides per region)."""
    idx = {n: norm[f"idx_{n}"] for n in NAMES}
    cen = {n: norm[f"center_{n}"] for n in NAMES if n != "central"}
    sca = {n: norm[f"scale_{n}"] for n in NAMES if n != "central"}
    regs = pkg.RegionNormalizer(idx, cen, sca)(_batch(norm))
    assert po.rel_max_err(regs[2][1].cpu().numpy(), norm["fallback_1_central"]) <= 2e-6
    assert po.rel_max_err(regs[0][1].cpu().numpy(), norm["robust_1_frontal"]) <= 1e-6


@pytest.mark.parametrize("B,C_in,C_out,T", [(5, 125, 48, 1651), (4, 64, 64, 2048), (2, 7, 3, 2), (1, 3, 3, 5)])
def test_dense_vs_oracle(B, C_in, C_out, T):
    rng = np.random.default_rng(B * 1000 + T)
    x = (25.0 * rng.standard_normal((B, C_in, T))).astype(np.float32)
    flat = x.reshape(-1)
    bad = rng.choice(flat.size, size=min(flat.size - flat.size % 3, max(3, flat.size // 1000 // 3 * 3)), replace=False)
    flat[bad[0::3]] = np.nan
    flat[bad[1::3]] = np.inf
    flat[bad[2::3]] = -np.inf
    idx = rng.choice(C_in, size=C_out, replace=False).astype(np.int32)
    center = rng.uniform(-5, 5, C_out).astype(np.float32)
    scale = rng.uniform(20, 90, C_out).astype(np.float32)
    got = pkg.normalize_dense(torch.from_numpy(x).cuda(), torch.from_numpy(idx).cuda(),
                              torch.from_numpy(center).cuda(), torch.from_numpy(scale).cuda())
    torch.cuda.synchronize()
    for b in range(B):
        ref = po.robust_scale(po.process_raw_eeg(x[b])[idx], center, scale)
        got_b = got[b].cpu().numpy()
        assert np.isfinite(got_b).all()
        assert po.rel_max_err(got_b, ref) <= 1e-6


def test_gather_only_is_bit_exact():
    x = torch.randn(3, 10, 256, device="cuda")
    x[0, 1, 5] = float("nan"); x[1, 2, 7] = float("inf"); x[2, 9, 0] = float("-inf")
    idx = torch.tensor([9, 2, 1, 0], dtype=torch.int32, device="cuda")
    got = pkg.normalize_dense(x, idx, None, None)
    ref = torch.nan_to_num(x[:, idx.long()], nan=0.0, posinf=10.0, neginf=-10.0)
    assert torch.equal(got, ref)


def test_empty_batch():
    out = pkg.normalize_dense(torch.empty(0, 4, 16, device="cuda"), None, None, None)
    assert out.shape == (0, 4, 16)


def test_rejects_cpu_and_wrong_dtype():
    with pytest.raises(pkg.EegxError):
        pkg.normalize_dense(torch.zeros(1, 2, 8), None, None, None)
    with pytest.raises(ValueError):
        pkg.normalize_dense(torch.zeros(1, 2, 8, device="cuda", dtype=torch.float64), None, None, None)
