"""Pins the CPU oracle (oracle/) against the committed golden vectors.

normalize_ref.npz was produced by the reference's own EEGDataset methods;
dsp_spec.npz by scipy.signal.firwin + F.conv1d + torch.stft in float64
(tests/golden/make_golden.py).  CPU only.
"""
import os

import numpy as np
import pytest

from oracle import preprocess_oracle as po

NAMES = po.REGION_ORDER


@pytest.fixture(scope="module")
def norm(golden_dir):
    return np.load(os.path.join(golden_dir, "normalize_ref.npz"))


@pytest.fixture(scope="module")
def dsp(golden_dir):
    return np.load(os.path.join(golden_dir, "dsp_spec.npz"))


def test_region_indices_match_reference_counts(norm):
    # frontal 16 / temporal 9 / central 11 / parietal 12 of 125
    assert [len(norm[f"idx_{n}"]) for n in NAMES] == [16, 9, 11, 12]


def test_robust_normalisation_matches_reference(norm):
    idx = {n: norm[f"idx_{n}"] for n in NAMES}
    cen = {n: norm[f"center_{n}"] for n in NAMES}
    sca = {n: norm[f"scale_{n}"] for n in NAMES}
    for i in range(norm["trials"].shape[0]):
        regs = po.normalize_regions(norm["trials"][i], idx, cen, sca)
        for r, n in enumerate(NAMES):
            ref = norm[f"robust_{i}_{n}"]
            assert regs[r].shape == ref.shape and regs[r].dtype == np.float32
            assert np.isfinite(regs[r]).all()
            # 3 float32 flops; sklearn rounds through float64 in places
            assert po.rel_max_err(regs[r], ref) <= 1e-6


def test_fallback_zscore_matches_reference(norm):
    idx = {n: norm[f"idx_{n}"] for n in NAMES}
    for i in range(norm["trials"].shape[0]):
        regs = po.normalize_regions(norm["trials"][i], idx)
        for r, n in enumerate(NAMES):
            assert po.rel_max_err(regs[r], norm[f"fallback_{i}_{n}"]) <= 2e-6


def test_robust_scaler_fit_matches_reference(norm):
    fit = norm["fit_samples"]                       # (n, 1, 125, T) as pickled
    for n in NAMES:
        region = np.stack([po.process_raw_eeg(s)[norm[f"idx_{n}"]] for s in fit])
        center, scale = po.robust_scaler_fit(region)
        np.testing.assert_allclose(center, norm[f"center_{n}"], rtol=1e-6)
        np.testing.assert_allclose(scale, norm[f"scale_{n}"], rtol=1e-6)


def test_nan_to_num_values():
    x = np.array([[np.nan, np.inf, -np.inf, 1.5]], dtype=np.float32)
    np.testing.assert_array_equal(po.process_raw_eeg(x[None]), [[0.0, 10.0, -10.0, 1.5]])


def test_firwin_restatement_matches_scipy(dsp):
    np.testing.assert_allclose(po.firwin_bandpass(65, 8.0, 30.0, 256.0), dsp["taps"], rtol=0, atol=1e-15)


@pytest.mark.parametrize("key", ["a", "b", "c"])
def test_dsp_oracle_matches_spec_calls(dsp, key):
    n_fft, hop = (int(v) for v in dsp[f"cfg_{key}"])
    x = dsp[f"x_{key}"]
    st = po.dsp_reference(x, dsp["taps"], n_fft=n_fft, hop=hop, return_stages=True)
    assert st["z"].shape == dsp[f"z_{key}"].shape
    assert po.rel_max_err(st["fir"][:, :1], dsp[f"fir_{key}"]) <= 1e-13
    assert po.rel_max_err(st["z"], dsp[f"z_{key}"]) <= 1e-10


def test_dsp_float32_cpu_within_stated_tolerance(dsp):
    """The <=1e-5 max-norm-relative bar is attainable in float32."""
    import torch
    x = torch.from_numpy(dsp["x_a"])
    h = torch.from_numpy(dsp["taps"]).float()
    z32 = po.dsp_torch_cpu_f32(x, h).numpy()
    assert po.rel_max_err(z32, dsp["z_a"]) <= 1e-5


def test_window_trials_and_layout():
    rec = np.arange(2 * 100, dtype=np.float64).reshape(2, 100)
    x = po.window_trials(rec, [0, 10, 50], 32)
    assert x.shape == (3, 2, 32) and x[1, 1, 0] == 110
    z = np.arange(2 * 4 * 3 * 5, dtype=np.float64).reshape(2, 4, 3, 5)
    regs = po.to_encoder_layout(z, [slice(0, 2), slice(2, 4)])
    assert regs[0].shape == (2, 6, 5) and regs[1][0, 0, 0] == z[0, 2, 0, 0]
