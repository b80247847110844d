"""CPU oracle for the preprocessing half of the hot path (TEST INFRASTRUCTURE).

Two groups of functions:

1. The preprocessing the reference really performs per trial
   (``main_model/src/data/dataset.py:172-225``): float32 cast, ``nan_to_num``,
   region gather, per-channel robust scaling, and the per-channel z-score
   fallback.  Pinned against the reference's own ``EEGDataset`` (golden file
   ``tests/golden/normalize_ref.npz``).

2. The DSP stages named by ``BASELINE.json: north_star`` (trial windowing,
   band-pass FIR, STFT log-power spectrogram, per-channel z-score).  The
   reference contains no such code, so this is a restatement of the written
   spec in SURVEY.md section 8(c), in float64 numpy, pinned against the library
   calls the spec names (``scipy.signal.firwin``, ``F.conv1d``, ``torch.stft``;
   golden file ``tests/golden/dsp_spec.npz``).  PARITY UNPINNED by the
   reference for this group.

Only numpy is needed for the checker functions; ``dsp_torch_cpu_f32`` (the CPU
baseline that bench.py times) uses torch on the host cores.
"""
from __future__ import annotations

import numpy as np

REGION_ORDER = ("frontal", "temporal", "central", "parietal")  # dataset.py:203


# --------------------------------------------------------------------------
# group 1: what the reference does today
# --------------------------------------------------------------------------
def process_raw_eeg(eeg_data) -> np.ndarray:
    """Restates ``EEGDataset._process_raw_eeg`` (dataset.py:172-191).

    float32 cast, squeeze, rank fix-up, then nan -> 0, +inf -> 10, -inf -> -10.
    """
    eeg = np.array(eeg_data, dtype=np.float32).squeeze()
    if eeg.ndim == 1:
        eeg = eeg.reshape(1, -1)
    elif eeg.ndim > 2:
        eeg = eeg.reshape(eeg.shape[0], -1)
    out = eeg.copy()
    out[np.isnan(eeg)] = np.float32(0.0)
    out[np.isposinf(eeg)] = np.float32(10.0)
    out[np.isneginf(eeg)] = np.float32(-10.0)
    return out


def robust_scale(region: np.ndarray, center: np.ndarray, scale: np.ndarray) -> np.ndarray:
    """``RobustScaler.transform(x.T).T`` in closed form (dataset.py:211).

    sklearn does ``X -= center_; X /= scale_`` in place on a float32 array, so
    every intermediate is rounded to float32; this does the same.
    """
    c = np.asarray(center, dtype=np.float32)[:, None]
    s = np.asarray(scale, dtype=np.float32)[:, None]
    return ((region.astype(np.float32) - c) / s).astype(np.float32)


def zscore_time(region: np.ndarray) -> np.ndarray:
    """The scaler-less fallback branch (dataset.py:213-216).

    ``(x - mean_t) / (std_t + 1e-8)`` per channel over time, population std.
    Computed in float64 and rounded once: this is the checker, numpy's own
    float32 pairwise summation is just one of many valid float32 answers.
    """
    x = region.astype(np.float64)
    mean = x.mean(axis=1, keepdims=True)
    std = x.std(axis=1, keepdims=True) + 1e-8
    return ((x - mean) / std).astype(np.float32)


def normalize_regions(eeg_data, region_indices, centers=None, scales=None):
    """Restates ``EEGDataset._normalize_eeg_sample`` (dataset.py:193-225).

    ``region_indices``: dict region -> list of channel rows (dataset.py:339-346).
    ``centers`` / ``scales``: dict region -> (C_r,) arrays (``RobustScaler``'s
    ``center_`` / ``scale_``); a region missing from them takes the z-score
    fallback exactly as the reference does when ``region_name not in scalers``.
    """
    eeg = process_raw_eeg(eeg_data)
    out = []
    for name in REGION_ORDER:
        idx = np.asarray(region_indices[name], dtype=np.int64)
        region = eeg[idx].astype(np.float32)
        if centers is not None and name in centers:
            out.append(robust_scale(region, centers[name], scales[name]))
        else:
            out.append(zscore_time(region))
    return out


def robust_scaler_fit(samples: np.ndarray):
    """Restates the fit of dataset.py:139-147 for one region.

    ``samples``: (n_samples, C_r, T).  Concatenate over time, then per channel
    ``center_`` = median, ``scale_`` = q95 - q5 (zeros in scale become 1, as
    sklearn's ``_handle_zeros_in_scale`` does).
    """
    flat = np.concatenate(list(samples), axis=1).T  # (n*T, C_r)
    center = np.nanmedian(flat, axis=0)
    q = np.nanpercentile(flat, (5.0, 95.0), axis=0)
    scale = q[1] - q[0]
    scale = np.where(scale < 10 * np.finfo(scale.dtype).eps, 1.0, scale)
    return center, scale


# --------------------------------------------------------------------------
# group 2: north_star DSP stages (spec: SURVEY.md section 8(c))
# --------------------------------------------------------------------------
def firwin_bandpass(numtaps=65, low=8.0, high=30.0, fs=256.0) -> np.ndarray:
    """Windowed-sinc band-pass, Hamming window, unit gain at the band centre.

    Restates ``scipy.signal.firwin(numtaps, [low, high], pass_zero=False,
    fs=fs, window='hamming')`` (the call SURVEY.md 8(c) names); pinned against
    scipy in tests/test_oracle.py.  Returns float64 taps.
    """
    nyq = 0.5 * fs
    left, right = low / nyq, high / nyq
    alpha = 0.5 * (numtaps - 1)
    m = np.arange(numtaps, dtype=np.float64) - alpha
    h = right * np.sinc(right * m) - left * np.sinc(left * m)
    n = np.arange(numtaps, dtype=np.float64)
    h *= 0.54 - 0.46 * np.cos(2.0 * np.pi * n / (numtaps - 1))
    centre = 0.5 * (left + right)
    h /= np.sum(h * np.cos(np.pi * m * centre))
    return h


def hann_periodic(n_fft: int) -> np.ndarray:
    """``torch.hann_window(n_fft, periodic=True)`` in float64."""
    n = np.arange(n_fft, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * n / n_fft)


def window_trials(rec: np.ndarray, onsets, T: int) -> np.ndarray:
    """Trial windowing: ``x[b] = rec[:, onsets[b] : onsets[b] + T]``."""
    return np.stack([rec[:, int(o):int(o) + T] for o in onsets], axis=0)


def fir_same(x: np.ndarray, h: np.ndarray) -> np.ndarray:
    """``y[..., t] = sum_k h[k] * x[..., t + (K-1)/2 - k]`` with zero padding.

    Equals ``F.conv1d(x, h.flip(0), padding=K//2)`` (delay-compensated "same").
    """
    K = len(h)
    assert K % 2 == 1
    P = (K - 1) // 2
    T = x.shape[-1]
    xp = np.zeros(x.shape[:-1] + (T + 2 * P,), dtype=np.float64)
    xp[..., P:P + T] = x
    y = np.zeros(x.shape, dtype=np.float64)
    for k in range(K):
        # index into xp: t + P + P - k
        y += h[k] * xp[..., 2 * P - k: 2 * P - k + T]
    return y


def stft_power(y: np.ndarray, n_fft: int, hop: int) -> np.ndarray:
    """|STFT|^2 with ``center=True, pad_mode='reflect'``, periodic Hann,
    one-sided, not normalised.  (..., T) -> (..., F=n_fft//2+1, N_f=1+T//hop)."""
    T = y.shape[-1]
    half = n_fft // 2
    assert T > half, "reflect padding needs T > n_fft/2"
    idx = np.arange(-half, T + half)
    idx = np.where(idx < 0, -idx, idx)
    idx = np.where(idx >= T, 2 * (T - 1) - idx, idx)
    yp = y[..., idx]
    n_frames = 1 + T // hop
    w = hann_periodic(n_fft)
    frames = np.stack([yp[..., m * hop: m * hop + n_fft] for m in range(n_frames)], axis=-2)
    spec = np.fft.rfft(frames * w, axis=-1)            # (..., N_f, F)
    power = spec.real ** 2 + spec.imag ** 2
    return np.swapaxes(power, -1, -2)                  # (..., F, N_f)


def dsp_reference(x, h, n_fft=256, hop=64, log_eps=1.0, z_eps=1e-8,
                  return_stages=False):
    """The whole DSP chain in float64.  x: (B, C, T) -> (B, C, F, N_f) float64.

    FIR -> STFT power -> ``log(P + log_eps)`` -> per-(trial, channel) z-score
    over all F*N_f values, population std, ``(L - mu) / (sigma + z_eps)``
    (eps placement and ddof=0 follow the reference fallback, dataset.py:213-216).
    """
    x = np.asarray(x, dtype=np.float64)
    h = np.asarray(h, dtype=np.float64)
    y = fir_same(x, h)
    P = stft_power(y, n_fft, hop)
    L = np.log(P + log_eps)
    mu = L.mean(axis=(-1, -2), keepdims=True)
    sd = L.std(axis=(-1, -2), keepdims=True)
    z = (L - mu) / (sd + z_eps)
    if return_stages:
        return {"fir": y, "power": P, "logp": L, "z": z}
    return z


def to_encoder_layout(z: np.ndarray, region_slices):
    """(B, C, F, N_f) -> list of (B, C_r*F, N_f), channel-major then frequency
    (SURVEY.md 8(c) "Layout to encoder")."""
    B, C, F, Nf = z.shape
    return [z[:, sl].reshape(B, -1, Nf) for sl in region_slices]


def dsp_torch_cpu_f32(x, h, n_fft=256, hop=64, log_eps=1.0, z_eps=1e-8):
    """CPU baseline that bench.py times: the same chain with the stock library
    calls the spec names, float32, all host threads.

    x: torch float32 (B, C, T) on CPU; h: torch float32 (K,).
    """
    import torch
    import torch.nn.functional as F

    B, C, T = x.shape
    K = h.numel()
    y = F.conv1d(x.reshape(B * C, 1, T), h.flip(0).view(1, 1, K), padding=K // 2)
    spec = torch.stft(y.reshape(B * C, T), n_fft=n_fft, hop_length=hop, win_length=n_fft,
                      window=torch.hann_window(n_fft, periodic=True, dtype=x.dtype),
                      center=True, pad_mode="reflect", normalized=False, onesided=True,
                      return_complex=True)
    L = torch.log(spec.real ** 2 + spec.imag ** 2 + log_eps)
    mu = L.mean(dim=(-1, -2), keepdim=True)
    sd = L.std(dim=(-1, -2), keepdim=True, unbiased=False)
    z = (L - mu) / (sd + z_eps)
    return z.reshape(B, C, L.shape[-2], L.shape[-1])


def rel_max_err(a, b) -> float:
    """The tolerance criterion of SURVEY.md section 7: max|a-b| / max|b|."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def synth_eeg(B, C, T, seed=1234, tones=True):
    """Seeded synthetic trials of SURVEY.md 8(d): 20 uV Gaussian noise plus
    10 Hz / 20 Hz sinusoids of amplitude 10 (fs = 256).  float32 (B, C, T)."""
    rng = np.random.default_rng(seed)
    x = 20.0 * rng.standard_normal((B, C, T))
    if tones:
        t = np.arange(T) / 256.0
        ph = rng.uniform(0, 2 * np.pi, size=(B, C, 2))
        x += 10.0 * np.sin(2 * np.pi * 10.0 * t + ph[..., 0:1])
        x += 10.0 * np.sin(2 * np.pi * 20.0 * t + ph[..., 1:2])
    return x.astype(np.float32)
