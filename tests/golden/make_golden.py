"""Generate the committed golden fixtures.  Run in the BUILD container only
(needs /root/reference; the GPU box never runs this):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Outputs (small, committed):
  tests/golden/normalize_ref.npz  produced by the REFERENCE's own EEGDataset
      methods (_build_region_indices, _initialize_scalers_efficiently,
      _process_raw_eeg, _normalize_eeg_sample; main_model/src/data/dataset.py)
  tests/golden/dsp_spec.npz       produced by the library calls SURVEY.md 8(c)
      names (scipy.signal.firwin, F.conv1d, torch.stft) in float64; the
      reference has no DSP code, so this pins the oracle to the written spec.
"""
import os
import pickle
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/main_model"


def make_normalize_golden():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    import pandas as pd
    from src.data.dataset import EEGDataset  # the reference itself

    rng = np.random.default_rng(7)
    T, n_fit = 96, 10
    tmp = tempfile.mkdtemp(prefix="eegx_golden_")
    fit_samples = []
    for f in range(2):
        items = []
        for i in range(n_fit // 2):
            # channel-dependent offset/scale so center_/scale_ are non-trivial
            off = rng.uniform(-30, 30, size=(1, 125, 1))
            amp = rng.uniform(5, 40, size=(1, 125, 1))
            arr = (off + amp * rng.standard_normal((1, 125, T))).astype(np.float32)
            items.append({"input_features": arr, "text": "x"})
            fit_samples.append(arr)
        with open(os.path.join(tmp, f"run{f}.pkl"), "wb") as fh:
            pickle.dump(items, fh)

    ds = EEGDataset.__new__(EEGDataset)            # skip tokenizer / network set-up
    ds.max_samples = None
    ds.ch_names = pd.read_csv(os.path.join(REF, "data/montage.csv"))["label"].to_numpy()
    ds.region_indices = ds._build_region_indices()
    ds.data_files = sorted(ds._get_validated_data_files(tmp))
    ds.sample_index = ds._build_sample_index()
    np.random.seed(42)
    ds._initialize_scalers_efficiently()

    # trials to normalise: clean, and with NaN / +-inf sprinkled in
    trials = []
    for i in range(3):
        arr = (rng.uniform(-30, 30, size=(1, 125, 1))
               + 25.0 * rng.standard_normal((1, 125, T))).astype(np.float32)
        if i > 0:
            flat = arr.reshape(-1)
            pos = rng.choice(flat.size, size=60, replace=False)
            flat[pos[:20]] = np.nan
            flat[pos[20:40]] = np.inf
            flat[pos[40:]] = -np.inf
        trials.append(arr)
    trials = np.stack(trials)                       # (3, 1, 125, T)

    out = {"trials": trials, "fit_samples": np.stack(fit_samples)}
    names = ["frontal", "temporal", "central", "parietal"]
    for r, name in enumerate(names):
        out[f"idx_{name}"] = np.asarray(ds.region_indices[name], dtype=np.int32)
        out[f"center_{name}"] = np.asarray(ds.scalers[name].center_)
        out[f"scale_{name}"] = np.asarray(ds.scalers[name].scale_)
    for i in range(trials.shape[0]):
        regs = ds._normalize_eeg_sample(trials[i])
        for r, name in enumerate(names):
            out[f"robust_{i}_{name}"] = np.asarray(regs[r])
    # fallback branch (dataset.py:213-216): no scaler for any region
    ds.scalers = {}
    for i in range(trials.shape[0]):
        regs = ds._normalize_eeg_sample(trials[i])
        for r, name in enumerate(names):
            out[f"fallback_{i}_{name}"] = np.asarray(regs[r])
    # which fit samples np.random.choice picked (all of them here: size == len)
    np.savez_compressed(os.path.join(HERE, "normalize_ref.npz"), **out)
    print("normalize_ref.npz:", {k: v.shape for k, v in out.items() if k.startswith(("idx", "center"))})


def make_dsp_golden():
    import scipy.signal as ss
    import torch
    import torch.nn.functional as F

    out = {}
    h = ss.firwin(65, [8.0, 30.0], pass_zero=False, fs=256.0, window="hamming")
    out["taps"] = h
    cases = {"a": (2, 4, 2048, 256, 64), "b": (1, 2, 4096, 1024, 256), "c": (1, 3, 640, 128, 32)}
    rng = np.random.default_rng(11)
    for key, (B, C, T, n_fft, hop) in cases.items():
        t = np.arange(T) / 256.0
        x = 20.0 * rng.standard_normal((B, C, T)) + 10.0 * np.sin(2 * np.pi * 10.0 * t) \
            + 10.0 * np.sin(2 * np.pi * 20.0 * t + 0.3)
        x = x.astype(np.float32)
        xt = torch.from_numpy(x).double()
        ht = torch.from_numpy(h)
        y = F.conv1d(xt.reshape(B * C, 1, T), ht.flip(0).view(1, 1, -1), padding=32)
        spec = torch.stft(y.reshape(B * C, T), n_fft=n_fft, hop_length=hop, win_length=n_fft,
                          window=torch.hann_window(n_fft, periodic=True, dtype=torch.float64),
                          center=True, pad_mode="reflect", normalized=False, onesided=True,
                          return_complex=True)
        L = torch.log(spec.real ** 2 + spec.imag ** 2 + 1.0)
        mu = L.mean(dim=(-1, -2), keepdim=True)
        sd = L.std(dim=(-1, -2), keepdim=True, unbiased=False)
        z = ((L - mu) / (sd + 1e-8)).reshape(B, C, L.shape[-2], L.shape[-1])
        out[f"x_{key}"] = x
        out[f"cfg_{key}"] = np.asarray([n_fft, hop], dtype=np.int64)
        out[f"fir_{key}"] = y.reshape(B, C, T).numpy()[:, :1].astype(np.float64)   # first channel only
        out[f"z_{key}"] = z.numpy().astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "dsp_spec.npz"), **out)
    print("dsp_spec.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    which = sys.argv[1:] or ["normalize", "dsp"]
    if "normalize" in which:
        make_normalize_golden()
    if "dsp" in which:
        make_dsp_golden()
