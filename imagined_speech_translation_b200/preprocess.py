"""Batched GPU preprocessing behind the reference's data-transform semantics.

``RegionNormalizer`` is the batched, on-device equivalent of
``EEGDataset._process_raw_eeg`` + ``EEGDataset._normalize_eeg_sample``
(reference ``main_model/src/data/dataset.py:172-225``): same region order, same
``nan_to_num`` constants, same ``(x - center_) / scale_`` arithmetic and the
same z-score fallback when a region has no scaler -- but for a whole
``(B, C, T)`` batch in one launch instead of one trial at a time on the host.

``SpectrogramFrontEnd`` is the DSP chain BASELINE.json's north_star adds
(trial windowing -> band-pass FIR -> STFT log-power -> per-channel z-score; spec
in SURVEY.md section 8(c)).  The reference has no
counterpart; its output feeds ``Conv1DWithAttention(n_channels=C_r*F,
n_timepoints=N_f)`` per region.

Both call the C ABI of libeegx.so through ctypes; neither has a CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib

REGION_ORDER = ("frontal", "temporal", "central", "parietal")  # dataset.py:203

DSP_CONFIG = {
    "fs": 256.0,          # Hz
    "band": (8.0, 30.0),  # Hz, the band the paper names
    "numtaps": 65,
    "n_fft": 256,
    "hop": 64,
    "log_eps": 1.0,       # uV^2 floor inside the log
    "z_eps": 1e-8,        # dataset.py:215
}

DSP_CONFIG_LONG = dict(DSP_CONFIG, n_fft=1024, hop=256)  # BASELINE config 4


def design_bandpass_fir(numtaps: int, band: Sequence[float], fs: float) -> np.ndarray:
    """Hamming-windowed sinc band-pass with unit gain at the band centre.

    Host-side, float64 math, returned as float32 taps (what the kernel uses).
    Same design rule as ``scipy.signal.firwin(numtaps, band, pass_zero=False,
    fs=fs, window='hamming')``, the call SURVEY.md section 8(c) names.
    """
    if numtaps % 2 != 1:
        raise ValueError("numtaps must be odd (type-I linear phase, integer delay)")
    lo, hi = (2.0 * float(f) / fs for f in band)
    if not 0.0 < lo < hi < 1.0:
        raise ValueError(f"band {band} must satisfy 0 < low < high < fs/2")
    n = np.arange(numtaps, dtype=np.float64)
    m = n - 0.5 * (numtaps - 1)
    taps = hi * np.sinc(hi * m) - lo * np.sinc(lo * m)
    taps *= 0.54 - 0.46 * np.cos(2.0 * np.pi * n / (numtaps - 1))
    taps /= np.sum(taps * np.cos(np.pi * m * 0.5 * (lo + hi)))
    return taps.astype(np.float32)


def _require_cuda_f32(t: torch.Tensor, name: str, ndim: int) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.EegxError(f"{name} must be a CUDA tensor (this path has no CPU fallback)")
    if t.dtype != torch.float32 or t.dim() != ndim or not t.is_contiguous():
        raise ValueError(f"{name} must be a contiguous float32 tensor with {ndim} dims, got "
                         f"{t.dtype} {tuple(t.shape)} contiguous={t.is_contiguous()}")


class SpectrogramFrontEnd:
    """FIR band-pass -> STFT log-power -> per-channel z-score, one fused launch.

    ``fe(x)`` with ``x`` (B, C, T) float32 CUDA returns (B, C, F, N_f) float32,
    F = n_fft/2 + 1, N_f = 1 + T // hop.
    """

    def __init__(self, n_channels: int, n_timepoints: int, config: Optional[dict] = None,
                 taps: Optional[np.ndarray] = None):
        cfg = dict(DSP_CONFIG)
        if config:
            cfg.update(config)
        self.config = cfg
        self.n_channels = int(n_channels)
        self.n_timepoints = int(n_timepoints)
        if taps is None:
            taps = design_bandpass_fir(cfg["numtaps"], cfg["band"], cfg["fs"])
        self.taps = np.ascontiguousarray(taps, dtype=np.float32)
        self._plan = C.c_void_p()
        lib = _lib.lib()
        _lib.check(lib.eegx_dsp_plan_create(
            C.byref(self._plan), self.n_channels, self.n_timepoints, int(cfg["n_fft"]),
            int(cfg["hop"]), self.taps.ctypes.data_as(_lib.c_f32p), int(self.taps.size),
            float(cfg["log_eps"]), float(cfg["z_eps"])), "eegx_dsp_plan_create")
        f, nf = C.c_int(), C.c_int()
        _lib.check(lib.eegx_dsp_plan_dims(self._plan, C.byref(f), C.byref(nf)), "eegx_dsp_plan_dims")
        self.n_freqs, self.n_frames = f.value, nf.value

    def __del__(self):
        plan = getattr(self, "_plan", None)
        if plan is not None and plan.value:
            try:
                _lib.lib().eegx_dsp_plan_destroy(plan)
            except Exception:
                pass
            self._plan = None

    @property
    def kernel_name(self) -> str:
        return ("generic", "tuned", "long", "precise")[_lib.lib().eegx_dsp_plan_kernel(self._plan)]

    def force_generic(self, on: bool = True) -> None:
        _lib.check(_lib.lib().eegx_dsp_plan_force_generic(self._plan, int(on)))

    def set_precise(self, on: bool = True) -> None:
        """float64 arithmetic between the float32 input and output (about 10x slower; the bound of the spec at
        n_fft = 1024, where float32 stops at 1.1e-5 .. 1.9e-5)."""
        _lib.check(_lib.lib().eegx_dsp_plan_set_precise(self._plan, int(on)))

    def out_shape(self, batch: int):
        return (batch, self.n_channels, self.n_freqs, self.n_frames)

    def __call__(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        _require_cuda_f32(x, "x", 3)
        B, Cc, T = x.shape
        if Cc != self.n_channels or T != self.n_timepoints:
            raise ValueError(f"plan is for (C={self.n_channels}, T={self.n_timepoints}), got {tuple(x.shape)}")
        if out is None:
            out = torch.empty(self.out_shape(B), dtype=torch.float32, device=x.device)
        else:
            _require_cuda_f32(out, "out", 4)
            if tuple(out.shape) != self.out_shape(B):
                raise ValueError(f"out must have shape {self.out_shape(B)}")
        _lib.check(_lib.lib().eegx_dsp_forward(self._plan, _lib.ptr(x), None, 0, _lib.ptr(out), B,
                                               _lib.stream_ptr()), "eegx_dsp_forward")
        return out

    def from_recording(self, rec: torch.Tensor, onsets: torch.Tensor,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Trial windowing fused into the load: trial b = rec[:, onsets[b]:onsets[b]+T]."""
        _require_cuda_f32(rec, "rec", 2)
        if rec.shape[0] != self.n_channels:
            raise ValueError("rec must be (C, rec_len)")
        if onsets.dtype != torch.int64 or not onsets.is_cuda or onsets.dim() != 1:
            raise ValueError("onsets must be a 1-D int64 CUDA tensor")
        B = onsets.numel()
        if out is None:
            out = torch.empty(self.out_shape(B), dtype=torch.float32, device=rec.device)
        _lib.check(_lib.lib().eegx_dsp_forward(self._plan, _lib.ptr(rec), _lib.ptr(onsets),
                                               rec.shape[1], _lib.ptr(out), B, _lib.stream_ptr()),
                   "eegx_dsp_forward")
        return out

    def split_regions(self, z: torch.Tensor, region_channel_counts: Dict[str, int]) -> List[torch.Tensor]:
        """(B, C, F, N_f) -> list of 4 views (B, C_r*F, N_f): the layout
        ``Conv1DWithAttention(n_channels=C_r*F, n_timepoints=N_f)`` consumes."""
        B = z.shape[0]
        outs, c0 = [], 0
        for name in REGION_ORDER:
            cr = int(region_channel_counts[name])
            outs.append(z[:, c0:c0 + cr].reshape(B, cr * self.n_freqs, self.n_frames))
            c0 += cr
        if c0 != self.n_channels:
            raise ValueError("region_channel_counts must sum to n_channels")
        return outs


class RegionNormalizer:
    """Batched ``_normalize_eeg_sample``: gather 4 regions, nan_to_num, robust-scale.

    ``region_indices``: dict region -> channel rows (``EEGDataset.region_indices``).
    ``centers`` / ``scales``: dict region -> (C_r,) arrays (``RobustScaler.center_``
    / ``.scale_``).  A region absent from them takes the reference's fallback
    branch (per-channel z-score over time, dataset.py:213-216).
    Returns a list of four dense (B, C_r, T) tensors in the reference's order.
    """

    def __init__(self, region_indices: Dict[str, Sequence[int]], centers=None, scales=None,
                 device="cuda"):
        self.device = torch.device(device)
        self.region_sizes = [len(region_indices[n]) for n in REGION_ORDER]
        centers = centers or {}
        scales = scales or {}
        self._groups = []  # (mode, idx, center, scale, region positions)
        robust = [n for n in REGION_ORDER if n in centers]
        fallback = [n for n in REGION_ORDER if n not in centers]
        self._robust_names, self._fallback_names = robust, fallback
        self._idx, self._center, self._scale = {}, {}, {}
        for n in REGION_ORDER:
            self._idx[n] = torch.as_tensor(np.asarray(region_indices[n], dtype=np.int32),
                                           device=self.device)
            if n in centers:
                self._center[n] = torch.as_tensor(np.asarray(centers[n], dtype=np.float32),
                                                  device=self.device)
                self._scale[n] = torch.as_tensor(np.asarray(scales[n], dtype=np.float32),
                                                 device=self.device)
        # one launch per mode: concatenate the regions that share a mode
        self._plan_cache = {}

    @classmethod
    def from_dataset(cls, dataset, device="cuda"):
        """Build from an object with the reference ``EEGDataset`` attributes
        ``region_indices`` and ``scalers`` (dataset.py:41, :140-147)."""
        centers = {n: s.center_ for n, s in dataset.scalers.items()}
        scales = {n: s.scale_ for n, s in dataset.scalers.items()}
        return cls(dataset.region_indices, centers, scales, device=device)

    @classmethod
    def fit(cls, samples: torch.Tensor, region_indices: Dict[str, Sequence[int]],
            quantile_range=(5.0, 95.0), device="cuda"):
        """The scaler fit of ``EEGDataset._initialize_scalers_efficiently`` (dataset.py:102-151) on the
        GPU: ``samples`` (n, C_in, T) float32 -- the fit subset, already chosen by the caller -- are
        nan_to_num'ed and gathered per region, every channel's n*T values go through an exact radix
        select for the median and the two percentiles (``RobustScaler.center_`` / ``scale_``)."""
        samples = torch.as_tensor(samples, dtype=torch.float32, device=device)
        if samples.dim() == 4 and samples.shape[1] == 1:          # pickled as (1, C, T)
            samples = samples[:, 0]
        _require_cuda_f32(samples, "samples", 3)
        n, _, T = samples.shape
        centers, scales = {}, {}
        lib = _lib.lib()
        for name in REGION_ORDER:
            idx = torch.as_tensor(np.asarray(region_indices[name], dtype=np.int32), device=samples.device)
            reg = normalize_dense(samples.contiguous(), idx, None, None)          # gather + nan_to_num
            rows = reg.permute(1, 0, 2).reshape(idx.numel(), n * T).contiguous()   # channel-major, time-concatenated
            cen = torch.empty(idx.numel(), dtype=torch.float32, device=samples.device)
            sca = torch.empty_like(cen)
            _lib.check(lib.eegx_robust_fit_f32(_lib.ptr(rows), idx.numel(), n * T, float(quantile_range[0]),
                                               float(quantile_range[1]), _lib.ptr(cen), _lib.ptr(sca),
                                               _lib.stream_ptr()), "eegx_robust_fit_f32")
            centers[name], scales[name] = cen.cpu().numpy(), sca.cpu().numpy()
        return cls(region_indices, centers, scales, device=device)

    @property
    def centers(self):
        return {n: self._center[n] for n in self._robust_names}

    @property
    def scales(self):
        return {n: self._scale[n] for n in self._robust_names}

    def _layout(self, names, B, T):
        key = (tuple(names), B, T)
        hit = self._plan_cache.get(key)
        if hit is not None:
            return hit
        sizes = [self._idx[n].numel() for n in names]
        offs, bstr, base = [], [], 0
        for cr in sizes:
            offs += [base + j * T for j in range(cr)]
            bstr += [cr * T] * cr
            base += B * cr * T
        idx = torch.cat([self._idx[n] for n in names])
        off_t = torch.tensor(offs, dtype=torch.int64, device=self.device)
        bstr_t = torch.tensor(bstr, dtype=torch.int64, device=self.device)
        cen = torch.cat([self._center[n] for n in names]) if names and names[0] in self._center else None
        sca = torch.cat([self._scale[n] for n in names]) if cen is not None else None
        hit = (idx, off_t, bstr_t, cen, sca, sizes, base)
        self._plan_cache[key] = hit
        return hit

    def __call__(self, x: torch.Tensor) -> List[torch.Tensor]:
        _require_cuda_f32(x, "x", 3)
        B, C_in, T = x.shape
        lib = _lib.lib()
        result = {}
        for names, robust in ((self._robust_names, True), (self._fallback_names, False)):
            if not names:
                continue
            idx, off_t, bstr_t, cen, sca, sizes, total = self._layout(names, B, T)
            buf = torch.empty(total, dtype=torch.float32, device=x.device)
            if robust:
                _lib.check(lib.eegx_normalize_f32(_lib.ptr(x), _lib.ptr(idx), _lib.ptr(cen),
                                                  _lib.ptr(sca), _lib.ptr(buf), _lib.ptr(off_t),
                                                  _lib.ptr(bstr_t), B, C_in, idx.numel(), T,
                                                  _lib.stream_ptr()), "eegx_normalize_f32")
            else:
                _lib.check(lib.eegx_zscore_time_f32(_lib.ptr(x), _lib.ptr(idx), _lib.ptr(buf),
                                                    _lib.ptr(off_t), _lib.ptr(bstr_t), B, C_in,
                                                    idx.numel(), T, _lib.stream_ptr()),
                           "eegx_zscore_time_f32")
            base = 0
            for n, cr in zip(names, sizes):
                result[n] = buf[base:base + B * cr * T].view(B, cr, T)
                base += B * cr * T
        return [result[n] for n in REGION_ORDER]


def normalize_dense(x: torch.Tensor, ch_idx: Optional[torch.Tensor], center: Optional[torch.Tensor],
                    scale: Optional[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(B, C_in, T) -> dense (B, C_out, T): gather + nan_to_num + (x - center) / scale."""
    _require_cuda_f32(x, "x", 3)
    B, C_in, T = x.shape
    C_out = ch_idx.numel() if ch_idx is not None else C_in
    if out is None:
        out = torch.empty((B, C_out, T), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().eegx_normalize_f32(_lib.ptr(x), _lib.ptr(ch_idx), _lib.ptr(center),
                                             _lib.ptr(scale), _lib.ptr(out), None, None, B, C_in,
                                             C_out, T, _lib.stream_ptr()), "eegx_normalize_f32")
    return out
