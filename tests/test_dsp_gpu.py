"""GPU parity of the fused DSP chain (through the C ABI) against the golden
spec vectors, the float64 oracle on seeded inputs, and size-independent
properties at BASELINE config sizes.

Tolerance (stated, SURVEY.md section 7 / north_star "<=1e-5 relative"):
    max|a - b| / max|b| <= 1e-5   against the float64 oracle.
"""
import os

import numpy as np
import pytest
import torch

import imagined_speech_translation_b200 as pkg
from oracle import preprocess_oracle as po

pytestmark = pytest.mark.gpu
TOL = 1e-5        # north_star bar; met with margin at n_fft <= 256 (measured 5.5e-6 on B200)
# Long-window (n_fft = 1024) and non-default tap sets: float32 itself does not reach 1e-5
# against float64 -- torch CPU float32 (F.conv1d + torch.stft) measures 1.0e-5 / 1.7e-5 on the
# same inputs and rounding the FIR output to float32 alone costs 6.7e-6.
# Stated bound for those cases:
TOL_LONG = 3e-5


def tol_for(n_fft, default_taps=True):
    return TOL if (n_fft <= 256 and default_taps) else TOL_LONG


@pytest.fixture(scope="module")
def dsp(golden_dir):
    return np.load(os.path.join(golden_dir, "dsp_spec.npz"))


def _run(x_np, cfg=None, generic=False, taps=None):
    B, C, T = x_np.shape
    fe = pkg.SpectrogramFrontEnd(C, T, cfg, taps=taps)
    if generic:
        fe.force_generic(True)
    out = fe(torch.from_numpy(x_np).cuda())
    torch.cuda.synchronize()
    return out.cpu().numpy(), fe


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("key", ["a", "b", "c"])
def test_golden_spec_vectors(dsp, key, generic):
    n_fft, hop = (int(v) for v in dsp[f"cfg_{key}"])
    got, fe = _run(dsp[f"x_{key}"], {"n_fft": n_fft, "hop": hop}, generic=generic)
    assert got.shape == dsp[f"z_{key}"].shape
    assert np.array_equal(fe.taps, dsp["taps"].astype(np.float32))
    assert po.rel_max_err(got, dsp[f"z_{key}"]) <= tol_for(n_fft)


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("B,C,T,n_fft,hop", [
    (3, 5, 2048, 256, 64),      # config-2 shape, small batch
    (2, 3, 2000, 256, 64),      # T not a multiple of hop
    (1, 1, 1651, 256, 64),      # the real-data length, odd
    (2, 2, 4096, 1024, 256),    # config-4 long window
    (2, 4, 512, 64, 16),
    (1, 2, 300, 256, 100),      # hop does not divide n_fft
])
def test_vs_float64_oracle(B, C, T, n_fft, hop, generic):
    x = po.synth_eeg(B, C, T, seed=B * 100 + T)
    got, fe = _run(x, {"n_fft": n_fft, "hop": hop}, generic=generic)
    ref = po.dsp_reference(x, fe.taps.astype(np.float64), n_fft=n_fft, hop=hop)
    assert got.shape == ref.shape == (B, C, n_fft // 2 + 1, 1 + T // hop)
    assert po.rel_max_err(got, ref) <= tol_for(n_fft)


def test_tuned_kernel_is_selected_for_config2():
    fe = pkg.SpectrogramFrontEnd(64, 2048)
    assert fe.kernel_name == "tuned"
    fe2 = pkg.SpectrogramFrontEnd(8, 4096, pkg.DSP_CONFIG_LONG)
    assert fe2.kernel_name == "long"                    # BASELINE configs[3] has its own tuned kernel
    fe3 = pkg.SpectrogramFrontEnd(8, 4000, pkg.DSP_CONFIG_LONG)
    assert fe3.kernel_name == "generic"


def test_tuned_and_generic_agree():
    x = torch.from_numpy(po.synth_eeg(4, 64, 2048, seed=5)).cuda()
    fe = pkg.SpectrogramFrontEnd(64, 2048)
    assert fe.kernel_name == "tuned"
    a = fe(x).clone()
    fe.force_generic(True)
    assert fe.kernel_name == "generic"
    b = fe(x)
    # two float32 implementations, each within TOL of float64, agree within 2*TOL
    assert po.rel_max_err(a.cpu().numpy(), b.cpu().numpy()) <= 2 * TOL
    ref = po.dsp_reference(x[:1, :8].cpu().numpy(), fe.taps.astype(np.float64))
    assert po.rel_max_err(a[:1, :8].cpu().numpy(), ref) <= TOL
    assert po.rel_max_err(b[:1, :8].cpu().numpy(), ref) <= TOL


def test_long_and_generic_agree():
    """configs[3] shape: the warp-per-frame radix-8 kernel against the generic kernel and the float64 oracle."""
    x = torch.from_numpy(po.synth_eeg(3, 5, 4096, seed=11)).cuda()
    fe = pkg.SpectrogramFrontEnd(5, 4096, pkg.DSP_CONFIG_LONG)
    assert fe.kernel_name == "long"
    a = fe(x).clone()
    fe.force_generic(True)
    b = fe(x)
    ref = po.dsp_reference(x.cpu().numpy(), fe.taps.astype(np.float64), n_fft=1024, hop=256)
    ea, eb = po.rel_max_err(a.cpu().numpy(), ref), po.rel_max_err(b.cpu().numpy(), ref)
    print(f"configs[3] DSP vs float64: long kernel {ea:.3e}, generic kernel {eb:.3e}")
    assert ea <= TOL_LONG and eb <= TOL_LONG
    assert po.rel_max_err(a.cpu().numpy(), b.cpu().numpy()) <= 2 * TOL_LONG


@pytest.mark.parametrize("T,cfg", [(4096, {"n_fft": 1024, "hop": 256}), (2048, None), (1651, None)])
def test_float64_mode_meets_the_spec_bound_at_every_shape(T, cfg):
    """set_precise(): FIR accumulation and FFT butterflies in float64.  At n_fft = 1024 float32 arithmetic stops at
    1.1e-5 .. 1.9e-5 against the float64 oracle (TOL_LONG above; pocketfft in float32 fed a float64-exact FIR output
    measures 1.9e-5 on these inputs); with float64 arithmetic the spec's 1e-5 is met with two orders of margin."""
    x = po.synth_eeg(3, 5, T, seed=11)
    fe = pkg.SpectrogramFrontEnd(5, T, cfg)
    fe.set_precise(True)
    assert fe.kernel_name == "precise"
    got = fe(torch.from_numpy(x).cuda()).cpu().numpy()
    kw = {"n_fft": cfg["n_fft"], "hop": cfg["hop"]} if cfg else {}
    ref = po.dsp_reference(x, fe.taps.astype(np.float64), **kw)
    err = po.rel_max_err(got, ref)
    print(f"float64 DSP kernel at T={T}: {err:.3e}")
    assert err <= 1e-6
    fe.set_precise(False)
    assert fe.kernel_name in ("long", "tuned", "generic")


@pytest.mark.parametrize("B,C", [(1, 1), (1, 3), (2, 7), (37, 9)])
def test_long_ragged_row_counts(B, C):
    """row counts below, and not a multiple of, the persistent grid (2 CTAs per SM)."""
    x = po.synth_eeg(B, C, 4096, seed=B * 10 + C)
    got, fe = _run(x, pkg.DSP_CONFIG_LONG)
    assert fe.kernel_name == "long"
    pick = slice(0, min(B, 2))
    ref = po.dsp_reference(x[pick], fe.taps.astype(np.float64), n_fft=1024, hop=256)
    assert po.rel_max_err(got[pick], ref) <= TOL_LONG
    fe.force_generic(True)
    other = fe(torch.from_numpy(x).cuda()).cpu().numpy()
    assert po.rel_max_err(got, other) <= 2 * TOL_LONG


def test_long_full_size_properties():
    """BASELINE configs[3] (256 x 128 x 4096): properties that need no oracle run."""
    B, C, T = 256, 128, 4096
    g = torch.Generator(device="cuda").manual_seed(4321)
    x = 20.0 * torch.randn(B, C, T, generator=g, device="cuda")
    fe = pkg.SpectrogramFrontEnd(C, T, pkg.DSP_CONFIG_LONG)
    assert fe.kernel_name == "long"
    z = fe(x)
    assert z.shape == (B, C, 513, 17) and torch.isfinite(z).all()
    flat = z.reshape(B * C, -1).double()
    assert flat.mean(dim=1).abs().max().item() <= 1e-5
    assert (flat.std(dim=1, unbiased=False) - 1.0).abs().max().item() <= 1e-5
    assert torch.equal(z, fe(x))                                   # bit-stable
    assert torch.equal(fe(x[17:19].contiguous()), z[17:19])        # no cross-trial state
    assert torch.equal(fe(-x), z)                                  # sign flip leaves the power spectrum alone
    xs = x[[0, 255]][:, [0, 127]].cpu().numpy()
    ref = po.dsp_reference(xs, fe.taps.astype(np.float64), n_fft=1024, hop=256)
    assert po.rel_max_err(z[[0, 255]][:, [0, 127]].cpu().numpy(), ref) <= TOL_LONG


@pytest.mark.parametrize("B,C", [(1, 1), (1, 3), (3, 5), (2, 7)])
def test_tuned_ragged_row_counts(B, C):
    """rows = B*C not a multiple of the 4-row tile: the last tile is partial."""
    x = po.synth_eeg(B, C, 2048, seed=B * 10 + C)
    got, fe = _run(x)
    assert fe.kernel_name == "tuned"
    ref = po.dsp_reference(x, fe.taps.astype(np.float64))
    assert po.rel_max_err(got, ref) <= TOL


def test_other_taps_and_identity_filter():
    x = po.synth_eeg(2, 3, 1024, seed=9)
    for taps in (np.array([1.0], dtype=np.float32),
                 pkg.design_bandpass_fir(33, (4.0, 40.0), 256.0),
                 pkg.design_bandpass_fir(129, (1.0, 45.0), 256.0)):
        got, fe = _run(x, taps=taps)
        ref = po.dsp_reference(x, taps.astype(np.float64))
        assert po.rel_max_err(got, ref) <= tol_for(256, default_taps=False)


def test_windowed_mode_equals_cut_trials():
    C, T, L = 6, 2048, 9000
    rec = torch.from_numpy(po.synth_eeg(1, C, L, seed=3)[0]).cuda().contiguous()
    onsets = torch.tensor([0, 64, 1001, 4097, L - T], dtype=torch.int64, device="cuda")
    fe = pkg.SpectrogramFrontEnd(C, T)
    cut = torch.stack([rec[:, o:o + T] for o in onsets.tolist()]).contiguous()
    a = fe.from_recording(rec, onsets)          # windowed loads always take the generic kernel
    b_tuned = fe(cut)
    fe.force_generic(True)
    b_generic = fe(cut)
    assert torch.equal(a, b_generic)            # same kernel, same arithmetic: bit-exact
    assert po.rel_max_err(a.cpu().numpy(), b_tuned.cpu().numpy()) <= 2 * TOL
    ref = po.dsp_reference(cut.cpu().numpy(), fe.taps.astype(np.float64))
    assert po.rel_max_err(a.cpu().numpy(), ref) <= TOL


def test_full_size_properties():
    """BASELINE config 2 (256 x 64 x 2048): properties that need no oracle run."""
    B, C, T = 256, 64, 2048
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = 20.0 * torch.randn(B, C, T, generator=g, device="cuda")
    fe = pkg.SpectrogramFrontEnd(C, T)
    z = fe(x)
    assert z.shape == (B, C, 129, 33) and torch.isfinite(z).all()
    flat = z.reshape(B * C, -1).double()
    assert flat.mean(dim=1).abs().max().item() <= 1e-5          # z-score: mean 0
    assert (flat.std(dim=1, unbiased=False) - 1.0).abs().max().item() <= 1e-5   # std 1
    # determinism (bit-stable) and batch independence (no cross-trial state)
    assert torch.equal(z, fe(x))
    sub = fe(x[17:19].contiguous())
    assert torch.equal(sub, z[17:19])
    # spot-check a handful of rows against the oracle at full size
    xs = x[[0, 255]][:, [0, 63]].cpu().numpy()
    ref = po.dsp_reference(xs, fe.taps.astype(np.float64))
    got = z[[0, 255]][:, [0, 63]].cpu().numpy()
    assert po.rel_max_err(got, ref) <= TOL
    # scale covariance breaks under log, but a sign flip must not change anything
    assert torch.equal(fe(-x), z)


def test_degenerate_inputs():
    fe = pkg.SpectrogramFrontEnd(2, 2048)
    z = fe(torch.zeros(1, 2, 2048, device="cuda"))
    assert torch.equal(z, torch.zeros_like(z))           # log(0 + 1) = 0, sigma = 0 -> 0 / eps
    empty = fe(torch.empty(0, 2, 2048, device="cuda"))
    assert empty.shape == (0, 2, 129, 33)


def test_shape_and_argument_errors():
    with pytest.raises(pkg.EegxError):
        pkg.SpectrogramFrontEnd(2, 100, {"n_fft": 256})          # T <= n_fft/2
    with pytest.raises(pkg.EegxError):
        pkg.SpectrogramFrontEnd(2, 2048, {"n_fft": 200})         # not a power of two
    fe = pkg.SpectrogramFrontEnd(2, 2048)
    with pytest.raises(ValueError):
        fe(torch.zeros(1, 3, 2048, device="cuda"))
    with pytest.raises(pkg.EegxError):
        fe(torch.zeros(1, 2, 2048))
