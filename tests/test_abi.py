"""CPU-side checks of the drop-in boundary: libeegx.so loads, exports every
symbol include/eegx.h declares, the ctypes table covers them, and compute calls
fail loudly (no CPU fallback) when no B200 is present."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import imagined_speech_translation_b200 as pkg
from imagined_speech_translation_b200 import _lib
from imagined_speech_translation_b200.build import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def handle():
    build()
    return _lib.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "eegx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(eegx_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound(handle):
    syms = declared_symbols()
    assert "eegx_dsp_forward" in syms and "eegx_normalize_f32" in syms
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in eegx.h but not exported by libeegx.so"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == syms


def test_version(handle):
    assert handle.eegx_version() == 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(handle):
    assert handle.eegx_device_check() != 0
    assert len(handle.eegx_last_error()) > 0
    with pytest.raises(pkg.EegxError):
        pkg.SpectrogramFrontEnd(4, 2048)
    with pytest.raises(pkg.EegxError):
        pkg.normalize_dense(torch.zeros(1, 2, 8), None, None, None)


def test_fir_design_matches_oracle():
    from oracle import preprocess_oracle as po
    for numtaps, band in ((65, (8.0, 30.0)), (33, (4.0, 40.0)), (129, (1.0, 45.0))):
        h = pkg.design_bandpass_fir(numtaps, band, 256.0)
        ref = po.firwin_bandpass(numtaps, band[0], band[1], 256.0)
        assert h.dtype == np.float32
        np.testing.assert_allclose(h, ref.astype(np.float32), rtol=0, atol=1e-9)
    with pytest.raises(ValueError):
        pkg.design_bandpass_fir(64, (8.0, 30.0), 256.0)
