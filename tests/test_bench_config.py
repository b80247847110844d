"""bench.py's workload constants against SURVEY.md section 8(d) / BASELINE.json (CPU only, no GPU work)."""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    return importlib.import_module("bench")


def test_algorithmic_bytes_per_trial_match_the_survey():
    b = _bench()
    try:
        b.select_config("stft")
        assert (b.B_PER_GPU, b.C, b.T, b.N_FFT, b.HOP) == (256, 64, 2048, 256, 64)
        assert (b.F, b.NF) == (129, 33)
        assert b.DSP_BYTES_PER_TRIAL == 1_614_080                     # 64 x (8192 B read + 17028 B written)
        b.select_config("long")
        assert (b.C, b.T, b.N_FFT, b.HOP) == (128, 4096, 1024, 256)
        assert (b.F, b.NF) == (513, 17)
        assert b.DSP_BYTES_PER_TRIAL == 6_562_304                     # SURVEY 8(d), configs[3]
        assert b.COUNTS == {k: 32 for k in ("frontal", "temporal", "central", "parietal")}
    finally:
        b.select_config("stft")


def test_workload_names_quote_the_baseline_configs():
    b = _bench()
    with open(os.path.join(ROOT, "BASELINE.json")) as fh:
        base = json.load(fh)
    assert len(base["configs"]) == 5
    assert b.WORKLOADS["stft"]["title"] == "BASELINE configs[2]" and b.WORKLOADS["long"]["title"] == "BASELINE configs[3]"
    assert b.CPU_SAMPLE_B == 32                                        # configs[0]: batch 32 on the host cores
    cfg = b.workload_config("train", {})
    assert cfg["workload"].startswith("BASELINE configs[2]") and cfg["batch_per_gpu"] == 256
    assert "model" not in cfg                                          # the contract: a workload description, no model keys
