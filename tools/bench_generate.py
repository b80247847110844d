"""Evaluation-time generation throughput (row f3): BARTDecoder.generate_from_eeg on our kernels against the same
call through transformers.generate (fp32 and bf16 autocast) on the same GPU and weights.

    python tools/bench_generate.py [--batch 256] [--iters 5]
"""
import argparse
import json
import os
os.environ.setdefault("EEGX_BART_RANDOM_INIT", "1")   # synthetic benchmark: reference architecture, random weights (no HF cache here)
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagined_speech_translation_b200.model import BARTDecoder  # noqa: E402


def timed(fn, iters):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    torch.manual_seed(0)
    dec = BARTDecoder(hidden_dim=768).cuda().eval()
    with torch.no_grad():
        dec.bart.model.shared.weight.mul_(6.0)
        dec.bart.final_logits_bias.copy_(torch.randn_like(dec.bart.final_logits_bias) * 1.5)
    feat = torch.randn(a.batch, 768, device="cuda")
    gen = dict(max_length=16, min_length=4, num_beams=3, early_stopping=True)       # training_config.py:32-39
    res = {}
    dec.native_generate = True
    for mode in ("reencode", "cache", "graph"):
        dec.generate_mode = mode
        res[f"native_{mode}_ms"], got = timed(lambda: dec.generate_from_eeg(feat, **gen), a.iters)
    res["native_ms"] = res["native_graph_ms"]
    dec.native_generate = False
    dec.autocast_dtype = None
    res["transformers_fp32_ms"], want = timed(lambda: dec.generate_from_eeg(feat, **gen), a.iters)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        res["transformers_bf16_autocast_ms"], _ = timed(lambda: dec.generate_from_eeg(feat, **gen), a.iters)
    n = min(got.shape[1], want.shape[1])
    print(json.dumps({"batch": a.batch, "beams": 3, "max_length": 16, "steps_taken": int(got.shape[1]) - 1,
                      **{k: round(v, 2) for k, v in res.items()},
                      "native_trials_per_s": round(a.batch / res["native_ms"] * 1e3, 1),
                      "transformers_fp32_trials_per_s": round(a.batch / res["transformers_fp32_ms"] * 1e3, 1),
                      "identical_sequences": round(sum(torch.equal(x[:n], y[:n]) for x, y in zip(got, want)) / a.batch, 4),
                      "identical_tokens": round((got[:, :n] == want[:, :n]).float().mean().item(), 4)}))


if __name__ == "__main__":
    main()
