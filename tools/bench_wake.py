"""Throughput of the wake_model dense head (BASELINE config 5 shape) on the GPU, next to the reference's own C++
(oracle/_ref, compiled from wake_model/layers/linear.cpp) timed on the host.

    python tools/bench_wake.py [--in 4096] [--hidden 1024] [--classes 120] [--samples 2000]

Prints one JSON object: samples/s through eegx_wake_dense_f64 (CUDA events, 3 warm-up launches), the algorithmic
W1 traffic rate (16 B * hidden * in per sample: one read + one write of W1), and the CPU reference rate on a
bounded sample.  The oracle/reference is used here only as the timed CPU baseline and as the parity checker.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagined_speech_translation_b200.wake import DenseHead  # noqa: E402
from oracle import wake_oracle  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--in", dest="n_in", type=int, default=4096)
    ap.add_argument("--hidden", type=int, default=1024)
    ap.add_argument("--classes", type=int, default=120)
    ap.add_argument("--samples", type=int, default=2000)
    ap.add_argument("--cpu-samples", type=int, default=64)
    a = ap.parse_args()
    g = torch.Generator().manual_seed(1)
    head = DenseHead(a.n_in, a.hidden, a.classes, generator=g)
    x = (torch.randn(a.samples, a.n_in, dtype=torch.float64, generator=g) * 0.05).cuda()
    y = torch.randint(0, a.classes, (a.samples,), generator=g).int().cuda()
    init = [t.cpu().numpy().copy() for t in (head.w1, head.b1, head.w2, head.b2)]

    # parity of the timed configuration on a prefix, before timing
    k = min(8, a.samples)
    probe = DenseHead(a.n_in, a.hidden, a.classes).load(*init)
    probe.train_samples(x[:k], y[:k], 0.1)
    want = wake_oracle.run(wake_oracle.oracle(), *init, x[:k].cpu().numpy(), y[:k].cpu().numpy(), lr=0.1)
    err = float(np.abs(probe.w1.cpu().numpy() - want["w1"]).max() / np.abs(want["w1"]).max())

    times = {}
    for want_dx in (False, True):
        for _ in range(3):
            head.train_samples(x, y, 0.1, want_dx=want_dx)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            head.train_samples(x, y, 0.1, want_dx=want_dx)
        e1.record()
        torch.cuda.synchronize()
        times[want_dx] = e0.elapsed_time(e1) / 3 / 1e3
    sps = a.samples / times[False]

    cpu = {}
    for name, fn in (("reference", wake_oracle.reference()), ("port", wake_oracle.oracle())):
        if fn is None:
            continue
        xs, ys = x[:a.cpu_samples].cpu().numpy(), y[:a.cpu_samples].cpu().numpy()
        t0 = time.perf_counter()
        wake_oracle.run(fn, *init, xs, ys, lr=0.1)
        cpu[name] = a.cpu_samples / (time.perf_counter() - t0)
    print(json.dumps({
        "metric": "wake_model dense-head SGD samples/s", "value": round(sps, 1), "unit": "samples/s",
        "with_dx_samples_per_s": round(a.samples / times[True], 1),
        "us_per_sample": round(1e6 / sps, 2), "dtype": "f64",
        "config": {"workload": "Linear(in,hidden,relu)->Linear(hidden,classes,softmax)->CCE, per-sample SGD lr 0.1",
                   "in": a.n_in, "hidden": a.hidden, "classes": a.classes, "samples_per_launch": a.samples},
        "w1_traffic_GBps": round(16.0 * a.hidden * a.n_in * sps / 1e9, 1),
        "parity_rel_err_w1_after_8_samples": err,
        "cpu_baseline": {"samples_per_s": {k: round(v, 2) for k, v in cpu.items()}, "cores": 1,
                         "sample": f"{a.cpu_samples} samples of the same problem, single thread (the reference is serial)"},
    }))


if __name__ == "__main__":
    main()
