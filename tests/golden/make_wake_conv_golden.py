"""Generates tests/golden/wake_conv_ref.npz from the REFERENCE's own Convolution / MaxPool classes.

Runs oracle/_ref/libwake_ref.so (wake_model/layers/convolution.cpp + maxpool.cpp compiled from /root/reference by
oracle/Makefile, driven by oracle/wake_ref_harness.cpp) on small seeded problems -- the train.cpp:26-33 shapes
(input height 2, kernels 32 / 64 / 128 x 1, pools 2 x 1 stride 1), 2-D kernels, strided and overlapping pools, ties,
and the H > W case in which the reference's row bound (maxpool.h:15) bites.  Run where /root/reference exists:

    make -C oracle && python tests/golden/make_wake_conv_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import wake_oracle  # noqa: E402

CONV_CASES = [(2, 120, 1, 32), (2, 88, 1, 64), (2, 140, 1, 128), (5, 40, 3, 7), (2, 64, 2, 3), (1, 9, 1, 9)]
POOL_CASES = [(2, 89, 2, 1, 1), (6, 20, 2, 2, 2), (30, 8, 3, 2, 1), (4, 9, 2, 1, 1), (7, 7, 3, 3, 2)]


def main():
    conv, pool = wake_oracle.conv_reference()
    assert conv is not None and pool is not None, "oracle/_ref/libwake_ref.so missing: run `make -C oracle`"
    rng = np.random.default_rng(20241018)
    out = {}
    for ci, (H, W, kh, kw) in enumerate(CONV_CASES):
        x = rng.normal(0, 1.0, (H, W))
        k = rng.uniform(-1, 1, (kh, kw)) * np.sqrt(6.0 / (kh * kw))
        b = np.array([rng.uniform(-0.05, 0.05)])
        d = rng.normal(0, 0.5, (H - kh + 1, W - kw + 1))
        r = wake_oracle.run_conv(conv, k, b, x, d, lr=0.1)
        pre = f"conv{ci}_"
        for name, v in dict(x=x, kernel=k, bias=b, dout=d).items():
            out[pre + "in_" + name] = v
        for name in ("y", "dx", "kernel", "bias"):
            out[pre + "out_" + name] = r[name]
    for ci, (H, W, pw, ph, s) in enumerate(POOL_CASES):
        x = np.round(rng.normal(0, 1.0, (H, W)), 1)          # one decimal: plenty of exact ties
        OH, OW = (H - ph) // s + 1, (W - pw) // s + 1
        d = rng.normal(0, 0.5, (OH, OW))
        r = wake_oracle.run_maxpool(pool, x, pw, ph, s, d)
        pre = f"pool{ci}_"
        out[pre + "cfg"] = np.array([pw, ph, s])
        out[pre + "in_x"], out[pre + "in_dout"] = x, d
        for name in ("y", "argmax", "dx"):
            out[pre + "out_" + name] = r[name]
    out["n_conv"], out["n_pool"] = np.array(len(CONV_CASES)), np.array(len(POOL_CASES))
    path = os.path.join(ROOT, "tests", "golden", "wake_conv_ref.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
