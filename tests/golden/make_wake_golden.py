"""Generates tests/golden/wake_dense_ref.npz from the REFERENCE's own wake_model code.

Runs oracle/_ref/libwake_ref.so (wake_model/layers/linear.cpp + activations.h + losses.h compiled from
/root/reference by oracle/Makefile, driven by oracle/wake_ref_harness.cpp) on small seeded problems for every
hidden activation and stores inputs and outputs.  Run in the container that has /root/reference:

    make -C oracle && python tests/golden/make_wake_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import wake_oracle  # noqa: E402


def main():
    ref = wake_oracle.reference()
    assert ref is not None, "oracle/_ref/libwake_ref.so missing: run `make -C oracle` where /root/reference exists"
    rng = np.random.default_rng(20240917)
    out = {}
    cases = [("relu", 24, 20, 5, 12), ("sigmoid", 17, 9, 4, 10), ("tanh", 8, 33, 3, 10), ("", 5, 6, 2, 8),
             ("relu", 40, 300, 7, 6)]
    for ci, (act, n_in, hidden, ncls, n) in enumerate(cases):
        w1 = rng.normal(0, np.sqrt(2.0 / n_in), (hidden, n_in))
        b1 = rng.normal(0, np.sqrt(2.0 / n_in), hidden)
        w2 = rng.normal(0, np.sqrt(2.0 / hidden), (ncls, hidden))
        b2 = rng.normal(0, np.sqrt(2.0 / hidden), ncls)
        x = rng.normal(0, 1.0, (n, n_in))
        label = rng.integers(0, ncls, n).astype(np.int32)
        r = wake_oracle.run(ref, w1, b1, w2, b2, x, label, lr=0.1, activation=act, train=True, want_dx=True)
        f = wake_oracle.run(ref, w1, b1, w2, b2, x, label, lr=0.1, activation=act, train=False)
        pre = f"c{ci}_"
        out[pre + "act"] = np.array(act)
        for k, v in dict(w1=w1, b1=b1, w2=w2, b2=b2, x=x, label=label).items():
            out[pre + "in_" + k] = v
        for k in ("w1", "b1", "w2", "b2", "loss", "probs", "dx"):
            out[pre + "out_" + k] = r[k]
        out[pre + "fwd_probs"] = f["probs"]
        out[pre + "fwd_loss"] = f["loss"]
    out["n_cases"] = np.array(len(cases))
    path = os.path.join(ROOT, "tests", "golden", "wake_dense_ref.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
