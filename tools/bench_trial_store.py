"""Host-side item rate: per-item pickle loading (the reference's way, dataset.py:153-170) against the binary trial
store (data.TrialStore), on a synthetic set of the real shape (1, 125, 1651).  CPU only.

    python tools/bench_trial_store.py [--files 6] [--per-file 40] [--batch 64]
"""
import argparse
import json
import os
import pickle
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import imagined_speech_translation_b200 as pkg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=6)
    ap.add_argument("--per-file", type=int, default=40)
    ap.add_argument("--batch", type=int, default=64)
    a = ap.parse_args()
    from transformers import BertTokenizer
    fix = os.path.join(ROOT, "tests", "golden", "dataset")
    tok = BertTokenizer(os.path.join(fix, "vocab.txt"), bos_token="[CLS]", eos_token="[SEP]")
    rng = np.random.default_rng(0)
    with tempfile.TemporaryDirectory() as d:
        for f in range(a.files):
            run = [{"input_features": rng.normal(0, 20, (1, 125, 1651)).astype(np.float32), "text": "数据 样本"}
                   for _ in range(a.per_file)]
            with open(os.path.join(d, f"run{f}.pkl"), "wb") as fh:
                pickle.dump(run, fh)
        ds = pkg.EEGDataset(d, os.path.join(fix, "montage.csv"), tok, max_length=16, data_augmentation=False,
                            device="cpu")
        n = len(ds)
        order = rng.permutation(n)
        ds._load_file.cache_clear()
        t0 = time.perf_counter()
        for s in range(0, n - a.batch + 1, a.batch):
            ds.collate_raw([ds[int(i)] for i in order[s:s + a.batch]])
        t_pickle = time.perf_counter() - t0
        done = (n // a.batch) * a.batch
        # the reference keeps no LRU across workers and re-reads whole files; with our LRU(32) the pickles above are
        # unpickled once each -- time the uncached cost of one item as well
        t0 = time.perf_counter()
        for i in order[:20]:
            ds._load_file.cache_clear()
            ds[int(i)]
        t_item_cold = (time.perf_counter() - t0) / 20
        ds.build_trial_store(os.path.join(d, "trials.eegx"))
        for s in range(0, n - a.batch + 1, a.batch):               # first epoch: tokenises and caches
            ds.fetch(order[s:s + a.batch])
        t0 = time.perf_counter()
        for s in range(0, n - a.batch + 1, a.batch):
            ds.fetch(order[s:s + a.batch])
        t_store = time.perf_counter() - t0
        t0 = time.perf_counter()
        for s in range(0, n - a.batch + 1, a.batch):
            ds.store.batch(order[s:s + a.batch], pin=False)
        t_gather = time.perf_counter() - t0
    print(json.dumps({"trials": n, "shape": [125, 1651], "batch": a.batch,
                      "pickle_lru_items_per_s": round(done / t_pickle, 1),
                      "pickle_cold_ms_per_item": round(1e3 * t_item_cold, 2),
                      "store_fetch_items_per_s": round(done / t_store, 1),
                      "store_gather_only_items_per_s": round(done / t_gather, 1),
                      "store_gather_GBps": round(done * 125 * 1651 * 4 / t_gather / 1e9, 2)}))


if __name__ == "__main__":
    main()
