"""Golden fixture for the dataset mirror, produced by the REFERENCE's own EEGDataset
(main_model/src/data/dataset.py) imported from /root/reference (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_dataset_golden.py

Writes tests/golden/dataset/{run0.pkl, run1.pkl, vocab.txt, montage.csv} (the synthetic input) and
tests/golden/dataset_ref.npz (what the reference returns for it: region indices, fitted scalers,
per-item regions with augmentation off, token ids)."""
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/main_model"
OUT = os.path.join(HERE, "dataset")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

import pandas as pd  # noqa: E402
from transformers import BertTokenizer  # noqa: E402
from src.data.dataset import EEGDataset  # noqa: E402  (the reference itself)

TEXTS = ["我想喝水", "今天天气很好", "请帮我开灯", "", "谢谢你", "我有点冷"]
CHARS = sorted(set("".join(TEXTS) + "数据样本"))

rng = np.random.default_rng(11)
T = 64
labels = pd.read_csv(os.path.join(REF, "data/montage.csv"))["label"]
pd.DataFrame({"label": labels}).to_csv(os.path.join(OUT, "montage.csv"), index=False)
with open(os.path.join(OUT, "vocab.txt"), "w", encoding="utf-8") as fh:
    fh.write("\n".join(["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + CHARS) + "\n")
k = 0
for f in range(2):
    items = []
    for i in range(3):
        off = rng.uniform(-20, 20, size=(1, 125, 1))
        amp = rng.uniform(5, 30, size=(1, 125, 1))
        arr = (off + amp * rng.standard_normal((1, 125, T))).astype(np.float32)
        if k == 1:
            arr[0, 3, 5] = np.nan
            arr[0, 40, 7] = np.inf
        items.append({"input_features": arr, "text": TEXTS[k]})
        k += 1
    with open(os.path.join(OUT, f"run{f}.pkl"), "wb") as fh:
        pickle.dump(items, fh)

tok = BertTokenizer(os.path.join(OUT, "vocab.txt"), bos_token="[CLS]", eos_token="[SEP]")
np.random.seed(3)
ds = EEGDataset(OUT, os.path.join(OUT, "montage.csv"), tok, max_length=16, data_augmentation=False)
order = [os.path.basename(s["file"]) + ":" + str(s["index"]) for s in ds.sample_index]
rec = {"order": np.array(order), "n": len(ds)}
for name, idx in ds.region_indices.items():
    rec[f"idx_{name}"] = np.asarray(idx, dtype=np.int32)
    rec[f"center_{name}"] = ds.scalers[name].center_.astype(np.float32)
    rec[f"scale_{name}"] = ds.scalers[name].scale_.astype(np.float32)
for i in range(len(ds)):
    it = ds[i]
    for r, name in enumerate(["frontal", "temporal", "central", "parietal"]):
        rec[f"eeg_{i}_{name}"] = np.asarray(it["eeg"][r], dtype=np.float32)
    rec[f"ids_{i}"] = it["decoder_input_ids"].numpy()
    rec[f"labels_{i}"] = it["labels"].numpy()
    rec[f"mask_{i}"] = it["attention_mask"].numpy()
np.savez_compressed(os.path.join(HERE, "dataset_ref.npz"), **rec)
print("samples", len(ds), "order", order, "bytes", os.path.getsize(os.path.join(HERE, "dataset_ref.npz")))
