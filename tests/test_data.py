"""The data-side rows of the hot path against fixtures produced by the
REFERENCE's own EEGDataset (tests/golden/make_dataset_golden.py, make_golden.py):

  CPU:  indexing, validation and tokenisation of the dataset mirror (token ids are integer: bit-exact)
  GPU:  RobustScaler fit by radix select (rtol 1e-6 vs sklearn's center_/scale_), the batched
        region normalisation of whole items (<= 1e-6 inf-norm-relative), augmentation semantics
        (exact for scaling + roll; the Gaussian noise and the Bernoulli decisions statistically).
"""
import os

import numpy as np
import pytest
import torch

import imagined_speech_translation_b200 as pkg

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "golden", "dataset")
NAMES = ("frontal", "temporal", "central", "parietal")


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(HERE, "golden", "dataset_ref.npz"))


@pytest.fixture(scope="module")
def tokenizer():
    from transformers import BertTokenizer
    return BertTokenizer(os.path.join(FIX, "vocab.txt"), bos_token="[CLS]", eos_token="[SEP]")


def _dataset(tokenizer, **kw):
    return pkg.EEGDataset(FIX, os.path.join(FIX, "montage.csv"), tokenizer, max_length=16,
                          data_augmentation=False, **kw)


def _by_key(ds):
    return {os.path.basename(s["file"]) + ":" + str(s["index"]): i for i, s in enumerate(ds.sample_index)}


def test_dataset_mirror_index_and_tokens(ref, tokenizer):
    ds = _dataset(tokenizer, device="cpu")
    assert len(ds) == int(ref["n"])
    for n in NAMES:
        assert list(ds.region_indices[n]) == list(ref[f"idx_{n}"])
    assert ds.region_channel_counts == {"frontal": 16, "temporal": 9, "central": 11, "parietal": 12}
    mine = _by_key(ds)
    for j, key in enumerate(ref["order"]):
        item = ds[mine[str(key)]]
        assert item["raw"].shape == (125, 64) and item["raw"].dtype == torch.float32
        assert torch.equal(item["decoder_input_ids"], torch.from_numpy(ref[f"ids_{j}"]))
        assert torch.equal(item["labels"], torch.from_numpy(ref[f"labels_{j}"]))
        assert torch.equal(item["attention_mask"], torch.from_numpy(ref[f"mask_{j}"]))
    batch = ds.collate_raw([ds[0], ds[1]])
    assert batch["raw"].shape == (2, 125, 64) and batch["labels"].shape == (2, 16)


def _same_bits(a, b):
    """Bit-pattern equality (the fixture trials contain NaN / inf on purpose)."""
    if a.dtype.is_floating_point:
        return a.shape == b.shape and torch.equal(a.contiguous().view(torch.int32), b.contiguous().view(torch.int32))
    return torch.equal(a, b)


def test_trial_store_serves_the_same_items_as_the_pickles(tokenizer, tmp_path):
    """Binary trial store (row f2): built from the pickles, memory-mapped, item-for-item identical (bit-exact,
    floats and token ids), malformed slots preserved, batches gathered in one go."""
    import pickle
    ds = _dataset(tokenizer, device="cpu")
    path = str(tmp_path / "trials.eegx")
    store = ds.build_trial_store(path)
    assert len(store) == len(ds) and store.valid.all() and (store.C, store.T) == (125, 64)
    plain = _dataset(tokenizer, device="cpu")
    again = _dataset(tokenizer, device="cpu", trial_store=path)              # re-opened from disk
    for i in range(len(ds)):
        a, b = plain[i], again[i]
        for k in a:
            assert _same_bits(a[k], b[k]), (i, k)
    idx = [5, 0, 3, 3]
    got = again.fetch(idx)
    want = plain.collate_raw([plain[i] for i in idx])
    assert set(got) == set(want)
    for k in want:
        assert _same_bits(got[k], want[k]), k
    assert _same_bits(plain.fetch(idx)["raw"], want["raw"])                  # fetch() without a store: same result
    # malformed entries keep their slot (indices must not shift) and are refused like the pickle path refuses them
    good = {"input_features": np.ones((1, 125, 8), np.float32), "text": "a"}
    bad = {"input_features": np.ones((1, 7, 8), np.float32), "text": "b"}
    st = pkg.TrialStore.build([good, bad, None, good], str(tmp_path / "m.eegx"), 125)
    assert st.valid.tolist() == [True, False, False, True] and st.trial(1) is None
    assert st.batch([3, 0], pin=False).shape == (2, 125, 8)
    with pytest.raises(ValueError):
        st.batch([0, 1], pin=False)
    with pytest.raises(ValueError):
        pkg.TrialStore(os.path.join(FIX, "vocab.txt"))                       # not a store
    with pytest.raises(ValueError):
        _dataset(tokenizer, device="cpu", trial_store=str(tmp_path / "m.eegx"))   # 4 trials vs the pickles' count


def test_prefetch_loader_batches_in_seeded_order(tokenizer, tmp_path):
    """PrefetchLoader: whole batches through dataset.fetch from a background thread -- seeded order per epoch, every
    sample exactly once, drop_last, identical content to direct fetches, errors raised in the consumer."""
    ds = _dataset(tokenizer, device="cpu")
    ds.build_trial_store(str(tmp_path / "t.eegx"))
    n = len(ds)
    loader = pkg.PrefetchLoader(ds, batch_size=4, shuffle=True, seed=7)
    plan = loader.batches()
    assert sorted(np.concatenate(plan).tolist()) == list(range(n)) and len(loader) == len(plan) == (n + 3) // 4
    got = [{k: v.clone() for k, v in b.items()} for b in loader]      # staging buffers are recycled: copy on receipt
    assert len(got) == len(plan)
    for b, idx in zip(got, plan):
        want = ds.fetch(idx)
        assert b["raw"].shape[0] == len(idx)
        assert torch.equal(b["labels"], want["labels"]) and _same_bits(b["raw"].clone(), want["raw"].clone())
    assert [p.tolist() for p in loader.batches()] == [p.tolist() for p in plan]          # same epoch: same order
    loader.set_epoch(1)
    assert [p.tolist() for p in loader.batches()] != [p.tolist() for p in plan]          # next epoch: reshuffled
    seq = pkg.PrefetchLoader(ds, batch_size=4, shuffle=False, drop_last=True)
    assert [p.tolist() for p in seq.batches()] == [list(range(s, s + 4)) for s in range(0, n - n % 4, 4)]
    assert sum(b["raw"].shape[0] for b in seq) == n - n % 4
    it = iter(pkg.PrefetchLoader(ds, batch_size=2, shuffle=False))                        # abandoning an iterator is fine
    next(it)
    it.close()
    bad = pkg.PrefetchLoader(ds, batch_size=2, shuffle=False, indices=[0, n + 5])
    with pytest.raises(IndexError):
        list(bad)
    with pytest.raises(ValueError):
        pkg.PrefetchLoader(ds, batch_size=2, depth=0)


def test_prefetch_loader_never_rewrites_a_batch_the_consumer_still_holds(tokenizer, tmp_path):
    """A slow consumer that keeps each yielded batch (uncloned) while the producer runs ahead: the batch it holds,
    and the `keep` batches before it, must still hold their own trials when it finally reads them (round-1 bug:
    depth + 2 live buffers against a ring of 3 -> batch k silently carried the trials of batch k + 3)."""
    import time
    ds = _dataset(tokenizer, device="cpu")
    ds.build_trial_store(str(tmp_path / "t.eegx"))
    for depth, keep in ((1, 1), (2, 1), (3, 0), (2, 3)):
        loader = pkg.PrefetchLoader(ds, batch_size=2, shuffle=False, depth=depth, keep=keep)
        plan = loader.batches()
        window = []
        for k, batch in enumerate(loader):
            time.sleep(0.02)                                   # let the producer fill the queue and block
            window.append((k, batch))
            window = window[-(keep + 1):]
            for j, old in window:                              # everything still inside the validity window
                want = np.stack([ds.store.trial(int(i)) for i in plan[j]])
                assert _same_bits(old["raw"], torch.from_numpy(want)), (depth, keep, k, j)
        assert k == len(plan) - 1


def test_dataset_mirror_rejects_bad_input(tokenizer, tmp_path):
    with pytest.raises(FileNotFoundError):
        pkg.EEGDataset(str(tmp_path / "missing"), os.path.join(FIX, "montage.csv"), tokenizer)
    with pytest.raises(ValueError):
        pkg.EEGDataset(str(tmp_path), os.path.join(FIX, "montage.csv"), tokenizer)      # no .pkl files


@pytest.mark.gpu
def test_robust_fit_matches_reference():
    norm = np.load(os.path.join(HERE, "golden", "normalize_ref.npz"))
    idx = {n: norm[f"idx_{n}"] for n in NAMES}
    rn = pkg.RegionNormalizer.fit(torch.from_numpy(norm["fit_samples"]), idx)
    for n in NAMES:
        np.testing.assert_allclose(rn.centers[n].cpu().numpy(), norm[f"center_{n}"], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(rn.scales[n].cpu().numpy(), norm[f"scale_{n}"], rtol=1e-6, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("n,T,q", [(7, 33, (5.0, 95.0)), (100, 1651, (5.0, 95.0)), (3, 10, (25.0, 75.0)), (1, 1, (5.0, 95.0))])
def test_robust_fit_vs_numpy_percentile(n, T, q):
    rng = np.random.default_rng(n * T)
    x = (rng.standard_normal((n, 6, T)) * rng.uniform(1, 50, (1, 6, 1)) + rng.uniform(-9, 9, (1, 6, 1))).astype(np.float32)
    x[:, 5] = 3.25                                                    # constant channel: scale -> 1
    x[0, 2, 0] = np.nan
    idx = {"frontal": [0, 1], "temporal": [2], "central": [3, 4], "parietal": [5]}
    rn = pkg.RegionNormalizer.fit(torch.from_numpy(x), idx, quantile_range=q)
    clean = np.nan_to_num(x, nan=0.0, posinf=10.0, neginf=-10.0)
    for name, rows in idx.items():
        flat = clean[:, rows].transpose(1, 0, 2).reshape(len(rows), -1).astype(np.float64)
        cen = np.median(flat, axis=1)
        lo, hi = np.percentile(flat, q, axis=1)
        sca = hi - lo
        sca[sca < 10 * np.finfo(np.float32).eps] = 1.0
        np.testing.assert_allclose(rn.centers[name].cpu().numpy(), cen, rtol=2e-6, atol=1e-6)
        np.testing.assert_allclose(rn.scales[name].cpu().numpy(), sca, rtol=2e-6, atol=1e-6)


@pytest.mark.gpu
def test_dataset_mirror_regions_match_reference(ref, tokenizer):
    np.random.seed(0)
    ds = _dataset(tokenizer)
    rn = ds.normalizer()                       # all 6 samples are in the fit subset, as in the reference run
    for n in NAMES:
        np.testing.assert_allclose(rn.centers[n].cpu().numpy(), ref[f"center_{n}"], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(rn.scales[n].cpu().numpy(), ref[f"scale_{n}"], rtol=1e-6, atol=1e-6)
    mine = _by_key(ds)
    keys = [str(k) for k in ref["order"]]
    batch = ds.collate_raw([ds[mine[k]] for k in keys])
    regions = ds.to_regions(batch)
    for r, n in enumerate(NAMES):
        got = regions[r].cpu().numpy()
        for j in range(len(keys)):
            want = ref[f"eeg_{j}_{n}"]
            assert np.abs(got[j] - want).max() / np.abs(want).max() <= 2e-6


@pytest.mark.gpu
def test_augmentation_semantics():
    g = torch.Generator().manual_seed(4)
    x = torch.randn(6, 9, 50, generator=g).cuda()
    scale = torch.tensor([1.0, 1.07, 0.93, 1.0, 1.1, 0.9]).cuda()
    shift = torch.tensor([0, 2, -2, 1, -1, 0], dtype=torch.int32).cuda()
    zero = torch.zeros(6).cuda()
    out = pkg.apply_augmentation(x, zero, scale, shift).cpu().numpy()
    for b in range(6):
        want = np.roll(x[b].cpu().numpy(), int(shift[b]), axis=1) * float(scale[b])
        np.testing.assert_allclose(out[b], want, rtol=1e-6, atol=0)
    # noise: N(0, sigma^2), applied before the scaling and rolled with the signal
    big = torch.zeros(4, 16, 4096).cuda()
    sigma = torch.tensor([0.5, 0.0, 2.0, 1.0]).cuda()
    res = pkg.apply_augmentation(big, sigma, torch.tensor([1.0, 1.0, 1.0, 2.0]).cuda(),
                                 torch.zeros(4, dtype=torch.int32).cuda())
    assert float(res[1].abs().max()) == 0.0
    for b, s in ((0, 0.5), (2, 2.0), (3, 2.0)):
        assert abs(float(res[b].mean())) < 0.03 * s and abs(float(res[b].std()) / s - 1) < 0.02
    # Bernoulli decisions of the batched front door
    regs = pkg.augment_regions([torch.ones(4000, 2, 8).cuda()], generator=torch.Generator().manual_seed(1))[0].cpu()
    changed_scale = ((regs.mean(dim=(1, 2)) - 1).abs() > 1e-3).float().mean().item()   # noise or scaling changed the mean
    assert 0.15 < changed_scale < 0.55
