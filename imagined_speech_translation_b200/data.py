"""Data side of the hot path on the B200: the reference ``EEGDataset``
(``main_model/src/data/dataset.py:18-553``) with the arithmetic moved to the GPU.

``EEGDataset`` keeps the reference constructor and the per-item dict, but the per-trial numpy /
sklearn work -- nan_to_num, region gather, RobustScaler fit and transform, augmentation -- runs
batched on the device: ``collate_raw`` hands the trainer one ``(B, 125, T)`` tensor per batch and
``RegionNormalizer`` / ``augment_regions`` do the rest in a handful of launches.
No CPU fallback for the arithmetic: ``__getitem__`` returns the *raw* trial plus the token ids.
"""
from __future__ import annotations

import os
import pickle
from functools import lru_cache
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib, fused
from .preprocess import REGION_ORDER, RegionNormalizer

# Electrode names per region: the reference's montage split (main_model/src/data/utils.py:13-28).
ELECTRODE_REGIONS = {
    'frontal': ('FC5', 'F5', 'F7', 'F3', 'FC1', 'F1', 'AF3', 'Fz', 'FC2', 'F2', 'AF4', 'Fp2', 'F4', 'F6', 'F8', 'FC6'),
    'temporal': ('T9', 'FT9', 'T7', 'TP7', 'FT8', 'T10', 'FT10', 'T8', 'TP8'),
    'central': ('C5', 'C3', 'FC3', 'C1', 'CP1', 'Cz', 'CP2', 'C2', 'C4', 'FC4', 'C6'),
    'parietal': ('P7', 'P5', 'CP3', 'P3', 'PO3', 'PO1', 'PO2', 'P4', 'PO4', 'P6', 'CP4', 'P8'),
}
DEFAULT_TEXT = "数据样本"      # dataset.py:312, 427


def build_region_indices(ch_names: Sequence[str]) -> Dict[str, List[int]]:
    """Channel rows per region, in montage order (``_build_region_indices``, dataset.py:339-353)."""
    return {r: [i for i, ch in enumerate(ch_names) if ch in ELECTRODE_REGIONS[r]] for r in REGION_ORDER}


# ------------------------------------------------------------------------------------------ augmentation
def augment_regions(regions: List[torch.Tensor], p_noise: float = 0.3, p_scale: float = 0.2, p_shift: float = 0.15,
                    generator: Optional[torch.Generator] = None) -> List[torch.Tensor]:
    """``_augment_eeg_regions`` (dataset.py:227-261) for a batch: per (trial, region) and independently,
    Gaussian noise of 5 % of the region's std with probability 0.3, an amplitude factor U(0.9, 1.1) with
    probability 0.2, a circular time shift of -2..2 samples with probability 0.15.  The decisions are drawn
    per trial on the host RNG ``generator`` (the reference uses numpy's global stream, so parity is
    distributional), the arithmetic is one fused pass per region."""
    lib = _lib.lib()
    out = []
    for reg in regions:
        if not reg.is_cuda or reg.dtype != torch.float32:
            raise _lib.EegxError("augment_regions needs CUDA float32 regions (no CPU fallback)")
        reg = reg.contiguous()
        B, C, T = reg.shape
        u = torch.rand(B, 5, generator=generator)
        std = torch.empty(B, dtype=torch.float32, device=reg.device)
        _lib.check(lib.eegx_region_std_f32(_lib.ptr(reg), B, C * T, _lib.ptr(std), _lib.stream_ptr()),
                   "eegx_region_std_f32")
        noise_on = (u[:, 0] < p_noise).to(reg.device)
        sigma = torch.where(noise_on, (std * 0.05).clamp_min(1e-6), torch.zeros_like(std))
        scale = torch.where(u[:, 1] < p_scale, 0.9 + 0.2 * u[:, 2], torch.ones(B)).to(reg.device)
        shift = torch.where(u[:, 3] < p_shift, torch.floor(u[:, 4] * 5).to(torch.int32) - 2,
                            torch.zeros(B, dtype=torch.int32)).to(reg.device)
        out.append(apply_augmentation(reg, sigma, scale, shift))
    return out


def apply_augmentation(reg: torch.Tensor, sigma: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor) -> torch.Tensor:
    """out[b, c, t] = scale[b] * (x[b, c, (t - shift[b]) mod T] + sigma[b] * N(0, 1)) -- noise, then scaling,
    then np.roll, exactly the reference's order."""
    B, C, T = reg.shape
    res = torch.empty_like(reg)
    rng, site = fused.rng_state(reg.device), fused._next_site()
    _lib.check(_lib.lib().eegx_augment_f32(_lib.ptr(reg), _lib.ptr(res), B, C, T,
                                           _lib.ptr(sigma.float().contiguous()), _lib.ptr(scale.float().contiguous()),
                                           _lib.ptr(shift.to(torch.int32).contiguous()), _lib.ptr(rng), site,
                                           _lib.stream_ptr()), "eegx_augment_f32")
    return res


# ------------------------------------------------------------------------------------------ dataset
# ------------------------------------------------------------------------------------------ binary trial store
class TrialStore:
    """Flat binary store of the trials.

    The reference unpickles a whole run file for every ``__getitem__`` (``dataset.py:153-170``, ~4 ms per sample,
    LRU of 32 files).  The store is written once from the same pickles -- same order as ``sample_index``, malformed
    entries kept as invalid slots so indices do not shift -- and then memory-mapped: an item is a (C, T) float32
    slice, a batch is one gather into a pinned buffer followed by one H2D copy.

    Layout: ``b"EEGXTS01"`` | u64 header length | JSON header {n, C, T, data_offset, valid[], texts[]} | padding to
    4096 | float32 data (n, C, T), little endian, C-contiguous."""

    MAGIC = b"EEGXTS01"

    def __init__(self, path: str):
        import json
        with open(path, 'rb') as fh:
            if fh.read(8) != self.MAGIC:
                raise ValueError(f"{path}: not a trial store")
            hlen = int(np.frombuffer(fh.read(8), dtype='<u8')[0])
            hdr = json.loads(fh.read(hlen).decode('utf-8'))
        self.path, self.n, self.C, self.T = path, hdr['n'], hdr['C'], hdr['T']
        self.valid = np.asarray(hdr['valid'], dtype=bool)
        self.texts = hdr['texts']
        self.data = np.memmap(path, dtype='<f4', mode='r', offset=hdr['data_offset'], shape=(self.n, self.C, self.T))
        self._offset, self._trial_bytes = hdr['data_offset'], 4 * self.C * self.T
        self._fd = os.open(path, os.O_RDONLY)

    def __del__(self):
        fd = getattr(self, '_fd', None)
        if fd is not None:
            try:
                os.close(fd)
            except OSError:
                pass
            self._fd = None

    @classmethod
    def build(cls, samples, path: str, n_channels: int) -> "TrialStore":
        """``samples``: iterable of the reference's sample dicts ({'input_features': (1, C, T), 'text': str}) or
        None / malformed entries, in ``sample_index`` order."""
        import json
        arrs, valid, texts = [], [], []
        shape = None
        for smp in samples:
            ok = isinstance(smp, dict) and 'input_features' in smp and 'text' in smp
            arr = np.asarray(smp['input_features'], dtype=np.float32) if ok else None
            ok = ok and arr.ndim >= 2 and arr.shape[1] == n_channels
            if ok:
                arr = arr.squeeze()
                ok = arr.ndim == 2 and (shape is None or arr.shape == shape)
            if ok:
                shape = arr.shape
            arrs.append(arr if ok else None)
            valid.append(bool(ok))
            texts.append(smp.get('text', '') if ok else '')
        if shape is None:
            raise ValueError("no valid trial to store")
        hdr = {'n': len(arrs), 'C': int(shape[0]), 'T': int(shape[1]), 'valid': valid, 'texts': texts, 'data_offset': 0}
        for _ in range(2):                                         # the offset is part of the header it depends on
            raw = json.dumps(hdr, ensure_ascii=False).encode('utf-8')
            hdr['data_offset'] = (16 + len(raw) + 64 + 4095) // 4096 * 4096
        raw = json.dumps(hdr, ensure_ascii=False).encode('utf-8')
        zero = np.zeros(shape, dtype='<f4')
        with open(path, 'wb') as fh:
            fh.write(cls.MAGIC)
            fh.write(np.asarray([len(raw)], dtype='<u8').tobytes())
            fh.write(raw)
            fh.write(b'\0' * (hdr['data_offset'] - 16 - len(raw)))
            for arr in arrs:
                fh.write((zero if arr is None else np.ascontiguousarray(arr, dtype='<f4')).tobytes())
        return cls(path)

    def __len__(self):
        return self.n

    def trial(self, idx: int) -> Optional[np.ndarray]:
        return np.asarray(self.data[idx]) if self.valid[idx] else None

    RING = 3
    THREADS = max(1, min(8, (os.cpu_count() or 2) // 2))

    def batch(self, indices, pin: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(B, C, T) float32 host tensor (pinned when CUDA is present): one gather, ready for one H2D copy.
        Without ``out`` the result lives in a ring of RING reusable staging buffers per batch size (fresh host
        allocations cost more in first-touch page faults than the copy itself): it stays valid until RING later
        calls with the same batch size."""
        idx = np.asarray(indices, dtype=np.int64)
        if not self.valid[idx].all():
            raise ValueError(f"trial(s) {idx[~self.valid[idx]].tolist()} are malformed")
        pin = pin and torch.cuda.is_available()
        if out is None:
            ring = self.__dict__.setdefault('_ring', {})
            slot = ring.setdefault((len(idx), pin), {'bufs': [], 'next': 0})
            if len(slot['bufs']) < self.RING:
                slot['bufs'].append(torch.empty((len(idx), self.C, self.T), dtype=torch.float32, pin_memory=pin))
            out = slot['bufs'][slot['next'] % len(slot['bufs'])]
            slot['next'] += 1
        elif tuple(out.shape) != (len(idx), self.C, self.T) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("out must be a contiguous float32 (B, C, T) host tensor")
        order = np.argsort(idx, kind='stable')                     # ascending file offsets
        dst = out.numpy()

        def read(js):                                              # positional reads straight into the batch buffer
            for j in js:
                got = os.preadv(self._fd, [memoryview(dst[j]).cast('B')],
                                self._offset + int(idx[j]) * self._trial_bytes)
                if got != self._trial_bytes:
                    raise IOError(f"{self.path}: short read for trial {int(idx[j])}")

        nthr = min(self.THREADS, max(1, len(idx) // 8))
        if nthr <= 1:
            read(order)
        else:                                                      # preadv releases the GIL: the copies run in parallel
            if getattr(self, '_pool', None) is None:
                from concurrent.futures import ThreadPoolExecutor
                self._pool = ThreadPoolExecutor(max_workers=self.THREADS)
            list(self._pool.map(read, np.array_split(order, nthr)))
        return out


class EEGDataset(torch.utils.data.Dataset):
    """Same constructor as the reference (dataset.py:23-24).  ``ds[i]`` returns
    ``{'raw': (125, T) float32, 'decoder_input_ids', 'labels', 'attention_mask'}``; ``collate_raw`` stacks
    a batch; ``normalizer()`` is the GPU ``RegionNormalizer`` fitted the reference's way (a random subset
    of min(100, max(10, N // 10)) samples, RobustScaler(5, 95)); ``to_regions(batch)`` yields the list of
    four ``(B, C_r, T)`` tensors the reference's ``ds[i]['eeg']`` holds (augmented if enabled)."""

    def __init__(self, data_dir, csv_path, tokenizer, max_length=64, eps=1e-6, max_samples=None,
                 data_augmentation=True, precompute_stats=False, device="cuda", trial_store: Optional[str] = None):
        import pandas as pd
        self.tokenizer = tokenizer
        self.max_length = max_length
        self.eps = eps
        self.max_samples = max_samples
        self.data_augmentation = data_augmentation
        self.precompute_stats = precompute_stats
        self.device = torch.device(device)
        self.vocab_size = len(tokenizer.get_vocab())
        self.ch_names = pd.read_csv(csv_path)['label'].to_numpy()
        self.region_indices = build_region_indices(self.ch_names)
        self.region_channel_counts = {r: len(ix) for r, ix in self.region_indices.items()}
        for r, ix in self.region_indices.items():
            if not ix:
                raise ValueError(f"No channels found for {r} region!")
        if self.tokenizer.pad_token is None:                      # _setup_tokenizer_safe (dataset.py:365-385)
            self.tokenizer.pad_token = self.tokenizer.eos_token
        if not os.path.exists(data_dir):
            raise FileNotFoundError(f"Data directory not found: {data_dir}")
        self.data_files = [os.path.join(data_dir, f) for f in os.listdir(data_dir) if f.endswith('.pkl')]
        if not self.data_files:
            raise ValueError(f"No .pkl files found in {data_dir}")
        self.sample_index = self._build_sample_index()
        self._normalizer = None
        self._load_file = lru_cache(maxsize=32)(self._load_file_uncached)
        self.store = TrialStore(trial_store) if trial_store else None
        if self.store is not None and len(self.store) != len(self.sample_index):
            raise ValueError(f"trial store holds {len(self.store)} trials, the pickles {len(self.sample_index)}")

    # -- indexing / loading (dataset.py:71-100, 153-170) ------------------------------------------------
    def _build_sample_index(self):
        index = []
        for path in self.data_files:
            with open(path, 'rb') as fh:
                loaded = pickle.load(fh)
            n = len(loaded) if isinstance(loaded, list) else 1
            for i in range(n):
                index.append({'file': path, 'index': i})
                if self.max_samples and len(index) >= self.max_samples:
                    return index
        return index

    @staticmethod
    def _load_file_uncached(path):
        with open(path, 'rb') as fh:
            loaded = pickle.load(fh)
        return loaded if isinstance(loaded, list) else [loaded]

    def build_trial_store(self, path: str) -> TrialStore:
        """Write the binary store from the pickles (once) and switch this dataset to it."""
        def samples():
            for info in self.sample_index:
                yield self._load_file(info['file'])[info['index']]
        self.store = TrialStore.build(samples(), path, len(self.ch_names))
        return self.store

    def _raw(self, idx) -> Optional[dict]:
        if self.store is not None:
            arr = self.store.trial(idx)
            return None if arr is None else {'eeg': arr, 'text': self.store.texts[idx]}
        info = self.sample_index[idx]
        sample = self._load_file(info['file'])[info['index']]
        if not isinstance(sample, dict) or 'input_features' not in sample or 'text' not in sample:
            return None
        arr = np.asarray(sample['input_features'], dtype=np.float32)
        if arr.ndim < 2 or arr.shape[1] != len(self.ch_names):      # _validate_sample: (1, 125, T)
            return None
        return {'eeg': arr.squeeze(), 'text': sample.get('text', '')}

    def __len__(self):
        return len(self.sample_index)

    # -- tokenisation (dataset.py:422-494) ----------------------------------------------------------------
    def _safe_tokenize(self, text):
        if not text or not isinstance(text, str) or not text.strip():
            text = DEFAULT_TEXT
        enc = self.tokenizer(text.strip(), max_length=self.max_length, padding='max_length', truncation=True,
                             return_tensors='pt', add_special_tokens=True)
        input_ids = enc['input_ids'].squeeze(0).clamp(0, self.vocab_size - 1)
        start = self.tokenizer.bos_token_id
        if start is None:
            start = self.tokenizer.eos_token_id
        if start is None or start >= self.vocab_size:
            start = self.tokenizer.pad_token_id
        decoder_input_ids = torch.cat([torch.tensor([start]), input_ids[:-1]]).clamp(0, self.vocab_size - 1)
        labels = input_ids.clone()
        labels[input_ids == self.tokenizer.pad_token_id] = -100
        return {'decoder_input_ids': decoder_input_ids, 'labels': labels,
                'attention_mask': enc['attention_mask'].squeeze(0)}

    def __getitem__(self, idx):
        raw = self._raw(idx)
        if raw is None:
            raise ValueError(f"sample {idx} is malformed (the reference silently substitutes zeros here)")
        return {'raw': torch.from_numpy(np.ascontiguousarray(raw['eeg'])), **self._safe_tokenize(raw['text'])}

    @staticmethod
    def collate_raw(items):
        return {k: torch.stack([it[k] for it in items]) for k in items[0]}

    def fetch(self, indices, out: Optional[torch.Tensor] = None) -> dict:
        """A whole batch by index list: with a trial store the raw trials are ONE gather into a pinned buffer
        (no per-item tensors, no default_collate); tokens are computed once per sample and cached.  ``out``: the
        caller's own (B, C, T) staging buffer (``PrefetchLoader`` passes the slot it has acquired); without it the
        store's ring is used and the result is only valid until ``TrialStore.RING`` later calls."""
        if self.store is None:
            return self.collate_raw([self[int(i)] for i in indices])
        if not hasattr(self, '_tok_cache'):
            self._tok_cache = {}
        toks = []
        for i in indices:
            i = int(i)
            if i not in self._tok_cache:
                self._tok_cache[i] = self._safe_tokenize(self.store.texts[i])
            toks.append(self._tok_cache[i])
        res = {'raw': self.store.batch(indices, out=out)}
        res.update({k: torch.stack([t[k] for t in toks]) for k in toks[0]})
        return res

    # -- GPU normalisation / augmentation -----------------------------------------------------------------
    def normalizer(self) -> RegionNormalizer:
        if self._normalizer is None:
            n_all = len(self.sample_index)
            size = min(min(100, max(10, n_all // 10)), n_all)       # dataset.py:105-108
            chosen = np.random.choice(n_all, size=size, replace=False)
            fit = [self._raw(int(i)) for i in chosen]
            fit = torch.from_numpy(np.stack([f['eeg'] for f in fit if f is not None]))
            self._normalizer = RegionNormalizer.fit(fit, self.region_indices, device=self.device)
        return self._normalizer

    def to_regions(self, batch, generator: Optional[torch.Generator] = None) -> List[torch.Tensor]:
        regions = self.normalizer()(batch['raw'].to(self.device, non_blocking=True).float())
        return augment_regions(regions, generator=generator) if self.data_augmentation else regions


# ------------------------------------------------------------------------------------------ batch prefetcher
class PrefetchLoader:
    """Iterates an ``EEGDataset`` in whole batches, built ``depth`` batches ahead by one background thread.

    Stands where ``DataLoader(dataset, batch_size, shuffle, num_workers=0)`` stands in the reference
    (``scripts/train.py:160-196``), but a batch is ONE ``dataset.fetch(indices)`` (with a trial store: one threaded
    gather into a pinned staging buffer) instead of ``batch_size`` ``__getitem__`` calls + ``default_collate``, and the
    next batches are assembled while the GPU runs the current step.  Order: a seeded permutation per epoch
    (``set_epoch``), sequential without ``shuffle``.

    Staging-buffer lifetime (trial-store datasets).  The loader owns ``depth + keep + 2`` pinned buffers per batch
    size: ``depth`` queued, one in the producer's hands, the one just yielded, and ``keep`` older ones.  A yielded
    batch's ``'raw'`` tensor stays valid while the consumer works on it AND on the next ``keep`` batches; its slot
    goes back to the producer when batch ``k + keep + 1`` is requested.  At that moment a CUDA event is recorded on
    the consumer's current stream and the producer waits for it before overwriting the buffer, so an asynchronous
    ``.to(device, non_blocking=True)`` (or a copy into the static inputs of a captured CUDA graph) that was
    enqueued before the next batch was requested always reads the data it was given -- also when the host runs
    several steps ahead of the GPU."""

    def __init__(self, dataset, batch_size: int, shuffle: bool = True, drop_last: bool = False, seed: int = 0,
                 depth: int = 2, indices: Optional[Sequence[int]] = None, keep: int = 1):
        if batch_size < 1:
            raise ValueError("batch_size must be >= 1")
        if depth < 1 or keep < 0:
            raise ValueError("depth must be >= 1 and keep >= 0")
        self.dataset, self.batch_size, self.shuffle, self.drop_last = dataset, int(batch_size), shuffle, drop_last
        self.seed, self.depth, self.keep, self.epoch = int(seed), int(depth), int(keep), 0
        self.indices = np.arange(len(dataset)) if indices is None else np.asarray(indices, dtype=np.int64)
        self._slots: Dict[int, list] = {}          # batch size -> staging buffers owned by this loader

    def set_epoch(self, epoch: int) -> None:
        self.epoch = int(epoch)

    def __len__(self):
        n = len(self.indices)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def batches(self) -> List[np.ndarray]:
        order = self.indices
        if self.shuffle:
            order = order[np.random.default_rng((self.seed, self.epoch)).permutation(len(order))]
        out = [order[s:s + self.batch_size] for s in range(0, len(order), self.batch_size)]
        if self.drop_last and out and len(out[-1]) < self.batch_size:
            out.pop()
        return out

    def _staging(self, n: int) -> list:
        store = getattr(self.dataset, 'store', None)
        if store is None:
            return []
        bufs = self._slots.get(n)
        if bufs is None:
            pin = torch.cuda.is_available()
            bufs = [torch.empty((n, store.C, store.T), dtype=torch.float32, pin_memory=pin)
                    for _ in range(self.depth + self.keep + 2)]
            self._slots[n] = bufs
        return bufs

    def __iter__(self):
        import collections
        import queue
        import threading
        plan = self.batches()
        q: "queue.Queue" = queue.Queue(maxsize=self.depth)
        stop = threading.Event()
        use_slots = getattr(self.dataset, 'store', None) is not None
        free: Dict[int, "queue.Queue"] = {}
        if use_slots:
            for n in sorted({len(b) for b in plan}):
                free[n] = queue.Queue()
                for buf in self._staging(n):
                    free[n].put((buf, None))

        def work():
            try:
                for idx in plan:
                    buf = None
                    if use_slots:
                        while buf is None:                              # wait for a buffer the consumer has released
                            if stop.is_set():
                                return
                            try:
                                buf, event = free[len(idx)].get(timeout=0.1)
                            except queue.Empty:
                                continue
                        if event is not None:
                            event.synchronize()                        # copies the consumer enqueued from it are done
                    if stop.is_set():
                        return
                    item = self.dataset.fetch(idx, out=buf) if use_slots else self.dataset.fetch(idx)
                    while not stop.is_set():
                        try:
                            q.put((item, buf), timeout=0.1)
                            break
                        except queue.Full:
                            continue
                q.put(None)
            except BaseException as exc:                              # surfaces in the consumer, never swallowed
                q.put(exc)

        t = threading.Thread(target=work, name="eegx-prefetch", daemon=True)
        t.start()
        held: "collections.deque" = collections.deque()
        try:
            while True:
                got = q.get()
                if got is None:
                    return
                if isinstance(got, BaseException):
                    raise got
                item, buf = got
                if buf is not None:
                    held.append(buf)
                yield item
                # the consumer is asking for the next batch: the oldest held buffer goes back, fenced by an event
                while len(held) > self.keep:
                    old = held.popleft()
                    event = None
                    if torch.cuda.is_available() and torch.cuda.is_initialized():
                        event = torch.cuda.Event()
                        event.record()
                    free[old.shape[0]].put((old, event))
        finally:
            stop.set()
            t.join(timeout=5.0)
