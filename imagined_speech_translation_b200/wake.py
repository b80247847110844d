"""wake_model's dense head on the GPU (BASELINE config 5).

Mirror of the two ``Linear`` layers at the end of ``wake_model/train.cpp:38-39`` and of the per-sample loop
``train.cpp:68-117`` restricted to them: ``Linear(in, hidden, "relu")`` -> ``Linear(hidden, n_cls, "softmax", true)``
-> categorical cross-entropy, SGD with the update applied inside ``backward`` (``layers/linear.cpp:47-72``), fp64,
one sample at a time.  All arithmetic runs in ``eegx_wake_dense_f64`` (``csrc/wake_dense.cu``); there is no CPU
path -- without libeegx.so / a B200 the calls raise.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib

ACTIVATIONS = {"": 0, None: 0, "relu": 1, "sigmoid": 2, "tanh": 3}


class DenseHead:
    """``weights`` / ``biases`` follow ``Linear``'s layout (``layers/linear.h:25-33``): ``w1`` (hidden, in),
    ``w2`` (n_cls, hidden), He-normal initialisation ``N(0, sqrt(2 / fan_in))`` for weights AND biases."""

    def __init__(self, input_size: int, hidden_size: int = 1024, n_classes: int = 2, activation: str = "relu",
                 device="cuda", generator: Optional[torch.Generator] = None):
        if activation not in ACTIVATIONS:
            raise ValueError(f"Unknown activation function: {activation}")       # activations.h:76
        self.input_size, self.hidden_size, self.n_classes = int(input_size), int(hidden_size), int(n_classes)
        self.activation = activation
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.EegxError("DenseHead runs on a CUDA device only (no CPU fallback)")
        s1, s2 = math.sqrt(2.0 / input_size), math.sqrt(2.0 / hidden_size)
        rn = lambda *shape: torch.randn(*shape, dtype=torch.float64, generator=generator).to(dev)
        self.w1, self.b1 = rn(hidden_size, input_size) * s1, rn(hidden_size) * s1
        self.w2, self.b2 = rn(n_classes, hidden_size) * s2, rn(n_classes) * s2
        self._ws = None

    def load(self, w1, b1, w2, b2) -> "DenseHead":
        dev = self.w1.device
        for name, t, shape in (("w1", w1, self.w1.shape), ("b1", b1, self.b1.shape), ("w2", w2, self.w2.shape),
                               ("b2", b2, self.b2.shape)):
            t = torch.as_tensor(t, dtype=torch.float64).to(dev).contiguous().clone()
            if t.shape != shape:
                raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
            setattr(self, name, t)
        return self

    def _run(self, x: torch.Tensor, label: torch.Tensor, lr: float, train: bool, want_dx: bool):
        lib = _lib.lib()
        if not (x.is_cuda and x.dtype == torch.float64 and x.dim() == 2 and x.shape[1] == self.input_size):
            raise _lib.EegxError(f"x must be a CUDA float64 (n, {self.input_size}) tensor")
        x = x.contiguous()
        label = label.to(device=x.device, dtype=torch.int32).contiguous()
        n = x.shape[0]
        if label.shape != (n,):
            raise _lib.EegxError("label must have shape (n,)")
        loss = torch.empty(n, dtype=torch.float64, device=x.device)
        probs = torch.empty(n, self.n_classes, dtype=torch.float64, device=x.device)
        dx = torch.empty(n, self.input_size, dtype=torch.float64, device=x.device) if want_dx else None
        need = lib.eegx_wake_dense_workspace_bytes(self.input_size, self.hidden_size, self.n_classes, int(want_dx))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
        _lib.check(lib.eegx_wake_dense_f64(
            _lib.ptr(self.w1), _lib.ptr(self.b1), _lib.ptr(self.w2), _lib.ptr(self.b2), _lib.ptr(x), _lib.ptr(label),
            n, self.input_size, self.hidden_size, self.n_classes, float(lr), ACTIVATIONS[self.activation], int(train),
            _lib.ptr(loss), _lib.ptr(probs), _lib.ptr(dx) if dx is not None else None, _lib.ptr(self._ws),
            self._ws.numel(), _lib.stream_ptr()), "eegx_wake_dense_f64")
        return loss, probs, dx

    def train_samples(self, x: torch.Tensor, label: torch.Tensor, learning_rate: float = 0.1, want_dx: bool = False):
        """One pass of ``train.cpp:68-117`` over the samples, IN ORDER (sample s+1 sees the weights sample s wrote).
        Returns (per-sample loss, per-sample probabilities[, d loss / d x per sample = what ``Linear::backward`` of the
        hidden layer returns to the layer below])."""
        loss, probs, dx = self._run(x, label, learning_rate, True, want_dx)
        return (loss, probs, dx) if want_dx else (loss, probs)

    @torch.no_grad()
    def forward(self, x: torch.Tensor, label: Optional[torch.Tensor] = None):
        """``Linear::forward`` of both layers without updates: probabilities (and the loss when labels are given)."""
        lab = label if label is not None else torch.zeros(x.shape[0], dtype=torch.int32, device=x.device)
        loss, probs, _ = self._run(x, lab, 0.0, False, False)
        return (probs, loss) if label is not None else probs
