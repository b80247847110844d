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


# ------------------------------------------------------------------------------------------ convolution / max-pool front
def _grid64(t: torch.Tensor, name: str) -> torch.Tensor:
    if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float64 and t.dim() == 2):
        raise _lib.EegxError(f"{name} must be a 2-D CUDA float64 tensor (no CPU fallback)")
    return t.contiguous()


class Convolution:
    """``Convolution(input_width, input_height, kernel_width, kernel_height, activation)`` of
    ``wake_model/layers/convolution.h:11-13``: one ``(kernel_height, kernel_width)`` kernel and one bias, valid
    cross-correlation; the activation string is stored and -- as in the reference -- never applied.  ``forward``
    keeps the input for ``backward``, which returns the layer's input gradient and applies the SGD update in place
    (``convolution.cpp:60-112``).  Initialisation: kernel uniform(-l, l) with l = sqrt(6 / (kw * kh)), bias
    uniform(-0.05, 0.05) (``convolution.cpp:14-32``; the reference draws from ``std::random_device`` / ``rand()``)."""

    def __init__(self, input_width: int, input_height: int, kernel_width: int, kernel_height: int, activation: str = "",
                 device="cuda", generator: Optional[torch.Generator] = None):
        if activation not in ACTIVATIONS:
            raise ValueError(f"Unknown activation function: {activation}")
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.EegxError("Convolution runs on a CUDA device only (no CPU fallback)")
        self.input_width, self.input_height = int(input_width), int(input_height)
        self.kernel_width, self.kernel_height = int(kernel_width), int(kernel_height)
        self.output_width, self.output_height = self.input_width - self.kernel_width + 1, self.input_height - self.kernel_height + 1
        if self.output_width <= 0 or self.output_height <= 0:
            raise ValueError("kernel larger than the input")
        self.activation = activation
        limit = math.sqrt(6.0 / (kernel_width * kernel_height))
        u = lambda *shape: torch.rand(*shape, dtype=torch.float64, generator=generator)
        self.kernel = ((2.0 * u(kernel_height, kernel_width) - 1.0) * limit).to(dev)
        self.biases = ((u(1) - 0.5) * 0.1).to(dev)
        self._input = None

    def load(self, kernel, bias) -> "Convolution":
        dev = self.kernel.device
        k = torch.as_tensor(kernel, dtype=torch.float64).to(dev).contiguous().clone()
        if k.shape != self.kernel.shape:
            raise ValueError(f"kernel: expected shape {tuple(self.kernel.shape)}, got {tuple(k.shape)}")
        self.kernel = k
        self.biases = torch.as_tensor(bias, dtype=torch.float64).reshape(-1)[:1].to(dev).contiguous().clone()
        return self

    def forward(self, input_neurons: torch.Tensor) -> torch.Tensor:
        x = _grid64(input_neurons, "input_neurons")
        if tuple(x.shape) != (self.input_height, self.input_width):
            raise ValueError(f"expected a ({self.input_height}, {self.input_width}) input, got {tuple(x.shape)}")
        self._input = x
        y = torch.empty(self.output_height, self.output_width, dtype=torch.float64, device=x.device)
        _lib.check(_lib.lib().eegx_wake_conv2d_f64(
            _lib.ptr(self.kernel), _lib.ptr(self.biases), _lib.ptr(x), self.input_height, self.input_width,
            self.kernel_height, self.kernel_width, None, 0.0, _lib.ptr(y), None, _lib.stream_ptr()), "eegx_wake_conv2d_f64")
        return y

    def backward(self, output_gradient: torch.Tensor, learning_rate: float) -> torch.Tensor:
        if self._input is None:
            raise _lib.EegxError("Convolution.backward called before forward")
        d = _grid64(output_gradient, "output_gradient")
        if tuple(d.shape) != (self.output_height, self.output_width):
            raise ValueError(f"expected a ({self.output_height}, {self.output_width}) gradient, got {tuple(d.shape)}")
        dx = torch.empty_like(self._input)
        _lib.check(_lib.lib().eegx_wake_conv2d_f64(
            _lib.ptr(self.kernel), _lib.ptr(self.biases), _lib.ptr(self._input), self.input_height, self.input_width,
            self.kernel_height, self.kernel_width, _lib.ptr(d), float(learning_rate), None, _lib.ptr(dx),
            _lib.stream_ptr()), "eegx_wake_conv2d_f64")
        return dx


class MaxPool:
    """``MaxPool(input_width, input_height, pool_width, pool_height, stride=1)`` of ``wake_model/layers/maxpool.h:11-20``:
    first strict maximum of every window, its position kept for ``backward`` (``maxpool.cpp:6-69``)."""

    def __init__(self, input_width: int, input_height: int, pool_width: int, pool_height: int, stride: int = 1):
        self.input_width, self.input_height = int(input_width), int(input_height)
        self.pool_width, self.pool_height, self.stride = int(pool_width), int(pool_height), int(stride)
        if self.stride < 1:
            raise ValueError("stride must be >= 1")
        self.output_height = (self.input_height - self.pool_height) // self.stride + 1
        self.output_width = (self.input_width - self.pool_width) // self.stride + 1
        if self.output_width <= 0 or self.output_height <= 0:
            raise ValueError("pooling window larger than the input")
        self.max_indices = None
        self._shape = None

    def forward(self, input_neurons: torch.Tensor) -> torch.Tensor:
        x = _grid64(input_neurons, "input_neurons")
        if tuple(x.shape) != (self.input_height, self.input_width):
            raise ValueError(f"expected a ({self.input_height}, {self.input_width}) input, got {tuple(x.shape)}")
        y = torch.empty(self.output_height, self.output_width, dtype=torch.float64, device=x.device)
        self.max_indices = torch.empty(self.output_height, self.output_width, 2, dtype=torch.int32, device=x.device)
        self._x = x
        _lib.check(_lib.lib().eegx_wake_maxpool_f64(
            _lib.ptr(x), self.input_height, self.input_width, self.pool_width, self.pool_height, self.stride, None,
            _lib.ptr(y), _lib.ptr(self.max_indices), None, _lib.stream_ptr()), "eegx_wake_maxpool_f64")
        return y

    def backward(self, output_gradient: torch.Tensor) -> torch.Tensor:
        if self.max_indices is None:
            raise _lib.EegxError("MaxPool.backward called before forward")
        d = _grid64(output_gradient, "output_gradient")
        if tuple(d.shape) != (self.output_height, self.output_width):
            raise ValueError(f"expected a ({self.output_height}, {self.output_width}) gradient, got {tuple(d.shape)}")
        dx = torch.empty_like(self._x)
        _lib.check(_lib.lib().eegx_wake_maxpool_f64(
            _lib.ptr(self._x), self.input_height, self.input_width, self.pool_width, self.pool_height, self.stride,
            _lib.ptr(d), None, _lib.ptr(self.max_indices), _lib.ptr(dx), _lib.stream_ptr()), "eegx_wake_maxpool_f64")
        return dx
