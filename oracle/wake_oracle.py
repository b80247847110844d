"""TEST INFRASTRUCTURE -- ctypes loaders for the wake_model dense-path checkers.

``oracle()``  -> oracle/_build/libwake_oracle.so  (C restatement, oracle/wake_dense_oracle.c)
``reference()`` -> oracle/_ref/libwake_ref.so     (the reference's own linear.cpp compiled here; None when absent)

Only tests/, ``__graft_entry__.smoke()`` and CPU-baseline timing legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ACTIVATIONS = {"": 0, None: 0, "relu": 1, "sigmoid": 2, "tanh": 3}

_SIG = [C.c_void_p] * 6 + [C.c_long, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int] + [C.c_void_p] * 3


def build(quiet: bool = True) -> None:
    """``make -C oracle``: the C restatement always, the reference build only where /root/reference exists."""
    subprocess.run(["make", "-C", HERE], check=True, capture_output=quiet)


def _load(path: str, symbol: str):
    if not os.path.exists(path):
        return None
    fn = getattr(C.CDLL(path), symbol)
    fn.argtypes = _SIG
    fn.restype = C.c_int
    return fn


def oracle():
    path = os.path.join(HERE, "_build", "libwake_oracle.so")
    if not os.path.exists(path):
        build()
    return _load(path, "wake_dense_oracle")


def reference():
    return _load(os.path.join(HERE, "_ref", "libwake_ref.so"), "wake_dense_ref")


def run(fn, w1, b1, w2, b2, x, label, lr=0.1, activation="relu", train=True, want_dx=False):
    """Runs ``fn`` (oracle() or reference()) on COPIES of the parameters.  Returns a dict with the updated
    parameters, per-sample loss, probabilities and (optionally) the input gradients."""
    w1, b1, w2, b2 = (np.array(a, dtype=np.float64, order="C", copy=True) for a in (w1, b1, w2, b2))
    x = np.ascontiguousarray(x, dtype=np.float64)
    label = np.ascontiguousarray(label, dtype=np.int32)
    n, n_in = x.shape
    hidden, ncls = w1.shape[0], w2.shape[0]
    loss = np.zeros(n)
    probs = np.zeros((n, ncls))
    dx = np.zeros((n, n_in)) if want_dx else None
    p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    rc = fn(p(w1), p(b1), p(w2), p(b2), p(x), p(label), n, n_in, hidden, ncls, float(lr), ACTIVATIONS[activation],
            int(train), p(loss), p(probs), p(dx))
    assert rc == 0
    return dict(w1=w1, b1=b1, w2=w2, b2=b2, loss=loss, probs=probs, dx=dx)


# ---- convolution / max-pool front (oracle/wake_conv_oracle.c, wake_ref_harness.cpp) ----
_CONV_SIG = [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
_POOL_SIG = [C.c_void_p] + [C.c_int] * 5 + [C.c_void_p] * 4


def _load_sym(path: str, symbol: str, sig):
    if not os.path.exists(path):
        return None
    fn = getattr(C.CDLL(path), symbol, None)
    if fn is None:
        return None
    fn.argtypes = sig
    fn.restype = C.c_int
    return fn


def conv_oracle():
    path = os.path.join(HERE, "_build", "libwake_oracle.so")
    if not os.path.exists(path):
        build()
    return _load_sym(path, "wake_conv2d_oracle", _CONV_SIG), _load_sym(path, "wake_maxpool_oracle", _POOL_SIG)


def conv_reference():
    """(conv, maxpool) entry points of the reference's own Convolution / MaxPool classes, or (None, None)."""
    path = os.path.join(HERE, "_ref", "libwake_ref.so")
    return _load_sym(path, "wake_conv2d_ref", _CONV_SIG), _load_sym(path, "wake_maxpool_ref", _POOL_SIG)


def run_conv(fn, kernel, bias, x, dout=None, lr=0.1):
    """Convolution forward (+ backward and SGD update when ``dout`` is given) on COPIES of kernel / bias."""
    kernel = np.array(kernel, dtype=np.float64, order="C", copy=True)
    bias = np.array([float(np.asarray(bias).reshape(-1)[0])], dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    H, W = x.shape
    kh, kw = kernel.shape
    y = np.zeros((H - kh + 1, W - kw + 1))
    dx = np.zeros((H, W)) if dout is not None else None
    d = np.ascontiguousarray(dout, dtype=np.float64) if dout is not None else None
    p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    rc = fn(p(kernel), p(bias), p(x), H, W, kh, kw, p(d), float(lr), p(y), p(dx))
    assert rc == 0
    return dict(y=y, dx=dx, kernel=kernel, bias=bias)


def run_maxpool(fn, x, pool_w, pool_h, stride=1, dout=None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    H, W = x.shape
    OH, OW = (H - pool_h) // stride + 1, (W - pool_w) // stride + 1
    y = np.zeros((OH, OW))
    arg = np.zeros((OH, OW, 2), dtype=np.int32)
    dx = np.zeros((H, W)) if dout is not None else None
    d = np.ascontiguousarray(dout, dtype=np.float64) if dout is not None else None
    p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    rc = fn(p(x), H, W, pool_w, pool_h, stride, p(d), p(y), p(arg), p(dx))
    assert rc == 0
    return dict(y=y, argmax=arg, dx=dx)
