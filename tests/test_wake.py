"""wake_model dense head (BASELINE config 5): the C oracle against the reference-generated golden vectors and the
compiled reference itself (CPU), and the CUDA path against the oracle (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import wake_oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "wake_dense_ref.npz")
KEYS = ("w1", "b1", "w2", "b2", "loss", "probs", "dx")


def _cases():
    g = np.load(GOLDEN)
    for ci in range(int(g["n_cases"])):
        pre = f"c{ci}_"
        yield str(g[pre + "act"]), {k[len(pre):]: g[k] for k in g.files if k.startswith(pre)}


def _problem(rng, n_in, hidden, ncls, n, scale=1.0):
    w1 = rng.normal(0, np.sqrt(2.0 / n_in), (hidden, n_in))
    b1 = rng.normal(0, np.sqrt(2.0 / n_in), hidden)
    w2 = rng.normal(0, np.sqrt(2.0 / hidden), (ncls, hidden))
    b2 = rng.normal(0, np.sqrt(2.0 / hidden), ncls)
    x = rng.normal(0, scale, (n, n_in))
    label = rng.integers(0, ncls, n).astype(np.int32)
    return w1, b1, w2, b2, x, label


# ------------------------------------------------------------------------------------------------- CPU
def test_oracle_matches_reference_golden_bit_exact():
    fn = wake_oracle.oracle()
    n = 0
    for act, c in _cases():
        r = wake_oracle.run(fn, c["in_w1"], c["in_b1"], c["in_w2"], c["in_b2"], c["in_x"], c["in_label"], lr=0.1,
                            activation=act, train=True, want_dx=True)
        for k in KEYS:
            assert np.array_equal(r[k], c["out_" + k]), (act, k)
        f = wake_oracle.run(fn, c["in_w1"], c["in_b1"], c["in_w2"], c["in_b2"], c["in_x"], c["in_label"],
                            activation=act, train=False)
        assert np.array_equal(f["probs"], c["fwd_probs"]) and np.array_equal(f["loss"], c["fwd_loss"])
        assert np.array_equal(f["w1"], c["in_w1"])                      # forward-only leaves the parameters alone
        n += 1
    assert n == 5


def test_oracle_matches_compiled_reference_on_random_problems():
    ref = wake_oracle.reference()
    if ref is None:
        pytest.skip("oracle/_ref/libwake_ref.so not built here (needs /root/reference); golden vectors cover it")
    fn = wake_oracle.oracle()
    rng = np.random.default_rng(7)
    for act in ("relu", "sigmoid", "tanh", ""):
        for shape in ((3, 4, 2, 5), (64, 50, 9, 7), (1, 1, 1, 3)):
            prob = _problem(rng, *shape)
            a = wake_oracle.run(fn, *prob, lr=0.1, activation=act, want_dx=True)
            b = wake_oracle.run(ref, *prob, lr=0.1, activation=act, want_dx=True)
            for k in KEYS:
                assert np.array_equal(a[k], b[k]), (act, shape, k)


def test_oracle_learns():
    """Sanity: repeating one separable sample drives its loss down (train.cpp's loop does the same per epoch)."""
    rng = np.random.default_rng(3)
    w1, b1, w2, b2, x, label = _problem(rng, 16, 32, 4, 1)
    x = np.repeat(x, 30, 0)
    label = np.repeat(label, 30)
    r = wake_oracle.run(wake_oracle.oracle(), w1, b1, w2, b2, x, label, lr=0.05)
    assert r["loss"][-1] < 0.1 * r["loss"][0]


# ------------------------------------------------------------------------------------------------- GPU
def _gpu_run(prob, act, lr=0.1, train=True, want_dx=True):
    from imagined_speech_translation_b200.wake import DenseHead
    w1, b1, w2, b2, x, label = prob
    head = DenseHead(w1.shape[1], w1.shape[0], w2.shape[0], activation=act).load(w1, b1, w2, b2)
    xt, lt = torch.from_numpy(x).cuda(), torch.from_numpy(label).cuda()
    if train:
        loss, probs, dx = head.train_samples(xt, lt, lr, want_dx=True)
    else:
        probs, loss = head.forward(xt, lt)
        dx = None
    torch.cuda.synchronize()
    return dict(w1=head.w1.cpu().numpy(), b1=head.b1.cpu().numpy(), w2=head.w2.cpu().numpy(), b2=head.b2.cpu().numpy(),
                loss=loss.cpu().numpy(), probs=probs.cpu().numpy(), dx=None if dx is None else dx.cpu().numpy())


def _close(a, b, tol):
    scale = max(float(np.abs(b).max()), 1e-300)
    return float(np.abs(a - b).max()) / scale <= tol


@pytest.mark.gpu
def test_cuda_matches_reference_golden():
    """Tolerance 1e-10 (inf-norm relative per tensor): the SGD update is rounded as the C++ does, only the dot
    products are summed in a different order."""
    for act, c in _cases():
        prob = (c["in_w1"], c["in_b1"], c["in_w2"], c["in_b2"], c["in_x"], c["in_label"])
        r = _gpu_run(prob, act)
        for k in KEYS:
            assert _close(r[k], c["out_" + k], 1e-10), (act, k)
        f = _gpu_run(prob, act, train=False)
        assert _close(f["probs"], c["fwd_probs"], 1e-12) and _close(f["loss"], c["fwd_loss"], 1e-12)
        assert np.array_equal(f["w1"], c["in_w1"])


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(4096, 1024, 100, 6), (1000, 1024, 64, 5), (33, 7, 3, 9), (1, 1, 1, 4),
                                   (130, 1500, 11, 4), (257, 148, 2, 3)])
def test_cuda_matches_oracle(shape):
    rng = np.random.default_rng(shape[0] + shape[1])
    prob = _problem(rng, *shape, scale=0.3)
    a = _gpu_run(prob, "relu", lr=0.01)
    b = wake_oracle.run(wake_oracle.oracle(), *prob, lr=0.01, activation="relu", want_dx=True)
    for k in KEYS:
        assert _close(a[k], b[k], 1e-10), (shape, k)


@pytest.mark.gpu
def test_cuda_single_update_is_bit_exact_in_the_rounding_of_the_step():
    """With ONE sample and a hidden layer whose dot products are exact (small integers), every quantity the update
    is built from is exact, so the mul-mul-sub rounding of the step itself is compared bit for bit."""
    rng = np.random.default_rng(11)
    n_in, hidden, ncls = 16, 12, 3
    w1 = rng.integers(-3, 4, (hidden, n_in)).astype(np.float64)
    b1 = rng.integers(-2, 3, hidden).astype(np.float64)
    w2 = np.zeros((ncls, hidden)); b2 = np.zeros(ncls)              # logits 0 -> p = 1/3 exactly representable? no, but equal in both
    x = rng.integers(-2, 3, (1, n_in)).astype(np.float64)
    label = np.array([1], dtype=np.int32)
    prob = (w1, b1, w2, b2, x, label)
    a = _gpu_run(prob, "relu", lr=0.1)
    b = wake_oracle.run(wake_oracle.oracle(), *prob, lr=0.1, activation="relu", want_dx=True)
    for k in ("w2", "b2", "w1", "b1", "probs"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.gpu
def test_cuda_edge_cases_and_errors():
    from imagined_speech_translation_b200 import _lib
    from imagined_speech_translation_b200.wake import DenseHead
    head = DenseHead(8, 4, 3)
    w_before = head.w1.clone()
    loss, probs = head.train_samples(torch.zeros(0, 8, dtype=torch.float64, device="cuda"),
                                     torch.zeros(0, dtype=torch.int32, device="cuda"))
    assert loss.numel() == 0 and probs.shape == (0, 3) and torch.equal(head.w1, w_before)      # empty input: no-op
    with pytest.raises(_lib.EegxError):
        head.train_samples(torch.zeros(2, 8, device="cuda"), torch.zeros(2, dtype=torch.int32, device="cuda"))   # fp32
    with pytest.raises(_lib.EegxError):
        head.train_samples(torch.zeros(2, 9, dtype=torch.float64, device="cuda"),
                           torch.zeros(2, dtype=torch.int32, device="cuda"))
    with pytest.raises(ValueError):
        DenseHead(8, 4, 3, activation="gelu")
    big = DenseHead(20000, 4, 3)                                     # 2 * 20000 doubles of x staging > 227 KB
    with pytest.raises(_lib.EegxError):
        big.train_samples(torch.zeros(1, 20000, dtype=torch.float64, device="cuda"),
                          torch.zeros(1, dtype=torch.int32, device="cuda"))


@pytest.mark.gpu
def test_cuda_training_reduces_loss_at_config_shape():
    """Size-independent property at the BASELINE config-5 shape (in = 4096, hidden = 1024): repeated passes over a
    small sample set drive the loss down, and probabilities are a distribution."""
    from imagined_speech_translation_b200.wake import DenseHead
    g = torch.Generator().manual_seed(5)
    head = DenseHead(4096, 1024, 120, generator=g)
    x = (torch.randn(16, 4096, dtype=torch.float64, generator=g) * 0.05).cuda()
    y = torch.randint(0, 120, (16,), generator=g).int().cuda()
    first = None
    for epoch in range(12):
        loss, probs = head.train_samples(x, y, 0.1)
        if first is None:
            first = float(loss.mean())
        assert torch.allclose(probs.sum(1), torch.ones(16, dtype=torch.float64, device="cuda"), atol=1e-12)
    assert float(loss.mean()) < 0.2 * first
