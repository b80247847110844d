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


# ------------------------------------------------------------------------------------------ convolution / max-pool front
CONV_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "wake_conv_ref.npz")


def _conv_cases():
    z = np.load(CONV_GOLDEN)
    for ci in range(int(z["n_conv"])):
        pre = f"conv{ci}_"
        yield ({k: z[pre + "in_" + k] for k in ("x", "kernel", "bias", "dout")},
               {k: z[pre + "out_" + k] for k in ("y", "dx", "kernel", "bias")})


def _pool_cases():
    z = np.load(CONV_GOLDEN)
    for ci in range(int(z["n_pool"])):
        pre = f"pool{ci}_"
        yield (tuple(int(v) for v in z[pre + "cfg"]), z[pre + "in_x"], z[pre + "in_dout"],
               {k: z[pre + "out_" + k] for k in ("y", "argmax", "dx")})


def test_conv_oracle_matches_reference_golden_bit_exact():
    """oracle/wake_conv_oracle.c against vectors the reference's own Convolution / MaxPool classes produced
    (tests/golden/make_wake_conv_golden.py): every double and every argmax identical."""
    conv, pool = wake_oracle.conv_oracle()
    for inp, want in _conv_cases():
        got = wake_oracle.run_conv(conv, inp["kernel"], inp["bias"], inp["x"], inp["dout"], lr=0.1)
        for k in want:
            assert np.array_equal(got[k], want[k]), k
    for (pw, ph, s), x, d, want in _pool_cases():
        got = wake_oracle.run_maxpool(pool, x, pw, ph, s, d)
        for k in want:
            assert np.array_equal(got[k], want[k]), k


def test_conv_oracle_matches_compiled_reference_on_random_problems():
    rconv, rpool = wake_oracle.conv_reference()
    if rconv is None or rpool is None:
        pytest.skip("oracle/_ref/libwake_ref.so not built here (needs /root/reference); golden vectors cover it")
    conv, pool = wake_oracle.conv_oracle()
    rng = np.random.default_rng(5)
    for _ in range(12):
        H, kh = int(rng.integers(1, 6)), 1
        kh = int(rng.integers(1, H + 1))
        W = int(rng.integers(8, 90))
        kw = int(rng.integers(1, W + 1))
        x, k = rng.normal(0, 1, (H, W)), rng.normal(0, 0.3, (kh, kw))
        d = rng.normal(0, 1, (H - kh + 1, W - kw + 1))
        a, b = wake_oracle.run_conv(conv, k, 0.01, x, d, 0.05), wake_oracle.run_conv(rconv, k, 0.01, x, d, 0.05)
        assert all(np.array_equal(a[n], b[n]) for n in ("y", "dx", "kernel", "bias"))
        ph, pw, s = int(rng.integers(1, H + 1)), int(rng.integers(1, 5)), int(rng.integers(1, 4))
        xq = np.round(x, 1)
        dd = rng.normal(0, 1, ((H - ph) // s + 1, (W - pw) // s + 1))
        a, b = wake_oracle.run_maxpool(pool, xq, pw, ph, s, dd), wake_oracle.run_maxpool(rpool, xq, pw, ph, s, dd)
        assert all(np.array_equal(a[n], b[n]) for n in ("y", "argmax", "dx"))


@pytest.mark.gpu
def test_cuda_conv_front_is_bit_exact_against_the_reference_golden():
    from imagined_speech_translation_b200.wake import Convolution, MaxPool
    for inp, want in _conv_cases():
        H, W = inp["x"].shape
        kh, kw = inp["kernel"].shape
        conv = Convolution(W, H, kw, kh, "relu").load(inp["kernel"], inp["bias"])
        y = conv.forward(torch.from_numpy(inp["x"]).cuda())
        dx = conv.backward(torch.from_numpy(inp["dout"]).cuda(), 0.1)
        assert np.array_equal(y.cpu().numpy(), want["y"])
        assert np.array_equal(dx.cpu().numpy(), want["dx"])
        assert np.array_equal(conv.kernel.cpu().numpy(), want["kernel"])
        assert np.array_equal(conv.biases.cpu().numpy(), want["bias"])
    for (pw, ph, s), x, d, want in _pool_cases():
        H, W = x.shape
        pool = MaxPool(W, H, pw, ph, s)
        y = pool.forward(torch.from_numpy(x).cuda())
        dx = pool.backward(torch.from_numpy(d).cuda())
        assert np.array_equal(y.cpu().numpy(), want["y"])
        assert np.array_equal(pool.max_indices.cpu().numpy(), want["argmax"])
        assert np.array_equal(dx.cpu().numpy(), want["dx"])


@pytest.mark.gpu
def test_cuda_conv_front_at_the_train_cpp_shapes_matches_the_oracle():
    """wake_model/train.cpp:26-33 on a 2 x 4096 input: conv 32x1 -> pool 2x1 -> conv 64x1 -> pool -> conv 128x1 -> pool,
    forward and backward through the chain, every tensor bit-identical to the C oracle run on the same data."""
    from imagined_speech_translation_b200.wake import Convolution, MaxPool
    oconv, opool = wake_oracle.conv_oracle()
    rng = np.random.default_rng(11)
    H, W = 2, 4096
    x = rng.normal(0, 1, (H, W))
    layers, cur_w = [], W
    for kw in (32, 64, 128):
        c = Convolution(cur_w, H, kw, 1, "relu", generator=torch.Generator().manual_seed(kw))
        p = MaxPool(c.output_width, c.output_height, 2, 1)
        layers += [c, p]
        cur_w = p.output_width
    acts, h = [], torch.from_numpy(x).cuda()
    ref_acts, hr = [], x
    params = [(l.kernel.cpu().numpy().copy(), l.biases.cpu().numpy().copy()) if isinstance(l, Convolution) else None
              for l in layers]
    for l, prm in zip(layers, params):
        h = l.forward(h)
        acts.append(h)
        if prm is not None:
            hr_out = wake_oracle.run_conv(oconv, prm[0], prm[1], hr)["y"]
        else:
            hr_out = wake_oracle.run_maxpool(opool, hr, 2, 1, 1)["y"]
        ref_acts.append((hr, hr_out))
        hr = hr_out
        assert np.array_equal(h.cpu().numpy(), hr)
    g = rng.normal(0, 0.1, hr.shape)
    gd = torch.from_numpy(g).cuda()
    for l, prm, (xin, _) in zip(reversed(layers), reversed(params), reversed(ref_acts)):
        if prm is not None:
            gd = l.backward(gd, 0.1)
            r = wake_oracle.run_conv(oconv, prm[0], prm[1], xin, g, 0.1)
            assert np.array_equal(l.kernel.cpu().numpy(), r["kernel"]) and np.array_equal(l.biases.cpu().numpy(), r["bias"])
        else:
            gd = l.backward(gd)
            r = wake_oracle.run_maxpool(opool, xin, 2, 1, 1, g)
        g = r["dx"]
        assert np.array_equal(gd.cpu().numpy(), g)


@pytest.mark.gpu
def test_cuda_conv_front_rejects_bad_arguments():
    from imagined_speech_translation_b200 import _lib
    from imagined_speech_translation_b200.wake import Convolution, MaxPool
    with pytest.raises(ValueError):
        Convolution(8, 2, 9, 1)
    with pytest.raises(ValueError):
        MaxPool(8, 2, 2, 3)
    c = Convolution(16, 2, 4, 1)
    with pytest.raises(_lib.EegxError):
        c.forward(torch.zeros(2, 16, dtype=torch.float64))               # CPU tensor: no fallback
    with pytest.raises(_lib.EegxError):
        c.backward(torch.zeros(2, 13, dtype=torch.float64, device="cuda"), 0.1)   # before forward
    with pytest.raises(ValueError):
        c.forward(torch.zeros(2, 15, dtype=torch.float64, device="cuda"))
