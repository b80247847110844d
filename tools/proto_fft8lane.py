"""numpy prototype of the tuned kernel's 8-lane real-FFT-256 decomposition
(index math only; validates step order, twiddles, exchange, pairing, split)."""
import numpy as np
rng = np.random.default_rng(0)
seg = rng.standard_normal(256)           # windowed segment
ref = np.abs(np.fft.rfft(seg)) ** 2

W = lambda N, e: np.exp(-2j * np.pi * e / N)
A = np.zeros((8, 2, 8), complex)         # [lane g][eps][k1]
for g in range(8):
    for eps in range(2):
        b = 2 * g + eps
        z = np.array([seg[32 * a + 4 * g + 2 * eps] + 1j * seg[32 * a + 4 * g + 2 * eps + 1] for a in range(8)])
        for k1 in range(8):
            A[g, eps, k1] = sum(z[a] * W(8, a * k1) for a in range(8)) * W(128, b * k1)
# exchange: lane g' gets B[b] = A[b//2, b%2, g']
X = np.zeros((8, 16), complex)           # [lane g'][k2] = Z[g' + 8 k2]
for gp in range(8):
    B = np.array([A[b // 2, b % 2, gp] for b in range(16)])
    for k2 in range(16):
        X[gp, k2] = sum(B[b] * W(16, b * k2) for b in range(16))
# check complex FFT
z_all = seg[0::2] + 1j * seg[1::2]
Zref = np.fft.fft(z_all)
assert np.allclose([[X[gp, k2] for k2 in range(16)] for gp in range(8)],
                   [[Zref[gp + 8 * k2] for k2 in range(16)] for gp in range(8)])
# pairing + split
P = np.full(129, np.nan)
for gp in range(8):
    partner = (8 - gp) & 7
    own = X[gp]
    src = X[partner]
    send = {j: (src[(j + 1) & 15] if partner == 0 else src[j]) for j in range(8, 16)}
    for k2 in range(8):
        k = gp + 8 * k2
        Zk, Zm = own[k2], send[15 - k2]
        E = Zk + np.conj(Zm); D = Zk - np.conj(Zm)
        O = complex(D.imag, -D.real)
        T = W(256, k) * O
        P[k] = abs(E + T) ** 2 / 4
        P[128 - k] = abs(E - T) ** 2 / 4
    if gp == 0:
        P[64] = abs(own[8]) ** 2
assert not np.isnan(P).any()
print("max rel err", np.max(np.abs(P - ref) / ref.max()))
assert np.allclose(P, ref, rtol=1e-9, atol=1e-9)
print("ok")
