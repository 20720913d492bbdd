"""Rounding error of the FP16 GCC-PHAT lag transform (csrc/gcc_tc.cu::gcc_ph16_kernel), emulated with numpy on the CPU:
half2 unit phasors, packed-half cross spectra (HMUL2 + HFMA2: two roundings per component), FP16 twiddle table, FP32
accumulation -- against the float64 evaluation of the same sum.  The GPU gate (1e-3 absolute) is tests/test_gpu_features.py."""
import numpy as np

rng = np.random.default_rng(0)
N, K, F = 1200, 601, 64
pairs = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
k = np.arange(K)
X = rng.standard_normal((F, K, 4)) + 1j * rng.standard_normal((F, K, 4))
X[:, :, 1] = X[:, :, 0] * np.exp(-2j * np.pi * k * 7 / N) + 0.05 * (rng.standard_normal((F, K)) + 1j * rng.standard_normal((F, K)))
u = X / np.abs(X)
ang = 2 * np.pi * np.outer(k, np.arange(-32, 32)) / N
ck = np.ones(K); ck[0] = ck[600] = 0.5
C, S = ck[:, None] * np.cos(ang), -ck[:, None] * np.sin(ang)
h = lambda a: a.astype(np.float16)
ur, ui = h(u.real), h(u.imag)
C16, S16 = h(C).astype(np.float32), h(S).astype(np.float32)
e_all = e_ph = 0.0
for m, n in pairs:
    R = np.conj(u[:, :, m]) * u[:, :, n]
    ref = (2 / N) * (R.real @ C + R.imag @ S)
    t = h(ui[:, :, m].astype(np.float32) * ui[:, :, n].astype(np.float32))
    re = h(ur[:, :, m].astype(np.float64) * ur[:, :, n].astype(np.float64) + t.astype(np.float64))
    t2 = h(ui[:, :, m].astype(np.float32) * ur[:, :, n].astype(np.float32))
    im = h(ur[:, :, m].astype(np.float64) * ui[:, :, n].astype(np.float64) - t2.astype(np.float64))
    e_all = max(e_all, np.abs((2 / N) * (re.astype(np.float32) @ C16 + im.astype(np.float32) @ S16) - ref).max())
    Rp = (ur[:, :, m].astype(np.float64) - 1j * ui[:, :, m]) * (ur[:, :, n].astype(np.float64) + 1j * ui[:, :, n])
    e_ph = max(e_ph, np.abs((2 / N) * (Rp.real @ C + Rp.imag @ S) - ref).max())
print("max abs error: FP16 pipeline %.2e ; half2 phasors alone (exact products, float64 twiddles) %.2e ; gate 1e-3" % (e_all, e_ph))
