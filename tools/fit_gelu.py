"""Reproduces the degree-5 coefficients of gelu_erf() in effocr_b200/csrc/gemm_sm100.cuh:
weighted near-minimax fit of P(u) = log2(0.5 erfc(u / sqrt 2)) on [0, 6]; gelu(x) = relu(x) - |x| 2^P(min(|x|, 6))."""
import numpy as np
from numpy.polynomial import polynomial as Pn
from scipy.special import erf, erfc

U, DEG = 6.0, 5
phi_neg = lambda u: 0.5 * erfc(u / np.sqrt(2))
u = np.cos(np.pi * (np.arange(4000) + 0.5) / 4000) * U / 2 + U / 2
target = np.log2(phi_neg(u))
wgt = u * phi_neg(u) * np.log(2) + 1e-9
w = wgt.copy()
for _ in range(60):
    c = Pn.polyfit(u, target, DEG, w=w)
    err = np.abs((Pn.polyval(u, c) - target) * wgt)
    w = w * (1 + 2 * err / err.max())
c32 = c.astype(np.float32)
x = np.linspace(-10, 10, 400001).astype(np.float32)
a = np.abs(x)
p = np.full_like(a, c32[-1])
for k in range(DEG - 1, -1, -1):
    p = (p * np.minimum(a, np.float32(U)) + c32[k]).astype(np.float32)
g = np.maximum(x, 0) - a * np.exp2(p).astype(np.float32)
exact = x.astype(np.float64) * 0.5 * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
print("coefficients (low -> high):", [float(v) for v in c32])
print("max |gelu - exact| =", np.abs(g - exact).max())
