"""Reproduces the coefficients of gelu_erf() in vista_slam_b200/csrc/common.cuh.

gelu(x) = max(x, 0) - |x| * Phi(-|x|) with Phi(-a) = 2^q(a); q = degree-6 polynomial fit of log2(0.5 * erfc(a / sqrt 2))
on [0, 6], re-weighted towards the minimax of the absolute error of a * 2^q.  Prints the coefficients (lowest order
first) and the max |error| of the fp32 evaluation against the erf form on [-9, 9]."""
import numpy as np
from scipy.special import erf, erfc

A, DEG = 6.0, 6
a = np.cos(np.linspace(0, np.pi, 4001)) * A / 2 + A / 2
tgt = np.log2(0.5 * erfc(a / np.sqrt(2)))
w = np.ones_like(a)
for _ in range(30):
    coef = np.polynomial.polynomial.polyfit(a, tgt, DEG, w=w)
    err = np.abs(a * (2.0 ** np.polynomial.polynomial.polyval(a, coef) - 2.0 ** tgt))
    w = w * (1 + err / err.max()) ** 0.5
    w /= w.mean()
c32 = coef.astype(np.float32)
x = np.linspace(-9, 9, 400001)
ax = np.minimum(np.abs(x), A).astype(np.float32)
q = np.full_like(ax, c32[-1])
for k in range(DEG - 1, -1, -1):
    q = q * ax + c32[k]
out = np.maximum(x, 0).astype(np.float32) - np.abs(x).astype(np.float32) * np.exp2(q)
ref = 0.5 * x * (1 + erf(x / np.sqrt(2)))
print("coefficients:", [float(v) for v in c32])
print("max |gelu_fit - gelu_erf| on [-9, 9]: %.3e" % np.abs(out - ref).max())
