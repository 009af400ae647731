"""Fit the polynomial of the fast exact-class GELU used by the tensor-core epilogues:
   erfc(t) ~= exp2(t * (c1 + c2 t + ... + cn t^(n-1))),  t = |v| / sqrt(2) clamped to T
   gelu(v) = v < 0 ? h : v - h,  h = 0.5 * v * erfc(t)            (no cancellation for v < 0)
Prints coefficients and the fp32-evaluated max abs / rel error against the exact erf GELU."""
import numpy as np
from scipy.special import erfc, erf

T = 4.0
for deg in (6, 7, 8, 9):
    # Chebyshev nodes on [0, T], weighted least squares on g(t) = log2(erfc(t)) / t
    k = np.arange(4000)
    t = 0.5 * T * (1 - np.cos(np.pi * (k + 0.5) / 4000))
    t = t[t > 1e-6]
    g = np.log2(erfc(t)) / t
    # weight: error in P = t * dg  ->  erfc rel error ln2 * t * dg; GELU abs error ~ 0.5 |v| erfc * that
    w = t * t * erfc(t) + 1e-3 * t
    V = np.vander(t, deg, increasing=True)
    c, *_ = np.linalg.lstsq(V * w[:, None], g * w, rcond=None)
    c32 = c.astype(np.float32)
    v = np.linspace(-9, 9, 2000001).astype(np.float32)
    tt = np.minimum(np.abs(v) * np.float32(0.70710678118654752), np.float32(T)).astype(np.float32)
    p = np.zeros_like(tt)
    for ci in c32[::-1]:
        p = (p * tt + ci).astype(np.float32)
    p = (p * tt).astype(np.float32)
    e = np.exp2(p.astype(np.float64)).astype(np.float32)
    h = (np.float32(0.5) * v * e).astype(np.float32)
    gl = np.where(v < 0, h, (v - h).astype(np.float32))
    exact = 0.5 * v.astype(np.float64) * (1 + erf(v.astype(np.float64) / np.sqrt(2)))
    abs_err = np.abs(gl - exact)
    rel = abs_err / np.maximum(np.abs(exact), 1e-6)
    print("deg %d: max abs %.3e (at v=%.3f)  max rel(|exact|>1e-6) %.3e" % (deg, abs_err.max(), v[abs_err.argmax()], rel[np.abs(exact) > 1e-6].max()))
    print("   coeffs:", ", ".join("%.9ef" % x for x in c32))
