#!/usr/bin/env python3
"""Coefficients of the branch-free FP32 sincos of models.cuh (sincos_r): reduction by multiples of pi
(a = j*pi + r, |r| <= pi/2, sin a = (-1)^j sin r, cos a = (-1)^j cos r — no swap of the two polynomials), odd degree-9
sine and even degree-10 cosine on [-pi/2, pi/2].  Weighted least squares on Chebyshev nodes, coefficients rounded to
FP32, error measured with the kernel's own FP32 FMA sequence (emulated: products of floats are exact in f64).

    python tools/fit_sincos.py        # prints the coefficients and the max abs error against f64 sin/cos
"""
import numpy as np

f32 = np.float32


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def cheb_nodes(n, lo, hi):
    k = np.arange(n)
    return 0.5 * (lo + hi) + 0.5 * (hi - lo) * np.cos(np.pi * (k + 0.5) / n)


def fit():
    h = np.pi / 2 * 1.0005
    x = cheb_nodes(4000, 0.0, h)
    x2 = x * x
    # sin r = r + r^3 * (s0 + s1 r^2 + s2 r^4 + s3 r^6)
    A = np.stack([x2 ** i for i in range(4)], axis=1) * (x ** 3)[:, None]
    s = np.linalg.lstsq(A, np.sin(x) - x, rcond=None)[0]
    # cos r = 1 + r^2 * (c0 + c1 r^2 + ... + c4 r^8) with c0 free (not pinned to -1/2: better minimax)
    B = np.stack([x2 ** i for i in range(5)], axis=1) * x2[:, None]
    c = np.linalg.lstsq(B, np.cos(x) - 1.0, rcond=None)[0]
    return s.astype(f32), c.astype(f32)


def sincos_f32(a, s, c):
    """The kernel's instruction sequence in FP32."""
    a = a.astype(f32)
    magic = f32(12582912.0)
    j = fma(a, np.full_like(a, f32(0.31830988618379067154)), np.full_like(a, magic))
    q = j.view(np.uint32)
    sgn = (q << np.uint32(31)).astype(np.uint32)
    j = (j - magic).astype(f32)
    r = fma(j, np.full_like(a, f32(-3.1415925025939941406)), a)
    r = fma(j, np.full_like(a, f32(-1.5099578831723193e-07)), r)
    r2 = (r * r).astype(f32)
    ps = fma(np.full_like(a, s[3]), r2, np.full_like(a, s[2]))
    ps = fma(ps, r2, np.full_like(a, s[1]))
    ps = fma(ps, r2, np.full_like(a, s[0]))
    sp = fma(ps, (r2 * r).astype(f32), r)
    pc = fma(np.full_like(a, c[4]), r2, np.full_like(a, c[3]))
    pc = fma(pc, r2, np.full_like(a, c[2]))
    pc = fma(pc, r2, np.full_like(a, c[1]))
    pc = fma(pc, r2, np.full_like(a, c[0]))
    cp = fma(pc, r2, np.full_like(a, f32(1.0)))
    so = (sp.view(np.uint32) ^ sgn).view(f32)
    co = (cp.view(np.uint32) ^ sgn).view(f32)
    return so, co


def main():
    s, c = fit()
    print("sin:", ", ".join(f"{v:.10e}f" for v in s))
    print("cos:", ", ".join(f"{v:.10e}f" for v in c))
    hi, lo = f32(3.1415925025939941406), np.float64(np.pi) - np.float64(f32(3.1415925025939941406))
    print(f"pi split: hi {hi:.19f} lo {lo:.19e} (as f32 {f32(lo):.19e})")
    rng = np.random.default_rng(0)
    for span in (0.5 * np.pi, 4.0, 100.0, 1.0e4, 1.0e5):
        a = np.concatenate([rng.uniform(-span, span, 2_000_000), np.linspace(-span, span, 200_001)]).astype(f32)
        so, co = sincos_f32(a, s, c)
        es = np.abs(so.astype(np.float64) - np.sin(a.astype(np.float64))).max()
        ec = np.abs(co.astype(np.float64) - np.cos(a.astype(np.float64))).max()
        print(f"|a| <= {span:9.1f}: max abs err sin {es:.3e} cos {ec:.3e}")


if __name__ == "__main__":
    main()
