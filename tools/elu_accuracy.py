#!/usr/bin/env python
"""Accuracy of the branch-free expm1 used by every ELU in libsurfnet_b200 (csrc/common.cuh: expm1_nonpos), emulated in
numpy float32 with fused multiply-adds, against float64 expm1.  Prints the maximum error in ulps over [-100, 0]."""
import numpy as np

f, d = np.float32, np.float64


def fma(a, b, c):
    return (np.asarray(a, d) * np.asarray(b, d) + np.asarray(c, d)).astype(f)


def expm1_nonpos(x):
    xn = np.maximum(x, f(-30))
    magic = f(12582912.0)
    t = fma(xn, f(1.4426950408889634), magic)
    n = (t - magic).astype(f)
    r = fma(n, f(-0.693145751953125), xn)
    r = fma(n, f(-1.428606765330187045e-06), r)
    p = np.full_like(x, f(1.0 / 5040))
    for c in (720.0, 120.0, 24.0, 6.0, 2.0):
        p = fma(p, r, f(1.0 / c))
    pm1 = fma(p, (r * r).astype(f), r)
    s = np.exp2(n.astype(d)).astype(f)
    return fma(s, pm1, (s - f(1)).astype(f))


def main():
    x = np.concatenate([-np.logspace(-8, 2, 1000001), -np.linspace(0, 3, 2000001), -np.linspace(0, 30, 1000001)]).astype(f)
    ref = np.expm1(x.astype(d))
    got = expm1_nonpos(x).astype(d)
    with np.errstate(all="ignore"):
        err = np.abs(got - ref) / np.spacing(np.abs(ref).astype(f)).astype(d)
    m = np.isfinite(err) & (ref != 0)
    print("max error %.2f ulp, mean %.3f ulp over %d points" % (err[m].max(), err[m].mean(), int(m.sum())))
    return float(err[m].max())


if __name__ == "__main__":
    main()
