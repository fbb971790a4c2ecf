#!/usr/bin/env python
"""Constants of the C-layer ziggurat for the standard normal (Marsaglia & Tsang 2000; Doornik's ZIGNOR layout): the rightmost
edge R and the common layer area V (unnormalised density f(x) = exp(-x^2/2)).

    V(R) = R f(R) + int_R^inf f,   x_1 = R,   x_{i+1} = f^-1(V / x_i + f(x_i)),   closing condition  V / x_{C-1} + f(x_{C-1}) = 1

Solved by bisection in 60-digit arithmetic (mpmath); reproduces the published pairs for C = 128 (R = 3.442619855899,
V = 9.91256303526217e-3) and C = 256 (R = 3.6541528853610088, V = 4.92867323399e-3).

    python tools/zig_constants.py 128 256 1024 4096
"""
import sys
from mpmath import mp, mpf, exp, sqrt, log, erfc, pi

mp.dps = 60


def f(x):
    return exp(-x * x / 2)


def residual(R, C):
    V = R * f(R) + sqrt(pi / 2) * erfc(R / sqrt(mpf(2)))
    x = R
    for _ in range(2, C):
        t = V / x + f(x)
        if t >= 1:
            return mpf(1), V  # ran past the mode: R too small
        x = sqrt(-2 * log(t))
    return V / x + f(x) - 1, V


def solve(C):
    lo, hi = mpf(2), mpf(8)
    for _ in range(200):
        mid = (lo + hi) / 2
        r, _ = residual(mid, C)
        if r > 0:
            lo = mid
        else:
            hi = mid
    R = (lo + hi) / 2
    return R, residual(R, C)[1]


if __name__ == "__main__":
    for C in [int(a) for a in sys.argv[1:]] or [128, 256, 1024, 4096]:
        R, V = solve(C)
        print(f"C = {C:5d}  R = {float(R)!r}  V = {float(V)!r}   (R = {mp.nstr(R, 25)}, V = {mp.nstr(V, 25)})")
