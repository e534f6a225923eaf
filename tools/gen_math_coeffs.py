#!/usr/bin/env python3
"""Derives the fp32 polynomial coefficients used by include/pt_math.h (weighted least squares on Chebyshev
nodes in float64, then rounded to fp32) and prints them with their worst-case error on a dense grid.
Nothing here runs at build time; the printed constants are pasted into pt_math.h."""
import numpy as np
import mpmath as mp
mp.mp.dps = 50
def hp(f):
    return lambda x: np.array([float(f(mp.mpf(float(v)))) for v in x])
np.set_printoptions(precision=17)

def cheb_nodes(a, b, n):
    k = np.arange(n)
    return 0.5 * (a + b) + 0.5 * (b - a) * np.cos(np.pi * (2 * k + 1) / (2 * n))

def fit(fun, a, b, powers, weight=None, n=4000):
    x = cheb_nodes(a, b, n)
    A = np.stack([x ** p for p in powers], axis=1)
    y = fun(x)
    w = np.ones_like(x) if weight is None else weight(x)
    c, *_ = np.linalg.lstsq(A * w[:, None], y * w, rcond=None)
    return c

def show(name, c):
    c32 = np.asarray(c, dtype=np.float32)
    print(name, ', '.join('%.9ef' % v for v in c32))
    return c32.astype(np.float64)

# sin(r) = r + r^3 * S(r^2), r in [-pi/4, pi/4]; fit S(z) = (sin(r)/r - 1)/z
z = lambda r: r * r
eps = 1e-6
S = fit(hp(lambda r: (mp.sin(r) / r - 1) / (r * r)), eps, np.pi / 4 * 1.001, [0, 2, 4, 6])
S = show('SIN', S)
# cos(r) = 1 - z/2 + z^2 * C(z)
C = fit(hp(lambda r: (mp.cos(r) - 1 + r * r / 2) / r ** 4), eps, np.pi / 4 * 1.001, [0, 2, 4, 6])
C = show('COS', C)
r = np.linspace(-np.pi / 4, np.pi / 4, 200001)
zz = r * r
s = r + r * zz * (S[0] + zz * (S[1] + zz * (S[2] + zz * S[3])))
c = 1 - zz / 2 + zz * zz * (C[0] + zz * (C[1] + zz * (C[2] + zz * C[3])))
print('  sin max rel err', np.max(np.abs(s - np.sin(r)) / np.maximum(np.abs(np.sin(r)), 1e-30)), 'cos max abs err', np.max(np.abs(c - np.cos(r))))

# asin(x) = x + x^3 * A(x^2), |x| <= 0.5
A = fit(hp(lambda x: (mp.asin(x) / x - 1) / (x * x)), eps, 0.5, [0, 2, 4, 6, 8, 10])
A = show('ASIN', A)
x = np.linspace(1e-9, 0.5, 100001); zz = x * x
a = x + x * zz * np.polyval(A[::-1], zz)
print('  asin max rel err', np.max(np.abs(a - np.arcsin(x)) / np.arcsin(x)))

# 2^f = 1 + f * E(f), f in [-0.5, 0.5]
E = fit(hp(lambda f: (mp.power(2, f) - 1) / f), -0.5, 0.5, [0, 1, 2, 3, 4, 5, 6], n=4001)
E = show('EXP2', E)
f = np.linspace(-0.5, 0.5, 100001); f = f[f != 0]
e = 1 + f * np.polyval(E[::-1], f)
print('  exp2 max rel err', np.max(np.abs(e - np.exp2(f)) / np.exp2(f)))

# log2(m) = s * L(s^2), s = (m-1)/(m+1), m in [sqrt(.5), sqrt(2)] -> |s| <= 0.17157288
smax = (np.sqrt(2) - 1) / (np.sqrt(2) + 1)
L = fit(hp(lambda s: mp.log((1 + s) / (1 - s), 2) / s), eps, smax * 1.001, [0, 2, 4, 6, 8])
L = show('LOG2', L)
s = np.linspace(1e-9, smax, 100001); zz = s * s
l = s * np.polyval(L[::-1], zz)
ref = np.log2((1 + s) / (1 - s))
print('  log2 max rel err', np.max(np.abs(l - ref) / ref))

pio2 = np.pi / 2
hi = np.float32(pio2); mid = np.float32(pio2 - np.float64(hi)); lo = np.float32(pio2 - np.float64(hi) - np.float64(mid))
print('PIO2 hi/mid/lo %.9ef %.9ef %.9ef' % (hi, mid, lo), ' 2/pi %.9ef' % np.float32(2 / np.pi))
print('LOG2E %.9ef LN2 %.9ef PI %.9ef PIO2 %.9ef' % (np.float32(np.log2(np.e)), np.float32(np.log(2)), np.float32(np.pi), np.float32(np.pi/2)))
