#!/usr/bin/env python
"""Generates the polynomial coefficients used by mosmicrosim_b200/csrc/mms_fastmath.cuh
(run once; the output is pasted into the header).  Chebyshev-node interpolation in
50-digit arithmetic, coefficients rounded to double, error re-measured with the
rounded coefficients."""
import mpmath as mp

mp.mp.dps = 60


def cheb_fit(f, a, b, deg):
    n = deg + 1
    xs = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n)
    y = mp.matrix(n, 1)
    for i, x in enumerate(xs):
        for j in range(n):
            A[i, j] = x ** j
        y[i] = f(x)
    c = mp.lu_solve(A, y)
    return [float(ci) for ci in c]


def max_err(f, approx, a, b, rel=True, n=4000):
    worst = mp.mpf(0)
    for k in range(n + 1):
        x = a + (b - a) * k / n
        t = f(x)
        e = abs(approx(x) - t)
        if rel and t != 0:
            e /= abs(t)
        worst = max(worst, e)
    return worst


def horner(c, x):
    s = mp.mpf(0)
    for ci in reversed(c):
        s = s * x + mp.mpf(ci)
    return s


# ---- exp(r) = 1 + r + r^2 q(r) on |r| <= ln2/2 (slightly widened)
half = mp.log(2) / 2 * mp.mpf("1.0001")
for deg in (8, 9, 10):
    q = lambda r: (mp.exp(r) - 1 - r) / (r * r) if r != 0 else mp.mpf(1) / 2
    c = cheb_fit(q, -half, half, deg)
    e = max_err(mp.exp, lambda r: 1 + r + r * r * horner(c, r), -half, half)
    print(f"// exp: q degree {deg} (exp degree {deg + 2}) max rel err {mp.nstr(e, 3)}")
    if deg == 9:
        cexp = c
print("static const double kExpQ[%d] = {" % len(cexp))
for v in cexp:
    print(f"    {v!r},")
print("};")

# ---- log(m) = 2 f + f s Q(s), f = (m-1)/(m+1), s = f^2, m in [sqrt(1/2), sqrt(2)]
fmax = (mp.sqrt(2) - 1) / (mp.sqrt(2) + 1)
smax = fmax * fmax * mp.mpf("1.001")


def Q(s):
    if s == 0:
        return mp.mpf(2) / 3
    f = mp.sqrt(s)
    return (2 * mp.atanh(f) - 2 * f) / (f * s)


for deg in (6, 7, 8):
    c = cheb_fit(Q, mp.mpf(0), smax, deg)

    def approx(s, c=c):
        f = mp.sqrt(s)
        return 2 * f + f * s * horner(c, s)

    e = max_err(lambda s: 2 * mp.atanh(mp.sqrt(s)) if s else mp.mpf(0), approx, mp.mpf(0), smax)
    print(f"// log: Q degree {deg} max rel err of log(m) {mp.nstr(e, 3)}")
    if deg == 6:
        clog = c
print("static const double kLogQ[%d] = {" % len(clog))
for v in clog:
    print(f"    {v!r},")
print("};")
ln2 = mp.log(2)
hi = float(ln2)
# split ln2 so that n * hi is exact for |n| < 2^11: clear the low 11+ bits... use 32-bit split
import struct
bits = struct.unpack("<Q", struct.pack("<d", hi))[0] & ~((1 << 21) - 1)
hi_s = struct.unpack("<d", struct.pack("<Q", bits))[0]
lo_s = float(ln2 - mp.mpf(hi_s))
print(f"// ln2 split: hi {hi_s!r} lo {lo_s!r}  (hi has 21 trailing zero bits)")
print(f"// log2(e) {float(1 / ln2)!r}")
