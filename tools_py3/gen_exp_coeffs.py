"""Derives the polynomial used by the kernels' exp(): exp(r) ~ 1 + r + r^2 * q(r) on |r| <= ln2/2, with q of
degree 9 interpolating (exp(r)-1-r)/r^2 at Chebyshev nodes (near-minimax), coefficients rounded to double.
Prints C initialisers and the approximation error.  Run: python tools_py3/gen_exp_coeffs.py"""
import mpmath as mp

mp.mp.dps = 60
DEG = 9
h = mp.log(2) / 2 * mp.mpf("1.0001")


def g(r):
    if r == 0:
        return mp.mpf(1) / 2
    return (mp.e ** r - 1 - r) / (r * r)


nodes = [h * mp.cos(mp.pi * (2 * k + 1) / (2 * (DEG + 1))) for k in range(DEG + 1)]
A = mp.matrix(DEG + 1, DEG + 1)
b = mp.matrix(DEG + 1, 1)
for i, x in enumerate(nodes):
    for j in range(DEG + 1):
        A[i, j] = x ** j
    b[i] = g(x)
c = mp.lu_solve(A, b)
coef = [float(ci) for ci in c]  # q(r) = sum coef[j] r^j


def approx(r):
    q = mp.mpf(0)
    for cj in reversed(coef):
        q = q * r + mp.mpf(cj)
    return 1 + r + r * r * q


worst = 0
for k in range(4001):
    r = -h + 2 * h * k / 4000
    e = abs(approx(r) / mp.e ** r - 1)
    worst = max(worst, e)
print("max relative approximation error: 2^%.2f" % float(mp.log(worst, 2)))
# highest degree first, as the Horner loop consumes them; then the two exact ones
print("static __constant__ double c_exp_poly[%d] = {" % (DEG + 1))
for cj in reversed(coef):
    print("    %s,  // %s" % (float.hex(cj), repr(cj)))
print("};")
print("log2e  =", float.hex(float(mp.log(mp.e, 2))))
ln2 = mp.log(2)
hi = float(ln2)
# ln2_hi with trailing zeros is not needed with fma; use hi = RN(ln2), lo = RN(ln2 - hi)
print("ln2_hi =", float.hex(hi))
print("ln2_lo =", float.hex(float(ln2 - mp.mpf(hi))))
