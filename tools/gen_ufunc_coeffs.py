"""Polynomial coefficients for arraylib_b200/csrc/fast_ufunc.h.

Chebyshev-node interpolation in 60-digit arithmetic (mpmath), coefficients rounded to double, and the error of the
ROUNDED polynomial measured on a dense grid. Prints C initialisers. Run: python tools/gen_ufunc_coeffs.py
"""
import sys

import mpmath as mp

mp.mp.dps = 150


def cheb_fit(g, lo, hi, deg):
    n = deg + 1
    xs = [(lo + hi) / 2 + (hi - lo) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n)
    b = mp.matrix(n, 1)
    for i, x in enumerate(xs):
        for j in range(n):
            A[i, j] = x ** j
        b[i] = g(x)
    c = mp.lu_solve(A, b)
    return [float(c[j]) for j in range(n)]  # rounded to double


def horner(c, x):
    p = mp.mpf(0)
    for a in reversed(c):
        p = p * x + mp.mpf(a)
    return p


def report(name, coeffs, final, exact, lo, hi, pts=4000):
    worst = mp.mpf(0)
    for i in range(pts + 1):
        x = lo + (hi - lo) * mp.mpf(i) / pts
        if x == 0:
            continue
        e = exact(x)
        if e == 0:
            continue
        err = abs(final(coeffs, x) / e - 1)
        worst = max(worst, err)
    print(f"// {name}: degree {len(coeffs) - 1}, max rel error of the rounded polynomial 2^{float(mp.log(worst, 2)):.1f}")
    print("//   { " + ", ".join(repr(c) for c in reversed(coeffs)) + " }   (highest power first)")
    return worst


def main():
    out = {}
    # exp(r) = 1 + r + r^2 q(r), |r| <= ln2/2
    a = mp.mpf("0.34658")
    for deg in (7, 8):
        c = cheb_fit(lambda r: (mp.exp(r) - 1 - r) / r ** 2 if abs(r) > mp.mpf(10) ** -40 else mp.mpf(1) / 2, -a, a, deg)
        report(f"exp q deg{deg}", c, lambda c, r: 1 + r + r * r * horner(c, r), mp.exp, -a, a)
        out[f"exp{deg}"] = c
    # log(w) = 2 s + 2 s z P(z), z = s^2 <= 0.03
    zm = mp.mpf("0.0300")
    for deg in (4, 5, 6):
        c = cheb_fit(lambda z: (mp.atanh(mp.sqrt(z)) / mp.sqrt(z) - 1) / z if z != 0 else mp.mpf(1) / 3, mp.mpf(0), zm, deg)
        report(f"log P deg{deg}", c, lambda c, z: 2 * mp.sqrt(z) * (1 + z * horner(c, z)), lambda z: 2 * mp.atanh(mp.sqrt(z)), mp.mpf(0), zm)
        out[f"log{deg}"] = c
    # sin(r) = r + r z S(z), |r| <= pi/2
    zm = (mp.pi / 2 * mp.mpf("1.00001")) ** 2
    for deg in (5, 6, 7):
        c = cheb_fit(lambda z: (mp.sin(mp.sqrt(z)) / mp.sqrt(z) - 1) / z if z != 0 else -mp.mpf(1) / 6, mp.mpf(0), zm, deg)
        report(f"sin S deg{deg}", c, lambda c, z: mp.sqrt(z) * (1 + z * horner(c, z)), lambda z: mp.sin(mp.sqrt(z)), mp.mpf(0), zm)
        out[f"sin{deg}"] = c
    # cosh(r) = C(z), sinh(r)/r = Sh(z), z = r^2 <= 0.34658^2
    zm = a * a
    for deg in (3, 4, 5):
        c = cheb_fit(lambda z: mp.cosh(mp.sqrt(z)), mp.mpf(0), zm, deg)
        report(f"cosh C deg{deg}", c, lambda c, z: horner(c, z), lambda z: mp.cosh(mp.sqrt(z)), mp.mpf(0), zm)
        out[f"ch{deg}"] = c
        c = cheb_fit(lambda z: mp.sinh(mp.sqrt(z)) / mp.sqrt(z) if z != 0 else mp.mpf(1), mp.mpf(0), zm, deg)
        report(f"sinh Sh deg{deg}", c, lambda c, z: horner(c, z), lambda z: mp.sinh(mp.sqrt(z)) / mp.sqrt(z), mp.mpf(0), zm)
        out[f"sh{deg}"] = c
    # tan: sin and cos on |r| <= pi/4
    zm = (mp.pi / 4 * mp.mpf("1.00001")) ** 2
    for deg in (3, 4, 5):
        c = cheb_fit(lambda z: (mp.sin(mp.sqrt(z)) / mp.sqrt(z) - 1) / z if z != 0 else -mp.mpf(1) / 6, mp.mpf(0), zm, deg)
        report(f"sin4 S deg{deg}", c, lambda c, z: mp.sqrt(z) * (1 + z * horner(c, z)), lambda z: mp.sin(mp.sqrt(z)), mp.mpf(0), zm)
        out[f"s4_{deg}"] = c
        c = cheb_fit(lambda z: (mp.cos(mp.sqrt(z)) - 1) / z if z != 0 else -mp.mpf(1) / 2, mp.mpf(0), zm, deg)
        report(f"cos4 C deg{deg}", c, lambda c, z: 1 + z * horner(c, z), lambda z: mp.cos(mp.sqrt(z)), mp.mpf(0), zm)
        out[f"c4_{deg}"] = c
    # atan(t) = t + t z T(z), |t| <= 0.225
    zm = mp.mpf("0.225") ** 2
    for deg in (4, 5, 6):
        c = cheb_fit(lambda z: (mp.atan(mp.sqrt(z)) / mp.sqrt(z) - 1) / z if z != 0 else -mp.mpf(1) / 3, mp.mpf(0), zm, deg)
        report(f"atan T deg{deg}", c, lambda c, z: mp.sqrt(z) * (1 + z * horner(c, z)), lambda z: mp.atan(mp.sqrt(z)), mp.mpf(0), zm)
        out[f"at{deg}"] = c
    # constants
    def hilo(x):
        hi = float(x)
        lo = float(x - mp.mpf(hi))
        return hi, lo
    for nm, v in (("ln2", mp.log(2)), ("pi", mp.pi), ("pi/2", mp.pi / 2), ("1/pi", 1 / mp.pi), ("2/pi", 2 / mp.pi), ("log2e", 1 / mp.log(2)),
                  ("atan(0.4375)", mp.atan(mp.mpf("0.4375"))), ("atan(1)", mp.atan(1)), ("ln2/2", mp.log(2) / 2)):
        print(f"// {nm}: hi {hilo(v)[0]!r} lo {hilo(v)[1]!r}")


if __name__ == "__main__":
    main()
