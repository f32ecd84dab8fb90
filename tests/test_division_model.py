"""CPU model of the table division used by the forward kernels (mmh_common.cuh: div_fast / div_rare), in exact rational arithmetic.

The kernels replace the reference's `/ SQRT[k]` (vanilla/core.py:104) by a Markstein-corrected multiplication with r = RN(1/s) and, for
numerators outside [2^-899, 2^900), by the same sequence on a power-of-two-scaled numerator.  This test restates those sequences with
every rounding done by `fractions.Fraction` -> float (round-to-nearest-even, Python's exact conversion) and checks them against the
correctly rounded quotient float(Fraction(x) / Fraction(s)) on random numerators over all three ranges.  It pins the arithmetic claim of
DESIGN.md section 3 without a GPU; the GPU tests check the compiled kernels bit for bit against the oracle."""
import math
import random
from fractions import Fraction

import numpy as np


def rn(fr: Fraction) -> float:
    return float(fr)            # exact rational -> nearest double, ties to even


def fma(a: float, b: float, c: float) -> float:
    return rn(Fraction(a) * Fraction(b) + Fraction(c))


def div_fast(x: float, s: float, r: float) -> float:
    q = rn(Fraction(x) * Fraction(r))
    e = fma(-s, q, x)
    q = fma(e, r, q)
    e = fma(-s, q, x)
    return fma(e, r, q)


def div_rare(x: float, s: float, r: float) -> float:
    ex = math.frexp(abs(x))[1] - 1                      # floor(log2 |x|)
    if -1008 <= ex < -899:
        return rn(Fraction(div_fast(rn(Fraction(x) * Fraction(2) ** 600), s, r)) * Fraction(1, 2 ** 600))
    if 900 <= ex < 1024:
        return rn(Fraction(div_fast(rn(Fraction(x) * Fraction(1, 2 ** 200)), s, r)) * Fraction(2) ** 200)
    raise AssertionError("IEEE path")


def test_table_division_is_correctly_rounded():
    rnd = random.Random(7)
    ks = [1, 2, 3, 5, 7, 10, 49, 50, 99, 1000, 4095, 65537, 99999, (1 << 24) - 1] + [rnd.randrange(1, 100000) for _ in range(40)]
    for k in ks:
        s = float(np.sqrt(np.float64(k)))               # SQRT = np.sqrt(np.arange(...)) (core.py:22)
        r = 1.0 / s                                      # host table: correctly rounded reciprocal
        for _ in range(60):
            m = rnd.uniform(1.0, 2.0) * rnd.choice((-1.0, 1.0))
            for lo, hi, fn in ((-899, 899, div_fast), (-1008, -900, div_rare), (900, 1023, div_rare)):
                x = math.ldexp(m, rnd.randint(lo, hi))
                want = rn(Fraction(x) / Fraction(s))
                assert fn(x, s, r) == want, (k, x)
