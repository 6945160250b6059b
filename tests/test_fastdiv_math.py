"""The arithmetic behind `csrc/fastdiv.cuh`, checked with exact rational arithmetic on the CPU (no GPU needed):

  rcp_fast(a)    seed r0 with relative error <= 1e-6 (MUFU.RCP64H; measured 9.9e-7 on B200), then two Newton steps
                 e = fma(-a, r, 1), r = fma(r, e, r)
  div_fast(a, b) q = a*r, then Markstein's correction q' = fma(fma(-q, b, a), r, q) with r = rcp_fast(b)

Every fma / product is rounded ONCE, as on the device (emulated through fractions.Fraction -> float, which rounds to
nearest even).  The device-side measurement over 2^24 operands is profiles/tools/rcp_check.cu (tests/test_gpu_fastdiv.py)."""
from fractions import Fraction

import numpy as np


def _rn(x: Fraction) -> float:
    return float(x)  # correctly rounded (round-half-even) conversion


def _fma(a: float, b: float, c: float) -> float:
    return _rn(Fraction(a) * Fraction(b) + Fraction(c))


def _rcp_fast(a: float, seed_rel_err: float) -> float:
    r = _rn(Fraction(1) / Fraction(a) * (1 + Fraction(seed_rel_err)))
    for _ in range(2):
        e = _fma(-a, r, 1.0)
        r = _fma(r, e, r)
    return r


def _div_fast(a: float, b: float, seed_rel_err: float) -> float:
    r = _rcp_fast(b, seed_rel_err)
    q = _rn(Fraction(a) * Fraction(r))
    return _fma(_fma(-q, b, a), r, q)


def _operands(n, seed):
    rng = np.random.Generator(np.random.Philox(seed))
    mant = 1.0 + rng.random(n)
    expo = rng.integers(-40, 40, n)
    sign = np.where(rng.random(n) < 0.5, -1.0, 1.0)
    return sign * np.ldexp(mant, expo)


def test_reciprocal_is_correctly_rounded():
    a = _operands(3000, 1)
    rng = np.random.Generator(np.random.Philox(2))
    err = rng.uniform(-1e-6, 1e-6, a.size)  # the hardware seed's error band
    bad = 0
    for x, e in zip(a, err):
        ref = _rn(Fraction(1) / Fraction(float(x)))
        bad += _rcp_fast(float(x), float(e)) != ref
    assert bad == 0


def test_markstein_quotient_is_correctly_rounded():
    a, b = _operands(3000, 3), _operands(3000, 4)
    rng = np.random.Generator(np.random.Philox(5))
    err = rng.uniform(-1e-6, 1e-6, a.size)
    bad = 0
    for x, y, e in zip(a, b, err):
        ref = _rn(Fraction(float(x)) / Fraction(float(y)))
        bad += _div_fast(float(x), float(y), float(e)) != ref
    assert bad == 0


def test_clamp_divisors_of_differentiate_solution():
    """One fixed divisor, many numerators: kappa_tol * gamma_reg for the tolerances of the reference's examples."""
    rng = np.random.Generator(np.random.Philox(6))
    for kt in (1e-8, 1e-4, 2e-4, 1e-5):
        d = kt * 0.1
        y = np.ldexp(1.0 + rng.random(400), rng.integers(-40, 10, 400))
        for v in y:
            assert _div_fast(float(v), d, 7e-7) == _rn(Fraction(float(v)) / Fraction(d))


def test_over_n_matches_division():
    """over_n<24>: q = v*c, q' = fma(fma(-q, 24, v), c, q) with c = RN(1/24)."""
    rng = np.random.Generator(np.random.Philox(7))
    c = 1.0 / 24.0
    for v in np.ldexp(1.0 + rng.random(1500), rng.integers(-40, 10, 1500)):
        v = float(v)
        q = _rn(Fraction(v) * Fraction(c))
        assert _fma(_fma(-q, 24.0, v), c, q) == _rn(Fraction(v) / 24)
