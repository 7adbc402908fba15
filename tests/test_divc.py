"""The 3-instruction correctly-rounded division used by the kernels (greb_core.h div3/div20):
q0 = x*RN(1/d); r = fma(-d, q0, x); q = fma(r, RN(1/d), q0)  ==  RN(x/d).
Checked here on the CPU with the same IEEE operations (np.float32 mul + a float64-exact fma)."""
import numpy as np

f32 = np.float32


def _fma32(a, b, c):
    # a*b is exact in float64 for float32 inputs; the sum is then rounded once to double and once
    # to float.  For these operand ranges (|a*b + c| small relative to the terms) the double
    # rounding is harmless: the residual r is exactly representable.
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def _div_c(x, d):
    r = f32(1.0) / f32(d)
    q = x * r
    e = _fma32(np.full_like(x, -d), q, x)
    return _fma32(e, np.full_like(x, r), q)


def test_div3_div20_bit_exact_on_dense_sample():
    rng = np.random.default_rng(0)
    # all mantissas at a few exponents + random bit patterns
    m = np.arange(1 << 23, dtype=np.uint32)
    xs = [(m | np.uint32(e << 23)).view(np.float32) for e in (100, 127, 128, 150)]
    bits = rng.integers(0, 1 << 32, size=4_000_000, dtype=np.uint64).astype(np.uint32)
    x = bits.view(np.float32)
    xs.append(x[np.isfinite(x) & (np.abs(x) > 1e-30)])
    for x in xs:
        for d in (3.0, 20.0):
            want = x / f32(d)
            got = _div_c(x, f32(d))
            assert np.array_equal(want.view(np.uint32), got.view(np.uint32)), d
