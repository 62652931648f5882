"""CPU check of the arithmetic behind glow_b200/csrc/mask8.cu: the masked sum of squared residuals
sum_s mask[s] r[s]^2 computed from r^2 in 55-bit fixed point (relative to max r^2), split into 7 balanced
int8 digits with the byte-bias trick the kernel uses, contracted with the 0/1 mask in exact integer
arithmetic and recombined in FP64."""
from fractions import Fraction

import numpy as np

BIAS = 0x0080808080808080


def digits_by_bias(r2, mx):
    """rsq4_kernel PASS 1: q = rint(r^2 * 2^54 / max); the bytes of q + 0x00808080808080 are digit + 128."""
    inv = 2.0**54 / mx
    q = [int(np.rint(v * inv)) for v in r2]
    out = np.empty((7, len(q)), dtype=np.int64)
    for j, qq in enumerate(q):
        b = qq + BIAS
        assert 0 <= b < 2**56
        for i in range(7):
            out[i, j] = (((b >> (8 * i)) & 0xFF) ^ 0x80) - (256 if (((b >> (8 * i)) & 0xFF) ^ 0x80) >= 128 else 0)
    return q, out


def test_bias_trick_gives_balanced_digits_of_the_fixed_point_value():
    rg = np.random.default_rng(3)
    r2 = rg.standard_normal(4000)**2 * rg.uniform(1e-8, 1.0, 4000)
    r2[0] = 0.0
    mx = float(r2.max())
    q, d = digits_by_bias(r2, mx)
    assert d.min() >= -128 and d.max() <= 127
    assert abs(max(q) - 2**54) <= 4 and min(q) == 0  # the maximum maps to 2^54 (+- the rounding of 2^54 / max)
    for j, qq in enumerate(q):
        assert sum(int(d[i, j]) * 256**i for i in range(7)) == qq


def test_masked_sum_matches_the_exact_sum():
    rg = np.random.default_rng(4)
    N = 30000
    r = rg.standard_normal(N) * rg.uniform(0.1, 1.5, N)
    r2 = r * r
    mask = (rg.random(N) > 0.1).astype(np.int64)
    mx = float(r2.max())
    _, d = digits_by_bias(r2, mx)
    z = [int(np.dot(d[i], mask)) for i in range(7)]          # what the int32 accumulators hold
    assert max(abs(v) for v in z) <= 128 * N < 2**31
    acc = 0.0
    for i in range(6, -1, -1):                                # recombine_kernel
        acc = acc * 256.0 + float(z[i])
    got = acc * (mx * 2.0**-54)
    exact = sum(Fraction(float(v)) for v, m in zip(r2, mask) if m)
    err = abs(Fraction(got) - exact)
    assert float(err) <= N * mx * 2.0**-55 + float(exact) * 2.0**-51   # a-priori bound: N quantisation errors + FP64 Horner
    assert float(err) <= 1e-14 * float(exact)                          # in practice far below the 1e-9 parity bar
    err_fp64 = abs(Fraction(float(np.dot(mask.astype(np.float64), r2))) - exact)
    assert float(err) <= 64 * float(err_fp64) + 1e-16 * float(exact)
