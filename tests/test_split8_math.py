"""CPU check of the arithmetic behind glow_b200/csrc/split8.cu: an FP64 column quantised to a 55-bit
integer, split into 7 balanced radix-256 int8 digits, contracted with integer genotypes in exact
integer arithmetic and recombined in FP64 reproduces the exact dot product to better than an FP64
GEMM's own rounding -- this is why the int8 tensor-core path may replace the FP64 one."""
from fractions import Fraction

import numpy as np


def split_digits(w):
    """Mirrors pack_split8_kernel: q = rint(w / max|w| * 2^54); 7 digits d_i in [-128, 127], q = sum d_i 256^i."""
    mx = float(np.max(np.abs(w)))
    q = np.rint(w / mx * 2.0**54).astype(object)  # exact Python ints
    digits = []
    for _ in range(7):
        d = [((int(v) + 128) & 255) - 128 for v in q]
        q = np.array([(int(v) - dd) >> 8 for v, dd in zip(q, d)], dtype=object)
        digits.append(np.array(d, dtype=np.int64))
    assert all(int(v) == 0 for v in q), 'seven digits must cover the 55-bit integer'
    return mx, digits


def test_digits_are_int8_and_reconstruct_the_quantised_value():
    rg = np.random.default_rng(0)
    w = rg.standard_normal(5000) * rg.uniform(1e-6, 1.0, 5000)
    w[0], w[1] = np.abs(w).max() * 1.0000001, -np.abs(w).max() * 1.0000001  # +/- column maximum
    mx, digits = split_digits(w)
    for d in digits:
        assert d.min() >= -128 and d.max() <= 127
    q = sum(int(256**i) * digits[i].astype(object) for i in range(7))
    wq = np.array([float(Fraction(int(v)) * Fraction(mx) / Fraction(2**54)) for v in q])
    # quantum 2^-54 of the column maximum (round to nearest) plus the FP64 rounding of w / max|w| itself
    assert np.all(np.abs(wq - w) <= mx * 2.0**-55 + np.abs(w) * 2.0**-52)


def test_integer_contraction_with_recombination_matches_exact_dot_product():
    rg = np.random.default_rng(1)
    N = 20000
    w = rg.standard_normal(N) / np.sqrt(N)
    xi = rg.integers(0, 3, N)                      # genotypes 0/1/2 (missing -> 0)
    e = (rg.random(N) < 0.01).astype(np.int64)     # missing-call indicator
    xi[e == 1] = 0
    mu = float(xi[e == 0].mean())                  # mean imputation
    mx, digits = split_digits(w)
    # exact int32-range dot products, as the tensor cores produce them
    zx = [int(np.dot(d, xi)) for d in digits]
    ze = [int(np.dot(d, e)) for d in digits]
    assert max(abs(v) for v in zx + ze) < 2**31
    acc = 0.0
    for i in range(6, -1, -1):                      # FP64 Horner sum of the epilogue
        acc = acc * 256.0 + (mu * float(ze[i]) + float(zx[i]))
    got = acc * (mx * 2.0**-54)
    exact = sum(Fraction(float(wi)) * (Fraction(int(a)) + Fraction(mu) * Fraction(int(b))) for wi, a, b in zip(w, xi, e))
    err_split = abs(Fraction(got) - exact)
    err_fp64_gemm = abs(Fraction(float(np.dot(w, xi + mu * e))) - exact)  # what a floating-point dot product gives
    scale = float(sum(abs(Fraction(float(wi))) * (int(a) + mu * int(b)) for wi, a, b in zip(w, xi, e)))
    assert float(err_split) <= 4 * N * 2.0**-53 * mx            # a-priori bound: N operand errors of ~1 ulp(max|w|) * |x| <= 2
    assert float(err_split) <= 1e-15 * scale                    # in practice: at the FP64 rounding level of the result
    assert float(err_split) <= 8 * float(err_fp64_gemm) + 1e-17 * scale


def test_reduced_digit_certificate_bounds_the_true_error():
    """The certificate of aux.cu (xx_full_kernel): with the covariate columns quantised to d digits, the error of
    XdotX = sum x^2 - sum_c a_c^2 is at most sum_c (2 |a_c| + t_c) t_c with t_c = 0.51 u_c sum|x|, u_c = max|q_c| 2^-(8d-2).
    Checked in exact rational arithmetic for d = 5 and 6, including a column with an outlier."""
    rg = np.random.default_rng(4)
    N, K = 4000, 5
    C = rg.standard_normal((N, K))
    C[17, 2] = 1e4
    Q, _ = np.linalg.qr(C - C.mean(0))
    x = rg.integers(0, 3, N).astype(np.float64)
    x[rg.random(N) < 0.02] = 0.37                       # mean-substituted entries (>= 0)
    xf = [Fraction(float(v)) for v in x]
    sabs = float(sum(xf))
    for d in (5, 6):
        err_xx, cert = Fraction(0), 0.0
        for c in range(K):
            q = Q[:, c]
            mx = float(np.max(np.abs(q)))
            u = mx * 2.0**-(8 * d - 2)
            qi = np.rint(q / mx * 2.0**(8 * d - 2))     # what pack_split8_kernel stores
            a_hat = sum(Fraction(int(v)) * xv for v, xv in zip(qi, xf)) * Fraction(u)   # exact integer contraction, rescaled
            a = sum(Fraction(float(v)) * xv for v, xv in zip(q, xf))
            t = 0.51 * u * sabs
            assert abs(a_hat - a) <= Fraction(t)
            err_xx += a_hat * a_hat - a * a
            cert += (2.0 * abs(float(a_hat)) + t) * t
        assert abs(err_xx) <= Fraction(cert)
        assert cert > 0.0
