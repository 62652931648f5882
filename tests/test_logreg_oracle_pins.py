"""Pins of the logistic-regression oracle (oracle/logreg_oracle.py): against the reference's own log_reg.py executed
byte-for-byte (build container only), against the golden fixtures that code generated, against an independent
per-variant formulation of the score test (the reference tests' statsmodels_baseline, test_log_reg.py:35-65, at their
rtol 1e-3), and the restated IRLS against a plain Newton-Raphson solve of the score equations."""
import os

import numpy as np
import pandas as pd
import pytest

from oracle import logreg_oracle as lo
from oracle import refshim

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'logreg_cases.npz')


def load_cases():
    d = np.load(GOLDEN)
    names = sorted({k.split('/', 1)[0] for k in d.files})
    return {n: {k.split('/', 1)[1]: d[k] for k in d.files if k.startswith(n + '/')} for n in names}


def case_frames(c):
    P = c['phenotypes'].shape[1]
    phe = pd.DataFrame(c['phenotypes'], columns=[f't{i}' for i in range(P)])
    cov = pd.DataFrame(c['covariates'])
    off = pd.DataFrame(c['offsets'], columns=phe.columns) if 'offsets' in c else None
    return phe, cov, off, bool(int(c['intercept']))


@pytest.mark.parametrize('name', ['multiple', 'multiple_missing', 'no_intercept', 'simple_offset', 'hardcalls'])
def test_oracle_reproduces_reference_fixtures(name):
    c = load_cases()[name]
    phe, cov, off, icpt = case_frames(c)
    C, mask, st = lo.prepare_state(phe, cov, off, add_intercept=icpt)
    res = lo.block_score_test(c['X'], C, mask, st)
    np.testing.assert_allclose(res['chisq'], c['ref_chisq'], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(res['pvalue'], c['ref_pvalue'], rtol=1e-10, atol=1e-300)


@pytest.mark.skipif(not refshim.available(), reason='reference tree only exists in the build container')
def test_oracle_equals_live_reference(rg):
    N, G, P, K = 200, 11, 3, 4
    X = rg.integers(0, 3, (G, N)).astype(np.float64)
    cov = pd.DataFrame(rg.standard_normal((N, K)))
    Y = rg.integers(0, 2, (N, P)).astype(np.float64)
    Y[rg.random((N, P)) < 0.08] = np.nan
    phe = pd.DataFrame(Y, columns=['a', 'b', 'c'])
    off = pd.DataFrame(0.3 * rg.standard_normal((N, P)), columns=phe.columns)
    for o in (None, off):
        out, ref_state = refshim.reference_score_test(X, phe, cov, offset_df=o)
        C, mask, st = lo.prepare_state(phe, cov, o)
        res = lo.block_score_test(X, C, mask, st)
        np.testing.assert_allclose(res['chisq'].ravel(), out['chisq'].to_numpy(), rtol=1e-12)
        np.testing.assert_allclose(res['pvalue'].ravel(), out['pvalue'].to_numpy(), rtol=1e-12)
        np.testing.assert_allclose(st.inv_CtGammaC, ref_state.inv_CtGammaC, rtol=1e-12)
        assert out['phenotype'].tolist() == [p for p in 'abc' for _ in range(G)]  # phenotype-major


def test_irls_restatement_against_newton_raphson(rg):
    """The restated statsmodels fit converges to the root of the score equations C^T (y - mu) = 0, found here by a
    different algorithm (Newton-Raphson on the log-likelihood from beta = 0)."""
    N, K = 500, 5
    X = np.hstack([np.ones((N, 1)), rg.standard_normal((N, K - 1))])
    beta_true = rg.standard_normal(K) * 0.7
    y = (rg.random(N) < 1 / (1 + np.exp(-(X @ beta_true)))).astype(np.float64)
    off = 0.2 * rg.standard_normal(N)
    beta, mu = lo.glm_binomial_fit(y, X, off)
    b = np.zeros(K)
    for _ in range(50):
        m = 1 / (1 + np.exp(-(X @ b + off)))
        step = np.linalg.solve(X.T @ ((m * (1 - m))[:, None] * X), X.T @ (y - m))
        b += step
        if np.abs(step).max() < 1e-14:
            break
    np.testing.assert_allclose(beta, b, rtol=1e-9, atol=1e-11)
    assert np.abs(X.T @ (y - mu)).max() < 1e-9


def test_score_test_against_direct_per_variant_formulation(rg):
    """test_log_reg.py::assert_glow_equals_golden restated: the block formulation vs the per-(variant, phenotype) Rao
    score test, at the reference tests' rtol 1e-3 (it holds to ~1e-12)."""
    N, G, P, K = 150, 6, 4, 3
    X = rg.random((G, N))
    cov = pd.DataFrame(rg.random((N, K)))
    Y = rg.integers(0, 2, (N, P)).astype(np.float64)
    Y[rg.random((N, P)) < 0.1] = np.nan
    phe = pd.DataFrame(Y)
    C, mask, st = lo.prepare_state(phe, cov)
    res = lo.block_score_test(X, C, mask, st)
    for p in range(P):
        for g in range(G):
            chisq, pv = lo.direct_score_test(X[g], Y[:, p], C)
            assert np.isclose(res['chisq'][p, g], chisq, rtol=1e-3)
            assert np.isclose(res['pvalue'][p, g], pv, rtol=1e-3)
            assert np.isclose(res['chisq'][p, g], chisq, rtol=1e-9)


def test_product_null_fit_matches_oracle(rg):
    """The host-side null fit of the product (numpy, no GPU needed) against the oracle's."""
    from glow_b200.gwas import log_reg as plr
    N, P, K = 300, 3, 4
    cov = pd.DataFrame(rg.standard_normal((N, K)))
    Y = rg.integers(0, 2, (N, P)).astype(np.float64)
    Y[rg.random((N, P)) < 0.1] = np.nan
    phe = pd.DataFrame(Y, columns=['a', 'b', 'c'])
    contigs = ['1', '2']
    off = pd.DataFrame(0.3 * rg.standard_normal((N * 2, P)), index=pd.MultiIndex.from_product([phe.index, contigs]), columns=phe.columns)
    C, mask, st = lo.prepare_state(phe, cov, off)
    got = plr._create_log_reg_state(phe, off, C, None)
    assert list(got) == contigs
    for c in contigs:
        np.testing.assert_allclose(got[c].gamma, st[c].gamma, rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(got[c].Y_res, st[c].Y_res, rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(got[c].inv_CtGammaC, st[c].inv_CtGammaC, rtol=1e-9)
    with pytest.raises(ValueError, match='0 / 1'):
        plr.logistic_regression(pd.DataFrame({'values': [np.zeros(N)]}), phe + 0.5, cov, correction='none')
    with pytest.raises(ValueError, match='correction methods'):
        plr.logistic_regression(pd.DataFrame({'values': [np.zeros(N)]}), phe, cov, correction='bogus')


@pytest.mark.parametrize('name', ['firth', 'firth_missing_offset'])
def test_host_firth_port_reproduces_reference_fixtures(name):
    """The product's host-side approximate Firth refit (glow_b200/gwas/approx_firth.py) against fixtures produced by the
    reference's own approx_firth.py: null fit offsets, per-variant refits, failure flags.  The score statistics that
    select the tests come from the oracle here (the GPU suite takes them from the device)."""
    from glow_b200.gwas import approx_firth as paf
    c = load_cases()[name]
    phe, cov, off, icpt = case_frames(c)
    thr = float(c['threshold'])
    C, mask, st = lo.prepare_state(phe, cov, off, add_intercept=icpt)
    Q = np.linalg.qr(C)[0]
    X = c['X']
    Xl = (X.T - Q @ (Q.T @ X.T)).T  # log_reg.py:312-313
    res = lo.block_score_test(Xl, C, mask, st)
    np.testing.assert_allclose(res['pvalue'], c['ref_score_pvalue'], rtol=1e-9)
    Y = np.nan_to_num(c['phenotypes'])
    P, G = res['chisq'].shape
    for p in range(P):
        o = None if off is None else off.iloc[:, p].to_numpy()
        fo = paf.perform_null_firth_fit(Y[:, p], C, mask[:, p], o, icpt)
        for g in range(G):
            if res['pvalue'][p, g] >= thr:
                assert c['ref_succeeded'][p, g] == -1
                np.testing.assert_allclose(res['chisq'][p, g], c['ref_chisq'][p, g], rtol=1e-9)
                continue
            fit = paf.correct_approx_firth(Xl[g], Y[:, p], fo, mask[:, p])
            assert (fit is not None) == bool(c['ref_succeeded'][p, g] == 1)
            if fit is not None:
                np.testing.assert_allclose([fit.effect, fit.stderror, fit.chisq, fit.pvalue],
                                           [c['ref_effect'][p, g], c['ref_stderror'][p, g], c['ref_chisq'][p, g], c['ref_pvalue'][p, g]],
                                           rtol=1e-8)
