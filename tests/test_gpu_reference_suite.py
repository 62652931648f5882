"""The reference's own lin_reg test cases (python/glow/gwas/tests/test_lin_reg.py) that the other GPU
modules do not already mirror, at the reference's shapes.  Each case runs through
glow_b200.gwas.linear_regression (pandas block path -> C ABI -> CUDA) and is compared

  * with the oracle (the shim-loaded reference kernel) at the north_star tolerances, and
  * with the reference tests' own golden, `statsmodels_baseline` (test_lin_reg.py:62-111), at the
    reference's rtol 1e-3 (statsmodels is not installed here: numpy.linalg.lstsq + scipy.stats.t
    follow that function step by step).
"""
import numpy as np
import pandas as pd
import pytest
from scipy import stats

from helpers import STATS, assert_stats_close
from oracle import linreg_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture
def rg():
    return np.random.default_rng(0)  # test_lin_reg.py conftest: default_rng(0)


def _residualize(Y, C):  # test_lin_reg.py:55-59
    coef = np.linalg.lstsq(C, np.nan_to_num(Y), rcond=None)[0]
    return Y - C @ coef


def ols_baseline(genotype_df, phenotype_df, covariate_df=None, offset_dfs=None, add_intercept=True):
    """statsmodels_baseline (test_lin_reg.py:62-111) restated without statsmodels: per (phenotype, variant) the
    standardised, covariate-residualised phenotype (minus the offset) is regressed on the residualised genotype
    alone (no intercept), with the residual degrees of freedom overridden to N - K - 1.  sm.OLS with one regressor
    is beta = x'y / x'x, bse = sqrt(ssr / df_resid / x'x), p = 2 t.sf(|t|, df_resid)."""
    n = phenotype_df.shape[0]
    C = covariate_df.to_numpy(np.float64) if covariate_df is not None and not covariate_df.empty else np.empty((n, 0))
    if add_intercept:
        C = np.column_stack([C, np.ones(n)]) if C.size else np.ones((n, 1))
    Y = phenotype_df.to_numpy(np.float64)
    X = genotype_df.to_numpy(np.float64)
    dof = C.shape[0] - C.shape[1] - 1
    rows = []
    for p in range(Y.shape[1]):
        for g in range(X.shape[1]):
            y = Y[:, p].copy()
            ok = ~np.isnan(y)
            y[~ok] = 0
            y -= y.mean()
            y = _residualize(y, C) * ok
            scale = np.sqrt((y**2).sum() / (ok.sum() - C.shape[1]))
            y /= scale
            if offset_dfs is not None:
                y = y - offset_dfs[g].reindex(phenotype_df.index).iloc[:, p].to_numpy(np.float64)
            x = _residualize(X[:, g], C)
            xo, yo = x[ok], y[ok]
            beta = (xo @ yo) / (xo @ xo)
            resid = yo - beta * xo
            bse = np.sqrt(resid @ resid / dof / (xo @ xo))
            t = beta / bse
            rows.append((beta * scale, bse * scale, t, 2 * stats.t.sf(abs(t), dof), str(phenotype_df.columns[p])))
    return pd.DataFrame(rows, columns=['effect', 'stderror', 'tvalue', 'pvalue', 'phenotype'])


def run_case(genotype_df, phenotype_df, covariate_df=None, offset_df=None, extra_cols=None, **kw):
    """genotype_df is samples x variants as in the reference tests; returns (ours, oracle frame)."""
    import glow_b200
    values_column = kw.get('values_column', 'values')
    pdf = pd.DataFrame({values_column: list(genotype_df.to_numpy().T)})
    if extra_cols is not None:
        pdf = pd.concat([pdf, extra_cols], axis=1)
    out = glow_b200.gwas.linear_regression(pdf, phenotype_df, covariate_df if covariate_df is not None else pd.DataFrame({}),
                                           offset_df if offset_df is not None else pd.DataFrame({}), **kw)
    okw = {k: v for k, v in kw.items() if k in ('contigs', 'dt', 'verbose_output', 'intersect_samples', 'genotype_sample_ids')}
    if 'add_intercept' in kw:
        okw['add_intercept_col'] = kw['add_intercept']
    st = orc.prepare_state(phenotype_df, covariate_df, offset_df, **okw)
    opdf = pdf.copy()
    opdf[values_column] = [np.asarray(v, dtype=np.float64) for v in opdf[values_column]]
    ref = orc.block_frame(opdf, values_column, st)
    return out.reset_index(drop=True), ref.reset_index(drop=True)


def check(out, ref, baseline=None, rtol_stat=1e-9, rtol_p=1e-6, rtol_base=1e-3):
    assert out['phenotype'].tolist() == ref['phenotype'].tolist()
    assert_stats_close({k: out[k].to_numpy() for k in STATS}, {k: ref[k].to_numpy() for k in STATS}, rtol_stat, rtol_p, label='vs oracle')
    if baseline is not None:
        for k in STATS:
            np.testing.assert_allclose(out[k].to_numpy(np.float64), baseline[k].to_numpy(), rtol=rtol_base, err_msg=k)
        assert out['phenotype'].tolist() == baseline['phenotype'].tolist()


def test_propagate_extra_cols(rg):  # test_lin_reg.py:272-284
    n = 10
    geno, phe, cov = pd.DataFrame(rg.random((n, 3))), pd.DataFrame(rg.random((n, 5))), pd.DataFrame(rg.random((n, 2)))
    extra = pd.DataFrame({'genotype_idx': range(3), 'animal': 'monkey'})
    out, ref = run_case(geno, phe, cov, extra_cols=extra)
    assert sorted(out['genotype_idx'].tolist()) == [0] * 5 + [1] * 5 + [2] * 5
    assert (out['animal'] == 'monkey').all()
    assert out.columns.tolist() == ['genotype_idx', 'animal', 'effect', 'stderror', 'tvalue', 'pvalue', 'phenotype']
    assert out['genotype_idx'].tolist() == ref['genotype_idx'].tolist()
    check(out, ref, ols_baseline(geno, phe, cov))


def test_different_values_column(rg):  # test_lin_reg.py:288-301
    n = 10
    geno, phe, cov = pd.DataFrame(rg.random((n, 3))), pd.DataFrame(rg.random((n, 5))), pd.DataFrame(rg.random((n, 2)))
    out, ref = run_case(geno, phe, cov, values_column='some_other_name')
    assert 'some_other_name' not in out.columns
    check(out, ref, ols_baseline(geno, phe, cov))


def test_intercept_no_covariates(rg):  # test_lin_reg.py:305-313
    n = 10
    geno, phe = pd.DataFrame(rg.random((n, 3))), pd.DataFrame(rg.random((n, 5)))
    out, ref = run_case(geno, phe)
    check(out, ref, ols_baseline(geno, phe))


def test_simple_offset(rg):  # test_lin_reg.py:340-354
    n, P, G = 25, 6, 10
    geno, phe, cov = pd.DataFrame(rg.random((n, G))), pd.DataFrame(rg.random((n, P))), pd.DataFrame(rg.random((n, 2)))
    off = pd.DataFrame(rg.random((n, P)))
    out, ref = run_case(geno, phe, cov, off)
    check(out, ref, ols_baseline(geno, phe, cov, [off] * G))


def test_simple_offset_out_of_order(rg):  # test_lin_reg.py:358-375
    n, P, G = 25, 6, 10
    samples = [f'sample_{i}' for i in range(n)]
    traits = [f'trait_{i}' for i in range(P)]
    geno = pd.DataFrame(rg.random((n, G)), samples, [f'snp_{i}' for i in range(G)])
    phe = pd.DataFrame(rg.random((n, P)), samples, traits)
    cov = pd.DataFrame(rg.random((n, 2)))
    off = pd.DataFrame(rg.random((n, P)), samples, traits)
    out, ref = run_case(geno, phe, cov, off.sample(frac=1, random_state=1))
    check(out, ref, ols_baseline(geno, phe, cov, [off] * G))


def test_simple_offset_out_of_order_with_intersect(rg):  # test_lin_reg.py:379-399
    n, P, G = 25, 6, 10
    samples = [f'sample_{i}' for i in range(n)]
    traits = [f'trait_{i}' for i in range(P)]
    geno = pd.DataFrame(rg.random((n, G)), samples, [f'snp_{i}' for i in range(G)])
    phe = pd.DataFrame(rg.random((n, P)), samples, traits)[0:-2]
    cov = pd.DataFrame(rg.random((n, 2)), samples)
    off = pd.DataFrame(rg.random((n, P)), samples, traits)
    out, ref = run_case(geno, phe, cov, off.sample(frac=1, random_state=2), intersect_samples=True, genotype_sample_ids=samples)
    check(out, ref, ols_baseline(geno[0:-2], phe, cov[0:-2], [off[0:-2]] * G))


def test_multi_offset(rg):  # test_lin_reg.py:403-422
    n, P, G = 25, 25, 10
    geno, phe, cov = pd.DataFrame(rg.random((n, G))), pd.DataFrame(rg.random((n, P))), pd.DataFrame(rg.random((n, 10)))
    off = pd.DataFrame(rg.random((n * 2, P)), index=pd.MultiIndex.from_product([phe.index, ['chr1', 'chr2']]))
    extra = pd.DataFrame({'contigName': ['chr1', 'chr2'] * 5})
    out, ref = run_case(geno, phe, cov, off, extra_cols=extra)
    assert out['contigName'].tolist() == ref['contigName'].tolist()
    # the reference's frame is grouped by contig (first appearance): compare its OLS golden per contig group
    order = [g for c in ('chr1', 'chr2') for g in range(G) if extra['contigName'][g] == c]
    base = pd.concat([
        ols_baseline(geno[[g for g in order if extra['contigName'][g] == c]], phe, cov,
                     [off.xs(c, level=1)] * 5) for c in ('chr1', 'chr2')
    ]).reset_index(drop=True)
    check(out, ref, base)


def test_multi_offset_with_missing(rg):  # test_lin_reg.py:426-450
    n, P, G = 25, 24, 18
    contigs = ['chr1', 'chr2', 'chr3']
    geno, phe = pd.DataFrame(rg.random((n, G))), pd.DataFrame(rg.random((n, P)))
    missing = np.zeros(phe.shape)
    np.fill_diagonal(missing, 1)
    missing[:, -1] = 0
    phe[missing.astype('bool')] = np.nan
    cov = pd.DataFrame(rg.random((n, 10)))
    off = pd.DataFrame(rg.random((n * 3, P)), index=pd.MultiIndex.from_product([phe.index, contigs]))
    extra = pd.DataFrame({'contigName': contigs * 6})
    out, ref = run_case(geno, phe, cov, off, extra_cols=extra)
    assert out['contigName'].tolist() == ref['contigName'].tolist()
    order = {c: [g for g in range(G) if extra['contigName'][g] == c] for c in contigs}
    base = pd.concat([ols_baseline(geno[order[c]], phe, cov, [off.xs(c, level=1)] * 6) for c in contigs]).reset_index(drop=True)
    check(out, ref, base)


def test_offset_different_index_order(rg):  # test_lin_reg.py:513-534
    n, P, G = 25, 6, 10
    geno, phe, cov = pd.DataFrame(rg.random((n, G))), pd.DataFrame(rg.random((n, P))), pd.DataFrame(rg.random((n, 2)))
    phe.columns = phe.columns.astype('str')
    off = pd.DataFrame(rg.random((n, P)))
    off.columns = off.columns.astype('str')
    base = ols_baseline(geno, phe, cov, [off] * G)
    shuffled = off.sort_index(axis=0, ascending=False).sort_index(axis=1, ascending=False)
    out, ref = run_case(geno, phe, cov, shuffled)
    check(out, ref, base)


def test_cast_genotypes(rg):  # test_lin_reg.py:538-546: integer genotype arrays are cast to dt
    n = 10
    geno = pd.DataFrame(rg.integers(0, 10, (n, 10)))
    phe, cov = pd.DataFrame(rg.random((n, 5))), pd.DataFrame(rg.random((n, 5)))
    out, ref = run_case(geno, phe, cov)
    assert out['effect'].dtype == np.float64
    check(out, ref, ols_baseline(geno, phe, cov))


def test_cast_genotypes_float32(rg):  # test_lin_reg.py:550-562
    n = 10
    geno = pd.DataFrame(rg.integers(0, 10, (n, 10)))
    phe, cov = pd.DataFrame(rg.random((n, 5))), pd.DataFrame(rg.random((n, 5)))
    out, ref = run_case(geno, phe, cov, dt=np.float32)
    assert out['effect'].dtype == np.float32
    base = ols_baseline(geno, phe, cov)
    for k in STATS:
        np.testing.assert_allclose(out[k].to_numpy(np.float64), base[k].to_numpy(), rtol=1e-3, err_msg=k)


def test_subset_contigs_no_loco(rg):  # test_lin_reg.py:623-634: contigs without LOCO offsets is not an error
    n, P = 50, 5
    geno, phe, off = pd.DataFrame(rg.random((n, 3))), pd.DataFrame(rg.random((n, P))), pd.DataFrame(rg.random((n, P)))
    out, ref = run_case(geno, phe, None, off, contigs=['chr1'])
    check(out, ref, ols_baseline(geno, phe, None, [off] * 3))
