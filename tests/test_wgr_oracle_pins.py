"""Pins oracle/wgr_oracle.py (GloWGR LOCO apply step) to the reference's own ``apply_model``
(python/glow/wgr/ridge_udfs.py:259-347), executed byte-for-byte through oracle/refshim.py::load_wgr.  The Spark
plumbing around it (apply_model_df's join of block and model rows, the header filter of transform_loco, the alpha
choice through cv_df, flatten_prediction_df) is restated here as pandas selections."""
import re

import numpy as np
import pandas as pd
import pytest

from oracle import refshim, wgr_oracle as wo


def make_case(rg, n_samples=37, n_chr=3, blocks_per_chr=2, n_alphas=3, n_sb=3, with_cov=True):
    sample_ids = [f's{i}' for i in range(n_samples)]
    sb_names = [str(i) for i in range(n_sb)]
    sample_blocks = {sb_names[b]: [sample_ids[i] for i in range(n_samples) if i % n_sb == b] for b in range(n_sb)}
    label = 'trait'
    headers, sort_keys = [], []
    for c in range(1, n_chr + 1):
        for blk in range(blocks_per_chr):
            for a in range(n_alphas):
                headers.append(f'chr_{c}_block_{blk}_alpha_{a}_label_{label}')
                sort_keys.append((c * 100 + blk) * 10 + a)
    H = len(headers)
    X_raw = rg.standard_normal((n_samples, H))
    mu, sig = X_raw.mean(axis=0), X_raw.std(axis=0)
    alphas = {f'alpha_{i}': np.float64(10.0**i) for i in range(2)}  # two shrinkage values at this level
    coef = rg.standard_normal((n_sb, H, len(alphas)))  # per sample block, header, alpha
    cov = np.column_stack([np.ones(n_samples), rg.standard_normal((n_samples, 2))]) if with_cov else None
    cov_coef = rg.standard_normal((n_sb, 3, len(alphas))) if with_cov else None
    return dict(sample_ids=sample_ids, sample_blocks=sample_blocks, label=label, headers=headers, sort_keys=sort_keys, X_raw=X_raw,
                mu=mu, sig=sig, alphas=alphas, coef=coef, cov=cov, cov_coef=cov_coef)


def reference_loco(case, best_alpha='alpha_1'):
    """transform_loco for one label through the reference's apply_model: {chromosome: prediction per sample}."""
    ru = refshim.load_wgr()
    c = case
    alpha_names = sorted(c['alphas'])
    labeldf = pd.DataFrame({c['label']: np.zeros(len(c['sample_ids']))}, index=c['sample_ids'])
    covdf = pd.DataFrame(c['cov'], index=c['sample_ids'], columns=['intercept', 'c1', 'c2']) if c['cov'] is not None else pd.DataFrame({})
    chroms = sorted({re.search(r'^chr_(.+?)_(alpha|block)', h).group(1) for h in c['headers']})
    out = {}
    for chrom in chroms:
        pred = pd.Series(np.nan, index=c['sample_ids'])
        for b, (sb, members) in enumerate(c['sample_blocks'].items()):
            rows_idx = [c['sample_ids'].index(s) for s in members]
            rows = []
            if c['cov'] is not None:  # model rows of the covariates: no block row joins them (values null), they sort first
                for k, name in enumerate(covdf.columns):
                    rows.append(dict(header=name, size=None, values=None, sample_block=sb, sort_key=k - 10, mu=None, sig=None,
                                     alphas=alpha_names, labels=[c['label']] * len(alpha_names), coefficients=c['cov_coef'][b, k]))
            for h, header in enumerate(c['headers']):
                if re.search(f'^chr_{chrom}_(alpha|block)', header):  # ridge_regression.py:221-223
                    continue
                rows.append(dict(header=header, size=len(members), values=c['X_raw'][rows_idx, h], sample_block=sb,
                                 sort_key=c['sort_keys'][h], mu=c['mu'][h], sig=c['sig'][h], alphas=alpha_names,
                                 labels=[c['label']] * len(alpha_names), coefficients=c['coef'][b, h]))
            pdf = pd.DataFrame(rows).sample(frac=1.0, random_state=b).reset_index(drop=True)  # group rows arrive unordered
            res = ru.apply_model((sb, c['label']), ['sample_block', 'label'], pdf, labeldf, c['sample_blocks'], c['alphas'], covdf)
            chosen = res[res['alpha'] == best_alpha]  # the inner join with cv_df (model_functions.py:770)
            assert len(chosen) == 1
            pred.loc[members] = chosen['values'].iloc[0]  # flatten_prediction_df: position inside the sample block
        out[chrom] = pred.to_numpy()
    return out


@pytest.mark.skipif(not refshim.available(), reason='reference tree only exists in the build container')
@pytest.mark.parametrize('with_cov', [True, False])
def test_oracle_matches_reference_apply_model(rg, with_cov):
    c = make_case(rg, with_cov=with_cov)
    ref = reference_loco(c, 'alpha_1')
    a = sorted(c['alphas']).index('alpha_1')
    sb = np.array([i % 3 for i in range(len(c['sample_ids']))])
    got = wo.loco_predict(c['headers'], c['X_raw'], c['mu'], c['sig'], c['coef'][:, :, a], sb, sorted(ref),
                          c['cov'], None if c['cov_coef'] is None else c['cov_coef'][:, :, a])
    for ci, chrom in enumerate(sorted(ref)):
        np.testing.assert_allclose(got[ci], ref[chrom], rtol=1e-12, atol=1e-13)
    assert not np.allclose(got[0], got[1])  # leaving out different chromosomes gives different offsets
