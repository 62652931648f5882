"""Shared helpers for the parity tests."""
import json
import os

import numpy as np
import pandas as pd

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
STATS = ('effect', 'stderror', 'tvalue', 'pvalue')

# north_star tolerances (BASELINE.json): rel 1e-9 on effect/stderror/tvalue, rel 1e-6 on pvalue
# with an absolute floor of 1e-300.
RTOL_STAT = 1e-9
RTOL_P = 1e-6
ATOL_P = 1e-300


def load_random_cases():
    arrays = np.load(os.path.join(GOLDEN, 'random_cases.npz'), allow_pickle=False)
    meta = json.load(open(os.path.join(GOLDEN, 'random_cases.json')))
    cases = {}
    for name, m in meta.items():
        c = {k.split('/', 1)[1]: arrays[k] for k in arrays.files if k.startswith(name + '/')}
        c['meta'] = m
        cases[name] = c
    return cases


def case_frames(name, c):
    """Rebuilds the pandas inputs of a stored random case exactly as oracle/make_golden.py made them."""
    m = c['meta']
    N, P = m['N'], m['P']
    kw = {}
    if name == 'intersect':
        samples = [f's{i}' for i in range(c['X'].shape[1])]
        keep = c['keep']
        phe = pd.DataFrame(c['phenotypes'], index=[samples[i] for i in keep])
        cov = pd.DataFrame(c['covariates'], index=samples)
        kw = dict(intersect_samples=True, genotype_sample_ids=samples)
    else:
        cols = ['t0', 't1', 't2'] if name == 'hardcalls_signal' else None
        phe = pd.DataFrame(c['phenotypes'], columns=cols)
        cov = pd.DataFrame(c['covariates']) if 'covariates' in c else None
    off = None
    if 'offsets' in c:
        if 'variant_contigs' in m:
            contigs = list(dict.fromkeys(m['variant_contigs']))
            off = pd.DataFrame(c['offsets'], index=pd.MultiIndex.from_product([phe.index, contigs]))
        else:
            off = pd.DataFrame(c['offsets'])
    for k in ('verbose_output', 'add_intercept'):
        if k in m:
            kw[k] = m[k]
    pdf = pd.DataFrame({'values': list(c['X'])})
    if 'variant_contigs' in m:
        pdf['contigName'] = m['variant_contigs']
    pdf['idx'] = np.arange(c['X'].shape[0])
    return pdf, phe, cov, off, kw


def degenerate_mask(ref_stderror):
    """Tests whose reference output is NaN/inf (monomorphic variants, SURVEY.md section 7 'hard parts')."""
    return ~np.isfinite(ref_stderror)


def assert_stats_close(got, ref, rtol_stat=RTOL_STAT, rtol_p=RTOL_P, label=''):
    """got/ref: dicts of equally shaped arrays.  NaN-ness must agree; finite entries within the
    north_star tolerances.  effect and tvalue additionally get an absolute floor tied to the
    statistic's own noise scale (1e-12 standard errors): a test statistic that is ~0 by
    cancellation cannot agree to 1e-9 *relative* between any two summation orders."""
    ref_se = np.asarray(ref['stderror'], dtype=np.float64)
    for k in STATS:
        g = np.asarray(got[k], dtype=np.float64)
        r = np.asarray(ref[k], dtype=np.float64)
        assert g.shape == r.shape, (label, k, g.shape, r.shape)
        nan_g, nan_r = np.isnan(g), np.isnan(r)
        assert np.array_equal(nan_g, nan_r), f'{label} {k}: NaN pattern differs ({nan_g.sum()} vs {nan_r.sum()})'
        ok = ~nan_r
        if k == 'pvalue':
            rtol, atol = rtol_p, ATOL_P
        elif k == 'effect':
            rtol, atol = rtol_stat, 1e-12 * np.abs(ref_se[ok])
        elif k == 'tvalue':
            rtol, atol = rtol_stat, 1e-12
        else:
            rtol, atol = rtol_stat, 0.0
        err = np.abs(g[ok] - r[ok])
        tol = atol + rtol * np.abs(r[ok])
        bad = err > tol
        if bad.any():
            i = np.argmax(err / np.maximum(tol, 1e-320))
            raise AssertionError(f'{label} {k}: {bad.sum()} of {ok.sum()} outside tolerance; worst got={g[ok][i]!r} '
                                 f'ref={r[ok][i]!r} rel={err[i] / max(abs(r[ok][i]), 1e-320):.3e}')


def load_regenie_case(name):
    """Inputs + regenie 1.0.6.7 outputs of one case of tests/golden/regenie.npz (oracle/make_golden.py::golden_regenie)."""
    d = np.load(os.path.join(GOLDEN, 'regenie.npz'))
    c = {k.split('/', 1)[1]: d[k] for k in d.files if k.startswith(name + '/')}
    samples = [str(s) for s in c['sample_ids']]
    phe = pd.DataFrame(c['phenotypes'], index=samples, columns=['Y1', 'Y2'])
    cov = pd.DataFrame(c['covariates'], index=samples, columns=['V1', 'V2', 'V3'])
    contigs = [str(x) for x in c['offset_contigs']]
    off = pd.concat([pd.DataFrame(c['offsets'][i], index=pd.MultiIndex.from_product([samples, [cn]]), columns=['Y1', 'Y2'])
                     for i, cn in enumerate(contigs)])
    pdf = pd.DataFrame({'contigName': [str(x) for x in c['contigs']], 'names': [str(x) for x in c['rsids']],
                        'values': list(c['states'].astype(np.float64))})
    return c, pdf, phe, cov, off


def regenie_compare(c, out, three_chr):
    """The reference's comparison (python/glow/gwas/tests/functions.py:81-102): BETA, SE, p within
    rtol 1e-2 / atol 1e-3 of regenie on the first 100 SNPs."""
    key = (out['contigName'].astype(str) + ':' + out['names']) if three_chr else out['names']
    ours = pd.DataFrame({'k': [f'{a}|{b}' for a, b in zip(key, out['phenotype'])], 'BETA': out['effect'].to_numpy(),
                         'SE': out['stderror'].to_numpy(), 'pvalue': out['pvalue'].to_numpy()}).set_index('k')
    ours = ours.reindex([str(k) for k in c['regenie_key']])
    assert np.isclose(ours[['BETA', 'SE', 'pvalue']].to_numpy(), c['regenie'], rtol=1e-2, atol=1e-3).all()
