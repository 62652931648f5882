"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):

    python oracle/make_golden.py

Every output array stored here was produced by the reference's own
``_linear_regression_inner`` (python/glow/gwas/lin_reg.py:203-254) loaded through
``oracle/refshim.py``; inputs are stored beside them so the GPU box (which has no reference
tree) can replay the cases.  The fixtures are small (a few hundred KB in total).
"""
import gzip
import io
import json
import os
import re
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import refshim  # noqa: E402

REF = refshim.REFERENCE_ROOT
OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
STATS = ('effect', 'stderror', 'tvalue', 'pvalue')


def read_vcf_states(path):
    """Tiny VCF text reader: returns (contigs, starts(0-based), names, states int8 [G, N], sample ids).
    states = number of alt alleles, -1 when any call is missing (genotype_states,
    core/.../sql/expressions/VariantUtilExprs.scala:37-59)."""
    opener = gzip.open if path.endswith('.gz') else open
    contigs, starts, names, rows, samples = [], [], [], [], None
    with opener(path, 'rt') as f:
        for line in f:
            if line.startswith('##'):
                continue
            parts = line.rstrip('\n').split('\t')
            if line.startswith('#CHROM'):
                samples = parts[9:]
                continue
            gt_idx = parts[8].split(':').index('GT')
            row = np.empty(len(parts) - 9, dtype=np.int8)
            for i, field in enumerate(parts[9:]):
                gt = field.split(':')[gt_idx]
                calls = re.split(r'[|/]', gt)
                row[i] = -1 if any(c == '.' for c in calls) else sum(int(c) for c in calls)
            contigs.append(parts[0])
            starts.append(int(parts[1]) - 1)  # Glow's `start` is 0-based
            names.append(parts[2])
            rows.append(row)
    return contigs, np.array(starts), names, np.stack(rows), samples


def stats_matrix(df, P, G):
    return {k: df[k].to_numpy(dtype=np.float64).reshape(P, G) for k in STATS if k in df}


def run_ref(X_variant_major, phenotype_df, covariate_df=None, offset_df=None, extra=None, **kw):
    pdf = pd.DataFrame({'values': list(np.asarray(X_variant_major, dtype=np.float64))})
    if extra is not None:
        for k, v in extra.items():
            pdf[k] = v
    out = refshim.reference_linear_regression(pdf, phenotype_df, covariate_df, offset_df, **kw)
    return out


def golden_docs():
    """G1: docs/source/tertiary/regression-tests.rst:53-83."""
    from oracle.linreg_oracle import mean_substitute
    contigs, starts, names, states, samples = read_vcf_states(os.path.join(REF, 'test-data/gwas/genotypes.vcf.gz'))
    cov = pd.read_csv(os.path.join(REF, 'test-data/gwas/covariates.csv.gz'), index_col=0)
    phe = pd.read_csv(os.path.join(REF, 'test-data/gwas/continuous-phenotypes.csv.gz'), index_col=0)
    off = pd.read_csv(os.path.join(REF, 'test-data/gwas/continuous-offsets.csv.gz'), index_col=0)
    assert list(cov.index) == samples and list(phe.index) == samples
    X = mean_substitute(states.astype(np.float64))
    out = run_ref(X, phe, cov)
    G, P = X.shape[0], phe.shape[1]
    res = stats_matrix(out, P, G)
    out_off = run_ref(X, phe, cov, off)
    res_off = stats_matrix(out_off, P, G)
    g = [i for i in range(G) if contigs[i] == '22' and starts[i] == 16050114][0]
    row = {k: float(res[k][0, g]) for k in STATS}
    expected = dict(effect=0.14722512852575978,
                    stderror=0.14155327969643167,
                    tvalue=1.0400686500623064,
                    pvalue=0.2984087428847886)
    for k in STATS:
        assert abs(row[k] / expected[k] - 1) < 1e-12, (k, row[k], expected[k])
    np.savez_compressed(os.path.join(OUT, 'docs_gwas.npz'),
                        states=states,
                        contigs=np.array(contigs),
                        starts=starts,
                        names=np.array(names),
                        samples=np.array(samples),
                        covariates=cov.to_numpy(np.float64),
                        covariate_names=np.array(cov.columns).astype(str),
                        phenotypes=phe.to_numpy(np.float64),
                        phenotype_names=np.array(phe.columns).astype(str),
                        offsets=off.to_numpy(np.float64),
                        golden_variant=g,
                        **{'ref_' + k: v
                           for k, v in res.items()},
                        **{'ref_off_' + k: v
                           for k, v in res_off.items()})
    print('docs_gwas: reference reproduces the docs row to',
          max(abs(row[k] / expected[k] - 1) for k in STATS))


def golden_doctest():
    """G2: python/glow/gwas/lin_reg.py:42-51 (np.random.seed(42), genotype_states, -1 kept)."""
    contigs, starts, names, states, samples = read_vcf_states(os.path.join(REF, 'test-data/1kg_sample.vcf'))
    np.random.seed(42)
    n_samples, n_phenotypes = 710, 3
    assert states.shape == (1, n_samples)
    phe = pd.DataFrame(np.random.random((n_samples, n_phenotypes)), columns=['p1', 'p2', 'p3'])
    cov = pd.DataFrame(np.random.random((n_samples, n_phenotypes)))
    out = run_ref(states.astype(np.float64), phe, cov)
    res = stats_matrix(out, 3, 1)
    assert f"{res['effect'][0, 0]:.4f}" == '0.0454' and f"{res['pvalue'][0, 0]:.4f}" == '0.0348'
    np.savez_compressed(os.path.join(OUT, 'onekg_doctest.npz'),
                        states=states,
                        contig=np.array(contigs),
                        start=starts,
                        phenotypes=phe.to_numpy(),
                        covariates=cov.to_numpy(),
                        **{'ref_' + k: v
                           for k, v in res.items()})
    print('onekg_doctest:', {k: float(v[0, 0]) for k, v in res.items()})


def golden_cars():
    """G3: R `cars`, data inline at core/src/test/scala/io/projectglow/tertiary/LinearRegressionSuite.scala:187-251."""
    src = open(os.path.join(REF, 'core/src/test/scala/io/projectglow/tertiary/LinearRegressionSuite.scala')).read()
    block = src[src.index('private val cars'):src.index('.stripMargin')]
    pairs = [(int(a), int(b)) for a, b in re.findall(r'\|\d+\s+(\d+)\s+(\d+)', block)]
    speed = np.array([p[0] for p in pairs], dtype=np.float64)
    dist = np.array([p[1] for p in pairs], dtype=np.float64)
    assert len(pairs) == 50 and speed.sum() == 770 and dist.sum() == 2149, (len(pairs), speed.sum(), dist.sum())
    phe = pd.DataFrame({'dist': dist})
    out = run_ref(speed[None, :], phe)
    res = stats_matrix(out, 1, 1)
    np.savez_compressed(os.path.join(OUT, 'cars.npz'), speed=speed, dist=dist, **{'ref_' + k: v for k, v in res.items()})
    print('cars:', {k: float(v[0, 0]) for k, v in res.items()})


def golden_plink():
    """G5: test-data/plink/five-samples-five-variants/bed-bim-fam/test.bed vs PlinkReaderSuite.scala:75-107."""
    raw = open(os.path.join(REF, 'test-data/plink/five-samples-five-variants/bed-bim-fam/test.bed'), 'rb').read()
    assert raw.hex() == '6c1b014b01bb02bb02bb02bb02'
    # calls listed in PlinkReaderSuite.scala:75-107 -> genotype_states
    expected = np.array([[0, 1, 2, -1, -1]] + [[0, 1, 0, 1, 1]] * 4, dtype=np.int8)
    np.savez_compressed(os.path.join(OUT, 'plink_kat.npz'), bed=np.frombuffer(raw, dtype=np.uint8), states=expected)
    print('plink_kat ok')


def golden_random():
    """Randomised cases in the style of python/glow/gwas/tests/test_lin_reg.py, run through the
    unmodified reference kernel.  Inputs + outputs are stored; meta.json describes each case."""
    rg = np.random.default_rng(20261019)
    arrays, meta = {}, {}

    def store(name, X, phe, cov, off, ref_df, kw, extra=None, contigs_per_variant=None):
        P, G = phe.shape[1], X.shape[0]
        arrays[f'{name}/X'] = X
        arrays[f'{name}/phenotypes'] = phe.to_numpy(np.float64)
        if cov is not None:
            arrays[f'{name}/covariates'] = cov.to_numpy(np.float64)
        if off is not None:
            arrays[f'{name}/offsets'] = off.to_numpy(np.float64)
        for k in STATS + ('n', 'sum_x', 'y_transpose_x'):
            if k in ref_df:
                arrays[f'{name}/ref_{k}'] = ref_df[k].to_numpy(np.float64)
        arrays[f'{name}/ref_phenotype'] = ref_df['phenotype'].to_numpy().astype(str)
        if 'idx' in ref_df:
            arrays[f'{name}/ref_idx'] = ref_df['idx'].to_numpy()
        m = dict(kw)
        m.update(P=P, G=G, N=int(phe.shape[0]))
        if contigs_per_variant is not None:
            m['variant_contigs'] = list(contigs_per_variant)
        meta[name] = m

    # basic: uniform genotypes, covariates (test_multiple, test_lin_reg.py:186-191)
    N, G, P = 100, 10, 25
    X = rg.random((G, N))
    phe = pd.DataFrame(rg.random((N, P)))
    cov = pd.DataFrame(rg.random((N, 5)))
    store('multiple', X, phe, cov, None, run_ref(X, phe, cov, extra={'idx': np.arange(G)}), {})

    # missing phenotypes + verbose (test_verbose_output, :239-257)
    N, G, P = 60, 7, 4
    X = rg.random((G, N)) * 2
    phe = pd.DataFrame(rg.random((N, P)))
    phe.iloc[0, 0] = np.nan
    phe.iloc[[1, 3, 5, 17], 1] = np.nan
    phe.iloc[10:30, 2] = np.nan
    cov = pd.DataFrame(rg.random((N, 3)))
    store('missing_verbose', X, phe, cov, None,
          run_ref(X, phe, cov, extra={'idx': np.arange(G)}, verbose_output=True), dict(verbose_output=True))

    # no covariates / no intercept
    N, G, P = 40, 5, 3
    X = rg.random((G, N))
    phe = pd.DataFrame(rg.random((N, P)))
    cov = pd.DataFrame(rg.random((N, 2)))
    store('no_intercept', X, phe, cov, None,
          run_ref(X, phe, cov, extra={'idx': np.arange(G)}, add_intercept=False), dict(add_intercept=False))
    store('intercept_only', X, phe, None, None, run_ref(X, phe, extra={'idx': np.arange(G)}), {})

    # single offset (test_simple_offset, :340-355)
    N, G, P = 25, 10, 6
    X = rg.random((G, N))
    phe = pd.DataFrame(rg.random((N, P)))
    cov = pd.DataFrame(rg.random((N, 2)))
    off = pd.DataFrame(rg.random((N, P)))
    store('simple_offset', X, phe, cov, off, run_ref(X, phe, cov, off, extra={'idx': np.arange(G)}), {})

    # LOCO offsets with missing phenotypes (test_multi_offset_with_missing, :422-447)
    N, G, P = 50, 18, 8
    contigs = ['chr1', 'chr2', 'chr3']
    X = rg.random((G, N))
    phe = pd.DataFrame(rg.random((N, P)))
    miss = np.zeros(phe.shape)
    np.fill_diagonal(miss, 1)
    miss[:, -1] = 0
    phe[miss.astype(bool)] = np.nan
    cov = pd.DataFrame(rg.random((N, 4)))
    off = pd.DataFrame(rg.random((N * 3, P)), index=pd.MultiIndex.from_product([phe.index, contigs]))
    vc = (contigs * 6)
    ref = run_ref(X, phe, cov, off, extra={'contigName': vc, 'idx': np.arange(G)})
    store('loco_missing', X, phe, cov, off, ref, {}, contigs_per_variant=vc)

    # intersect_samples: phenotype rows are a subset of genotype samples (:217-235)
    N, G, P = 30, 6, 3
    samples = [f's{i}' for i in range(N)]
    X = rg.random((G, N))
    keep = [i for i in range(N) if i not in (0, 7, 29)]
    phe = pd.DataFrame(rg.random((N, P)), index=samples).iloc[keep]
    phe.iloc[[1, 3], 1] = np.nan
    cov = pd.DataFrame(rg.random((N, 3)), index=samples)
    ref = run_ref(X, phe, cov, extra={'idx': np.arange(G)}, intersect_samples=True, genotype_sample_ids=samples)
    arrays['intersect/keep'] = np.array(keep)
    store('intersect', X, phe, cov, None, ref, dict(intersect_samples=True))

    # genotype-like integer states incl. strong signal (small p-values) and an all-zero variant
    N, G, P = 400, 24, 3
    f = rg.uniform(0.02, 0.5, G)
    X = rg.binomial(2, f[:, None], (G, N)).astype(np.float64)
    X[5] = 0.0  # monomorphic -> NaN in the reference
    cov = pd.DataFrame(rg.standard_normal((N, 4)))
    Y = rg.standard_normal((N, P))
    Y[:, 0] += 2.0 * X[0] + 0.5 * X[1]
    Y[:, 1] += 0.3 * X[2]
    phe = pd.DataFrame(Y, columns=['t0', 't1', 't2'])
    ref = run_ref(X, phe, cov, extra={'idx': np.arange(G)})
    store('hardcalls_signal', X, phe, cov, None, ref, {})

    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, 'random_cases.npz'), **arrays)
    with open(os.path.join(OUT, 'random_cases.json'), 'w') as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)
    print('random_cases:', sorted(meta))


def golden_regenie():
    """G4: python/glow/gwas/tests/test_lin_reg.py:632-666 + tests/functions.py:47-102 -- regenie 1.0.6.7
    outputs for the first 100 SNPs, Glow hard calls from the BGEN fixtures, LOCO offsets from the
    regenie step-1 .loco files.  The reference's own comparison (rtol 1e-2, atol 1e-3) is asserted
    here against the UNMODIFIED reference kernel before the fixture is written."""
    from oracle.bgen import read_bgen_states
    d = os.path.join(REF, 'test-data/regenie')

    def fid_iid(df):
        df['FID_IID'] = df['FID'].astype(str) + '_' + df['IID'].astype(str)
        return df.sort_values(by=['FID', 'IID']).drop(columns=['FID', 'IID']).set_index(['FID_IID'])

    def read_offset(file, trait):
        df = pd.melt(pd.read_table(file, sep=r'\s+'), id_vars=['FID_IID']) \
            .rename(columns={'FID_IID': 'contigName', 'variable': 'FID_IID', 'value': trait}) \
            .astype({'FID_IID': 'str', 'contigName': 'str'})
        df[['FID', 'IID']] = df.FID_IID.str.split('_', expand=True).astype(int)
        return df.sort_values(by=['FID', 'IID']).drop(columns=['FID', 'IID']).set_index(['FID_IID', 'contigName'])

    cases = {
        'lin': dict(bgen='example.bgen', loco=('fit_lin_out_1.loco', 'fit_lin_out_2.loco'), prefix='test_lin_out_',
                    missing=[]),
        'lin_missing': dict(bgen='example.bgen', loco=('fit_lin_out_1.loco', 'fit_lin_out_2.loco'),
                            prefix='test_lin_out_missing_',
                            missing=['35_35', '136_136', '77_77', '100_100', '204_204', '474_474']),
        'lin_3chr': dict(bgen='example_3chr.bgen', loco=('fit_lin_out_3chr_1.loco', 'fit_lin_out_3chr_2.loco'),
                         prefix='test_lin_out_3chr_', missing=[]),
    }
    arrays = {}
    for name, c in cases.items():
        contigs, pos, rsids, states = read_bgen_states(os.path.join(d, c['bgen']))
        sel = (pos - 1) < 100  # .filter('start < 100'), start = position - 1
        contigs = [x for x, k in zip(contigs, sel) if k]
        rsids = [x for x, k in zip(rsids, sel) if k]
        states = states[sel]
        phe = fid_iid(pd.read_table(os.path.join(d, 'phenotype.txt'), sep=r'\s+'))
        phe.loc[c['missing'], :] = np.nan
        cov = fid_iid(pd.read_table(os.path.join(d, 'covariates.txt'), sep=r'\s+'))
        off = pd.merge(read_offset(os.path.join(d, c['loco'][0]), 'Y1'), read_offset(os.path.join(d, c['loco'][1]), 'Y2'),
                       left_index=True, right_index=True)
        assert list(phe.index[:3]) == ['1_1', '2_2', '3_3']
        pdf = pd.DataFrame({'contigName': contigs, 'names': rsids, 'values': list(states.astype(np.float64))})
        out = refshim.reference_linear_regression(pdf, phe, cov, off)
        reg = pd.concat([
            pd.read_table(os.path.join(d, c['prefix'] + t + '.regenie'), sep=r'\s+').astype({'ID': 'str'}).assign(phenotype=t)
            for t in ('Y1', 'Y2')
        ], ignore_index=True)
        reg = reg[reg['GENPOS'] <= 100]
        reg['pvalue'] = np.power(10, -reg['LOG10P'])
        ours = out.rename(columns={'effect': 'BETA', 'stderror': 'SE', 'names': 'ID'})
        if name == 'lin_3chr':
            reg['key'] = reg['CHROM'].astype(str) + ':' + reg['ID']
            ours['key'] = ours['contigName'].astype(str) + ':' + ours['ID']
        else:
            reg['key'], ours['key'] = reg['ID'], ours['ID']
        reg = reg.set_index(['key', 'phenotype'])
        ours = ours.set_index(['key', 'phenotype']).reindex(reg.index)
        cols = ['BETA', 'SE', 'pvalue']
        ok = np.isclose(ours[cols].to_numpy(), reg[cols].to_numpy(), rtol=1e-2, atol=1e-3)
        assert ok.all(), (name, (~ok).sum())
        G = states.shape[0]
        contig_levels = list(dict.fromkeys(off.index.get_level_values(1)))
        arrays.update({
            f'{name}/states': states,
            f'{name}/contigs': np.array(contigs),
            f'{name}/rsids': np.array(rsids),
            f'{name}/phenotypes': phe.to_numpy(np.float64),
            f'{name}/covariates': cov.to_numpy(np.float64),
            f'{name}/sample_ids': np.array(phe.index).astype(str),
            f'{name}/offset_contigs': np.array(contig_levels),
            # offsets as [contig][sample][trait] in phenotype row order
            f'{name}/offsets': np.stack([off.xs(cn, level=1).reindex(phe.index).to_numpy(np.float64) for cn in contig_levels]),
            f'{name}/regenie_key': np.array([f'{k}|{t}' for k, t in reg.index]),
            f'{name}/regenie': reg[cols].to_numpy(np.float64),
        })
        for k in STATS:
            arrays[f'{name}/ref_{k}'] = out[k].to_numpy(np.float64)
        arrays[f'{name}/ref_key'] = np.array([f'{a}|{b}' for a, b in zip(
            (out['contigName'].astype(str) + ':' + out['names']) if name == 'lin_3chr' else out['names'], out['phenotype'])])
        print(f'regenie {name}: reference within rtol 1e-2 / atol 1e-3 of regenie on {ok.size} values')
    np.savez_compressed(os.path.join(OUT, 'regenie.npz'), **arrays)


def golden_logreg():
    """Logistic score test (log_reg.py, correction='none'): randomised cases in the style of
    python/glow/gwas/tests/test_log_reg.py (test_multiple, test_multiple_missing, test_spector_no_intercept,
    test_simple_offset), run through the UNMODIFIED reference block kernel via the reference tests' own
    ``run_score_test`` recipe (oracle/refshim.py::reference_score_test).  The null-model GLM fit is the restated IRLS
    (statsmodels is absent from this image)."""
    rg = np.random.default_rng(20261020)
    arrays = {}
    cases = {
        'multiple': dict(N=50, G=5, P=5, K=3, missing=0.0, intercept=True, offset=False),
        'multiple_missing': dict(N=60, G=5, P=5, K=3, missing=0.1, intercept=True, offset=False),
        'no_intercept': dict(N=64, G=6, P=2, K=2, missing=0.1, intercept=False, offset=False),
        'simple_offset': dict(N=80, G=9, P=3, K=3, missing=0.05, intercept=True, offset=True),
        'hardcalls': dict(N=400, G=40, P=2, K=4, missing=0.05, intercept=True, offset=False, hardcalls=True),
    }
    for name, c in cases.items():
        N, G, P, K = c['N'], c['G'], c['P'], c['K']
        if c.get('hardcalls'):
            X = rg.binomial(2, rg.uniform(0.1, 0.5, (G, 1)), (G, N)).astype(np.float64)
        else:
            X = rg.random((G, N))
        cov = pd.DataFrame(rg.random((N, K)))
        Y = rg.integers(0, 2, (N, P)).astype(np.float64)
        if c['missing']:
            Y[rg.random((N, P)) < c['missing']] = np.nan
        phe = pd.DataFrame(Y, columns=[f't{i}' for i in range(P)])
        off = pd.DataFrame(0.5 * rg.standard_normal((N, P)), columns=phe.columns) if c['offset'] else None
        out, _ = refshim.reference_score_test(X, phe, cov, add_intercept=c['intercept'], offset_df=off)
        arrays[f'{name}/X'] = X
        arrays[f'{name}/phenotypes'] = Y
        arrays[f'{name}/covariates'] = cov.to_numpy(np.float64)
        arrays[f'{name}/intercept'] = np.array(int(c['intercept']))
        if off is not None:
            arrays[f'{name}/offsets'] = off.to_numpy(np.float64)
        arrays[f'{name}/ref_chisq'] = out['chisq'].to_numpy(np.float64).reshape(P, G)
        arrays[f'{name}/ref_pvalue'] = out['pvalue'].to_numpy(np.float64).reshape(P, G)
        assert list(out['phenotype'][:G]) == ['t0'] * G
    # correction='approx-firth': score test on linearly residualised genotypes + Firth refits below the threshold
    firth_cases = {
        'firth': dict(N=120, G=24, P=2, K=3, missing=0.0, offset=False, threshold=0.2),
        'firth_missing_offset': dict(N=150, G=30, P=3, K=2, missing=0.08, offset=True, threshold=0.3),
    }
    for name, c in firth_cases.items():
        N, G, P, K = c['N'], c['G'], c['P'], c['K']
        X = rg.binomial(2, rg.uniform(0.1, 0.5, (G, 1)), (G, N)).astype(np.float64)
        cov = pd.DataFrame(rg.standard_normal((N, K)))
        eta = 0.8 * (X[:P].T - X[:P].T.mean(axis=0)) - 0.5
        Y = (rg.random((N, P)) < 1 / (1 + np.exp(-eta))).astype(np.float64)
        if c['missing']:
            Y[rg.random((N, P)) < c['missing']] = np.nan
        phe = pd.DataFrame(Y, columns=[f't{i}' for i in range(P)])
        off = pd.DataFrame(0.3 * rg.standard_normal((N, P)), columns=phe.columns) if c['offset'] else None
        out, score_p = refshim.reference_firth_test(X, phe, cov, pvalue_threshold=c['threshold'], offset_df=off)
        arrays[f'{name}/X'] = X
        arrays[f'{name}/phenotypes'] = Y
        arrays[f'{name}/covariates'] = cov.to_numpy(np.float64)
        arrays[f'{name}/intercept'] = np.array(1)
        arrays[f'{name}/threshold'] = np.array(c['threshold'])
        if off is not None:
            arrays[f'{name}/offsets'] = off.to_numpy(np.float64)
        for k in ('effect', 'stderror', 'chisq', 'pvalue'):
            arrays[f'{name}/ref_{k}'] = out[k].to_numpy(np.float64).reshape(P, G)
        arrays[f'{name}/ref_score_pvalue'] = score_p.reshape(P, G)
        arrays[f'{name}/ref_succeeded'] = np.array([-1 if v is None else int(v) for v in out['correctionSucceeded']]).reshape(P, G)
        n_corr = int((arrays[f'{name}/ref_succeeded'] >= 0).sum())
        assert 0 < n_corr < P * G, n_corr
        print(f'firth case {name}: {n_corr} of {P * G} tests corrected')
    np.savez_compressed(os.path.join(OUT, 'logreg_cases.npz'), **arrays)
    print('logreg golden cases:', ', '.join(list(cases) + list(firth_cases)))


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == 'logreg':
        golden_logreg()
        sys.exit(0)
    golden_logreg()
    golden_regenie()
    golden_docs()
    golden_doctest()
    golden_cars()
    golden_plink()
    golden_random()
