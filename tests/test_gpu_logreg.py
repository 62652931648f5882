"""Logistic-regression score test on the GPU against the oracle and the fixtures the
unmodified reference generated: every genotype format, missing phenotype values, offsets (single and LOCO),
intersect_samples, no intercept, many covariates, ragged sizes.  Tolerances: the north-star's rel 1e-9 on the
statistic and rel 1e-6 (abs 1e-300) on the p-value."""
import numpy as np
import pandas as pd
import pytest

from gpu_utils import pack_plink
from oracle import logreg_oracle as lo
from oracle import linreg_oracle as orc
from test_logreg_oracle_pins import case_frames, load_cases

pytestmark = pytest.mark.gpu


def _close(got, ref, label, rtol=1e-9):
    for k, rt, at in (('chisq', rtol, 1e-13), ('pvalue', max(rtol, 1e-6), 1e-300)):
        g, r = np.asarray(got[k], dtype=np.float64), np.asarray(ref[k], dtype=np.float64)
        assert g.shape == r.shape, (label, k)
        assert np.array_equal(np.isnan(g), np.isnan(r)), (label, k, 'NaN pattern')
        ok = ~np.isnan(r)
        err = np.abs(g[ok] - r[ok])
        tol = at + rt * np.abs(r[ok])
        assert (err <= tol).all(), f'{label} {k}: worst rel {np.max(err / np.maximum(np.abs(r[ok]), 1e-300)):.3e}'


@pytest.mark.parametrize('name', ['multiple', 'multiple_missing', 'no_intercept', 'simple_offset', 'hardcalls'])
def test_reference_fixtures(name):
    import glow_b200
    c = load_cases()[name]
    phe, cov, off, icpt = case_frames(c)
    G, P = c['X'].shape[0], phe.shape[1]
    pdf = pd.DataFrame({'values': list(c['X']), 'idx': np.arange(G)})
    out = glow_b200.gwas.logistic_regression(pdf, phe, cov, off if off is not None else pd.DataFrame({}), correction='none',
                                             add_intercept=icpt)
    assert out.columns.tolist() == ['idx', 'chisq', 'pvalue', 'phenotype']
    assert out['phenotype'].tolist() == [p for p in phe.columns for _ in range(G)]
    assert out['idx'].tolist() == list(range(G)) * P
    _close({k: out[k].to_numpy().reshape(P, G) for k in ('chisq', 'pvalue')}, {'chisq': c['ref_chisq'], 'pvalue': c['ref_pvalue']},
           f'fixture {name}')


@pytest.mark.parametrize('N,G,P,K', [(200, 33, 3, 4), (1000, 300, 1, 16), (515, 129, 11, 3), (4099, 260, 20, 16), (130, 9, 2, 40)])
def test_all_formats_against_oracle(rg, N, G, P, K):
    from glow_b200.gwas.log_reg import LogRegDeviceState
    from glow_b200.engine import GenotypeBlock
    states = rg.binomial(2, rg.uniform(0.05, 0.5, (G, 1)), (G, N)).astype(np.int64)
    states[:, 0] = [g % 3 for g in range(G)]
    miss = states.copy()
    miss[rg.random((G, N)) < 0.02] = -1
    cov = pd.DataFrame(rg.standard_normal((N, K)) * 0.3)
    Y = (rg.random((N, P)) < 0.4).astype(np.float64)
    Y[rg.random((N, P)) < 0.07] = np.nan
    phe = pd.DataFrame(Y)
    C, mask, st = lo.prepare_state(phe, cov)
    eng = LogRegDeviceState(C, [st])
    X = states.astype(np.float64) + 0.25 * rg.random((G, N))
    _close(eng.run(GenotypeBlock(X, 'f64')), lo.block_score_test(X, C, mask, st), 'f64')
    X32 = X.astype(np.float32)
    _close(eng.run(GenotypeBlock(X32, 'f32')), lo.block_score_test(X32.astype(np.float64), C, mask, st), 'f32')
    Xm = orc.mean_substitute(miss.astype(np.float64))
    ref_mean, ref_raw = lo.block_score_test(Xm, C, mask, st), lo.block_score_test(miss.astype(np.float64), C, mask, st)
    _close(eng.run(GenotypeBlock(pack_plink(miss), 'plink2', 'mean')), ref_mean, 'plink2 mean')
    _close(eng.run(GenotypeBlock(pack_plink(miss), 'plink2', 'minus_one')), ref_raw, 'plink2 raw')
    _close(eng.run(GenotypeBlock(miss.astype(np.int8), 'i8', 'mean')), ref_mean, 'i8 mean')
    eng.close()


@pytest.mark.parametrize('N', [320, 1500])  # below / above the sample count from which 2-bit blocks take the int8 path
def test_loco_offsets_intersect_and_sources(rg, N):
    import glow_b200
    G, P, K = 60, 2, 3
    samples = [f's{i}' for i in range(N)]
    states = rg.integers(0, 3, (G, N))
    states[rg.random((G, N)) < 0.03] = -1
    keep = [i for i in range(N) if i % 9 != 4]
    Y = (rg.random((N, P)) < 0.5).astype(np.float64)
    Y[rg.random((N, P)) < 0.05] = np.nan
    phe = pd.DataFrame(Y, index=samples, columns=['y1', 'y2']).iloc[keep]
    cov = pd.DataFrame(rg.standard_normal((N, K)), index=samples)
    contigs = ['1', '2']
    off = pd.DataFrame(0.3 * rg.standard_normal((len(keep) * 2, P)), index=pd.MultiIndex.from_product([phe.index, contigs]),
                       columns=phe.columns)
    meta = pd.DataFrame({'contigName': [contigs[(g // 7) % 2] for g in range(G)], 'idx': np.arange(G)})
    src = glow_b200.PlinkBedSource(bytes([0x6c, 0x1b, 0x01]) + pack_plink(states).tobytes(), N, meta, mean_substitute=True,
                                   block_variants=25)
    out = glow_b200.gwas.logistic_regression(src, phe, cov, off, correction='none', intersect_samples=True,
                                             genotype_sample_ids=samples)
    X = orc.mean_substitute(states.astype(np.float64))
    C, mask, st = lo.prepare_state(phe, cov.loc[phe.index], off)
    Xk = X[:, keep]
    for c in contigs:
        sel = (meta['contigName'] == c).to_numpy()
        ref = lo.block_score_test(Xk[sel], C, mask, st[c])
        for p, name in enumerate(phe.columns):
            rows = out[(out['contigName'] == c) & (out['phenotype'] == name)].sort_values('idx')
            assert rows['idx'].tolist() == meta['idx'][sel].tolist()
            _close({'chisq': rows['chisq'].to_numpy(), 'pvalue': rows['pvalue'].to_numpy()},
                   {'chisq': ref['chisq'][p], 'pvalue': ref['pvalue'][p]}, f'LOCO contig {c} {name}')
    # the same through pandas blocks (float64 arrays, the reference's own input shape)
    pdf = meta.copy()
    pdf['values'] = list(X)
    out2 = glow_b200.gwas.logistic_regression(iter([pdf.iloc[:31], pdf.iloc[31:]]), phe, cov, off, correction='none',
                                              intersect_samples=True, genotype_sample_ids=samples)
    key = lambda d: d.sort_values(['phenotype', 'idx'], kind='stable').reset_index(drop=True)
    a, b = key(out), key(out2)
    _close({k: a[k].to_numpy() for k in ('chisq', 'pvalue')}, {k: b[k].to_numpy() for k in ('chisq', 'pvalue')}, 'bed vs pandas')


def test_c_abi_logreg_errors(rg):
    import ctypes as C
    from glow_b200 import _capi
    lib = _capi.load()
    h = C.c_void_p()
    d = _capi.LogRegDesc(abi_version=_capi.GLR_ABI_VERSION, device=0, n_samples=10, n_geno_samples=10, n_covariates=0, n_phenotypes=1,
                         n_contigs=1)
    assert lib.glr_logreg_state_create(C.byref(d), C.byref(h)) == _capi.ERR_INVALID
    assert b'dimensions' in lib.glr_last_error()
    assert lib.glr_logreg_state_destroy(None) == 0


@pytest.mark.parametrize('name', ['firth', 'firth_missing_offset'])
def test_approx_firth_reference_fixtures(name):
    """correction='approx-firth' (the reference's default): score test on linearly residualised genotypes on the device,
    Firth refits of the tests below the threshold on the host; fixtures from the reference's log_reg.py + approx_firth.py."""
    import glow_b200
    c = load_cases()[name]
    phe, cov, off, icpt = case_frames(c)
    G, P = c['X'].shape[0], phe.shape[1]
    pdf = pd.DataFrame({'values': list(c['X']), 'idx': np.arange(G)})
    out = glow_b200.gwas.logistic_regression(pdf, phe, cov, off if off is not None else pd.DataFrame({}),
                                             pvalue_threshold=float(c['threshold']), add_intercept=icpt)
    assert out.columns.tolist() == ['idx', 'effect', 'stderror', 'correctionSucceeded', 'chisq', 'pvalue', 'phenotype']
    got_ok = np.array([-1 if v is None else int(v) for v in out['correctionSucceeded']]).reshape(P, G)
    assert np.array_equal(got_ok, c['ref_succeeded'])
    for k, rtol in (('chisq', 1e-8), ('pvalue', 1e-6), ('effect', 1e-8), ('stderror', 1e-8)):
        g, r = out[k].to_numpy(np.float64).reshape(P, G), c[f'ref_{k}']
        assert np.array_equal(np.isnan(g), np.isnan(r)), k
        ok = ~np.isnan(r)
        np.testing.assert_allclose(g[ok], r[ok], rtol=rtol, atol=1e-13, err_msg=k)


def test_approx_firth_on_packed_blocks_and_linear_residualisation(rg):
    """The score test of the approx-firth mode (genotypes residualised linearly first, folded into the device epilogue)
    against the oracle fed with explicitly residualised genotypes, on 2-bit and int8 blocks; the Firth refit decodes
    the flagged rows on the host.  Device-only blocks are refused."""
    import glow_b200
    from glow_b200.gwas.log_reg import LogRegDeviceState
    from glow_b200.engine import GenotypeBlock
    N, G, P, K = 700, 90, 2, 5
    states = rg.binomial(2, rg.uniform(0.1, 0.5, (G, 1)), (G, N)).astype(np.int64)
    states[rg.random((G, N)) < 0.02] = -1
    cov = pd.DataFrame(rg.standard_normal((N, K)) * 0.5)
    Y = (rg.random((N, P)) < 0.35).astype(np.float64)
    Y[rg.random((N, P)) < 0.05] = np.nan
    phe = pd.DataFrame(Y, columns=['a', 'b'])
    C, mask, st = lo.prepare_state(phe, cov)
    Q = np.linalg.qr(C)[0]
    X = orc.mean_substitute(states.astype(np.float64))
    Xl = (X.T - Q @ (Q.T @ X.T)).T
    ref = lo.block_score_test(Xl, C, mask, st)
    eng = LogRegDeviceState(C, [st], Q=Q)
    _close(eng.run(GenotypeBlock(pack_plink(states), 'plink2', 'mean')), ref, 'lin-residualised plink2')
    _close(eng.run(GenotypeBlock(states.astype(np.int8), 'i8', 'mean')), ref, 'lin-residualised i8')
    _close(eng.run(GenotypeBlock(X, 'f64')), ref, 'lin-residualised f64')
    eng.close()
    bed = bytes([0x6c, 0x1b, 0x01]) + pack_plink(states).tobytes()
    meta = pd.DataFrame({'idx': np.arange(G)})
    a = glow_b200.gwas.logistic_regression(glow_b200.PlinkBedSource(bed, N, meta, block_variants=40), phe, cov, pvalue_threshold=0.2)
    pdf = meta.copy()
    pdf['values'] = list(X)
    b = glow_b200.gwas.logistic_regression(pdf, phe, cov, pvalue_threshold=0.2)
    key = lambda d: d.sort_values(['phenotype', 'idx'], kind='stable').reset_index(drop=True)
    a, b = key(a), key(b)
    assert a['correctionSucceeded'].tolist() == b['correctionSucceeded'].tolist() and a['correctionSucceeded'].notna().sum() > 0
    for k in ('effect', 'stderror', 'chisq', 'pvalue'):
        np.testing.assert_allclose(a[k].to_numpy(np.float64), b[k].to_numpy(np.float64), rtol=1e-7, atol=1e-13, equal_nan=True)
    res = glow_b200.ResidentSource(glow_b200.PlinkBedSource(bed, N, meta), devices=[0])
    with pytest.raises(NotImplementedError, match='host'):
        glow_b200.gwas.logistic_regression(res, phe, cov, pvalue_threshold=0.9)
    res.close()


def test_int8_tensor_path_of_the_score_test(rg, monkeypatch):
    """2-bit blocks run both contractions of the score test on the int8 tensor cores (pass 2 = the squared-call table with
    mu^2 as the missing substitute).  At 100 000 samples: against the oracle on a subset, and against the FP64 DMMA path
    (GLR_LOGREG_SPLIT8=0) on the whole block, both missing policies, a small block (sample range split) included."""
    from glow_b200.gwas.log_reg import LogRegDeviceState
    from glow_b200.engine import GenotypeBlock
    N, G, P, K = 100_000, 700, 2, 8
    f = rg.uniform(0.02, 0.5, (G, 1))
    u = rg.random((G, N), dtype=np.float32)
    states = (u > (1 - f)**2).astype(np.int8) + (u > (1 - f)**2 + 2 * f * (1 - f)).astype(np.int8)
    states[rg.random((G, N), dtype=np.float32) < 0.01] = -1
    states[5] = -1                                  # an all-missing variant
    states[6] = 2                                   # a constant one
    cov = pd.DataFrame(rg.standard_normal((N, K)) * 0.3)
    Y = (rg.random((N, P)) < 0.3).astype(np.float64)
    Y[rg.random((N, P)) < 0.05] = np.nan
    C, mask, st = lo.prepare_state(pd.DataFrame(Y), cov)
    packed = pack_plink(states.astype(np.int64))
    eng = LogRegDeviceState(C, [st])
    monkeypatch.setenv('GLR_LOGREG_SPLIT8', '0')
    eng_dmma = LogRegDeviceState(C, [st])
    monkeypatch.delenv('GLR_LOGREG_SPLIT8')
    idx = np.array([0, 1, 127, 128, 300, 699])
    for policy in ('mean', 'minus_one'):
        got = eng.run(GenotypeBlock(packed, 'plink2', policy))
        want = eng_dmma.run(GenotypeBlock(packed, 'plink2', policy))
        ok = np.isfinite(want['chisq'])
        assert np.array_equal(np.isfinite(got['chisq']), ok)
        np.testing.assert_allclose(got['chisq'][ok], want['chisq'][ok], rtol=1e-9, atol=1e-13)
        np.testing.assert_allclose(got['pvalue'][ok], want['pvalue'][ok], rtol=1e-6, atol=1e-300)
        X = states[idx].astype(np.float64)
        ref = lo.block_score_test(orc.mean_substitute(X) if policy == 'mean' else X, C, mask, st)
        _close({k: v[:, idx] for k, v in got.items()}, ref, f'int8 path {policy}')
        small = eng.run(GenotypeBlock(packed[:100], 'plink2', policy))
        for k in ('chisq', 'pvalue'):
            assert np.array_equal(small[k], got[k][:, :100], equal_nan=True), k
        # the same calls as int8 genotype states: same kernels, same bits
        got8 = eng.run(GenotypeBlock(states, 'i8', policy))
        for k in ('chisq', 'pvalue'):
            assert np.array_equal(got8[k], got[k], equal_nan=True), k
    # int8 values that are not genotype states: detected in the kernel, the block is run again on the FP64 path
    wild = states[7:71].copy()                      # (not the degenerate rows 5 and 6)
    wild[3, 10] = 3
    wild[63, N - 1] = -2
    ref = lo.block_score_test(wild.astype(np.float64), C, mask, st)
    _close(eng.run(GenotypeBlock(wild, 'i8', 'minus_one')), ref, 'int8 values outside the state range')
    eng.close()
    eng_dmma.close()
