"""BASELINE.json shapes on the GPU: spot parity against the oracle on a subset of variants plus
size-independent properties (variant permutation / block-split invariance, duplicated variants)."""
import numpy as np
import pandas as pd
import pytest

from gpu_utils import pack_plink
from helpers import STATS, assert_stats_close
from oracle import linreg_oracle as orc

pytestmark = pytest.mark.gpu


def _engine_for(phe, cov, off=None, **kw):
    """DeviceState built from the oracle's own driver-side preparation (lin_reg.py:142-159)."""
    from glow_b200.engine import DeviceState
    st = orc.prepare_state(phe, cov, off, **kw)
    ys = list(st.y_state.values()) if isinstance(st.y_state, dict) else [st.y_state]
    eng = DeviceState(Q=st.Q, Y=np.stack([y.Y for y in ys]), Y_mask=st.Y_mask, YdotY=np.stack([y.YdotY for y in ys]),
                      Y_scale=st.Y_scale, dof=st.dof)
    return eng, st


def _hardcalls(rg, G, N, missing=0.01):
    f = rg.uniform(0.01, 0.5, (G, 1))
    u = rg.random((G, N), dtype=np.float32)
    s = (u > (1 - f)**2).astype(np.int8) + (u > (1 - f)**2 + 2 * f * (1 - f)).astype(np.int8)
    s[rg.random((G, N), dtype=np.float32) < missing] = -1
    return s


def test_config2_full_size_dosages(rg):
    """configs[1]: 10k samples x 100k variants x 10 phenotypes x 10 covariates, fp64 dosages."""
    from glow_b200.engine import GenotypeBlock
    N, G, P, K = 10_000, 100_000, 10, 10
    cov = pd.DataFrame(rg.standard_normal((N, K)))
    X = np.empty((G, N), dtype=np.float64)
    for g0 in range(0, G, 10_000):
        X[g0:g0 + 10_000] = _hardcalls(rg, 10_000, N, missing=0.0)
        X[g0:g0 + 10_000] += 0.1 * rg.random((10_000, N), dtype=np.float32)  # dense, non-integer dosages
    phe = pd.DataFrame(rg.standard_normal((N, P)) + 0.1 * (X[:5].T @ rg.standard_normal((5, P))))
    eng, st = _engine_for(phe, cov)
    res = eng.run(GenotypeBlock(X, 'f64'))
    # spot parity on a scattered subset of variants
    idx = np.sort(rg.choice(G, 256, replace=False))
    ref = orc.block_stats(X[idx], st)
    assert_stats_close({k: res[k][:, idx] for k in STATS}, ref, label='cfg2 subset')
    # variant permutation -> identically permuted outputs (indexing / ordering is bit exact)
    perm = rg.permutation(G)
    res_p = eng.run(GenotypeBlock(np.ascontiguousarray(X[perm]), 'f64'))
    for k in STATS:
        assert np.array_equal(res_p[k], res[k][:, perm], equal_nan=True), k
    # the strong-signal variants are found
    assert (res['pvalue'][:, :5] < 1e-6).any()
    eng.close()


def test_config3_shape_plink_subset(rg):
    """configs[2] shape: 500k samples, 16 covariates + intercept, 1 phenotype, PLINK 2-bit with 1 % missing."""
    from glow_b200.engine import GenotypeBlock
    N, G, K = 500_000, 1024, 16
    cov = pd.DataFrame(rg.standard_normal((N, K)))
    states = _hardcalls(rg, G, N)
    phe = pd.DataFrame(rg.standard_normal((N, 1)) + 0.02 * states[:1].T)
    packed = pack_plink(states)
    eng, st = _engine_for(phe, cov)
    res = eng.run(GenotypeBlock(packed, 'plink2', 'mean'))
    idx = np.sort(rg.choice(G, 12, replace=False))
    idx[0] = 0
    X = orc.mean_substitute(states[idx].astype(np.float64))
    ref = orc.block_stats(X, st)
    assert_stats_close({k: res[k][:, idx] for k in STATS}, ref, label='cfg3 subset')
    # block-split invariance: two half blocks == one block (different CTA split of the sample range)
    a = eng.run(GenotypeBlock(packed[:300], 'plink2', 'mean'))
    b = eng.run(GenotypeBlock(packed[300:], 'plink2', 'mean'))
    for k in STATS:
        np.testing.assert_allclose(np.concatenate([a[k], b[k]], axis=1), res[k], rtol=1e-10, atol=1e-13)
    # a duplicated variant gives identical statistics wherever it sits in the block
    dup = packed.copy()
    dup[777] = dup[3]
    r2 = eng.run(GenotypeBlock(dup, 'plink2', 'mean'))
    for k in STATS:
        assert r2[k][0, 777] == r2[k][0, 3]
    eng.close()


def test_config3_shape_with_missing_phenotype_values(rg, monkeypatch):
    """configs[2] shape with what phenotypes look like in practice: 2 % of the values unobserved.  The three forms of the
    masked pass (sparse complement = the default here, FP64 DMMA, int8 tensor cores) must agree with the oracle on a subset
    and with one another on the whole block."""
    from glow_b200.engine import GenotypeBlock
    N, G, K, P = 500_000, 512, 16, 2
    cov = pd.DataFrame(rg.standard_normal((N, K)))
    states = _hardcalls(rg, G, N)
    Y = rg.standard_normal((N, P)) + 0.02 * states[:P].T
    Y[rg.random((N, P)) < 0.02] = np.nan
    phe = pd.DataFrame(Y)
    packed = pack_plink(states)
    results = {}
    for form, env in (('sparse', {}), ('dmma', {'GLR_MASK_SPARSE': '0', 'GLR_MASK8': '0'}), ('int8', {'GLR_MASK8': '2'})):
        for k in ('GLR_MASK_SPARSE', 'GLR_MASK8'):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        eng, st = _engine_for(phe, cov)
        assert eng.unique_masks == P
        results[form] = eng.run(GenotypeBlock(packed, 'plink2', 'mean'))
        eng.close()
    idx = np.sort(rg.choice(G, 8, replace=False))
    ref = orc.block_stats(orc.mean_substitute(states[idx].astype(np.float64)), st)
    for form, res in results.items():
        assert_stats_close({k: res[k][:, idx] for k in STATS}, ref, label=f'cfg3 missing phenotype, {form} form')
    for form in ('dmma', 'int8'):
        assert_stats_close(results[form], results['sparse'], label=f'{form} vs sparse form')


def test_config4_shape_many_masked_phenotypes(rg):
    """configs[3] shape scaled in G: 50k samples x 2 000 phenotypes with 10 % missing each, 2-bit genotypes."""
    from glow_b200.engine import GenotypeBlock
    N, G, P, K = 50_000, 512, 2_000, 10
    cov = pd.DataFrame(rg.standard_normal((N, K)))
    states = _hardcalls(rg, G, N)
    Y = rg.standard_normal((N, P))
    Y[rg.random((N, P), dtype=np.float32) < 0.1] = np.nan
    phe = pd.DataFrame(Y)
    packed = pack_plink(states)
    eng, st = _engine_for(phe, cov)
    assert eng.unique_masks == P
    res = eng.run(GenotypeBlock(packed, 'plink2', 'mean'))
    idx = np.sort(rg.choice(G, 8, replace=False))
    ref = orc.block_stats(orc.mean_substitute(states[idx].astype(np.float64)), st)
    assert_stats_close({k: res[k][:, idx] for k in STATS}, ref, label='cfg4 subset')
    eng.close()


def test_config5_shape_loco(rg):
    """configs[4] shape scaled: 200k samples x 20 phenotypes, per-contig LOCO offsets (3 of 22 contigs)."""
    from glow_b200.engine import GenotypeBlock
    N, G, P, K = 200_000, 300, 20, 16
    contigs = ['1', '2', '3']
    cov = pd.DataFrame(rg.standard_normal((N, K)))
    states = _hardcalls(rg, G, N)
    phe = pd.DataFrame(rg.standard_normal((N, P)))
    off = pd.DataFrame(0.1 * rg.standard_normal((N * 3, P)), index=pd.MultiIndex.from_product([phe.index, contigs]))
    packed = pack_plink(states)
    eng, st = _engine_for(phe, cov, off)
    for ci, c in enumerate(contigs):
        res = eng.run(GenotypeBlock(packed[ci * 100:(ci + 1) * 100], 'plink2', 'mean'), contig=ci)
        idx = np.array([0, 57, 99])
        X = orc.mean_substitute(states[ci * 100 + idx].astype(np.float64))
        ref = orc.block_stats(X, st, c)
        assert_stats_close({k: res[k][:, idx] for k in STATS}, ref, label=f'cfg5 contig {c}')
    eng.close()
