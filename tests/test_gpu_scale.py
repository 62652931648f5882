"""BASELINE.json shapes on the GPU: spot parity against the oracle on a subset of variants plus
size-independent properties (variant permutation / block-split invariance, duplicated variants)."""
import numpy as np
import pandas as pd
import pytest

from gpu_utils import pack_plink
from helpers import STATS, assert_stats_close
from oracle import linreg_oracle as orc

pytestmark = pytest.mark.gpu


def _engine_from(st, options=None):
    from glow_b200.engine import DeviceState
    ys = list(st.y_state.values()) if isinstance(st.y_state, dict) else [st.y_state]
    return DeviceState(Q=st.Q, Y=np.stack([y.Y for y in ys]), Y_mask=st.Y_mask, YdotY=np.stack([y.YdotY for y in ys]),
                       Y_scale=st.Y_scale, dof=st.dof, options=options)


def _engine_for(phe, cov, off=None, options=None, **kw):
    """DeviceState built from the oracle's own driver-side preparation (lin_reg.py:142-159)."""
    st = orc.prepare_state(phe, cov, off, **kw)
    return _engine_from(st, options), st


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


# ------------------------------------------------------------------------------------------------
# The headline kernel (split8_kernel, tcgen05 int8 digit-split contraction) at the headline shapes, ALL FOUR statistics
# against the oracle.  N = 500 000 is 3 907 stages per tile: the ring parities wrap ~490 times and the int32 digit sums
# reach 1.3e8.
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def ukb_block():
    """4 096 variants x 500 000 samples, PLINK 2-bit, 1 % missing calls (configs[2] shape), with its state."""
    from gpu_utils import gen_packed_gpu
    N, G, K = 500_000, 4096, 16
    rg = np.random.default_rng(20261020)
    packed = gen_packed_gpu(G, N, seed=77, missing=0.01)
    cov = pd.DataFrame(rg.standard_normal((N, K)))
    from gpu_utils import unpack_states
    phe = pd.DataFrame(rg.standard_normal((N, 1)) + 0.02 * unpack_states(packed[:1], N).T.astype(np.float64))
    st = orc.prepare_state(phe, cov)
    return packed, st, N, G


def _subset_ref(packed, st, N, idx, contig=None):
    from gpu_utils import unpack_states
    X = orc.mean_substitute(unpack_states(packed[idx], N).astype(np.float64))
    return orc.block_stats(X, st) if contig is None else orc.block_stats(X, st, contig)


@pytest.mark.parametrize('pair', ['force', 'off'])
def test_split8_headline_shape_all_statistics(ukb_block, pair):
    """N = 500 000, K = 17, P = 1, 1 % missing, mean-imputed: split8 forced (CTA pairs / single CTAs) vs the oracle on a
    scattered subset that includes the tile and pair boundaries."""
    from glow_b200 import _capi
    from glow_b200.engine import GenotypeBlock
    packed, st, N, G = ukb_block
    eng = _engine_from(st, {'split8': 'force', 'split8_pair': pair, 'optimistic_calls': 'off'})
    res = eng.run(GenotypeBlock(packed, 'plink2', 'mean'))
    path = eng.timing(0)['path']
    assert path & _capi.PATH_SPLIT8 and bool(path & _capi.PATH_SPLIT8_PAIR) == (pair == 'force')
    assert not path & (_capi.PATH_PREPASS | _capi.PATH_DMMA)
    rg = np.random.default_rng(5)
    idx = np.unique(np.concatenate([[0, 127, 128, 255, 256, 2047, 2048, G - 129, G - 128, G - 1], rg.choice(G, 14, replace=False)]))
    assert_stats_close({k: res[k][:, idx] for k in STATS}, _subset_ref(packed, st, N, idx), label=f'split8 pair={pair} N=500k')
    assert np.isfinite(res['pvalue']).all()
    eng.close()


def test_default_knobs_reach_split8_at_headline_shape(ukb_block):
    """With default knobs the 4 096-variant block takes the int8 tensor-core path (no env, no options), a 1 024-variant
    block takes it too -- with the sample range split across CTAs -- and both agree with the oracle; the integer partial
    sums make the split form bit-identical to the unsplit one."""
    from glow_b200 import _capi
    from glow_b200.engine import GenotypeBlock
    packed, st, N, G = ukb_block
    eng = _engine_from(st)
    res = eng.run(GenotypeBlock(packed, 'plink2', 'mean'))
    assert eng.timing(0)['path'] & _capi.PATH_SPLIT8
    idx = np.array([1, 130, 1999, 4000])
    assert_stats_close({k: res[k][:, idx] for k in STATS}, _subset_ref(packed, st, N, idx), label='default knobs N=500k')
    small = eng.run(GenotypeBlock(packed[1000:2024], 'plink2', 'mean'))
    path = eng.timing(0)['path']
    assert path & _capi.PATH_SPLIT8 and path & _capi.PATH_SPLIT8_NSPLIT, path
    for k in STATS:
        assert np.array_equal(small[k], res[k][:, 1000:2024], equal_nan=True), k
    # every split factor gives the same bits
    for ns in (1, 3, 7):
        e2 = _engine_from(st, {'split8': 'force', 'sample_split': ns})
        r2 = e2.run(GenotypeBlock(packed[1000:1300], 'plink2', 'mean'))
        for k in STATS:
            assert np.array_equal(r2[k], res[k][:, 1000:1300], equal_nan=True), (k, ns)
        e2.close()
    eng.close()


def test_optimistic_form_without_missing_calls(ukb_block):
    """Fully called blocks: after one block without a missing call the next ones run without the missing-call operand
    (half the tensor work); a block that does have missing calls is detected in the kernel and re-run.  Identical bits
    either way."""
    from glow_b200 import _capi
    from glow_b200.engine import GenotypeBlock
    from gpu_utils import gen_packed_gpu
    packed_m, st, N, G = ukb_block
    full = gen_packed_gpu(1024, N, seed=99, missing=0.0)
    ref_eng = _engine_from(st, {'split8': 'force', 'optimistic_calls': 'off'})
    want_full = ref_eng.run(GenotypeBlock(full, 'plink2', 'mean'))
    want_miss = ref_eng.run(GenotypeBlock(packed_m[:1024], 'plink2', 'mean'))
    ref_eng.close()
    eng = _engine_from(st, {'split8': 'force'})
    r1 = eng.run(GenotypeBlock(full, 'plink2', 'mean'))
    assert not eng.timing(0)['path'] & _capi.PATH_OPTIMISTIC  # nothing known about the stream yet
    r2 = eng.run(GenotypeBlock(full, 'plink2', 'mean'))
    assert eng.timing(0)['path'] & _capi.PATH_OPTIMISTIC
    r3 = eng.run(GenotypeBlock(packed_m[:1024], 'plink2', 'mean'))  # has missing calls: optimistic run is void, re-run
    assert eng.timing(0)['path'] & _capi.PATH_REPLAYED
    r4 = eng.run(GenotypeBlock(full, 'plink2', 'mean'))
    assert not eng.timing(0)['path'] & _capi.PATH_OPTIMISTIC  # the stream has shown missing calls
    for k in STATS:
        assert np.array_equal(r1[k], want_full[k], equal_nan=True), k
        assert np.array_equal(r2[k], want_full[k], equal_nan=True), k
        assert np.array_equal(r3[k], want_miss[k], equal_nan=True), k
        assert np.array_equal(r4[k], want_full[k], equal_nan=True), k
    idx = np.array([0, 500, 1023])
    assert_stats_close({k: r2[k][:, idx] for k in STATS}, _subset_ref(full, st, N, idx), label='optimistic form')
    eng.close()


@pytest.mark.parametrize('pair', ['force', 'off'])
def test_split8_loco_200k_all_statistics(pair):
    """configs[4] shape: 200 000 samples x 20 phenotypes, K = 17, LOCO offsets for 3 contigs, split8 forced."""
    from glow_b200 import _capi
    from glow_b200.engine import GenotypeBlock
    from gpu_utils import gen_packed_gpu
    rg = np.random.default_rng(11)
    N, G, P, K = 200_000, 3 * 640, 20, 16
    contigs = ['1', '2', '3']
    cov = pd.DataFrame(rg.standard_normal((N, K)))
    phe = pd.DataFrame(rg.standard_normal((N, P)))
    off = pd.DataFrame(0.1 * rg.standard_normal((N * 3, P)), index=pd.MultiIndex.from_product([phe.index, contigs]))
    packed = gen_packed_gpu(G, N, seed=5, missing=0.01)
    eng, st = _engine_for(phe, cov, off, options={'split8': 'force', 'split8_pair': pair})
    for ci, c in enumerate(contigs):
        blk = packed[ci * 640:(ci + 1) * 640]
        res = eng.run(GenotypeBlock(blk, 'plink2', 'mean'), contig=ci)
        assert eng.timing(0)['path'] & _capi.PATH_SPLIT8
        idx = np.array([0, 127, 128, 300, 639])
        assert_stats_close({k: res[k][:, idx] for k in STATS}, _subset_ref(blk, st, N, idx, c), label=f'split8 LOCO contig {c}')
    eng.close()


@pytest.mark.parametrize('form', ['split8_pair', 'split8_single', 'dmma'])
def test_adversarial_dynamic_range(form):
    """Pins the quantisation bound of the digit split (2^-54 of each column's maximum): a one-hot covariate, a covariate
    with a 1e6 x outlier, a heavy-tailed phenotype (Student t, 1.5 degrees of freedom) and heavy-tailed LOCO offsets --
    columns whose maximum is far above their typical entry.  Every variant against the oracle at the north-star
    tolerances."""
    from glow_b200.engine import GenotypeBlock
    rg = np.random.default_rng(3)
    N, G, P = 65_537, 384, 3
    cov = rg.standard_normal((N, 6))
    cov[:, 0] = 0.0
    cov[12_345, 0] = 1.0                      # one-hot covariate
    cov[777, 1] = 1e6                         # outlier 1e6 x the column's scale
    cov[:, 2] = rg.standard_t(1.5, N)         # heavy-tailed covariate
    phe = rg.standard_normal((N, P))
    phe[:, 0] = rg.standard_t(1.5, N)
    phe[4242, 1] = 1e5
    contigs = ['1', '2']
    off = pd.DataFrame(rg.standard_t(1.5, (N * 2, P)) * 0.05, index=pd.MultiIndex.from_product([range(N), contigs]))
    f = rg.uniform(0.001, 0.5, (G, 1))
    states = rg.binomial(2, f, (G, N)).astype(np.int8)
    states[rg.random((G, N), dtype=np.float32) < 0.01] = -1
    states[:, 12_345] = 2                     # the one-hot sample carries the allele in every variant
    states[:3, 777] = [0, 1, 2]
    opts = {'split8_pair': {'split8': 'force', 'split8_pair': 'force'}, 'split8_single': {'split8': 'force', 'split8_pair': 'off'},
            'dmma': {'split8': 'off'}}[form]
    eng, st = _engine_for(pd.DataFrame(phe), pd.DataFrame(cov), off, options=opts)
    X = orc.mean_substitute(states.astype(np.float64))
    packed = pack_plink(states)
    for ci, c in enumerate(contigs):
        res = eng.run(GenotypeBlock(packed, 'plink2', 'mean'), contig=ci)
        assert_stats_close(res, orc.block_stats(X, st, c), label=f'adversarial range, {form}, contig {c}')
    eng.close()


def test_float64_blocks_are_reproducible_with_short_lived_ctas(rg):
    """Regression for a cross-proxy hazard found in round 2: stages of the shared operand are read with ordinary loads and
    refilled by cp.async.bulk; without a proxy fence before the release a refill occasionally overtook the last reads --
    wrong rows for one warp's 16 variants, a few per 10^4 warps, only with short-lived CTAs of the float64 kernels (few
    samples or a split sample range, many variants) and FMA tail columns.  The same block must give the same bits on every
    run, and permuted variants permuted results."""
    from glow_b200.engine import GenotypeBlock
    N, G, P, K = 3328, 60_000, 7, 10
    cov = pd.DataFrame(rg.standard_normal((N, K)))
    X = rg.integers(0, 3, (G, N)).astype(np.float64) + 0.1 * rg.random((G, N), dtype=np.float32)
    phe = pd.DataFrame(rg.standard_normal((N, P)))
    for opts in ({}, {'sample_split': 3}):
        eng, st = _engine_for(phe, cov, options=opts)
        first = eng.run(GenotypeBlock(X, 'f64'))
        for _ in range(5):
            again = eng.run(GenotypeBlock(X, 'f64'))
            for k in STATS:
                assert np.array_equal(again[k], first[k], equal_nan=True), (k, opts)
        idx = np.sort(rg.choice(G, 64, replace=False))
        assert_stats_close({k: first[k][:, idx] for k in STATS}, orc.block_stats(X[idx], st), label='short CTAs')
        eng.close()


def _max_rel(a, b):
    ok = np.isfinite(a) & np.isfinite(b) & (b != 0)
    return float(np.max(np.abs(a[ok] - b[ok]) / np.abs(b[ok]))) if ok.any() else 0.0


@pytest.mark.parametrize('missing_phenotype', [False, True])
def test_reduced_digit_covariates_are_certified(missing_phenotype):
    """The covariate columns enter the int8 contraction with 5 digits (6 when a masked pass follows) instead of 7.  Every
    test carries a certified bound of the error this causes in XdotX (aux.cu: xx_full_kernel); the default run must take
    the reduced image, stay within 1e-11 of the 7-digit results, and a run whose certificate fails (3 digits, test only)
    must be re-run on the 7-digit image and give its bits."""
    from glow_b200 import _capi
    from glow_b200.engine import GenotypeBlock
    from gpu_utils import gen_packed_gpu, unpack_states
    rg = np.random.default_rng(8)
    N, G, K, P = 200_000, 1024, 16, 2
    cov = rg.standard_normal((N, K))
    cov[:, 3] = rg.random(N) < 0.002            # rare binary covariate: a column whose maximum is far above its rms
    packed = gen_packed_gpu(G, N, seed=21, missing=0.01)
    Y = rg.standard_normal((N, P)) + 0.02 * unpack_states(packed[:P], N).T
    if missing_phenotype:
        Y[rg.random((N, P)) < 0.15] = np.nan
    st = orc.prepare_state(pd.DataFrame(Y), pd.DataFrame(cov))
    blk = GenotypeBlock(packed, 'plink2', 'mean')
    e7 = _engine_from(st, {'split8': 'force', 'q_digits': 7})
    want = e7.run(blk)
    assert not e7.timing(0)['path'] & _capi.PATH_Q_DIGITS_LO
    e7.close()
    eng = _engine_from(st, {'split8': 'force'})
    got = eng.run(blk)
    path = eng.timing(0)['path']
    assert path & _capi.PATH_Q_DIGITS_LO and not path & _capi.PATH_REPLAYED, path
    eng.close()
    for k, tol in (('effect', 1e-11), ('stderror', 1e-11), ('tvalue', 1e-11), ('pvalue', 1e-9)):
        assert _max_rel(got[k], want[k]) <= tol, (k, _max_rel(got[k], want[k]))
    idx = np.array([0, 1, 511, 1023])
    assert_stats_close({k: got[k][:, idx] for k in STATS}, _subset_ref(packed, st, N, idx), label='reduced digits')
    # failed certificate -> same block again on the 7-digit image; the state then stays on it for a while
    e3 = _engine_from(st, {'split8': 'force', 'q_digits': 3})
    r3 = e3.run(blk)
    path = e3.timing(0)['path']
    assert path & _capi.PATH_REPLAYED and not path & _capi.PATH_Q_DIGITS_LO, path
    r3b = e3.run(blk)
    assert not e3.timing(0)['path'] & (_capi.PATH_REPLAYED | _capi.PATH_Q_DIGITS_LO)
    e3.close()
    for k in STATS:
        assert np.array_equal(r3[k], want[k], equal_nan=True), k
        assert np.array_equal(r3b[k], want[k], equal_nan=True), k
