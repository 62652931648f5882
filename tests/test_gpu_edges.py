"""Edge cases of the CUDA path through the C ABI: tiny and ragged shapes, strides and alignment,
device-resident buffers, asynchronous lanes, error codes."""
import ctypes as C

import numpy as np
import pandas as pd
import pytest

from gpu_utils import pack_plink
from helpers import STATS, assert_stats_close
from oracle import linreg_oracle as orc

pytestmark = pytest.mark.gpu


def _state(phe, cov, **kw):
    from glow_b200.engine import DeviceState
    st = orc.prepare_state(phe, cov, **kw)
    ys = list(st.y_state.values()) if isinstance(st.y_state, dict) else [st.y_state]
    eng = DeviceState(Q=st.Q, Y=np.stack([y.Y for y in ys]), Y_mask=st.Y_mask, YdotY=np.stack([y.YdotY for y in ys]),
                      Y_scale=st.Y_scale, dof=st.dof)
    return eng, st


def _data(rg, N, G, P, K):
    X = rg.integers(0, 3, (G, N)).astype(np.float64)
    X[:, 0] = [g % 3 for g in range(G)]  # no constant rows even for tiny N
    X[:, -1] = [(g + 1) % 3 for g in range(G)]
    cov = pd.DataFrame(rg.standard_normal((N, K))) if K else None
    phe = pd.DataFrame(rg.standard_normal((N, P)))
    return X, phe, cov


@pytest.mark.parametrize('N', [5, 7, 8, 9, 31, 63, 64, 65, 127, 129])
def test_tiny_and_ragged_sample_counts(rg, N):
    from glow_b200.engine import GenotypeBlock
    G, P, K = 11, 2, 1
    X, phe, cov = _data(rg, N, G, P, K)
    eng, st = _state(phe, cov)
    ref = orc.block_stats(X, st)
    for blk in (GenotypeBlock(X, 'f64'), GenotypeBlock(X.astype(np.float32), 'f32'), GenotypeBlock(X.astype(np.int8), 'i8'),
                GenotypeBlock(pack_plink(X.astype(np.int64)), 'plink2')):
        got = eng.run(blk)
        assert_stats_close({k: got[k] for k in STATS}, ref, label=f'N={N} {blk.format}')
    eng.close()


@pytest.mark.parametrize('G', [1, 2, 7, 8, 9, 127, 128, 129, 255, 257, 1000])
def test_ragged_variant_counts(rg, G):
    from glow_b200.engine import GenotypeBlock
    N, P, K = 200, 3, 2
    X, phe, cov = _data(rg, N, G, P, K)
    eng, st = _state(phe, cov)
    ref = orc.block_stats(X, st)
    assert_stats_close(eng.run(GenotypeBlock(X, 'f64')), ref, label=f'G={G} f64')
    assert_stats_close(eng.run(GenotypeBlock(pack_plink(X.astype(np.int64)), 'plink2')), ref, label=f'G={G} plink')
    eng.close()


@pytest.mark.parametrize('K,P', [(0, 1), (1, 1), (3, 5), (7, 1), (8, 8), (15, 1), (16, 1), (17, 1), (30, 40), (40, 100)])
def test_column_count_shapes(rg, K, P):
    """Every packing shape: fewer than one m-tile, exact multiples of 8, 1-3 tail columns, several column groups."""
    from glow_b200.engine import GenotypeBlock
    N, G = 300, 40
    X, phe, cov = _data(rg, N, G, P, K)
    eng, st = _state(phe, cov)
    ref = orc.block_stats(X, st)
    assert_stats_close(eng.run(GenotypeBlock(X, 'f64')), ref, label=f'K={K} P={P} f64')
    assert_stats_close(eng.run(GenotypeBlock(pack_plink(X.astype(np.int64)), 'plink2')), ref, label=f'K={K} P={P} p2')
    eng.close()


def test_no_intercept_and_no_covariates(rg):
    # add_intercept=False with covariates, and no covariates at all (Q has zero columns)
    import glow_b200
    N, G, P = 120, 9, 2
    X, phe, cov = _data(rg, N, G, P, 3)
    out = glow_b200.gwas.linear_regression(pd.DataFrame({'values': list(X)}), phe, cov, add_intercept=False)
    ref = orc.block_stats(X, orc.prepare_state(phe, cov, add_intercept_col=False))
    assert_stats_close({k: out[k].to_numpy().reshape(P, G) for k in STATS}, ref, label='no intercept')
    # covariate matrix whose constant column is not the first one: no intercept shortcut, still exact
    cov2 = cov.copy()
    cov2[3] = 1.0
    out2 = glow_b200.gwas.linear_regression(pd.DataFrame({'values': list(X)}), phe, cov2, add_intercept=False)
    ref2 = orc.block_stats(X, orc.prepare_state(phe, cov2, add_intercept_col=False))
    assert_stats_close({k: out2[k].to_numpy().reshape(P, G) for k in STATS}, ref2, label='constant last')


@pytest.mark.parametrize('K', [33, 64, 65, 100])
@pytest.mark.parametrize('knobs', [{}, {'mask_sparse': 'off', 'mask_int8': 'off'}])
def test_masked_pass_any_covariate_count(rg, K, knobs):
    """Phenotypes with missing values and many covariates: the reference has no limit on K (lin_reg.py:241).  The FP64
    DMMA form takes K <= 32 and the int8 form K <= 64; beyond that (or with those forms switched off) the
    sparse-complement form runs, with its generic FP64 kernel for K > 64."""
    from glow_b200 import _capi
    from glow_b200.engine import DeviceState, GenotypeBlock
    N, G, P = 400, 40, 3
    cov = pd.DataFrame(rg.standard_normal((N, K - 1)))
    Yv = rg.standard_normal((N, P))
    Yv[rg.random((N, P)) < 0.2] = np.nan
    Yv[:, 2] = rg.standard_normal(N)  # one fully observed phenotype
    phe = pd.DataFrame(Yv)
    X = rg.integers(0, 3, (G, N)).astype(np.float64)
    st = orc.prepare_state(phe, cov)
    eng = DeviceState(Q=st.Q, Y=st.y_state.Y[None], Y_mask=st.Y_mask, YdotY=st.y_state.YdotY[None], Y_scale=st.Y_scale, dof=st.dof,
                      options=knobs)
    ref = orc.block_stats(X, st)
    assert_stats_close(eng.run(GenotypeBlock(X, 'f64')), ref, label=f'K={K} masked f64')
    path = eng.timing(0)['path']
    if K > 64:
        assert path & _capi.PATH_MASK_GENERIC
    assert_stats_close(eng.run(GenotypeBlock(pack_plink(X.astype(np.int64)), 'plink2')), ref, label=f'K={K} masked plink2')
    eng.close()


def test_dropped_sample_nan_does_not_propagate(rg):
    """intersect_samples: the reference deletes the dropped genotype samples before any arithmetic (np.delete,
    lin_reg.py:228-229), so a NaN / inf in a dropped sample must not reach the statistics."""
    import glow_b200
    N, G, P, K = 230, 21, 2, 3
    samples = [f's{i}' for i in range(N)]
    X = rg.integers(0, 3, (G, N)).astype(np.float64) + 0.25 * rg.random((G, N))
    keep = [i for i in range(N) if i % 7 != 3]
    dropped = [i for i in range(N) if i % 7 == 3]
    X[:, dropped[0]] = np.nan
    X[5, dropped[1]] = np.inf
    X[6, dropped[2]] = -np.inf
    Yv = rg.standard_normal((N, P))
    Yv[rg.random((N, P)) < 0.1] = np.nan  # masked second pass as well
    phe = pd.DataFrame(Yv, index=samples).iloc[keep]
    cov = pd.DataFrame(rg.standard_normal((N, K)), index=samples)
    for dt in (np.float64, np.float32):
        pdf = pd.DataFrame({'values': list(X.astype(dt)), 'idx': np.arange(G)})
        out = glow_b200.gwas.linear_regression(pdf, phe, cov, intersect_samples=True, genotype_sample_ids=samples, dt=dt,
                                               verbose_output=True)
        st = orc.prepare_state(phe, cov, intersect_samples=True, genotype_sample_ids=samples, verbose_output=True, dt=dt)
        ref = orc.block_frame(pd.DataFrame({'values': list(X.astype(dt).astype(np.float64)), 'idx': np.arange(G)}), 'values', st)
        assert np.isfinite(out['effect'].to_numpy()).all()
        tol = dict(rtol_stat=1e-9, rtol_p=1e-6) if dt == np.float64 else dict(rtol_stat=1e-3, rtol_p=1e-2)
        assert_stats_close({k: out[k].to_numpy() for k in STATS}, {k: ref[k].to_numpy() for k in STATS},
                           label=f'NaN in dropped sample {dt.__name__}', **tol)
        np.testing.assert_allclose(out['sum_x'].to_numpy(), ref['sum_x'].to_numpy(), rtol=1e-5)


def test_masks_differing_in_one_sample_are_not_merged(rg):
    """Mask de-duplication compares the columns on a hash hit; masks with equal counts that differ in a single sample
    stay distinct, identical masks share one."""
    from glow_b200.engine import DeviceState, GenotypeBlock
    N, G, P = 300, 16, 4
    Yv = rg.standard_normal((N, P))
    Yv[10, 0] = np.nan
    Yv[11, 1] = np.nan
    Yv[10, 2] = np.nan   # same mask as phenotype 0
    phe, cov = pd.DataFrame(Yv), pd.DataFrame(rg.standard_normal((N, 2)))
    st = orc.prepare_state(phe, cov)
    eng = DeviceState(Q=st.Q, Y=st.y_state.Y[None], Y_mask=st.Y_mask, YdotY=st.y_state.YdotY[None], Y_scale=st.Y_scale, dof=st.dof)
    assert eng.unique_masks == 2
    X = rg.integers(0, 3, (G, N)).astype(np.float64)
    assert_stats_close(eng.run(GenotypeBlock(X, 'f64')), orc.block_stats(X, st), label='mask dedup')
    eng.close()


def test_fully_observed_phenotypes_take_any_covariate_count(rg):
    from glow_b200.engine import DeviceState
    N, P = 200, 2
    Y = rg.standard_normal((N, P))
    Q65 = np.linalg.qr(rg.standard_normal((N, 65)))[0]
    # fully observed phenotypes have no such limit
    eng = DeviceState(Q=Q65, Y=Y, Y_mask=np.ones((N, P)), YdotY=np.sum(Y**2, axis=0), Y_scale=np.ones(P), dof=N - 65 - 1)
    eng.close()


def test_strided_and_unaligned_rows(rg):
    """row_stride larger than a row, and rows that start at odd addresses (PLINK) / 8-byte-only aligned (f64, odd N)."""
    from glow_b200 import _capi
    from glow_b200.engine import GenotypeBlock
    N, G, P, K = 131, 37, 2, 2  # odd N: f64 rows alternate 16-byte alignment
    X, phe, cov = _data(rg, N, G, P, K)
    eng, st = _state(phe, cov)
    ref = orc.block_stats(X, st)
    assert_stats_close(eng.run(GenotypeBlock(X, 'f64')), ref, label='odd N f64')
    lib = _capi.load()
    # strided f64: a [G, N + 5] buffer of which the first N entries of every row are the block
    wide = np.full((G, N + 5), np.nan)
    wide[:, :N] = X
    out = {k: np.empty((P, G)) for k in STATS}
    bd = _capi.BlockDesc(format=_capi.FMT_F64, memspace=_capi.MEM_HOST, genotypes=wide.ctypes.data,
                         row_stride_bytes=wide.strides[0], n_variants=G, contig=0, missing_policy=0)
    bo = _capi.BlockOut(memspace=_capi.MEM_HOST, **{k: out[k].ctypes.data for k in STATS})
    _capi.check(lib.glr_run_block(eng._handle, C.byref(bd), C.byref(bo)))
    assert_stats_close(out, ref, label='strided f64')
    # PLINK rows at every byte alignment: 33-byte rows inside a buffer offset by 1..3 bytes
    packed = pack_plink(X.astype(np.int64))
    rb = packed.shape[1]
    for shift in (1, 2, 3):
        buf = np.zeros(G * (rb + 3) + 8, dtype=np.uint8)
        view = np.lib.stride_tricks.as_strided(buf[shift:], shape=(G, rb), strides=(rb + 3, 1))
        view[:] = packed
        bd = _capi.BlockDesc(format=_capi.FMT_PLINK2, memspace=_capi.MEM_HOST, genotypes=buf.ctypes.data + shift,
                             row_stride_bytes=rb + 3, n_variants=G, contig=0, missing_policy=0)
        _capi.check(lib.glr_run_block(eng._handle, C.byref(bd), C.byref(bo)))
        assert_stats_close(out, ref, label=f'plink shift {shift}')
    eng.close()


def test_device_resident_buffers_and_async_lanes(rg):
    import torch
    from glow_b200.engine import GenotypeBlock
    N, G, P, K = 1000, 300, 3, 4
    X, phe, cov = _data(rg, N, G, P, K)
    eng, st = _state(phe, cov)
    ref = orc.block_stats(X, st)
    packed = pack_plink(X.astype(np.int64))
    d_in = torch.from_numpy(packed).cuda()
    d_out = {k: torch.empty((P, G), dtype=torch.float64, device='cuda') for k in STATS}
    eng.run_device('plink2', d_in.data_ptr(), packed.shape[1], G, {k: v.data_ptr() for k, v in d_out.items()},
                   missing='minus_one')
    assert_stats_close({k: v.cpu().numpy() for k, v in d_out.items()}, ref, label='device buffers')
    t = eng.timing(0)
    assert t['kernel_launches'] >= 3 and t['gram_ms'] > 0
    # two lanes in flight, results in submission order
    blocks = [(GenotypeBlock(packed[i:i + 64], 'plink2'), 0) for i in range(0, G, 64)]
    outs = list(eng.run_pipelined(iter(blocks)))
    for k in STATS:
        np.testing.assert_allclose(np.concatenate([o[k] for o in outs], axis=1), ref[k], rtol=1e-9, atol=1e-12)
    eng.close()


def test_error_codes(rg):
    from glow_b200 import _capi
    from glow_b200.engine import GenotypeBlock
    N, G, P, K = 64, 4, 1, 1
    X, phe, cov = _data(rg, N, G, P, K)
    eng, _ = _state(phe, cov)
    with pytest.raises(ValueError):  # wrong row length
        eng.run(GenotypeBlock(X[:, :-1].copy(), 'f64'))
    with pytest.raises(ValueError):  # contig out of range
        eng.run(GenotypeBlock(X, 'f64'), contig=3)
    lib = _capi.load()
    assert lib.glr_block_wait(eng._handle, 0) == _capi.ERR_STATE  # nothing in flight
    assert b'no block in flight' in lib.glr_last_error()
    with pytest.raises(ValueError):  # lane out of range
        eng.submit(9, GenotypeBlock(X, 'f64'))
    eng.close()
    from glow_b200.engine import DeviceState
    st = orc.prepare_state(phe, cov)
    with pytest.raises(ValueError, match='q_digits'):  # digit count outside 3..7
        DeviceState(Q=st.Q, Y=st.y_state.Y[None], Y_mask=st.Y_mask, YdotY=st.y_state.YdotY[None], Y_scale=st.Y_scale, dof=st.dof,
                    options={'q_digits': 9})


def test_tensor_issue_rate_probe():
    """The measurement aid bench.py uses as the roofline ceiling of the int8 first pass."""
    from glow_b200 import _capi
    r = C.c_double()
    _capi.check(_capi.load().glr_tensor_i8_issue_rate(0, 1.0, C.byref(r)))
    assert 1000.0 < r.value < 9000.0  # dense int8 on a B200: 4.5 POP/s nominal
    with pytest.raises(ValueError):
        _capi.check(_capi.load().glr_tensor_i8_issue_rate(0, 0.0, C.byref(r)))


def test_smoke_entry_point():
    import __graft_entry__ as ge
    ge.smoke()


@pytest.mark.parametrize('N,G,P,K', [(1000, 300, 1, 16), (515, 129, 2, 3), (4099, 1000, 20, 16), (127, 64, 1, 0), (130, 5, 3, 30),
                                     (700, 260, 45, 16), (300, 129, 19, 0)])
@pytest.mark.parametrize('pair', ['0', '2'])
def test_int8_split_path_matches_dmma_path_and_oracle(rg, N, G, P, K, pair, monkeypatch):
    """The tcgen05 int8 digit-split contraction (split8.cu) against the FP64 DMMA path and the oracle:
    ragged sample / variant counts, unaligned rows, both imputation policies, one to four column blocks,
    single-CTA and CTA-pair kernels."""
    from glow_b200 import _capi
    from glow_b200.engine import GenotypeBlock
    X = rg.integers(0, 3, (G, N)).astype(np.float64)
    X[:, 0] = [g % 3 for g in range(G)]
    X[:, -1] = [(g + 1) % 3 for g in range(G)]
    states = X.astype(np.int64)
    states[rg.random((G, N)) < 0.03] = -1
    states[1, :] = np.where(states[1] == -1, 0, states[1])
    cov = pd.DataFrame(rg.standard_normal((N, K))) if K else None
    phe = pd.DataFrame(rg.standard_normal((N, P)))
    packed = pack_plink(states)
    monkeypatch.setenv('GLR_S8_PAIR', pair)  # single-CTA kernel / CTA-pair (cta_group::2) kernel
    monkeypatch.setenv('GLR_SPLIT8', '2')
    eng8, st = _state(phe, cov)
    monkeypatch.setenv('GLR_SPLIT8', '0')
    engd, _ = _state(phe, cov)
    for policy in ('mean', 'minus_one'):
        Xo = orc.mean_substitute(states.astype(np.float64)) if policy == 'mean' else states.astype(np.float64)
        ref = orc.block_stats(Xo, st)
        r8 = eng8.run(GenotypeBlock(packed, 'plink2', policy))
        rd = engd.run(GenotypeBlock(packed, 'plink2', policy))
        assert_stats_close(r8, ref, label=f'split8 {policy}')
        assert_stats_close(rd, ref, label=f'dmma {policy}')
        assert_stats_close(r8, rd, label=f'split8 vs dmma {policy}')
    # rows at an odd byte offset with a padded stride
    rb = packed.shape[1]
    buf = np.zeros(G * (rb + 5) + 8, dtype=np.uint8)
    view = np.lib.stride_tricks.as_strided(buf[3:], shape=(G, rb), strides=(rb + 5, 1))
    view[:] = packed
    out = {k: np.empty((P, G)) for k in STATS}
    bd = _capi.BlockDesc(format=_capi.FMT_PLINK2, memspace=_capi.MEM_HOST, genotypes=buf.ctypes.data + 3,
                         row_stride_bytes=rb + 5, n_variants=G, contig=0, missing_policy=_capi.MISSING_MEAN)
    bo = _capi.BlockOut(memspace=_capi.MEM_HOST, **{k: out[k].ctypes.data for k in STATS})
    _capi.check(_capi.load().glr_run_block(eng8._handle, C.byref(bd), C.byref(bo)))
    assert_stats_close(out, orc.block_stats(orc.mean_substitute(states.astype(np.float64)), st), label='split8 unaligned')
    eng8.close()
    engd.close()


def test_int8_split_path_with_dropped_samples_and_loco(rg, monkeypatch):
    import glow_b200
    monkeypatch.setenv('GLR_SPLIT8', '2')
    N, G, P, K = 260, 40, 2, 4
    samples = [f's{i}' for i in range(N)]
    states = rg.integers(-1, 3, (G, N))
    states[:, :3] = [0, 1, 2]
    keep = [i for i in range(N) if i % 11 != 5]
    phe = pd.DataFrame(rg.standard_normal((N, P)), index=samples).iloc[keep]
    cov = pd.DataFrame(rg.standard_normal((N, K)), index=samples)
    contigs = ['1', '2']
    off = pd.DataFrame(0.1 * rg.standard_normal((len(keep) * 2, P)), index=pd.MultiIndex.from_product([phe.index, contigs]))
    meta = pd.DataFrame({'contigName': [contigs[g % 2] for g in range(G)], 'idx': np.arange(G)})
    src = glow_b200.PlinkBedSource(bytes([0x6c, 0x1b, 0x01]) + pack_plink(states).tobytes(), N, meta, mean_substitute=True)
    out = glow_b200.gwas.linear_regression(src, phe, cov, off, intersect_samples=True, genotype_sample_ids=samples)
    X = orc.mean_substitute(states.astype(np.float64))
    st = orc.prepare_state(phe, cov, off, intersect_samples=True, genotype_sample_ids=samples)
    pdf = meta.copy()
    pdf['values'] = list(X)
    ref = orc.block_frame(pdf, 'values', st)
    assert out['idx'].tolist() == ref['idx'].tolist()
    assert_stats_close({k: out[k].to_numpy() for k in STATS}, {k: ref[k].to_numpy() for k in STATS}, label='split8 drop+loco')


@pytest.mark.parametrize('bits,compress', [(8, True), (16, True), (8, False)])
def test_bgen_source_runs_on_the_device(rg, tmp_path, bits, compress):
    """BGEN file -> inflated probabilities -> glr_bgen_to_plink2 on the device -> 2-bit blocks in HBM -> regression on
    the int8 tensor path, against the oracle fed with the oracle-side reader's hard calls (Glow's threshold 0.9)."""
    import glow_b200
    from glow_b200 import _capi
    from gpu_utils import write_bgen
    from oracle import bgen as obgen
    G, N, P, K = 300, 1003, 2, 3
    hi = 2**bits - 1
    pick = rg.integers(0, 3, (G, N))
    a = np.where(pick == 0, hi, 0)
    b = np.where(pick == 1, hi, 0)
    fuzzy = rg.random((G, N)) < 0.1  # uncertain genotypes: some clear the threshold, some do not
    fa = rg.integers(0, hi + 1, (G, N))
    fb = np.minimum(rg.integers(0, hi + 1, (G, N)), hi - fa)
    probs = np.stack([np.where(fuzzy, fa, a), np.where(fuzzy, fb, b)], axis=2)
    missing = rg.random((G, N)) < 0.02
    path = tmp_path / 'g.bgen'
    write_bgen(str(path), probs, missing, compress=compress, bits=bits, contig=[1 + g // 150 for g in range(G)])
    _, _, _, states = obgen.read_bgen_states(str(path))
    assert (states == -1).any() and (states == 2).any()
    phe = pd.DataFrame(rg.standard_normal((N, P)) + 0.1 * np.maximum(states[:P], 0).T, columns=['a', 'b'])
    cov = pd.DataFrame(rg.standard_normal((N, K)))
    contigs = ['1', '2']
    off = pd.DataFrame(0.1 * rg.standard_normal((N * 2, P)), index=pd.MultiIndex.from_product([phe.index, contigs]), columns=phe.columns)
    for o in (pd.DataFrame({}), off):
        src = glow_b200.BgenSource(str(path), block_variants=128, split_contigs=not o.empty)
        out = glow_b200.gwas.linear_regression(src, phe, cov, o, options={'split8': 'force'})
        st = orc.prepare_state(phe, cov, o if not o.empty else None)
        pdf = src.meta.copy()
        pdf['values'] = list(orc.mean_substitute(states.astype(np.float64)))
        ref = pd.concat([orc.block_frame(pdf.iloc[g0:g0 + 128], 'values', st) for g0 in range(0, G, 128)])
        assert out['names'].tolist() == ref['names'].tolist() and out['phenotype'].tolist() == ref['phenotype'].tolist()
        assert_stats_close({k: out[k].to_numpy() for k in STATS}, {k: ref[k].to_numpy() for k in STATS}, label=f'bgen {bits} bits')
    # logistic score test from the same file
    yb = pd.DataFrame((rg.random((N, P)) < 0.4).astype(np.float64), columns=['a', 'b'])
    lo_out = glow_b200.gwas.logistic_regression(glow_b200.BgenSource(str(path), block_variants=128), yb, cov, correction='none')
    from oracle import logreg_oracle as lo
    C, mask, lst = lo.prepare_state(yb, cov)
    want = lo.block_score_test(orc.mean_substitute(states.astype(np.float64)), C, mask, lst)
    got = lo_out.sort_values(['phenotype', 'start'], kind='stable')
    np.testing.assert_allclose(got['chisq'].to_numpy(), want['chisq'].ravel(), rtol=1e-9, atol=1e-13)


@pytest.mark.parametrize('N,G,P,K', [(1000, 300, 1, 16), (4100, 700, 3, 5), (2048, 129, 20, 16)])
@pytest.mark.parametrize('pair', ['force', 'off'])
def test_int8_states_on_the_int8_tensor_path(rg, N, G, P, K, pair):
    """genotype_states as int8 (GLR_FMT_I8) through split8_kernel: same digit-split contraction, 128 bytes of the row per
    stage.  The kernel checks every value; a block with anything outside {-1, 0, 1, 2} is re-run on the FP64 path, where
    any int8 value is legal."""
    from glow_b200 import _capi
    from glow_b200.engine import DeviceState, GenotypeBlock
    states = rg.integers(0, 3, (G, N)).astype(np.int8)
    states[rg.random((G, N)) < 0.03] = -1
    states[:, :3] = [0, 1, 2]
    cov = pd.DataFrame(rg.standard_normal((N, K)))
    Y = rg.standard_normal((N, P))
    Y[rg.random((N, P)) < 0.02] = np.nan
    phe = pd.DataFrame(Y)
    st = orc.prepare_state(phe, cov)
    eng = DeviceState(Q=st.Q, Y=st.y_state.Y[None], Y_mask=st.Y_mask, YdotY=st.y_state.YdotY[None], Y_scale=st.Y_scale, dof=st.dof,
                      options={'split8': 'force', 'split8_pair': pair})
    for policy in ('mean', 'minus_one'):
        X = states.astype(np.float64)
        ref = orc.block_stats(orc.mean_substitute(X) if policy == 'mean' else X, st)
        got = eng.run(GenotypeBlock(states, 'i8', policy))
        path = eng.timing(0)['path']
        assert path & _capi.PATH_SPLIT8 and not path & (_capi.PATH_DMMA | _capi.PATH_REPLAYED), path
        assert_stats_close(got, ref, label=f'i8 on split8 {policy}')
        # the same block as 2-bit rows gives the same bits (same kernel arithmetic, other decode)
        got2 = eng.run(GenotypeBlock(pack_plink(states.astype(np.int64)), 'plink2', policy))
        for k in STATS:
            assert np.array_equal(got[k], got2[k], equal_nan=True), k
    # rows that are not 4-byte aligned (odd sample count): the block takes the FP64 path
    st3 = orc.prepare_state(phe.iloc[:N - 1], cov.iloc[:N - 1])
    eng3 = DeviceState(Q=st3.Q, Y=st3.y_state.Y[None], Y_mask=st3.Y_mask, YdotY=st3.y_state.YdotY[None], Y_scale=st3.Y_scale,
                       dof=st3.dof, options={'split8': 'force', 'split8_pair': pair})
    odd = np.ascontiguousarray(states[:, :N - 1])
    got = eng3.run(GenotypeBlock(odd, 'i8', 'mean'))
    assert eng3.timing(0)['path'] & _capi.PATH_DMMA
    assert_stats_close(got, orc.block_stats(orc.mean_substitute(odd.astype(np.float64)), st3), label='i8 unaligned rows')
    eng3.close()
    # values outside {-1, 0, 1, 2} (triploid call sums, say): detected in the kernel, block re-run on the FP64 path
    wild = states.copy()
    wild[5, 7] = 3
    wild[G - 1, N - 1] = 4
    wild[9, 100] = -2
    got = eng.run(GenotypeBlock(wild, 'i8', 'minus_one'))
    path = eng.timing(0)['path']
    assert path & _capi.PATH_REPLAYED and path & _capi.PATH_DMMA, path
    assert_stats_close(got, orc.block_stats(wild.astype(np.float64), st), label='i8 out of range -> FP64 path')
    eng.close()
