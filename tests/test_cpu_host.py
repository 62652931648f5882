"""CPU-side tests: the C ABI exports what include/glow_b200.h declares, host logic mirrors the
reference's validation behaviour, and the product path fails loudly without a GPU."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, 'include', 'glow_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(glr_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from glow_b200 import _capi
    lib = ctypes.CDLL(_capi.LIB_PATH)
    declared = _header_functions()
    assert 'glr_state_create' in declared and 'glr_run_block' in declared
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/glow_b200.h but not exported'
    assert sorted(_capi.SYMBOLS) == declared  # the ctypes binding covers exactly the header
    assert _capi.load().glr_abi_version() == _capi.GLR_ABI_VERSION


def test_ctypes_structs_match_header_layout():
    # compile a tiny C program that prints sizeof/offsetof of the ABI structs and compare with ctypes
    from glow_b200 import _capi
    prog = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "glow_b200.h"
    int main(void) {
      printf("%zu %zu %zu %zu\n", sizeof(glr_state_desc), sizeof(glr_block_desc), sizeof(glr_block_out), sizeof(glr_timing));
      printf("%zu %zu %zu %zu\n", offsetof(glr_state_desc, Q), offsetof(glr_state_desc, dof), offsetof(glr_state_desc, memspace), offsetof(glr_block_desc, row_stride_bytes));
      printf("%zu %zu %zu %zu\n", sizeof(glr_options), offsetof(glr_state_desc, opt), offsetof(glr_options, q_digits), offsetof(glr_timing, path));
      return 0;
    }'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, 'a.c')
        open(c, 'w').write(prog)
        exe = os.path.join(d, 'a.out')
        subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), c, '-o', exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    sizes = [int(x) for x in out]
    assert sizes[:4] == [ctypes.sizeof(_capi.StateDesc), ctypes.sizeof(_capi.BlockDesc), ctypes.sizeof(_capi.BlockOut),
                         ctypes.sizeof(_capi.Timing)]
    assert sizes[4:8] == [_capi.StateDesc.Q.offset, _capi.StateDesc.dof.offset, _capi.StateDesc.memspace.offset,
                          _capi.BlockDesc.row_stride_bytes.offset]
    assert sizes[8:] == [ctypes.sizeof(_capi.Options), _capi.StateDesc.opt.offset, _capi.Options.q_digits.offset,
                         _capi.Timing.path.offset]
    o = _capi.make_options({'split8': 'force', 'split8_pair': 'off', 'sample_split': 3, 'q_digits': 7})
    assert (o.split8, o.split8_pair, o.mask_sparse, o.sample_split, o.q_digits) == (_capi.FORCE, _capi.OFF, _capi.AUTO, 3, 7)
    with pytest.raises(ValueError):
        _capi.make_options({'nope': 1})


def _no_gpu():
    from glow_b200 import _capi
    return _capi.load().glr_device_count() <= 0


@pytest.mark.skipif(not _no_gpu(), reason='only meaningful on a machine without a GPU')
def test_fails_loudly_without_gpu():
    import glow_b200
    from glow_b200._capi import GlowB200Error
    rg = np.random.default_rng(0)
    with pytest.raises(GlowB200Error, match='no CPU fallback'):
        glow_b200.gwas.linear_regression(pd.DataFrame({'values': list(rg.random((3, 10)))}),
                                         pd.DataFrame(rg.random((10, 2))))


def test_product_code_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'glow_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r'^\s*(from|import)\s+oracle', txt, flags=re.M) or 'linreg_oracle' in txt:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


# ---- validation errors raised before any device work (reference: test_lin_reg.py:311-336,454-509,565-586) ----
def _lr(*a, **k):
    import glow_b200
    return glow_b200.gwas.linear_regression(*a, **k)


def _frames(rg, n=10, g=3, p=5, k=2):
    return (pd.DataFrame({'values': list(rg.random((g, n)))}), pd.DataFrame(rg.random((n, p))),
            pd.DataFrame(rg.random((n, k))))


def test_validates_missing_covariates(rg):
    gdf, phe, cov = _frames(rg)
    cov.loc[0, 0] = np.nan
    with pytest.raises(ValueError):
        _lr(gdf, phe, cov)


def test_validate_same_number_of_rows(rg):
    gdf, phe, _ = _frames(rg, n=4)
    with pytest.raises(ValueError):
        _lr(gdf, phe, pd.DataFrame(rg.random((5, 2))))


def test_too_many_missing(rg):
    gdf, phe, cov = _frames(rg, n=25, p=24, k=10)
    missing = np.triu(np.ones(phe.shape))
    missing[:, -1] = 0
    phe[missing.astype('bool')] = np.nan
    with pytest.raises(ValueError):
        _lr(gdf, phe, cov)


def test_bad_datatype(rg):
    gdf, phe, cov = _frames(rg)
    with pytest.raises(ValueError):
        _lr(gdf, phe, cov, dt=np.int32)


def test_bad_column_name(rg):
    gdf, phe, cov = _frames(rg)
    with pytest.raises(ValueError):
        _lr(gdf.rename(columns={'values': 'genotypes'}), phe, cov, values_column='genotypes')


def test_offset_wrong_columns(rg):
    gdf, phe, _ = _frames(rg, n=25, p=10)
    off = pd.DataFrame(rg.random((25, 10)), columns=range(10, 20))
    with pytest.raises(ValueError):
        _lr(gdf, phe, offset_df=off)


def test_offset_wrong_index(rg):
    gdf, phe, _ = _frames(rg, n=25, p=10)
    off = pd.DataFrame(rg.random((25, 10)), index=range(1, 26))
    with pytest.raises(ValueError):
        _lr(gdf, phe, offset_df=off)


def test_offset_wrong_multi_index(rg):
    gdf, phe, _ = _frames(rg, n=25, p=10)
    off = pd.DataFrame(rg.random((50, 10)), pd.MultiIndex.from_product([range(1, 26), ['chr1', 'chr2']]))
    with pytest.raises(ValueError):
        _lr(gdf, phe, offset_df=off)


def test_intersect_requires_sample_ids(rg):
    gdf, phe, cov = _frames(rg)
    with pytest.raises(ValueError):
        _lr(gdf, phe, cov, intersect_samples=True)


def test_subset_contigs_states(rg):
    # test_lin_reg.py:607-620
    from glow_b200.gwas import lin_reg as lr
    n, p = 25, 25
    contigs = ['chr1', 'chr2', 'chr3']
    phe = pd.DataFrame(rg.random((n, p)))
    off = pd.DataFrame(rg.random((n * 3, p)), index=pd.MultiIndex.from_product([phe.index, contigs]))
    mask = (~np.isnan(phe.to_numpy())).astype(np.float64)
    st = lr._create_YState(phe.to_numpy(), phe, off, mask, np.float64, None)
    assert set(st.keys()) == set(contigs)
    st = lr._create_YState(phe.to_numpy(), phe, off, mask, np.float64, ['chr1', 'chr3'])
    assert set(st.keys()) == {'chr1', 'chr3'}
    # the per-contig states follow lin_reg.py:194-199
    from oracle import linreg_oracle as orc
    ref = orc.prepare_state(phe, None, off)
    ours = lr._create_YState(ref.y_state['chr1'].Y * 0 + _prepared_Y(phe), phe, off, mask, np.float64, None)
    for c in contigs:
        np.testing.assert_allclose(ours[c].Y, ref.y_state[c].Y, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(ours[c].YdotY, ref.y_state[c].YdotY, rtol=1e-12)


def _prepared_Y(phe):
    from glow_b200.gwas import functions as fx
    C = fx._add_intercept(np.empty((phe.shape[0], 0)), phe.shape[0])
    Q = np.linalg.qr(C)[0]
    Y = phe.to_numpy(np.float64, copy=True)
    m = (~np.isnan(Y)).astype(np.float64)
    Y = np.nan_to_num(Y)
    Y -= Y.mean(axis=0)
    Y = fx._residualize_in_place(Y, Q) * m
    return Y / np.sqrt(np.sum(Y**2, axis=0) / (m.sum(axis=0) - Q.shape[1]))[None, :]


def test_loco_dispatch_first_appearance_order():
    from glow_b200.gwas import functions as fx
    pdf = pd.DataFrame({'contigName': ['b', 'a', 'b', 'c', 'a'], 'i': range(5)})
    seen = []

    def f(sub, state):
        seen.append((state, sub['i'].tolist()))
        return sub

    out = fx._loco_dispatch(pdf, {'a': 'A', 'b': 'B', 'c': 'C'}, f)
    assert seen == [('B', [0, 2]), ('A', [1, 4]), ('C', [3])]
    assert out['i'].tolist() == [0, 2, 1, 4, 3]


def test_get_indices_to_drop_and_keep_index():
    from glow_b200.gwas import functions as fx
    from glow_b200.gwas.lin_reg import _keep_index
    phe = pd.DataFrame({'y': [1.0, 2.0, 3.0]}, index=['s1', 's3', 's4'])
    drop = fx._get_indices_to_drop(phe, ['s0', 's1', 's2', 's3', 's4'])
    assert drop.tolist() == [0, 2]
    assert _keep_index(5, drop).tolist() == [1, 3, 4]
    assert _keep_index(5, np.array([], dtype=np.int64)) is None


def test_plink_source_and_block_checks(golden_dir):
    import glow_b200
    d = np.load(os.path.join(golden_dir, 'plink_kat.npz'))
    src = glow_b200.PlinkBedSource(d['bed'].tobytes(), 5)
    metas = list(src.blocks())
    assert metas[0][1].format == 'plink2' and metas[0][1].data.shape == (5, 2)
    with pytest.raises(ValueError):
        glow_b200.PlinkBedSource(b'\x00\x00\x00' + d['bed'][3:].tobytes(), 5)
    with pytest.raises(ValueError):
        glow_b200.PlinkBedSource(d['bed'].tobytes(), 9)  # 10 payload bytes are not a multiple of ceil(9/4) = 3


def test_plink_file_is_streamed_block_by_block(tmp_path):
    """A .bed path is read block by block into a ring of staging buffers (parallel preadv); the blocks must be the file's
    rows in order, including the ragged last block, and a view must stay valid until three more blocks were read."""
    import glow_b200
    rg = np.random.default_rng(5)
    N, G = 1003, 700
    rb = (N + 3) // 4
    body = rg.integers(0, 256, (G, rb), dtype=np.uint8)
    path = tmp_path / 'chr.bed'
    path.write_bytes(bytes([0x6c, 0x1b, 0x01]) + body.tobytes())
    src = glow_b200.PlinkBedSource(str(path), N, block_variants=256, read_threads=3)
    seen, g0 = [], 0
    for meta, blk in src.blocks():
        assert blk.format == 'plink2' and len(meta) == blk.data.shape[0]
        assert np.array_equal(blk.data, body[g0:g0 + blk.data.shape[0]])
        seen.append((blk.data, g0))
        if len(seen) >= 3:  # the block two reads back still owns its buffer
            old, og = seen[-3]
            assert np.array_equal(old, body[og:og + old.shape[0]])
        g0 += blk.data.shape[0]
    assert g0 == G and len(seen) == 3
    assert meta.index[-1] == G - 1
    (tmp_path / 'bad.bed').write_bytes(b'\x00\x00\x00' + body.tobytes())
    with pytest.raises(ValueError):
        glow_b200.PlinkBedSource(str(tmp_path / 'bad.bed'), N)
    with pytest.raises(ValueError):
        glow_b200.PlinkBedSource(str(path), N + 4)


def test_arrow_list_columns_become_blocks_without_copy():
    import pyarrow as pa
    from glow_b200.engine import as_block
    X = np.random.default_rng(3).random((6, 9))
    col = pa.array(list(X), type=pa.list_(pa.float64()))
    blk = as_block(col)
    assert blk.format == 'f64' and np.array_equal(blk.data, X)
    # zero copy: the block aliases the Arrow child buffer
    assert blk.data.ctypes.data == col.flatten().buffers()[1].address
    fixed = pa.FixedSizeListArray.from_arrays(pa.array(X.ravel()), 9)
    assert np.array_equal(as_block(fixed).data, X)
    assert np.array_equal(as_block(pa.table({'v': col})['v']).data, X)
    i8 = as_block(pa.array([[0, 1, -1], [2, 2, 0]], type=pa.list_(pa.int8())))
    assert i8.format == 'i8' and i8.data.tolist() == [[0, 1, -1], [2, 2, 0]]
    assert as_block(pa.array(list(X.astype(np.float32)), type=pa.list_(pa.float32())), np.float32).format == 'f32'
    with pytest.raises(ValueError):
        as_block(pa.array([[1.0], [1.0, 2.0]]))
    with pytest.raises(ValueError):
        as_block(pa.array([[1.0], None], type=pa.list_(pa.float64())))
    # pandas Series backed by Arrow
    s = pd.Series(pd.arrays.ArrowExtensionArray(pa.chunked_array([col])))
    assert np.array_equal(as_block(s).data, X)


def test_assemble_matches_reference_layout():
    # lin_reg.py:231-252: pass-through columns replicated per phenotype, statistics raveled phenotype-major
    from glow_b200.gwas.lin_reg import _assemble
    pdf = pd.DataFrame({'contigName': ['1', '1', '2'], 'start': [5, 6, 7]}, index=[10, 11, 12])
    res = {k: np.arange(6, dtype=np.float64).reshape(2, 3) + i for i, k in enumerate(('effect', 'stderror', 'tvalue', 'pvalue'))}
    out = _assemble(pdf, res, ['a', 'b'], False, np.float64)
    ref = pd.concat([pdf] * 2)
    for i, k in enumerate(('effect', 'stderror', 'tvalue', 'pvalue')):
        ref[k] = list(np.ravel(res[k]))
    ref['phenotype'] = pd.Series(['a', 'b']).repeat(3).tolist()
    pd.testing.assert_frame_equal(out, ref, check_dtype=False)
    assert out.index.tolist() == [10, 11, 12, 10, 11, 12]


def test_bgen_source_parses_layout2_records(tmp_path):
    """Host half of the BGEN source: variant records and inflated probability blocks, checked
    against the oracle-side reader on synthetic files and, in the build container, on the reference's own fixture."""
    import glow_b200
    from gpu_utils import write_bgen
    from oracle import bgen as obgen
    rg = np.random.default_rng(2)
    G, N = 23, 101
    for bits, compress in ((8, True), (16, True), (8, False)):
        hi = 2**bits - 1
        a = rg.integers(0, hi + 1, (G, N))
        b = np.minimum(rg.integers(0, hi + 1, (G, N)), hi - a)
        sure = rg.random((G, N)) < 0.7  # most genotypes are confident calls
        pick = rg.integers(0, 3, (G, N))
        a = np.where(sure, np.where(pick == 0, hi, 0), a)
        b = np.where(sure, np.where(pick == 1, hi, 0), b)
        probs = np.stack([a, b], axis=2)
        missing = rg.random((G, N)) < 0.05
        path = tmp_path / f'x{bits}{int(compress)}.bgen'
        write_bgen(str(path), probs, missing, compress=compress, bits=bits)
        src = glow_b200.BgenSource(str(path), block_variants=7)
        assert (src.n_samples, src.n_variants, src.n_blocks()) == (N, G, 4)
        assert src.meta['names'].tolist() == [f'rs{g}' for g in range(G)] and src.meta['start'].tolist() == [999 + g for g in range(G)]
        pr, pl, got_bits = src._read_block(3, 11)
        assert got_bits == bits and np.array_equal(pr.reshape(8, N, 2), probs[3:11]) and np.array_equal((pl & 0x80) != 0, missing[3:11])
        _, _, _, states = obgen.read_bgen_states(str(path))
        p = probs / hi
        pa, pb = p[:, :, 0], p[:, :, 1]
        P3 = np.stack([pa, pb, 1.0 - pa - pb], axis=2)
        want = np.where((P3.max(axis=2) >= 0.9) & ~missing, P3.argmax(axis=2), -1)
        assert np.array_equal(states, want)
    ref = '/root/reference/test-data/regenie/example.bgen'
    if os.path.exists(ref):
        src = glow_b200.BgenSource(ref)
        contigs, pos, rsids, _ = obgen.read_bgen_states(ref)
        assert src.meta['names'].tolist() == rsids and src.meta['contigName'].tolist() == contigs
        assert (src.meta['start'].to_numpy() == pos - 1).all()
