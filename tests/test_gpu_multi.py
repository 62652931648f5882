"""Multi-GPU product path on hardware: `linear_regression(..., devices=[...])` shards the variant
blocks over the GPUs of one process, replicates the state with one NCCL broadcast (glr_comm_broadcast_state) and gathers
the results in order.  The sharded result must be IDENTICAL, bit for bit, to the single-GPU result -- rows, order and
values -- and agree with the oracle.  Needs >= 2 GPUs (`gpurun --gpus 2`); the single-GPU cases run everywhere."""
import ctypes as C

import numpy as np
import pandas as pd
import pytest

from gpu_utils import pack_plink
from helpers import STATS, assert_stats_close
from oracle import linreg_oracle as orc

pytestmark = pytest.mark.gpu


def _n_gpus():
    from glow_b200 import _capi
    return _capi.load().glr_device_count()


def _inputs(rg, N=3000, G=1100, P=3, K=5, loco=False):
    states = rg.integers(0, 3, (G, N)).astype(np.int8)
    states[rg.random((G, N)) < 0.02] = -1
    Yv = rg.standard_normal((N, P)) + 0.1 * states[:P].T
    Yv[rg.random((N, P)) < 0.03] = np.nan
    phe = pd.DataFrame(Yv, columns=[f'p{i}' for i in range(P)])
    cov = pd.DataFrame(rg.standard_normal((N, K)))
    contigs = ['1', '2', '3']
    meta = pd.DataFrame({'contigName': [contigs[min(2, g * 3 // G)] for g in range(G)], 'idx': np.arange(G)})
    off = None
    if loco:
        off = pd.DataFrame(0.1 * rg.standard_normal((N * 3, P)), index=pd.MultiIndex.from_product([phe.index, contigs]),
                           columns=phe.columns)
    return states, phe, cov, off, meta


def _frames_identical(a, b):
    assert list(a.columns) == list(b.columns)
    assert len(a) == len(b)
    for c in a.columns:
        x, y = a[c].to_numpy(), b[c].to_numpy()
        if x.dtype.kind == 'f':
            assert np.array_equal(x, y, equal_nan=True), c
        else:
            assert (x == y).all(), c


@pytest.mark.parametrize('loco', [False, True])
def test_device_group_of_one_equals_plain_state(rg, loco):
    import glow_b200
    states, phe, cov, off, meta = _inputs(rg, loco=loco)
    bed = bytes([0x6c, 0x1b, 0x01]) + pack_plink(states.astype(np.int64)).tobytes()
    src = lambda: glow_b200.PlinkBedSource(bed, states.shape[1], meta, block_variants=256)
    a = glow_b200.gwas.linear_regression(src(), phe, cov, off if loco else pd.DataFrame({}))
    b = glow_b200.gwas.linear_regression(src(), phe, cov, off if loco else pd.DataFrame({}), devices=[0])
    _frames_identical(a, b)


@pytest.mark.skipif(_n_gpus() < 2, reason='needs two GPUs')
@pytest.mark.parametrize('loco', [False, True])
def test_sharded_result_is_bit_identical_to_single_gpu(rg, loco, tmp_path):
    import glow_b200
    from glow_b200.engine import DeviceGroup
    n = min(_n_gpus(), 4)
    states, phe, cov, off, meta = _inputs(rg, loco=loco)
    N = states.shape[1]
    offs = off if loco else pd.DataFrame({})
    packed = pack_plink(states.astype(np.int64))
    path = tmp_path / 'g.bed'
    path.write_bytes(bytes([0x6c, 0x1b, 0x01]) + packed.tobytes())
    one = glow_b200.gwas.linear_regression(glow_b200.PlinkBedSource(str(path), N, meta, block_variants=128), phe, cov, offs)
    # (a) a file source: contiguous block ranges per GPU
    many = glow_b200.gwas.linear_regression(glow_b200.PlinkBedSource(str(path), N, meta, block_variants=128), phe, cov, offs,
                                            devices=list(range(n)))
    _frames_identical(one, many)
    # (b) a stream of pandas frames: round robin over the GPUs, gathered in order
    X = orc.mean_substitute(states.astype(np.float64))
    frames = []
    for g0 in range(0, len(meta), 128):
        f = meta.iloc[g0:g0 + 128].copy()
        f['values'] = list(X[g0:g0 + 128])
        frames.append(f)
    one_f = glow_b200.gwas.linear_regression(iter(frames), phe, cov, offs)
    many_f = glow_b200.gwas.linear_regression(iter(frames), phe, cov, offs, devices=list(range(n)))
    _frames_identical(one_f, many_f)
    # (c) resident genotypes: loaded once into the HBM of the GPUs, then run against two phenotype sets
    res = glow_b200.ResidentSource(glow_b200.PlinkBedSource(str(path), N, meta, block_variants=128), devices=list(range(n)),
                                   split_contigs=loco)
    many_r = glow_b200.gwas.linear_regression(res, phe, cov, offs, devices=list(range(n)))
    _frames_identical(one, many_r)
    phe2 = phe * 2.0 + 1.0
    one2 = glow_b200.gwas.linear_regression(glow_b200.PlinkBedSource(str(path), N, meta, block_variants=128), phe2, cov, offs)
    _frames_identical(one2, glow_b200.gwas.linear_regression(res, phe2, cov, offs, devices=list(range(n))))
    with pytest.raises(ValueError, match='resident on GPUs'):
        glow_b200.gwas.linear_regression(res, phe, cov, offs, devices=[0])
    res.close()
    # the oracle agrees
    st = orc.prepare_state(phe, cov, off if loco else None)
    pdf = meta.copy()
    pdf['values'] = list(X)
    ref = orc.block_frame(pdf, 'values', st) if not loco else pd.concat(
        [orc.block_frame(pdf.iloc[g0:g0 + 128], 'values', st) for g0 in range(0, len(pdf), 128)])
    if not loco:  # (single block in the oracle frame: phenotype-major over all variants; compare per (idx, phenotype))
        key = lambda d: d.sort_values(['phenotype', 'idx'], kind='stable').reset_index(drop=True)
        a, b = key(many), key(ref)
    else:
        a, b = many.reset_index(drop=True), ref.reset_index(drop=True)
    assert a['idx'].tolist() == b['idx'].tolist() and a['phenotype'].tolist() == b['phenotype'].tolist()
    assert_stats_close({k: a[k].to_numpy() for k in STATS}, {k: b[k].to_numpy() for k in STATS}, label='sharded vs oracle')
    # the state went over NCCL
    g = DeviceGroup(list(range(n)), Q=st.Q, Y=np.stack([y.Y for y in (st.y_state.values() if loco else [st.y_state])]),
                    Y_mask=st.Y_mask, YdotY=np.stack([y.YdotY for y in (st.y_state.values() if loco else [st.y_state])]),
                    Y_scale=st.Y_scale, dof=st.dof)
    assert g.transport == 'nccl'
    g.close()


@pytest.mark.skipif(_n_gpus() < 2, reason='needs two GPUs')
def test_c_abi_comm_exports_peer_copy_transport(rg, monkeypatch):
    """glr_comm_init / glr_comm_broadcast_state through ctypes, with NCCL switched off: the replicas are filled by
    cudaMemcpyPeerAsync and give the same bits."""
    from glow_b200 import _capi
    from glow_b200.engine import DeviceGroup, GenotypeBlock
    states, phe, cov, off, meta = _inputs(rg, N=2000, G=300)
    st = orc.prepare_state(phe, cov)
    kw = dict(Q=st.Q, Y=st.y_state.Y[None], Y_mask=st.Y_mask, YdotY=st.y_state.YdotY[None], Y_scale=st.Y_scale, dof=st.dof)
    monkeypatch.setenv('GLR_COMM_NO_NCCL', '1')
    g = DeviceGroup([0, 1], **kw)
    assert g.transport == 'p2p' and len(g) == 2
    blk = GenotypeBlock(pack_plink(states.astype(np.int64)), 'plink2', 'mean')
    r0, r1 = g.states[0].run(blk), g.states[1].run(blk)
    for k in STATS:
        assert np.array_equal(r0[k], r1[k], equal_nan=True)
    assert_stats_close(r1, orc.block_stats(orc.mean_substitute(states.astype(np.float64)), st), label='replica on GPU 1')
    g.close()
