"""World-size-2 gloo test of the N>1 host path: variant sharding + ONE state broadcast.  No GPU here,
so each rank checks the state it received and computes its shard with the oracle (checker only);
rank order concatenation must reproduce the single-process result and ordering exactly."""
import os
import socket
import sys

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from glow_b200 import dist as gdist
    from oracle import linreg_oracle as orc
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rg = np.random.default_rng(5)  # every rank can regenerate the inputs; only rank 0 prepares the state
        N, G, P, K = 120, 23, 3, 2
        X = rg.binomial(2, 0.3, (G, N)).astype(np.float64) + rg.uniform(0, 0.1, (G, N))
        phe = pd.DataFrame(rg.standard_normal((N, P)))
        phe.iloc[3:9, 1] = np.nan
        cov = pd.DataFrame(rg.standard_normal((N, K)))
        state = None
        if rank == 0:
            st = orc.prepare_state(phe, cov)
            state = dict(Q=st.Q, Y=st.y_state.Y, Y_mask=st.Y_mask, YdotY=st.y_state.YdotY, Y_scale=st.Y_scale,
                         dof=st.dof)
        flat, header = gdist.broadcast_state(state, src=0)
        got = gdist.unpack_state(flat.numpy(), header)
        ost = orc.OracleState(got['Q'], got['Y_mask'], got['Y_scale'], got['dof'], ['0', '1', '2'],
                              orc.OracleYState(got['Y'][0], got['YdotY'][0]), None, None)
        lo, hi = gdist.variant_shard(G, rank, world)
        res = orc.block_stats(X[lo:hi], ost)
        gathered = [None] * world
        dist.all_gather_object(gathered, (lo, hi, {k: v for k, v in res.items()}))
        if rank == 0:
            q.put(gathered)
    finally:
        dist.destroy_process_group()


def test_state_broadcast_and_variant_sharding_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from oracle import linreg_oracle as orc
    rg = np.random.default_rng(5)
    N, G, P, K = 120, 23, 3, 2
    X = rg.binomial(2, 0.3, (G, N)).astype(np.float64) + rg.uniform(0, 0.1, (G, N))
    phe = pd.DataFrame(rg.standard_normal((N, P)))
    phe.iloc[3:9, 1] = np.nan
    cov = pd.DataFrame(rg.standard_normal((N, K)))
    ref = orc.block_stats(X, orc.prepare_state(phe, cov))
    assert [(lo, hi) for lo, hi, _ in gathered] == [(0, 11), (11, 23)]
    for k in ('effect', 'stderror', 'tvalue', 'pvalue'):
        whole = np.concatenate([r[k] for _, _, r in gathered], axis=1)
        assert np.array_equal(whole, ref[k]), k  # identical state + identical order => bit-identical


def test_variant_shard_partition():
    from glow_b200.dist import variant_shard
    for G in (0, 1, 7, 1000003):
        for world in (1, 2, 4, 8):
            edges = [variant_shard(G, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == G
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        variant_shard(10, 2, 2)


def test_pack_unpack_state_roundtrip():
    from glow_b200.dist import pack_state, unpack_state, state_layout
    rg = np.random.default_rng(1)
    N, K, P, C = 17, 3, 4, 2
    Q, Y, M = rg.random((N, K)), rg.random((C, N, P)), (rg.random((N, P)) > 0.2).astype(float)
    ydy, sc, raw = rg.random((C, P)), rg.random(P), rg.random((N, P))
    flat, hdr = pack_state(Q, Y, M, ydy, sc, 13, raw)
    assert hdr.tolist() == [N, K, P, C, 1, 13]
    assert state_layout(hdr)['_total'][0] == flat.size
    got = unpack_state(flat, hdr)
    for name, ref in (('Q', Q), ('Y', Y), ('Y_mask', M), ('YdotY', ydy), ('Y_scale', sc), ('Y_raw', raw)):
        assert np.array_equal(got[name], ref)
    assert got['dof'] == 13
