"""GloWGR LOCO apply step on the device against the oracle, and end to end: the offsets
it produces feed ``linear_regression(offset_df=...)`` on the same GPU."""
import numpy as np
import pandas as pd
import pytest

from helpers import STATS, assert_stats_close
from oracle import linreg_oracle as orc
from oracle import wgr_oracle as wo
from test_wgr_oracle_pins import make_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('n_samples,n_chr,with_cov', [(37, 3, True), (1000, 22, True), (513, 2, False), (64, 1, True)])
def test_loco_offsets_against_oracle(rg, n_samples, n_chr, with_cov):
    from glow_b200 import wgr
    models, want = {}, {}
    cases = {lab: make_case(rg, n_samples=n_samples, n_chr=n_chr, blocks_per_chr=3, with_cov=with_cov) for lab in ('y1', 'y2')}
    first = cases['y1']
    sb = np.array([i % 3 for i in range(n_samples)])
    chroms = sorted({str(c) for c in range(1, n_chr + 1)})
    for lab, c in cases.items():
        models[lab] = wgr.LabelModel(c['headers'], c['X_raw'], c['mu'], c['sig'], c['coef'][:, :, 1],
                                     None if not with_cov else c['cov_coef'][:, :, 1])
        want[lab] = wo.loco_predict(c['headers'], c['X_raw'], c['mu'], c['sig'], c['coef'][:, :, 1], sb, chroms,
                                    first['cov'], None if not with_cov else c['cov_coef'][:, :, 1])
    off = wgr.loco_offsets(models, first['sample_ids'], first['sample_blocks'], first['cov'])
    assert off.index.names == ['sample_id', 'contigName'] and off.columns.tolist() == ['y1', 'y2']
    assert off.index.get_level_values(1).unique().tolist() == chroms  # chromosome-major, sorted
    for lab in models:
        for ci, chrom in enumerate(chroms):
            got = off[lab].xs(chrom, level=1).reindex(first['sample_ids']).to_numpy()
            np.testing.assert_allclose(got, want[lab][ci], rtol=1e-11, atol=1e-12)
    with pytest.raises(ValueError, match='Standard deviation cannot be 0'):
        bad = wgr.LabelModel(first['headers'], first['X_raw'], first['mu'], np.zeros_like(first['sig']), first['coef'][:, :, 1],
                             None if not with_cov else first['cov_coef'][:, :, 1])
        wgr.loco_offsets({'y1': bad}, first['sample_ids'], first['sample_blocks'], first['cov'])


def test_offsets_feed_linear_regression(rg):
    import glow_b200
    from glow_b200 import wgr
    N, G, n_chr = 400, 60, 3
    c = make_case(rg, n_samples=N, n_chr=n_chr, blocks_per_chr=2, with_cov=True)
    model = wgr.LabelModel(c['headers'], c['X_raw'], c['mu'], c['sig'], 0.05 * c['coef'][:, :, 0], 0.05 * c['cov_coef'][:, :, 0])
    off = wgr.loco_offsets({'trait': model}, c['sample_ids'], c['sample_blocks'], c['cov'])
    phe = pd.DataFrame({'trait': rg.standard_normal(N)}, index=c['sample_ids'])
    cov = pd.DataFrame(rg.standard_normal((N, 3)), index=c['sample_ids'])
    X = rg.integers(0, 3, (G, N)).astype(np.float64)
    pdf = pd.DataFrame({'contigName': [str(1 + g % n_chr) for g in range(G)], 'idx': np.arange(G), 'values': list(X)})
    out = glow_b200.gwas.linear_regression(pdf, phe, cov, off)
    ref = orc.block_frame(pdf, 'values', orc.prepare_state(phe, cov, off))
    assert out['idx'].tolist() == ref['idx'].tolist()
    assert_stats_close({k: out[k].to_numpy() for k in STATS}, {k: ref[k].to_numpy() for k in STATS}, label='offsets from the device')
