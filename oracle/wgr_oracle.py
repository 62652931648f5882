"""TEST INFRASTRUCTURE ONLY -- numpy restatement of GloWGR's LOCO apply step (RidgeRegression.transform_loco,
python/glow/wgr/ridge_regression.py:199-234; apply_model, python/glow/wgr/ridge_udfs.py:259-347; assemble_block,
python/glow/wgr/model_functions.py:118-158).  Parity oracle for ``glow_b200.wgr.loco_offsets``; only ``tests/`` import it.

Parity status: PINNED -- ``tests/test_wgr_oracle_pins.py`` runs the reference's own ``apply_model`` (loaded byte-for-byte
by ``oracle/refshim.py::load_wgr``) on the joined block/model rows of every (sample block, label, left-out chromosome)
group and compares.  The Spark joins / groupBys around it (apply_model_df, flatten_prediction_df) are restated as
plain pandas selections in that test driver."""
import re

import numpy as np

_CHR_RE = re.compile(r'^chr_(.+?)_(alpha|block)')


def loco_predict(headers, X_raw, mu, sig, B, sample_block, chromosomes, cov=None, B_cov=None):
    """out[c][s]: prediction for sample s with the headers of chromosome c removed (ridge_regression.py:221-223: model rows
    whose header matches ^chr_<c>_(alpha|block) are filtered out, everything else -- covariates included -- stays)."""
    X = (np.asarray(X_raw, dtype=np.float64) - mu) / sig  # model_functions.py:153
    hdr_chr = np.array([_CHR_RE.search(h).group(1) for h in headers])
    out = np.empty((len(chromosomes), X.shape[0]))
    for ci, c in enumerate(chromosomes):
        keep = hdr_chr != c
        for b in np.unique(sample_block):
            rows = sample_block == b
            Xb = X[rows][:, keep]
            Bb = B[b][keep]
            if cov is not None:  # ridge_udfs.py:318-323: covariates are prepended to the block
                Xb = np.column_stack((cov[rows], Xb))
                Bb = np.concatenate((B_cov[b], Bb))
            out[ci, rows] = Xb @ Bb  # ridge_udfs.py:325-326
    return out
