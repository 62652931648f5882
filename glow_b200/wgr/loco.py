"""Leave-one-chromosome-out prediction on the device.

Reference: ``RidgeRegression.transform_loco`` (python/glow/wgr/ridge_regression.py:199-234) calls ``transform`` once per
chromosome with the model rows of that chromosome filtered out (``~header.rlike('^chr_<c>_(alpha|block)')``); ``transform``
joins the reduced block matrix with the model of the cross-validated alpha and runs ``apply_model``
(python/glow/wgr/ridge_udfs.py:259-347) per (sample block, label): ``X @ B`` with
``X = [covariates | (X_raw - mu) / sig]`` (``assemble_block``, model_functions.py:118-158); ``flatten_prediction_df``
then lays the predictions out like the label frame, and the per-chromosome frames are concatenated with a
``contigName`` index level.

In the reference all of that is Spark joins and grouped pandas UDFs over the block-matrix DataFrames.  The numbers it
works on are, per label: the reduced matrix of the label's headers, their (mu, sig), the chromosome of every header,
and one coefficient vector per sample block.  ``loco_offsets`` takes exactly those arrays and does the apply step for
every chromosome in one kernel launch per label (``glr_loco_predict``).
"""
import ctypes as C
import re
from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import numpy as np
import pandas as pd

from .. import _capi

_CHR_RE = re.compile(r'^chr_(.+?)_(alpha|block)')  # model_functions.py:704 (infer_chromosomes)


@dataclass
class LabelModel:
    """Everything ``apply_model`` uses for one label.

    headers:       level-1 / level-2 header names of the reduced matrix columns (``chr_<c>_block_<i>_alpha_<a>_label_<l>``)
    X_raw:         [n_samples, n_headers] reduced block matrix, samples in the order of ``sample_ids``
    mu, sig:       [n_headers] centring / scaling stored with the reduced matrix rows
    coefficients:  [n_sample_blocks, n_headers] coefficients of the chosen alpha, one model per sample block
    cov_coefficients: [n_sample_blocks, n_covariates] coefficients of the covariate columns (or None)
    """
    headers: Sequence[str]
    X_raw: np.ndarray
    mu: np.ndarray
    sig: np.ndarray
    coefficients: np.ndarray
    cov_coefficients: Optional[np.ndarray] = None


def header_chromosome(header: str) -> str:
    m = _CHR_RE.search(header)
    if not m:
        raise ValueError(f'cannot infer the chromosome of header {header!r}')
    return m.group(1)


def loco_offsets(models: Dict[str, LabelModel], sample_ids: Sequence, sample_blocks: Dict[str, Sequence],
                 covariates: Optional[np.ndarray] = None, chromosomes: Optional[Sequence[str]] = None,
                 device: int = 0) -> pd.DataFrame:
    """LOCO offsets for every label: DataFrame indexed by (sample_id, contigName) with one column per label, sorted by
    chromosome first and label-frame sample order second -- the layout ``transform_loco`` returns
    (ridge_regression.py:229-232) and ``linear_regression(offset_df=...)`` consumes.

    ``sample_blocks``: sample block id -> sample ids (the dict the reference passes around as ``sample_blocks``);
    ``covariates``: [n_samples, n_covariates] standardised covariate matrix with its intercept column, as
    ``apply_model`` prepends it (or None); ``chromosomes``: restrict / order the output (default: inferred from the
    headers, sorted -- ridge_regression.py:215-216)."""
    lib = _capi.load()
    sample_ids = list(sample_ids)
    N = len(sample_ids)
    block_names = list(sample_blocks)
    block_of = {}
    for bi, name in enumerate(block_names):
        for sid in sample_blocks[name]:
            block_of[sid] = bi
    try:
        sb = np.array([block_of[s] for s in sample_ids], dtype=np.int32)
    except KeyError as exc:
        raise ValueError(f'sample {exc.args[0]!r} is in no sample block') from None
    inferred = sorted({header_chromosome(h) for m in models.values() for h in m.headers})
    chroms = sorted(chromosomes) if chromosomes else inferred
    cov = None if covariates is None else np.ascontiguousarray(covariates, dtype=np.float64)
    if cov is not None and cov.shape[0] != N:
        raise ValueError('covariates must have one row per sample')
    cols = {}
    for label, m in models.items():
        X = np.ascontiguousarray(m.X_raw, dtype=np.float64)
        H = X.shape[1]
        if X.shape[0] != N or len(m.headers) != H:
            raise ValueError(f'label {label!r}: X_raw must be [n_samples, n_headers]')
        B = np.ascontiguousarray(m.coefficients, dtype=np.float64)
        if B.shape != (len(block_names), H):
            raise ValueError(f'label {label!r}: coefficients must be [n_sample_blocks, n_headers]')
        # headers of chromosomes outside `chroms` are never left out: they go to an extra, never reported, group
        cidx = {c: i for i, c in enumerate(chroms)}
        hc = np.array([cidx.get(header_chromosome(h), len(chroms)) for h in m.headers], dtype=np.int32)
        n_chr = len(chroms) + (1 if (hc == len(chroms)).any() else 0)
        mu = np.ascontiguousarray(m.mu, dtype=np.float64)
        sig = np.ascontiguousarray(m.sig, dtype=np.float64)
        Bc = None
        if cov is not None:
            if m.cov_coefficients is None:
                raise ValueError(f'label {label!r}: covariates were given but the model has no covariate coefficients')
            Bc = np.ascontiguousarray(m.cov_coefficients, dtype=np.float64)
            if Bc.shape != (len(block_names), cov.shape[1]):
                raise ValueError(f'label {label!r}: cov_coefficients must be [n_sample_blocks, n_covariates]')
        out = np.empty((n_chr, N), dtype=np.float64)
        _capi.check(lib.glr_loco_predict(device, N, H, len(block_names), n_chr, X.ctypes.data, mu.ctypes.data, sig.ctypes.data,
                                         B.ctypes.data, hc.ctypes.data, sb.ctypes.data,
                                         None if cov is None else cov.ctypes.data, 0 if cov is None else cov.shape[1],
                                         None if Bc is None else Bc.ctypes.data, out.ctypes.data))
        cols[label] = out[:len(chroms)].reshape(-1)
    index = pd.MultiIndex.from_arrays([np.tile(np.asarray(sample_ids, dtype=object), len(chroms)),
                                       np.repeat(np.asarray(chroms, dtype=object), N)], names=['sample_id', 'contigName'])
    return pd.DataFrame(cols, index=index)
