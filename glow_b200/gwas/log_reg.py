"""B200-native drop-in for ``glow.gwas.logistic_regression`` (reference: python/glow/gwas/log_reg.py): the score test
on the device, the approximate Firth refit of the significant tests on the host (as in the reference).

Same signature, validation and output columns as the reference.  The driver-side preparation -- one null logistic
model per phenotype (and per contig with LOCO offsets), log_reg.py:160-200 -- stays on the host: the reference fits it
with statsmodels' GLM (iteratively reweighted least squares); here the same IRLS is written in numpy.  The block loop
(``_logistic_regression_inner``, log_reg.py:278-350: null-model residualisation of the genotypes and the score
statistic) runs on the device through the C ABI (``glr_logreg_*`` in include/glow_b200.h).  There is no CPU fallback
for the block loop.
"""
import ctypes as C_
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence, Union

import numpy as np
import pandas as pd

from . import approx_firth as af
from . import functions as gwas_fx
from .functions import _VALUES_COLUMN_NAME, _get_indices_to_drop
from .. import _capi
from ..engine import FORMATS, GenotypeBlock, ResidentBlock, as_block

__all__ = ['logistic_regression']

correction_none = 'none'
correction_approx_firth = 'approx-firth'


@dataclass
class LogRegState:
    """log_reg.py:153-158."""
    inv_CtGammaC: np.ndarray  # n_phenotypes x n_covariates x n_covariates
    gamma: np.ndarray  # n_samples x n_phenotypes
    Y_res: np.ndarray  # n_samples x n_phenotypes
    firth_offset: Optional[np.ndarray] = None  # n_samples x n_phenotypes (approx-firth only)


def _logistic_null_model_predictions(y, X, mask, offset, max_iter=100, tol=1e-8):
    """log_reg.py:160-175: binomial GLM with logit link fitted on the observed samples, predictions re-mapped to all
    samples with 0 for the unobserved ones.  The fit is the IRLS statsmodels' ``GLM.fit`` runs (starting values
    mu = (y + 0.5) / 2, weighted least squares on the working response, stop when the deviance changes by less than
    ``tol``); the maximum-likelihood estimate is unique, so the route to it does not show in the result."""
    yo = y[mask]
    Xo = X[mask, :]
    off = np.zeros(yo.shape[0]) if offset is None else np.asarray(offset, dtype=np.float64)[mask]
    mu = (yo + 0.5) / 2.0
    eta = np.log(mu / (1.0 - mu))
    dev_old = np.inf
    for _ in range(max_iter):
        w = mu * (1.0 - mu)
        z = eta - off + (yo - mu) / w
        sw = np.sqrt(w)
        beta = np.linalg.lstsq(Xo * sw[:, None], z * sw, rcond=None)[0]
        eta = Xo @ beta + off
        mu = 1.0 / (1.0 + np.exp(-eta))
        with np.errstate(divide='ignore', invalid='ignore'):
            dev = -2.0 * np.sum(np.where(yo > 0, np.log(mu), 0.0) * yo + np.where(yo < 1, np.log1p(-mu), 0.0) * (1.0 - yo))
        if abs(dev - dev_old) < tol:
            break
        dev_old = dev
    # two more Newton steps pin the estimate far below the tolerances of the comparison with the reference
    for _ in range(2):
        w = mu * (1.0 - mu)
        z = eta - off + (yo - mu) / w
        sw = np.sqrt(w)
        beta = np.linalg.lstsq(Xo * sw[:, None], z * sw, rcond=None)[0]
        eta = Xo @ beta + off
        mu = 1.0 / (1.0 + np.exp(-eta))
    out = np.zeros(y.shape)
    out[mask] = mu
    return out


def _prepare_one_phenotype(C, y, offset):
    """log_reg.py:178-200 for one (phenotype, offset) pair -> (y_res, gamma, inv_CtGammaC)."""
    mask = ~np.isnan(y)
    y_pred = _logistic_null_model_predictions(y, C, mask, offset)
    y_res = np.nan_to_num(y - y_pred)
    gamma = y_pred * (1 - y_pred)
    CtGammaC = C.T @ (gamma[:, None] * C)
    return y_res, gamma, np.linalg.inv(CtGammaC)


def _create_log_reg_state(phenotype_df, offset_df, C, contigs, correction=correction_none,
                          add_intercept=True) -> Union[LogRegState, Dict[str, LogRegState]]:
    """log_reg.py:223-275, without the Spark distribution of the null fits."""
    offset_type = gwas_fx._validate_offset(phenotype_df, offset_df)
    if offset_type == gwas_fx._OffsetType.LOCO_OFFSET and contigs is not None:
        offset_df = offset_df.loc[pd.IndexSlice[:, contigs], :]
    Y = phenotype_df.to_numpy(np.float64, copy=True)
    names = list(phenotype_df.columns)

    def one_state(off_df):
        ys, gs, invs, fos = [], [], [], []
        for j, name in enumerate(names):
            off = None if off_df is None else off_df[name].reindex(phenotype_df.index).to_numpy(np.float64)
            y_res, gamma, inv = _prepare_one_phenotype(C, Y[:, j], off)
            ys.append(y_res)
            gs.append(gamma)
            invs.append(inv)
            if correction == correction_approx_firth:  # log_reg.py:197-199
                fos.append(af.perform_null_firth_fit(Y[:, j], C, ~np.isnan(Y[:, j]), off, add_intercept))
        return LogRegState(np.stack(invs), np.column_stack(gs), np.column_stack(ys), np.column_stack(fos) if fos else None)

    if offset_type == gwas_fx._OffsetType.NO_OFFSET:
        return one_state(None)
    if offset_type == gwas_fx._OffsetType.SINGLE_OFFSET:
        return one_state(offset_df)
    all_contigs = gwas_fx._get_contigs_from_loco_df(offset_df)
    return {contig: one_state(offset_df.xs(contig, level=1)) for contig in all_contigs}


class LogRegDeviceState:
    """Replica of (C, gamma, Y_res, inv_CtGammaC per contig) on one GPU."""

    def __init__(self, Cmat, states: Sequence[LogRegState], keep_idx=None, n_geno_samples=None, device=0, Q=None):
        """``Q`` (orthonormal basis of ``Cmat``): the genotypes are first residualised linearly against it, as the
        reference does for correction='approx-firth' (log_reg.py:312-313)."""
        self._lib = _capi.load()
        self._handle = C_.c_void_p()
        Cmat = np.ascontiguousarray(Cmat, dtype=np.float64)
        gamma = np.ascontiguousarray(np.stack([s.gamma for s in states]), dtype=np.float64)
        y_res = np.ascontiguousarray(np.stack([s.Y_res for s in states]), dtype=np.float64)
        inv = np.ascontiguousarray(np.stack([s.inv_CtGammaC for s in states]), dtype=np.float64)
        n_contigs, N, P = gamma.shape
        K = Cmat.shape[1]
        if keep_idx is not None:
            keep_idx = np.ascontiguousarray(keep_idx, dtype=np.int64)
        if Q is not None:
            Q = np.ascontiguousarray(Q, dtype=np.float64)
            if Q.shape != Cmat.shape:
                raise ValueError('Q must have the shape of the covariate matrix')
        self.N, self.K, self.P, self.n_contigs = N, K, P, n_contigs
        self.Ng = int(n_geno_samples) if n_geno_samples is not None else N
        self.device = device
        d = _capi.LogRegDesc(abi_version=_capi.GLR_ABI_VERSION, device=device, n_samples=N, n_geno_samples=self.Ng, n_covariates=K,
                             n_phenotypes=P, n_contigs=n_contigs, C=Cmat.ctypes.data, gamma=gamma.ctypes.data,
                             Y_res=y_res.ctypes.data, inv_CtGammaC=inv.ctypes.data,
                             keep_idx=None if keep_idx is None else keep_idx.ctypes.data,
                             Q=None if Q is None else Q.ctypes.data)
        _capi.check(self._lib.glr_logreg_state_create(C_.byref(d), C_.byref(self._handle)))

    def run(self, block, contig: int = 0):
        resident = isinstance(block, ResidentBlock)
        if resident and block.device != self.device:
            raise ValueError(f'resident block lives on GPU {block.device}, this state on GPU {self.device}')
        G = int(block.n_variants) if resident else int(block.data.shape[0])
        chisq = np.empty((self.P, G), dtype=np.float64)
        pvalue = np.empty((self.P, G), dtype=np.float64)
        if G:
            if resident:
                addr, stride, space = block.addr, block.stride, _capi.MEM_DEVICE
            else:
                data = block.data
                if data.ndim != 2 or not data.flags['C_CONTIGUOUS']:
                    raise ValueError('genotype block must be a C-contiguous 2-D array')
                addr, stride, space = data.ctypes.data, data.strides[0], _capi.MEM_HOST
            bd = _capi.BlockDesc(format=FORMATS[block.format], memspace=space, genotypes=addr,
                                 row_stride_bytes=stride, n_variants=G, contig=contig,
                                 missing_policy=_capi.MISSING_MEAN if block.missing == 'mean' else _capi.MISSING_AS_MINUS_ONE)
            _capi.check(self._lib.glr_logreg_run_block(self._handle, C_.byref(bd), chisq.ctypes.data, pvalue.ctypes.data))
        return {'chisq': chisq, 'pvalue': pvalue}

    def close(self):
        if getattr(self, '_handle', None) is not None and self._handle.value:
            self._lib.glr_logreg_state_destroy(self._handle)
            self._handle = C_.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def logistic_regression(genotype_df,
                        phenotype_df: pd.DataFrame,
                        covariate_df: pd.DataFrame = pd.DataFrame({}),
                        offset_df: pd.DataFrame = pd.DataFrame({}),
                        correction: str = correction_approx_firth,
                        pvalue_threshold: float = 0.05,
                        contigs: Optional[List[str]] = None,
                        add_intercept: bool = True,
                        values_column: str = 'values',
                        dt: type = np.float64,
                        intersect_samples: bool = False,
                        genotype_sample_ids: Optional[List[str]] = None,
                        device: int = 0):
    """Score test of every variant against every binary phenotype (log_reg.py:24-148).

    Arguments, defaults and output columns are the reference's; ``device`` is the only addition.  ``genotype_df`` is a
    pandas DataFrame, an iterable of pandas DataFrames, or a :class:`glow_b200.sources.GenotypeSource`.
    The score test of every (variant, phenotype) runs on the device.  With ``correction='approx-firth'`` (the
    reference's default) the tests whose p-value is below ``pvalue_threshold`` are then refitted with the approximate
    Firth method on the host, one variant at a time, exactly as the reference does (log_reg.py:330-348,
    approx_firth.py); it needs the genotype values on the host, so it takes pandas / array / ``.bed`` input but not
    blocks that only live in device memory (``ResidentSource``, ``BgenSource``).
    """
    if correction not in (correction_none, correction_approx_firth):
        raise ValueError(f"Only supported correction methods are '{correction_none}' and '{correction_approx_firth}'")
    _validate_binary(covariate_df, phenotype_df, intersect_samples)
    np_dt = gwas_fx._regression_dtype(dt)
    if isinstance(values_column, str) and values_column == gwas_fx._GENOTYPES_COLUMN_NAME:
        raise ValueError(f'The values column should not be called "{gwas_fx._GENOTYPES_COLUMN_NAME}"')
    gt_indices_to_drop = None
    _covs = covariate_df
    if intersect_samples:  # log_reg.py:117-124
        if not genotype_sample_ids:
            raise ValueError('genotype_sample_ids required when intersect_samples enabled')
        if not covariate_df.empty:
            _covs = covariate_df.loc[phenotype_df.index, :]
        gt_indices_to_drop = _get_indices_to_drop(phenotype_df, genotype_sample_ids)
        if not offset_df.empty:
            if offset_df.index.nlevels == 1:
                offset_df = offset_df.reindex(phenotype_df.index)
            elif offset_df.index.nlevels == 2:
                offset_df = offset_df[offset_df.index.get_level_values(0).isin(phenotype_df.index)]

    Cmat = _covs.to_numpy(np.float64, copy=True)
    if add_intercept:
        Cmat = gwas_fx._add_intercept(Cmat, phenotype_df.shape[0])
    if Cmat.size == 0:
        raise ValueError('logistic_regression needs at least one covariate column (or add_intercept=True)')
    firth = correction == correction_approx_firth
    Q = np.linalg.qr(Cmat)[0] if firth else None  # log_reg.py:133-136
    Y = phenotype_df.to_numpy(np.float64, copy=True)
    Y_mask = ~np.isnan(Y)
    np.nan_to_num(Y, copy=False)
    state = _create_log_reg_state(phenotype_df, offset_df, Cmat, contigs, correction, add_intercept)
    phenotype_names = phenotype_df.columns.to_series().astype('str')

    states = list(state.values()) if isinstance(state, dict) else [state]
    contig_index = {c: i for i, c in enumerate(state)} if isinstance(state, dict) else None
    n_geno = len(genotype_sample_ids) if (intersect_samples and genotype_sample_ids) else phenotype_df.shape[0]
    keep = None
    if gt_indices_to_drop is not None and np.size(gt_indices_to_drop):
        k = np.ones(n_geno, dtype=bool)
        k[np.asarray(gt_indices_to_drop, dtype=np.int64)] = False
        keep = np.flatnonzero(k).astype(np.int64)
    engine = LogRegDeviceState(Cmat, states, keep_idx=keep, n_geno_samples=n_geno, device=device, Q=Q)

    from ..sources import GenotypeSource

    def work():
        if isinstance(genotype_df, GenotypeSource):
            for meta, block in genotype_df.blocks():
                yield meta, block
            return
        blocks = [genotype_df] if isinstance(genotype_df, pd.DataFrame) else genotype_df
        for pdf in blocks:
            pdf = gwas_fx._prepare_genotype_pdf(pdf, values_column, np_dt)
            values = pdf[_VALUES_COLUMN_NAME]
            yield pdf.drop(columns=[_VALUES_COLUMN_NAME]), as_block(values.array if hasattr(values, 'array') else values, np_dt)

    parts = []
    try:
        for meta, block in work():
            if contig_index is None:
                groups = [(meta, block, 0)]
            else:  # _loco_dispatch (functions.py:92-101): contigs in first-appearance order
                groups = []
                for contig in pd.unique(meta['contigName']):
                    sel = (meta['contigName'] == contig).to_numpy()
                    if sel.all():
                        sub = block
                    elif isinstance(block, ResidentBlock):
                        raise ValueError('a resident block may not span several contigs')
                    else:
                        sub = GenotypeBlock(np.ascontiguousarray(block.data[sel]), block.format, block.missing)
                    groups.append((meta[sel], sub, contig_index[contig]))
            for m, blk, ci in groups:
                res = engine.run(blk, ci)
                if firth:
                    _apply_firth(res, blk, states[ci], Y, Y_mask, Q, keep, n_geno, pvalue_threshold)
                parts.append(_assemble(m, res, phenotype_names, np_dt))
    finally:
        engine.close()
    result_cols = (['effect', 'stderror', 'correctionSucceeded'] if firth else []) + ['chisq', 'pvalue', 'phenotype']  # log_reg.py:95-108
    if not parts:
        return pd.DataFrame(columns=result_cols)
    out = pd.concat(parts) if len(parts) > 1 else parts[0]
    passthrough = [c for c in out.columns if c not in result_cols]
    return out[passthrough + result_cols]


def _apply_firth(res, block, state, Y, Y_mask, Q, keep, n_geno, pvalue_threshold):
    """log_reg.py:330-348: tests below the threshold are refitted with the approximate Firth method (host)."""
    P, G = res['chisq'].shape
    res['effect'] = np.full((P, G), np.nan)
    res['stderror'] = np.full((P, G), np.nan)
    res['correctionSucceeded'] = np.full((P, G), None, dtype=object)
    flagged = np.argwhere(res['pvalue'] < pvalue_threshold)
    if not len(flagged):
        return
    X = None
    cache = {}
    for p, g in flagged:
        if g not in cache:
            if X is None:
                X = _decode_for_firth(block, n_geno, keep)
            x = X[g]
            cache[g] = x - Q @ (Q.T @ x)  # the linear residualisation of log_reg.py:312-313
        fit = af.correct_approx_firth(cache[g], Y[:, p], state.firth_offset[:, p], Y_mask[:, p])
        if fit is None:
            res['correctionSucceeded'][p, g] = False
        else:
            res['correctionSucceeded'][p, g] = True
            res['effect'][p, g], res['stderror'][p, g] = fit.effect, fit.stderror
            res['chisq'][p, g], res['pvalue'][p, g] = fit.chisq, fit.pvalue


def _decode_for_firth(block, n_geno, keep) -> np.ndarray:
    """float64 [G, N] genotype values of a host block as the regression sees them, in phenotype-row order: 2-bit decode
    (PlinkRowToInternalRowConverter.scala:40-47 + genotype_states), mean substitution over ALL genotype samples
    (MeanSubstitute.scala:57-84; the values column is evaluated before np.delete), then the dropped samples removed
    (log_reg.py:303-304)."""
    if isinstance(block, ResidentBlock):
        raise NotImplementedError("correction='approx-firth' needs the genotype values on the host: it does not take blocks "
                                  "that live only in device memory (ResidentSource / BgenSource)")
    data = block.data
    if block.format == 'plink2':
        codes = np.empty((data.shape[0], data.shape[1] * 4), dtype=np.uint8)
        for j in range(4):
            codes[:, j::4] = (data >> (2 * j)) & 3
        X = np.array([2.0, -1.0, 1.0, 0.0])[codes[:, :n_geno]]
    else:
        X = np.array(data, dtype=np.float64)
    if block.format in ('plink2', 'i8') and block.missing == 'mean':
        for row in X:
            miss = row == -1.0
            if miss.any() and not miss.all():
                row[miss] = row[~miss].sum() / (row.size - miss.sum())
    return X if keep is None else X[:, keep]


def _validate_binary(covariate_df, phenotype_df, intersect_samples):
    """functions.py:30-44 with is_binary=True (model_functions._check_binary: every observed value is 0 or 1)."""
    for col in covariate_df:
        gwas_fx._assert_all_present(covariate_df, col, 'covariate')
    if not covariate_df.empty and not intersect_samples and phenotype_df.shape[0] != covariate_df.shape[0]:
        raise ValueError('phenotype_df and covariate_df must have the same number of rows '
                         f'({phenotype_df.shape[0]} != {covariate_df.shape[0]}')
    if not ((~phenotype_df.isna()).sum() > covariate_df.shape[1]).all():
        raise ValueError('There must be more non-missing samples than covariates')
    vals = phenotype_df.to_numpy(np.float64)
    if not np.isin(vals[~np.isnan(vals)], (0.0, 1.0)).all():
        raise ValueError('Binary phenotypes must be coded as 0 / 1 (missing values as NaN)')


def _assemble(meta: pd.DataFrame, res, phenotype_names, out_dt) -> pd.DataFrame:
    """log_reg.py:322-328: pass-through columns replicated once per phenotype, statistics raveled phenotype-major."""
    G = meta.shape[0]
    P = res['chisq'].shape[0]
    cols = {c: np.tile(meta[c].to_numpy(), P) for c in meta.columns}
    if 'effect' in res:  # approx-firth columns (log_reg.py:95-101)
        cols['effect'] = np.ravel(res['effect']).astype(out_dt, copy=False)
        cols['stderror'] = np.ravel(res['stderror']).astype(out_dt, copy=False)
        cols['correctionSucceeded'] = np.ravel(res['correctionSucceeded'])
    cols['chisq'] = np.ravel(res['chisq']).astype(out_dt, copy=False)
    cols['pvalue'] = np.ravel(res['pvalue']).astype(out_dt, copy=False)
    cols['phenotype'] = np.repeat(np.asarray(list(phenotype_names), dtype=object), G)
    return pd.DataFrame(cols, index=np.tile(meta.index.to_numpy(), P), copy=False)
