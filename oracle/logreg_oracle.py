"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy/scipy) of Glow's logistic-regression score test
(python/glow/gwas/log_reg.py, correction='none').  Parity oracle for ``glow_b200.gwas.logistic_regression``; only
``tests/`` may import it.

Parity status: PINNED, with one restated third-party step.
  * ``block_score_test`` keeps the reference's formulation -- the explicit N x G x P residual tensor of
    ``_logistic_residualize`` (log_reg.py:263-275) and the two einsums of log_reg.py:316-321 -- whereas the CUDA path
    works from Gram-type contractions, so agreement at 1e-9 is a real check.  ``tests/test_oracle_pins.py`` runs the
    reference's own ``log_reg.py`` byte-for-byte (``oracle/refshim.py::load_logreg``) against it.
  * The null-model fit is ``statsmodels.api.GLM(..., family=Binomial()).fit()`` in the reference (log_reg.py:160-175);
    statsmodels (python/environment.yml: statsmodels=0.13.5) is absent from this image, so its published algorithm --
    IRLS on the working response, start mu = (y + 0.5) / 2, stop when the deviance moves by < 1e-8 -- is restated in
    ``glm_binomial_fit``.  The estimate it converges to is the unique maximiser of the likelihood; it is checked
    against an independent Newton-Raphson solve of the score equations and, through the score statistic, against a
    direct per-variant restatement of the reference tests' ``statsmodels_baseline`` (test_log_reg.py:35-65) at the
    reference's own rtol 1e-3.
"""
from dataclasses import dataclass
from typing import Dict, Optional, Union

import numpy as np
from scipy import stats


def glm_binomial_fit(y, X, offset=None, max_iter=100, tol=1e-8):
    """statsmodels GLM.fit (IRLS) for the binomial family with logit link: returns (params, fitted mu)."""
    off = np.zeros(y.shape[0]) if offset is None else np.asarray(offset, dtype=np.float64)
    mu = (y + 0.5) / 2.0
    eta = np.log(mu / (1.0 - mu))
    dev_old = np.inf
    beta = np.zeros(X.shape[1])
    for it in range(max_iter + 2):
        w = mu * (1.0 - mu)
        z = eta - off + (y - mu) / w
        sw = np.sqrt(w)
        beta = np.linalg.lstsq(X * sw[:, None], z * sw, rcond=None)[0]
        eta = X @ beta + off
        mu = 1.0 / (1.0 + np.exp(-eta))
        with np.errstate(divide='ignore', invalid='ignore'):
            dev = -2.0 * np.sum(np.where(y > 0, np.log(mu), 0.0) * y + np.where(y < 1, np.log1p(-mu), 0.0) * (1.0 - y))
        if abs(dev - dev_old) < tol and it >= 1:
            # (two further steps: quadratic convergence puts the estimate at rounding level)
            for _ in range(2):
                w = mu * (1.0 - mu)
                z = eta - off + (y - mu) / w
                sw = np.sqrt(w)
                beta = np.linalg.lstsq(X * sw[:, None], z * sw, rcond=None)[0]
                eta = X @ beta + off
                mu = 1.0 / (1.0 + np.exp(-eta))
            break
        dev_old = dev
    return beta, mu


def null_model_predictions(y, X, mask, offset):
    """log_reg.py:160-175."""
    off = None if offset is None else np.asarray(offset, dtype=np.float64)[mask]
    _, mu = glm_binomial_fit(y[mask], X[mask, :], off)
    out = np.zeros(y.shape)
    out[mask] = mu
    return out


@dataclass
class OracleLogRegState:
    inv_CtGammaC: np.ndarray
    gamma: np.ndarray
    Y_res: np.ndarray


def prepare_one(C, y, offset):
    """log_reg.py:178-200."""
    mask = ~np.isnan(y)
    y_pred = null_model_predictions(y, C, mask, offset)
    y_res = np.nan_to_num(y - y_pred)
    gamma = y_pred * (1 - y_pred)
    return y_res, gamma, np.linalg.inv(C.T @ (gamma[:, None] * C))


def prepare_state(phenotype_df, covariate_df=None, offset_df=None, add_intercept=True, contigs=None):
    """log_reg.py:126-139, 223-275 -> (C, Y_mask, state or {contig: state})."""
    import pandas as pd
    covariate_df = pd.DataFrame({}) if covariate_df is None else covariate_df
    offset_df = pd.DataFrame({}) if offset_df is None else offset_df
    C = covariate_df.to_numpy(np.float64, copy=True)
    if add_intercept:
        ones = np.ones((phenotype_df.shape[0], 1))
        C = np.hstack((ones, C)) if C.size else ones
    Y = phenotype_df.to_numpy(np.float64, copy=True)
    Y_mask = ~np.isnan(Y)

    def one_state(off_df):
        rows = [prepare_one(C, Y[:, j], None if off_df is None else off_df[name].reindex(phenotype_df.index).to_numpy(np.float64))
                for j, name in enumerate(phenotype_df.columns)]
        return OracleLogRegState(np.stack([r[2] for r in rows]), np.column_stack([r[1] for r in rows]),
                                 np.column_stack([r[0] for r in rows]))

    if offset_df.empty:
        return C, Y_mask, one_state(None)
    if offset_df.index.nlevels == 1:
        return C, Y_mask, one_state(offset_df)
    cs = list(offset_df.index.get_level_values(1).unique()) if contigs is None else list(contigs)
    return C, Y_mask, {c: one_state(offset_df.xs(c, level=1)) for c in cs}


def block_score_test(X, C, Y_mask, state: OracleLogRegState):
    """log_reg.py:263-275, 316-321.  X: [G, N] (variant-major).  Returns P x G arrays."""
    Xs = np.ascontiguousarray(X.T, dtype=np.float64)  # N x G, as np.column_stack builds it (log_reg.py:301-302)
    m = Y_mask.astype(np.float64)
    X_hat = np.einsum('ic,pcd,ds,sp,sg,sp->igp', C, state.inv_CtGammaC, C.T, state.gamma, Xs, m, optimize=True)
    X_res = Xs[:, :, None] - X_hat
    num = np.einsum('sgp,sp->pg', X_res, state.Y_res, optimize=True)**2
    denom = np.einsum('sgp,sgp,sp->pg', X_res, X_res, state.gamma, optimize=True)
    with np.errstate(divide='ignore', invalid='ignore'):
        chisq = num / denom
    return {'chisq': chisq, 'pvalue': stats.chi2.sf(chisq, 1)}


def direct_score_test(x, y, C, offset=None):
    """Independent formulation for one (variant, phenotype): the Rao score test for adding x to the fitted null model,
    as ``statsmodels GLM.score_test(params, exog_extra=x)`` computes it (test_log_reg.py:52-56): U^2 / Var(U) with
    U = x^T (y - mu) and Var(U) the Schur complement of the information matrix."""
    mask = ~np.isnan(y)
    yo, Co, xo = y[mask], C[mask], x[mask]
    off = None if offset is None else offset[mask]
    _, mu = glm_binomial_fit(yo, Co, off)
    w = mu * (1 - mu)
    U = xo @ (yo - mu)
    Ixx = xo @ (w * xo)
    Ixc = Co.T @ (w * xo)
    Icc = Co.T @ (w[:, None] * Co)
    var = Ixx - Ixc @ np.linalg.solve(Icc, Ixc)
    chisq = U * U / var
    return chisq, stats.chi2.sf(chisq, 1)
