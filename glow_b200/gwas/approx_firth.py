"""Approximate Firth correction of the logistic score test: the per-significant-variant CPU refinement of the reference
(python/glow/gwas/approx_firth.py, after regenie).  Host numpy, as in the reference: the device scans every variant
with the score test, and only the tests below ``pvalue_threshold`` are refitted here, one N-vector at a time.

Firth's penalised likelihood: l*(beta) = l(beta) + 0.5 log det I(beta).  Newton iterations on the modified score
U*(beta) = X^T (y - pi + h (0.5 - pi)), h = diag of the hat matrix, with step capping and step halving
(approx_firth.py:54-124).  "Approximate": the covariate effects are fitted once per phenotype (null fit) and enter the
per-variant one-parameter fits as a fixed offset (approx_firth.py:128-172, 176-206)."""
from dataclasses import dataclass
from typing import Optional

import numpy as np
from scipy import stats
from scipy.special import expit


@dataclass
class FirthStatistics:
    effect: float
    stderror: float
    chisq: float
    pvalue: float


def _penalised(beta, X, y, offset):
    """(pi, information matrix, -2 x penalised log-likelihood): approx_firth.py:38-49."""
    pi = expit(X @ beta + offset)
    info = X.T @ ((pi * (1 - pi))[:, None] * X)
    _, logdet = np.linalg.slogdet(info)
    loglik = np.sum(y * np.log(pi) + (1 - y) * np.log(1 - pi))
    return pi, info, -2.0 * (loglik + 0.5 * logdet)


def fit_firth(beta_init, X, y, offset, convergence_limit=1e-5, deviance_tolerance=1e-6, max_iter=250, max_step_size=5,
              max_half_steps=25):
    """approx_firth.py:52-124.  Returns (beta, pi, info, deviance) or None when the iteration limit is reached."""
    beta = np.array(beta_init, dtype=np.float64, copy=True)
    pi, info, dev = _penalised(beta, X, y, offset)
    n_iter = 0
    while n_iter < max_iter:
        inv_info = np.linalg.pinv(info)
        root = np.sqrt(pi * (1 - pi))[:, None] * X
        h = np.sum((root @ inv_info) * root, axis=1)
        U = X.T @ (y - pi + h * (0.5 - pi))
        delta = inv_info @ U
        mx = np.amax(np.abs(delta)) / max_step_size
        if mx > 1:
            delta = delta / mx
        halvings = 0
        while halvings < max_half_steps:
            new = _penalised(beta + delta, X, y, offset)
            if new[2] < dev + deviance_tolerance:
                break
            delta = delta / 2
            halvings += 1
        beta = beta + delta
        pi, info, dev = new
        if np.amax(np.abs(U)) < convergence_limit:
            break
        n_iter += 1
    if n_iter == max_iter:
        return None
    return beta, pi, info, dev


def perform_null_firth_fit(y, C, mask, offset, includes_intercept):
    """approx_firth.py:128-172: covariate effects of the Firth null model as a per-sample offset (0 where unobserved)."""
    firth_offset = np.zeros(y.shape)
    off = np.zeros(y.shape) if offset is None else np.asarray(offset, dtype=np.float64)
    ym, Cm, om = y[mask], C[mask, :], off[mask]
    b0 = np.zeros(C.shape[1])
    if includes_intercept:
        p0 = (0.5 + ym.sum()) / (mask.sum() + 1)
        b0[0] = np.log(p0 / (1 - p0)) - om.mean()
    fit = fit_firth(b0, Cm, ym, om)
    if fit is None:
        raise ValueError('Null fit failed!')
    firth_offset[mask] = om + Cm @ fit[0]
    return firth_offset


def correct_approx_firth(x, y, firth_offset, mask) -> Optional[FirthStatistics]:
    """approx_firth.py:176-206: one-parameter Firth fit of the variant with the null-model offset, likelihood-ratio
    statistic against beta = 0, standard error from the information matrix at the fit."""
    ym, Xm, om = y[mask], x[mask, None], firth_offset[mask]
    b0 = np.zeros(1)
    fit = fit_firth(b0, Xm, ym, om)
    if fit is None:
        return None
    beta, _, info, dev = fit
    null_dev = _penalised(b0, Xm, ym, om)[2]
    chisq = null_dev - dev
    return FirthStatistics(effect=beta.item(), stderror=float(np.sqrt(np.linalg.pinv(info).diagonal()[-1])), chisq=float(chisq),
                           pvalue=float(stats.chi2.sf(chisq, 1)))
