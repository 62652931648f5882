"""TEST INFRASTRUCTURE ONLY -- loader for the UNMODIFIED reference hot path.

Loads ``/root/reference/python/glow/gwas/{functions,lin_reg}.py`` byte-for-byte with
``importlib`` after pre-seeding ``sys.modules`` with stubs for the third-party imports that are
absent from this image (pyspark, nptyping, opt_einsum, the ``glow`` package ``__init__``).
Nothing from the reference is copied into this repo: the files are executed where they lie.

Only the golden-vector generator (``oracle/make_golden.py``) and the ``not gpu`` tests that pin
``oracle/linreg_oracle.py`` use this module, and only in the build container --
``/root/reference`` does not exist on the GPU box, where :func:`available` returns False.

The single arithmetic substitution is ``opt_einsum.contract`` -> ``np.einsum(optimize=True)``
(reference call site ``python/glow/gwas/functions.py:78-82``): same contraction, the summation
order may differ (measured <= 1e-15 relative).
"""
import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get('GLOW_REFERENCE_ROOT', '/root/reference')
_GWAS_DIR = os.path.join(REFERENCE_ROOT, 'python', 'glow', 'gwas')

_loaded = None


def available() -> bool:
    return os.path.isfile(os.path.join(_GWAS_DIR, 'lin_reg.py'))


class _Subscriptable(type):
    def __getitem__(cls, item):
        return cls


def _dummy_class(name):
    return _Subscriptable(name, (), {'__init__': lambda self, *a, **k: None})


def _stub_modules():
    mods = {}

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        mods[name] = m
        return m

    # pyspark: only used for annotations / schema objects outside the block kernel
    sql_types = mod('pyspark.sql.types',
                    **{
                        n: _dummy_class(n)
                        for n in ('StringType', 'StructField', 'IntegerType', 'StructType', 'FloatType',
                                  'DoubleType', 'ArrayType', 'BooleanType', 'DataType')
                    })
    sql_functions = mod('pyspark.sql.functions')
    sql = mod('pyspark.sql',
              Column=_dummy_class('Column'),
              DataFrame=_dummy_class('DataFrame'),
              SparkSession=_dummy_class('SparkSession'),
              functions=sql_functions,
              types=sql_types)
    mod('pyspark', sql=sql)
    # nptyping: annotations only
    mod('nptyping', **{n: _dummy_class(n) for n in ('Float', 'NDArray', 'Int32', 'Int', 'Shape', 'Bool')})
    # opt_einsum (python/setup.py:32 `opt_einsum>=3.2.0`, un-vendored)
    import contextlib
    mod('opt_einsum', contract=lambda s, *ops, **kw: np.einsum(s, *ops, optimize=True),
        shared_intermediates=contextlib.nullcontext)  # (a cache of intermediates: no effect on the values)

    # bare glow packages so glow/__init__.py (py4j, JVM) is never executed
    def _assert_all_present(df, col_name, df_name):
        # behaviour of python/glow/wgr/model_functions.py:479-489
        if df[col_name].isnull().any():
            raise ValueError(f"Missing values are present in the {df_name} dataframe's {col_name} column")

    def _check_binary(df):
        raise NotImplementedError('binary phenotypes are not on the linear-regression path')

    def _get_contigs_from_loco_df(df):
        # python/glow/wgr/wgr_functions.py:32-33
        return df.index.get_level_values(1).unique()

    glow = mod('glow')
    glow.__path__ = []
    gwas = mod('glow.gwas')
    gwas.__path__ = [_GWAS_DIR]
    wgr = mod('glow.wgr')
    wgr.__path__ = []
    mod('glow.wgr.model_functions', _assert_all_present=_assert_all_present, _check_binary=_check_binary)
    mod('glow.wgr.wgr_functions', _get_contigs_from_loco_df=_get_contigs_from_loco_df)
    glow.gwas = gwas
    glow.wgr = wgr
    return mods


def load():
    """Returns ``(functions_module, lin_reg_module)`` = the reference files, executed unmodified."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f'reference tree not found under {REFERENCE_ROOT}')
    saved = {}
    stubs = _stub_modules()
    for name, m in stubs.items():
        saved[name] = sys.modules.get(name)
        sys.modules[name] = m
    try:
        out = []
        for short in ('functions', 'lin_reg'):
            full = f'glow.gwas.{short}'
            spec = importlib.util.spec_from_file_location(full, os.path.join(_GWAS_DIR, short + '.py'))
            module = importlib.util.module_from_spec(spec)
            sys.modules[full] = module
            saved.setdefault(full, None)
            spec.loader.exec_module(module)
            setattr(sys.modules['glow.gwas'], short, module)
            out.append(module)
        _loaded = tuple(out)
    finally:
        # leave no stubs behind: a real pyspark (if ever installed) must not be shadowed
        for name, old in saved.items():
            if old is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = old
    return _loaded


def reference_linear_regression(genotype_pdf,
                                phenotype_df,
                                covariate_df=None,
                                offset_df=None,
                                contigs=None,
                                add_intercept=True,
                                values_column='values',
                                dt=np.float64,
                                verbose_output=False,
                                intersect_samples=False,
                                genotype_sample_ids=None):
    """Driver that follows ``lin_reg.py:118-163`` statement by statement, calling the reference's
    own helpers, and then feeds ONE pandas block through the reference's
    ``_linear_regression_inner`` (``lin_reg.py:203-254``).

    Restated (cannot run under pandas 3 / without Spark):
      * Spark wiring (``_prepare_genotype_df``, ``mapInPandas``): the block is a pandas frame.
      * ``_create_one_YState`` offset branch (``lin_reg.py:194-198``): ``to_numpy`` returns a
        read-only view under copy-on-write, so the array is copied before ``*= Y_mask``.
      * ``_loco_dispatch`` dict branch (``functions.py:98-99``): explicit loop over contigs in
        first-appearance order (= ``groupby(sort=False)``), results concatenated in that order.
    """
    import pandas as pd
    fx, lr = load()
    covariate_df = pd.DataFrame({}) if covariate_df is None else covariate_df
    offset_df = pd.DataFrame({}) if offset_df is None else offset_df

    fx._validate_covariates_and_phenotypes(covariate_df,
                                           phenotype_df,
                                           is_binary=False,
                                           intersect_samples=intersect_samples)
    if dt not in (np.float32, np.float64):
        raise ValueError('dt must be np.float32 or np.float64')
    if isinstance(values_column, str) and values_column == 'genotypes':
        raise ValueError('The values column should not be called "genotypes"')

    pdf = genotype_pdf.copy()
    pdf[fx._VALUES_COLUMN_NAME] = [np.asarray(v, dtype=dt) for v in pdf[values_column]]
    pdf = pdf.drop(columns=[values_column])
    if 'genotypes' in pdf.columns:
        pdf = pdf.drop(columns=['genotypes'])

    gt_indices_to_drop = None
    _covs = covariate_df
    if intersect_samples:
        if not genotype_sample_ids:
            raise ValueError('genotype_sample_ids required when intersect_samples enabled')
        _covs = covariate_df.loc[phenotype_df.index, :]
        gt_indices_to_drop = fx._get_indices_to_drop(phenotype_df, genotype_sample_ids)
        if not offset_df.empty:
            if offset_df.index.nlevels == 1:
                offset_df = offset_df.reindex(phenotype_df.index)
            elif offset_df.index.nlevels == 2:
                offset_df = offset_df[offset_df.index.get_level_values(0).isin(phenotype_df.index)]

    C = _covs.to_numpy(dt, copy=True)
    if add_intercept:
        C = fx._add_intercept(C, phenotype_df.shape[0])
    Q = np.linalg.qr(C)[0]
    Y = phenotype_df.to_numpy(dt, copy=True)
    Y_mask = (~np.isnan(Y)).astype(dt)
    Y = np.nan_to_num(Y, copy=False)
    Y_for_verbose_output = np.copy(Y) if verbose_output else None
    Y -= Y.mean(axis=0)
    Y = fx._residualize_in_place(Y, Q) * Y_mask
    Y_scale = np.sqrt(np.sum(Y**2, axis=0) / (Y_mask.sum(axis=0) - Q.shape[1]))
    Y /= Y_scale[None, :]

    offset_type = fx._validate_offset(phenotype_df, offset_df)

    def one_state(off):
        Yc = Y
        if not off.empty:
            base_Y = pd.DataFrame(Y, phenotype_df.index, phenotype_df.columns)
            Yc = (base_Y - off).reindex(phenotype_df.index).to_numpy(dt).copy()
        else:
            Yc = Y.copy()
        Yc *= Y_mask
        return lr.YState(Yc, np.sum(Yc * Yc, axis=0))

    if offset_type != fx._OffsetType.LOCO_OFFSET:
        Y_state = one_state(offset_df)
    else:
        if contigs is None:
            contigs = offset_df.index.get_level_values(1).unique()
        Y_state = {c: one_state(offset_df.xs(c, level=1)) for c in contigs}

    dof = C.shape[0] - C.shape[1] - 1
    names = phenotype_df.columns.to_series().astype('str')
    args = (Y_mask, Y_scale, Q, dof, names, Y_for_verbose_output, verbose_output, gt_indices_to_drop)
    if isinstance(Y_state, dict):
        parts = []
        seen = []
        for c in pdf['contigName']:
            if c not in seen:
                seen.append(c)
        for c in seen:
            sub = pdf[pdf['contigName'] == c].copy()
            parts.append(lr._linear_regression_inner(sub, Y_state[c], *args))
        return pd.concat(parts)
    return lr._linear_regression_inner(pdf, Y_state, *args)


# ----------------------------------------------------------------------------------------------
# logistic regression (python/glow/gwas/log_reg.py), score test
# ----------------------------------------------------------------------------------------------
_loaded_logreg = None


def _statsmodels_stub():
    """``statsmodels.api`` as far as log_reg.py uses it (log_reg.py:160-175): GLM(y, X, family=Binomial(), offset=...,
    missing='ignore').fit() and model.predict(params).  The fit is the restated IRLS of oracle/logreg_oracle.py
    (statsmodels itself is absent from this image)."""
    from oracle import logreg_oracle as lo

    class _Binomial:
        pass

    class _Fit:
        def __init__(self, params):
            self.params = params

    class GLM:
        def __init__(self, endog, exog, family=None, offset=None, missing='none'):
            self.y = np.asarray(endog, dtype=np.float64)
            self.X = np.asarray(exog, dtype=np.float64)
            self.offset = None if offset is None else np.asarray(offset, dtype=np.float64)

        def fit(self):
            self._params, self._mu = lo.glm_binomial_fit(self.y, self.X, self.offset)
            return _Fit(self._params)

        def predict(self, params):
            eta = self.X @ params + (0.0 if self.offset is None else self.offset)
            return 1.0 / (1.0 + np.exp(-eta))

    api = types.ModuleType('statsmodels.api')
    api.GLM = GLM
    api.families = types.SimpleNamespace(Binomial=_Binomial)
    sm = types.ModuleType('statsmodels')
    sm.api = api
    return {'statsmodels': sm, 'statsmodels.api': api}


def load_logreg():
    """Returns the reference's ``log_reg`` module, executed unmodified (with approx_firth.py beside it)."""
    global _loaded_logreg
    if _loaded_logreg is not None:
        return _loaded_logreg
    fx, _ = load()
    stubs = _stub_modules()
    stubs.update(_statsmodels_stub())
    stubs['glow.wgr.wgr_functions'].reshape_for_gwas = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError('Spark'))
    try:
        import typeguard  # noqa: F401
    except Exception:  # noqa: BLE001
        tg = types.ModuleType('typeguard')
        tg.typechecked = lambda f: f
        stubs['typeguard'] = tg
    saved = {}
    for name, m in stubs.items():
        saved[name] = sys.modules.get(name)
        sys.modules[name] = m
    sys.modules['glow.gwas.functions'] = fx
    setattr(sys.modules['glow.gwas'], 'functions', fx)
    saved.setdefault('glow.gwas.functions', None)
    try:
        mods = {}
        for short in ('approx_firth', 'log_reg'):
            full = f'glow.gwas.{short}'
            spec = importlib.util.spec_from_file_location(full, os.path.join(_GWAS_DIR, short + '.py'))
            module = importlib.util.module_from_spec(spec)
            sys.modules[full] = module
            saved.setdefault(full, None)
            spec.loader.exec_module(module)
            setattr(sys.modules['glow.gwas'], short, module)
            mods[short] = module
        _loaded_logreg = mods['log_reg']
    finally:
        for name, old in saved.items():
            if old is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = old
    return _loaded_logreg


def reference_score_test(genotype_matrix, phenotype_df, covariate_df, add_intercept=True, offset_df=None):
    """The reference's own ``run_score_test`` recipe (python/glow/gwas/tests/test_log_reg.py:10-32): state rows from
    ``_prepare_one_phenotype``, ``_pdf_to_log_reg_state``, then ONE block through ``_logistic_regression_inner``.
    genotype_matrix: [G, N].  Returns the reference's output frame."""
    import pandas as pd
    fx, _ = load()
    lr = load_logreg()
    C = covariate_df.to_numpy(copy=True)
    if add_intercept:
        C = fx._add_intercept(C, phenotype_df.shape[0])
    Y = phenotype_df.to_numpy(copy=True)
    Y_mask = ~np.isnan(Y)
    Y[~Y_mask] = 0
    rows = []
    for p in phenotype_df:
        d = {'label': p, 'values': phenotype_df[p].to_numpy(copy=True)}
        if offset_df is not None:
            d['offset'] = offset_df[p].to_numpy(copy=True)
        rows.append(lr._prepare_one_phenotype(C, pd.Series(d), lr.correction_none, add_intercept))
    names = phenotype_df.columns.to_series().astype('str')
    state = lr._pdf_to_log_reg_state(pd.DataFrame(rows), names, C.shape[1])
    values_df = pd.DataFrame({fx._VALUES_COLUMN_NAME: list(np.asarray(genotype_matrix, dtype=np.float64))})
    out = lr._logistic_regression_inner(values_df, state, C, Y, Y_mask, None, lr.correction_none, 0.05, names, None)
    return out, state


# ----------------------------------------------------------------------------------------------
# GloWGR apply step (python/glow/wgr/ridge_udfs.py::apply_model, model_functions.py)
# ----------------------------------------------------------------------------------------------
_loaded_wgr = None


def load_wgr():
    """Returns the reference's ``glow.wgr.ridge_udfs`` module (with ``model_functions`` behind it), executed unmodified.
    Stubs: pyspark (type/SQL-function names only, none is called by apply_model), nptyping, typeguard (identity
    decorator: the nptyping stand-ins are not real types)."""
    global _loaded_wgr
    if _loaded_wgr is not None:
        return _loaded_wgr
    if not available():
        raise RuntimeError(f'reference tree not found under {REFERENCE_ROOT}')
    stubs = _stub_modules()
    stubs['pyspark.sql'].Row = _dummy_class('Row')
    stubs['pyspark.sql.window'] = types.ModuleType('pyspark.sql.window')
    stubs['pyspark.sql.window'].Window = _dummy_class('Window')
    tg = types.ModuleType('typeguard')
    tg.typechecked = lambda f=None, **kw: f if f is not None else (lambda g: g)
    stubs['typeguard'] = tg
    wgr_dir = os.path.join(REFERENCE_ROOT, 'python', 'glow', 'wgr')
    saved = {}
    for name, m in stubs.items():
        saved[name] = sys.modules.get(name)
        sys.modules[name] = m
    try:
        mods = {}
        for short in ('model_functions', 'ridge_udfs'):
            full = f'glow.wgr.{short}'
            spec = importlib.util.spec_from_file_location(full, os.path.join(wgr_dir, short + '.py'))
            module = importlib.util.module_from_spec(spec)
            saved.setdefault(full, sys.modules.get(full))
            sys.modules[full] = module
            spec.loader.exec_module(module)
            mods[short] = module
        _loaded_wgr = mods['ridge_udfs']
    finally:
        for name, old in saved.items():
            if old is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = old
    return _loaded_wgr


def reference_firth_test(genotype_matrix, phenotype_df, covariate_df, pvalue_threshold=0.05, add_intercept=True, offset_df=None):
    """correction='approx-firth' through the reference: state rows from ``_prepare_one_phenotype`` (null Firth fit included),
    ONE block through ``_logistic_regression_inner`` -- which residualises the genotypes against Q and runs the score
    test -- and then the correction loop of log_reg.py:330-348 restated: under pandas 3 its chained assignments
    (``out_df.effect.iloc[i] = ...``) no longer write into ``out_df`` (copy-on-write), so the reference's own
    ``approx_firth.correct_approx_firth`` is applied here to the same rows it would have been applied to."""
    import warnings
    import pandas as pd
    fx, _ = load()
    lr = load_logreg()
    af = sys.modules.get('glow.gwas.approx_firth') or lr.af
    C = covariate_df.to_numpy(copy=True)
    if add_intercept:
        C = fx._add_intercept(C, phenotype_df.shape[0])
    Y = phenotype_df.to_numpy(copy=True)
    Y_mask = ~np.isnan(Y)
    Y[~Y_mask] = 0
    Q = np.linalg.qr(C)[0]
    rows = []
    for p in phenotype_df:
        d = {'label': p, 'values': phenotype_df[p].to_numpy(copy=True)}
        if offset_df is not None:
            d['offset'] = offset_df[p].to_numpy(copy=True)
        rows.append(lr._prepare_one_phenotype(C, pd.Series(d), lr.correction_approx_firth, add_intercept))
    names = phenotype_df.columns.to_series().astype('str')
    state = lr._pdf_to_log_reg_state(pd.DataFrame(rows), names, C.shape[1])
    Xg = np.asarray(genotype_matrix, dtype=np.float64)
    values_df = pd.DataFrame({fx._VALUES_COLUMN_NAME: list(Xg.copy())})
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        out = lr._logistic_regression_inner(values_df, state, C, Y, Y_mask, Q, lr.correction_approx_firth, pvalue_threshold,
                                            names, None)
    out = out.reset_index(drop=True)
    G = Xg.shape[0]
    X = Xg.T.copy()
    X = fx._residualize_in_place(X, Q)
    eff, se, ok = np.full(len(out), np.nan), np.full(len(out), np.nan), np.full(len(out), None, dtype=object)
    chisq, pv = out['chisq'].to_numpy(dtype=np.float64).copy(), out['pvalue'].to_numpy(dtype=np.float64).copy()
    score_p = pv.copy()
    for i in np.where(score_p < pvalue_threshold)[0]:
        snp, ph = i % G, int(i / G)
        fit = af.correct_approx_firth(X[:, snp], Y[:, ph], state.firth_offset[:, ph], Y_mask[:, ph])
        if fit is None:
            ok[i] = False
        else:
            ok[i] = True
            eff[i], se[i], chisq[i], pv[i] = fit.effect, fit.stderror, fit.chisq, fit.pvalue
    return pd.DataFrame({'effect': eff, 'stderror': se, 'correctionSucceeded': ok, 'chisq': chisq, 'pvalue': pv,
                         'phenotype': out['phenotype'].to_numpy()}), score_p
