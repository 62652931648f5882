"""Host-side helpers of the GWAS regression path, mirroring python/glow/gwas/functions.py of the
reference (same names, argument meaning and error behaviour); the arithmetic they fed in the
reference now runs on the device (glow_b200/engine.py)."""
from enum import Enum
from typing import Any, Callable, Dict, Iterable, List, Optional, Sequence, Union

import numpy as np
import pandas as pd

_VALUES_COLUMN_NAME = '_glow_regression_values'  # functions.py:14
_GENOTYPES_COLUMN_NAME = 'genotypes'  # functions.py:15


def _is_spark_df(obj) -> bool:
    mod = type(obj).__module__ or ''
    return mod.startswith('pyspark.') and hasattr(obj, 'mapInPandas')


def _check_spark_version(spark) -> None:
    """functions.py:18-21."""
    if int(spark.version.split('.')[0]) < 3:
        raise AttributeError('Pandas based regression tests are only supported on Spark 3.0 or greater')


def _assert_all_present(df: pd.DataFrame, col_name: Any, df_name: str) -> None:
    """python/glow/wgr/model_functions.py:479-489."""
    if df[col_name].isnull().any():
        raise ValueError(f"Missing values are present in the {df_name} dataframe's {col_name} column")


def _validate_covariates_and_phenotypes(covariate_df, phenotype_df, is_binary, intersect_samples=False):
    """functions.py:30-44."""
    for col in covariate_df:
        _assert_all_present(covariate_df, col, 'covariate')
    if not covariate_df.empty:
        if not intersect_samples and phenotype_df.shape[0] != covariate_df.shape[0]:
            raise ValueError('phenotype_df and covariate_df must have the same number of rows '
                             f'({phenotype_df.shape[0]} != {covariate_df.shape[0]}')
    if not ((~phenotype_df.isna()).sum() > covariate_df.shape[1]).all():
        raise ValueError('There must be more non-missing samples than covariates')
    if is_binary:
        raise NotImplementedError('binary phenotypes are outside the linear-regression path')


def _regression_dtype(dt):
    """functions.py:47-53 (the reference maps dt to a Spark SQL type; here it stays a numpy dtype)."""
    if dt == np.float32:
        return np.dtype(np.float32)
    elif dt == np.float64:
        return np.dtype(np.float64)
    raise ValueError('dt must be np.float32 or np.float64')


def _add_intercept(C: np.ndarray, num_samples: int) -> np.ndarray:
    """functions.py:72-75: the intercept is the first column."""
    intercept = np.ones((num_samples, 1))
    return np.hstack((intercept, C)) if C.size else intercept


def _have_same_elements(idx1: pd.Index, idx2: pd.Index) -> bool:
    return idx1.sort_values().equals(idx2.sort_values())


class _OffsetType(Enum):
    NO_OFFSET = 0
    SINGLE_OFFSET = 1  # per-sample offset for all contigs
    LOCO_OFFSET = 2  # per-sample, per-contig offset


def _get_contigs_from_loco_df(df: pd.DataFrame) -> pd.Index:
    """python/glow/wgr/wgr_functions.py:32-33."""
    return df.index.get_level_values(1).unique()


def _validate_offset(phenotype_df: pd.DataFrame, offset_df: pd.DataFrame) -> _OffsetType:
    """functions.py:110-132."""
    if offset_df.empty:
        return _OffsetType.NO_OFFSET
    if not _have_same_elements(phenotype_df.columns, offset_df.columns):
        raise ValueError('phenotype_df and offset_df should have the same column names.')
    if offset_df.index.nlevels == 1:
        if not _have_same_elements(phenotype_df.index, offset_df.index):
            raise ValueError('phenotype_df and offset_df should have the same index.')
        return _OffsetType.SINGLE_OFFSET
    if offset_df.index.nlevels == 2:
        for contig in _get_contigs_from_loco_df(offset_df):
            if not _have_same_elements(phenotype_df.index, offset_df.xs(contig, level=1).index):
                raise ValueError('When using a multi-indexed offset_df, the offsets for each contig '
                                 'should have the same index as phenotype_df')
        return _OffsetType.LOCO_OFFSET
    raise ValueError('offset_df must have one or two index levels')


def _residualize_in_place(M: np.ndarray, Q: np.ndarray) -> np.ndarray:
    """functions.py:136-143.  Host copy used only for the once-per-call phenotype preparation
    (lin_reg.py:153); genotype blocks are residualised on the device."""
    M -= Q @ (Q.T @ M)
    return M


def _get_indices_to_drop(phe_pdf, sample_ids) -> np.ndarray:
    """functions.py:146-152."""
    have = set(phe_pdf.index.values.astype(str).tolist())
    return np.array([i for i, s in enumerate(sample_ids) if s not in have], dtype=np.int64)


def _loco_dispatch(genotype_pdf: pd.DataFrame, state, f: Callable, *args):
    """functions.py:92-101: one kernel call per contig present in the block, contigs in
    first-appearance order (``groupby(sort=False)``), each with that contig's state."""
    if isinstance(state, dict):
        parts = []
        for contig in pd.unique(genotype_pdf['contigName']):
            sub = genotype_pdf[genotype_pdf['contigName'] == contig]
            parts.append(f(sub.copy(), state[contig], *args))
        return pd.concat(parts) if parts else f(genotype_pdf, next(iter(state.values())), *args)
    return f(genotype_pdf, state, *args)


def _prepare_genotype_pdf(pdf: pd.DataFrame, values_column, dt) -> pd.DataFrame:
    """pandas analogue of _prepare_genotype_df (functions.py:56-68): the values column is renamed
    to ``_glow_regression_values`` and a ``genotypes`` column, if present, is dropped."""
    if isinstance(values_column, str):
        if values_column == _GENOTYPES_COLUMN_NAME:
            raise ValueError(f'The values column should not be called "{_GENOTYPES_COLUMN_NAME}"')
        if values_column not in pdf.columns:
            raise ValueError(f'genotype_df has no column named "{values_column}"')
        out = pdf.rename(columns={values_column: _VALUES_COLUMN_NAME})
    elif callable(values_column):
        out = pdf.copy()
        out[_VALUES_COLUMN_NAME] = list(values_column(pdf))
    else:
        raise ValueError('values_column must be a column name (or, for pandas input, a callable pdf -> arrays)')
    if _GENOTYPES_COLUMN_NAME in out.columns:
        out = out.drop(columns=[_GENOTYPES_COLUMN_NAME])
    return out
