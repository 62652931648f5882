"""B200-native drop-in for ``glow.gwas.linear_regression`` (reference: python/glow/gwas/lin_reg.py).

Same signature, validation, output schema and row order.  The driver-side preparation
(lin_reg.py:118-163: QR of the covariates, phenotype centring / residualisation / scaling, per-contig
offsets) stays on the host in numpy, exactly as in the reference; everything inside the block loop
(``_linear_regression_inner``, lin_reg.py:203-254) runs in hand-written sm_100a CUDA through the C
ABI in include/glow_b200.h.  There is no CPU fallback for the block loop.

``genotype_df`` may be
  * a pandas DataFrame (one block) or an iterable of pandas DataFrames (a stream of blocks): the
    result is one pandas DataFrame;
  * a pyspark DataFrame (when pyspark is installed): the same block function runs under
    ``mapInPandas`` exactly like the reference (lin_reg.py:277-284) and a Spark DataFrame is returned;
  * a :class:`glow_b200.sources.GenotypeSource` (e.g. a PLINK .bed file): packed blocks go to the
    device undecoded and the 2-bit decode / mean substitution is fused into the GEMM operand load.
"""
from dataclasses import dataclass, field
from typing import Any, Dict, Iterable, Iterator, List, Optional, Sequence, Union

import numpy as np
import pandas as pd

from . import functions as gwas_fx
from .functions import _VALUES_COLUMN_NAME, _get_indices_to_drop, _get_contigs_from_loco_df
from ..engine import (DeviceGroup, DeviceState, GenotypeBlock, as_block, contiguous_blocks, STAT_NAMES, VERBOSE_NAMES)

__all__ = ['linear_regression']


@dataclass
class YState:
    """State that varies per contig (lin_reg.py:166-172) plus the device replica it lives in."""
    Y: np.ndarray
    YdotY: np.ndarray
    _engine: Optional[DeviceState] = field(default=None, repr=False, compare=False)
    _contig_index: int = 0

    def __getstate__(self):
        # the device replica never crosses a process boundary (Spark closure pickling,
        # lin_reg.py:277-284): each worker rebuilds it lazily in _linear_regression_inner
        return {'Y': self.Y, 'YdotY': self.YdotY, '_engine': None, '_contig_index': 0}

    def __setstate__(self, state):
        self.__dict__.update(state)


def linear_regression(genotype_df,
                      phenotype_df: pd.DataFrame,
                      covariate_df: pd.DataFrame = pd.DataFrame({}),
                      offset_df: pd.DataFrame = pd.DataFrame({}),
                      contigs: Optional[List[str]] = None,
                      add_intercept: bool = True,
                      values_column: Union[str, Any] = 'values',
                      dt: type = np.float64,
                      verbose_output: bool = False,
                      intersect_samples: bool = False,
                      genotype_sample_ids: Optional[List[str]] = None,
                      device: int = 0,
                      devices: Optional[Sequence[int]] = None,
                      options: Optional[dict] = None):
    """Linear-regression association test of every variant against every phenotype.

    Arguments, defaults, validation errors and the output columns are those of the reference
    (lin_reg.py:18-116).  Additions: ``device`` (CUDA ordinal); ``devices`` (several ordinals: the variant blocks are
    sharded over those GPUs -- contiguous block ranges when the source knows its size, round robin for streams -- the
    state is replicated once by an NCCL broadcast and the results are gathered in the single-GPU order, bit for bit
    the same rows); ``options`` (kernel-selection knobs of the C ABI, ``glr_options``; default: the library decides).

    Output columns: all columns of ``genotype_df`` except the values column and ``genotypes``,
    then ``effect, stderror, tvalue, pvalue, phenotype`` [+ ``n, sum_x, y_transpose_x`` when
    ``verbose_output``]; within a block rows are phenotype-major, variant-minor.
    """
    spark_input = gwas_fx._is_spark_df(genotype_df)
    if spark_input:
        gwas_fx._check_spark_version(genotype_df.sparkSession)

    gwas_fx._validate_covariates_and_phenotypes(covariate_df,
                                                phenotype_df,
                                                is_binary=False,
                                                intersect_samples=intersect_samples)
    np_dt = gwas_fx._regression_dtype(dt)
    if isinstance(values_column, str) and values_column == gwas_fx._GENOTYPES_COLUMN_NAME:
        raise ValueError(f'The values column should not be called "{gwas_fx._GENOTYPES_COLUMN_NAME}"')

    gt_indices_to_drop = None
    _covs = covariate_df
    if intersect_samples:  # lin_reg.py:131-140
        if not genotype_sample_ids:
            raise ValueError('genotype_sample_ids required when intersect_samples enabled')
        _covs = covariate_df.loc[phenotype_df.index, :]
        gt_indices_to_drop = _get_indices_to_drop(phenotype_df, genotype_sample_ids)
        if not offset_df.empty:
            if offset_df.index.nlevels == 1:
                offset_df = offset_df.reindex(phenotype_df.index)
            elif offset_df.index.nlevels == 2:
                offset_df = offset_df[offset_df.index.get_level_values(0).isin(phenotype_df.index)]

    C = _covs.to_numpy(np_dt, copy=True)
    if add_intercept:
        C = gwas_fx._add_intercept(C, phenotype_df.shape[0]).astype(np_dt, copy=False)

    # covariate basis and phenotype residuals (lin_reg.py:146-155)
    Q = _thin_q(C)
    Y = phenotype_df.to_numpy(np_dt, copy=True)
    Y_mask = (~np.isnan(Y)).astype(np_dt)
    Y = np.nan_to_num(Y, copy=False)
    Y_for_verbose_output = np.copy(Y) if verbose_output else None
    Y -= Y.mean(axis=0)
    Y = gwas_fx._residualize_in_place(Y, Q) * Y_mask
    Y_scale = np.sqrt(np.sum(Y**2, axis=0) / (Y_mask.sum(axis=0) - Q.shape[1]))
    Y /= Y_scale[None, :]

    Y_state = _create_YState(Y, phenotype_df, offset_df, Y_mask, np_dt, contigs)
    dof = C.shape[0] - C.shape[1] - 1  # lin_reg.py:159

    n_geno = len(genotype_sample_ids) if (intersect_samples and genotype_sample_ids) else phenotype_df.shape[0]
    placement = _Placement(device=int(device), devices=None if devices is None else [int(d) for d in devices],
                           options=options)
    if not spark_input:
        # (under Spark the driver needs neither the library nor a GPU: every Python worker builds its own replica
        # lazily in _linear_regression_inner, once per process)
        _attach_engine(Y_state, Y_mask, Y_scale, Q, dof, Y_for_verbose_output, gt_indices_to_drop, n_geno, placement)

    return _generate_linreg_output(genotype_df, np_dt, values_column, Y_state, Y_mask, Y_scale, Q, dof, phenotype_df,
                                   Y_for_verbose_output, verbose_output, gt_indices_to_drop, placement)


def _thin_q(C: np.ndarray) -> np.ndarray:
    """``np.linalg.qr(C)[0]`` (lin_reg.py:147).  For tall matrices (UK-Biobank shape: 500 000 x 17, 0.22 s in LAPACK's
    panel factorisation on one core) the same Householder QR is done as a TSQR: row chunks factorised in parallel
    threads, then the stacked R factors; Q = blockdiag(Q_i) Q2.  Q spans the same space with orthonormal columns (the
    columns may differ from numpy's in sign, which no quantity of the regression depends on: Q only enters as Q Q^T)."""
    n, k = C.shape
    if n < 100_000 or k == 0 or n < 64 * k:
        return np.linalg.qr(C)[0]
    import os
    from concurrent.futures import ThreadPoolExecutor
    parts = min(16, max(2, os.cpu_count() or 2))
    bounds = [(n * i) // parts for i in range(parts + 1)]
    with ThreadPoolExecutor(parts) as pool:
        qs, rs = zip(*pool.map(lambda i: np.linalg.qr(C[bounds[i]:bounds[i + 1]]), range(parts)))
        q2 = np.linalg.qr(np.vstack(rs))[0]
        Q = np.empty_like(C)

        def fill(i):
            np.matmul(qs[i], q2[i * k:(i + 1) * k], out=Q[bounds[i]:bounds[i + 1]])

        list(pool.map(fill, range(parts)))
    return Q


def _create_YState(Y, phenotype_df, offset_df, Y_mask, dt, contigs) -> Union[YState, Dict[str, YState]]:
    """lin_reg.py:175-188."""
    offset_type = gwas_fx._validate_offset(phenotype_df, offset_df)
    if offset_type != gwas_fx._OffsetType.LOCO_OFFSET:
        return _create_one_YState(Y, phenotype_df, offset_df, Y_mask, dt)
    if contigs is None:
        contigs = _get_contigs_from_loco_df(offset_df)
    return {
        contig: _create_one_YState(Y, phenotype_df, offset_df.xs(contig, level=1), Y_mask, dt)
        for contig in contigs
    }


def _create_one_YState(Y, phenotype_df, offset_df, Y_mask, dt) -> YState:
    """lin_reg.py:191-199.  The offset is subtracted from the residualised, scaled phenotypes and
    the result re-masked; it is NOT residualised or rescaled."""
    if not offset_df.empty:
        base_Y = pd.DataFrame(Y, phenotype_df.index, phenotype_df.columns)
        Yc = (base_Y - offset_df).reindex(phenotype_df.index).to_numpy(dt).copy()
    else:
        Yc = np.array(Y, dtype=dt, copy=True)
    Yc *= Y_mask
    return YState(Yc, np.sum(Yc * Yc, axis=0))


def _keep_index(n_geno: int, gt_indices_to_drop) -> Optional[np.ndarray]:
    if gt_indices_to_drop is None or not np.size(gt_indices_to_drop):
        return None
    keep = np.ones(n_geno, dtype=bool)
    keep[np.asarray(gt_indices_to_drop, dtype=np.int64)] = False
    return np.flatnonzero(keep).astype(np.int64)


@dataclass
class _Placement:
    """Where the block loop runs: one CUDA ordinal, or several (sharded), plus the kernel knobs.  Travels with the
    closure to the Spark workers (``spark_worker``): there every Python worker drives ONE GPU -- the one Spark's GPU
    scheduling assigned to the task if any, else ``devices[partition % len(devices)]``, else ``device``."""
    device: int = 0
    devices: Optional[List[int]] = None
    options: Optional[dict] = None
    spark_worker: bool = False

    def resolve(self) -> List[int]:
        if not self.spark_worker:
            return list(self.devices) if self.devices else [self.device]
        try:
            from pyspark import TaskContext
            ctx = TaskContext.get()
            res = ctx.resources() if ctx is not None else {}
            if 'gpu' in res and res['gpu'].addresses:
                return [int(res['gpu'].addresses[0])]
            if self.devices and ctx is not None:
                return [self.devices[ctx.partitionId() % len(self.devices)]]
        except Exception:  # noqa: BLE001
            pass
        return [self.devices[0] if self.devices else self.device]


def _attach_engine(Y_state, Y_mask, Y_scale, Q, dof, Y_for_verbose_output, gt_indices_to_drop, n_geno, placement=None):
    """Builds the device replica(s) of (Q, Y per contig, mask, scale, dof): what the reference ships to
    every Spark worker by closure capture (lin_reg.py:277-284).  ONE state holds every contig's Y."""
    placement = placement or _Placement()
    states = list(Y_state.values()) if isinstance(Y_state, dict) else [Y_state]
    keep = _keep_index(n_geno, gt_indices_to_drop)
    if keep is not None and keep.size != states[0].Y.shape[0]:
        raise ValueError('after dropping the samples absent from phenotype_df the genotype arrays must have one '
                         'entry per phenotype row')
    kw = dict(Q=np.asarray(Q, dtype=np.float64),
              Y=np.stack([np.asarray(s.Y, dtype=np.float64) for s in states]),
              Y_mask=np.asarray(Y_mask, dtype=np.float64),
              YdotY=np.stack([np.asarray(s.YdotY, dtype=np.float64) for s in states]),
              Y_scale=np.asarray(Y_scale, dtype=np.float64),
              dof=int(dof),
              Y_raw=None if Y_for_verbose_output is None else np.asarray(Y_for_verbose_output, np.float64),
              keep_idx=keep,
              n_geno_samples=n_geno,
              options=placement.options)
    devs = placement.resolve()
    engine = DeviceGroup(devs, **kw) if len(devs) > 1 else DeviceState(device=devs[0], **kw)
    for i, s in enumerate(states):
        s._engine = engine
        s._contig_index = i
    return engine


def _assemble(genotype_pdf: pd.DataFrame, res: Dict[str, np.ndarray], phenotype_names, verbose_output, out_dt):
    """Result frame of one block (lin_reg.py:231-252): pass-through columns replicated once per
    phenotype (``pd.concat([genotype_pdf] * P)``), statistics raveled phenotype-major
    (``np.ravel`` of the P x G_b matrices).  Built column-wise with numpy instead of concatenating
    P frames and converting every matrix to a Python list."""
    num_genotypes = genotype_pdf.shape[0]
    P = res['effect'].shape[0]
    cols = {c: np.tile(genotype_pdf[c].to_numpy(), P) for c in genotype_pdf.columns}
    if verbose_output:
        cols['n'] = np.ravel(res['n']).astype(np.int32)
        cols['sum_x'] = np.ravel(res['sum_x']).astype(out_dt, copy=False)
        cols['y_transpose_x'] = np.ravel(res['y_transpose_x']).astype(out_dt, copy=False)
    for k in STAT_NAMES:
        cols[k] = np.ravel(res[k]).astype(out_dt, copy=False)
    cols['phenotype'] = np.repeat(np.asarray(list(phenotype_names), dtype=object), num_genotypes)
    return pd.DataFrame(cols, index=np.tile(genotype_pdf.index.to_numpy(), P), copy=False)


def _linear_regression_inner(genotype_pdf: pd.DataFrame, Y_state: YState, Y_mask, Y_scale, Q, dof: int,
                             phenotype_names: pd.Series, Y_for_verbose_output, verbose_output,
                             gt_indices_to_drop) -> pd.DataFrame:
    """Same contract as the reference's block kernel (lin_reg.py:203-254): one pandas block with a
    ``_glow_regression_values`` column in, the result frame out.  The arithmetic of lin_reg.py:226-249
    runs on the device; if ``Y_state`` has no device replica yet (callers that build the state by
    hand, like the reference's unit tests), one is created and cached on it."""
    values = genotype_pdf[_VALUES_COLUMN_NAME]
    first = values.iloc[0] if len(values) else None
    if isinstance(first, GenotypeBlock):
        block = first
    else:
        block = as_block(values.array if hasattr(values, 'array') else values, np.asarray(Y_scale).dtype)
    n_geno = block.data.shape[1] if block.format != 'plink2' else None
    if Y_state._engine is None:
        n_rows = Y_state.Y.shape[0]
        n_drop = 0 if gt_indices_to_drop is None else int(np.size(gt_indices_to_drop))
        _attach_engine(Y_state, Y_mask, Y_scale, Q, dof, Y_for_verbose_output if verbose_output else None,
                       gt_indices_to_drop, n_rows + n_drop)
    engine = Y_state._engine
    res = (engine.states[0] if isinstance(engine, DeviceGroup) else engine).run(block, Y_state._contig_index)
    rest = genotype_pdf.drop(columns=[_VALUES_COLUMN_NAME])
    out_dt = np.asarray(Y_scale).dtype
    return _assemble(rest, res, phenotype_names, bool(verbose_output), out_dt)


def _output_columns(input_columns, verbose_output) -> List[str]:
    """Column order of the result (lin_reg.py:260-275, functions.py:24-27)."""
    cols = [c for c in input_columns if c not in (_VALUES_COLUMN_NAME, gwas_fx._GENOTYPES_COLUMN_NAME)]
    if verbose_output:
        cols += ['n', 'sum_x', 'y_transpose_x']
    return cols + ['effect', 'stderror', 'tvalue', 'pvalue', 'phenotype']


_WORKER_ENGINES: Dict[str, Any] = {}  # Spark Python worker: token of a linear_regression call -> its device replica


def _generate_linreg_output(genotype_df, np_dt, values_column, Y_state, Y_mask, Y_scale, Q, dof, phenotype_df,
                            Y_for_verbose_output, verbose_output, gt_indices_to_drop, placement=None):
    """lin_reg.py:257-284."""
    phenotype_names = phenotype_df.columns.to_series().astype('str')
    args = (Y_mask, Y_scale, Q, dof, phenotype_names, Y_for_verbose_output, verbose_output, gt_indices_to_drop)

    if gwas_fx._is_spark_df(genotype_df):
        import uuid
        token = uuid.uuid4().hex
        pl = placement or _Placement()
        worker_placement = _Placement(device=pl.device, devices=pl.devices, options=pl.options, spark_worker=True)

        def map_func(pdf_iterator):
            # one replica per Python worker process and call: ONE state with every contig's Y, on the GPU of this task
            first = next(iter(Y_state.values())) if isinstance(Y_state, dict) else Y_state
            if first._engine is None:
                cached = _WORKER_ENGINES.get(token)
                if cached is None:
                    n_rows = first.Y.shape[0]
                    n_drop = 0 if gt_indices_to_drop is None else int(np.size(gt_indices_to_drop))
                    cached = _attach_engine(Y_state, Y_mask, Y_scale, Q, dof,
                                            Y_for_verbose_output if verbose_output else None, gt_indices_to_drop,
                                            n_rows + n_drop, worker_placement)
                    while len(_WORKER_ENGINES) >= 2:
                        _WORKER_ENGINES.pop(next(iter(_WORKER_ENGINES))).close()
                    _WORKER_ENGINES[token] = cached
                else:
                    for i, s in enumerate(Y_state.values() if isinstance(Y_state, dict) else [Y_state]):
                        s._engine, s._contig_index = cached, i
            for pdf in pdf_iterator:
                yield gwas_fx._loco_dispatch(pdf, Y_state, _linear_regression_inner, *args)

        return _spark_map(genotype_df, np_dt, values_column, map_func, verbose_output)

    from ..sources import GenotypeSource
    if isinstance(genotype_df, GenotypeSource):
        return _run_source(genotype_df, Y_state, phenotype_names, verbose_output, np_dt)
    if _is_arrow_input(genotype_df):
        return _run_arrow(genotype_df, values_column, Y_state, phenotype_names, verbose_output, np_dt)

    blocks = [genotype_df] if isinstance(genotype_df, pd.DataFrame) else genotype_df
    return _run_frames(blocks, values_column, Y_state, phenotype_names, verbose_output, np_dt)


def _engine_of(Y_state):
    return (next(iter(Y_state.values())) if isinstance(Y_state, dict) else Y_state)._engine


def _split_by_contig(meta: pd.DataFrame, block: GenotypeBlock, Y_state):
    """_loco_dispatch (functions.py:92-101) for a packed block: one sub-block per contig present, contigs in
    first-appearance order, each with the index of that contig's Y inside the device state."""
    if not isinstance(Y_state, dict):
        yield meta, block, 0
        return
    contigs = pd.unique(meta['contigName'])
    if len(contigs) == 1:  # the usual case (variants sorted by chromosome): no gather, the block goes as is
        yield meta, block, Y_state[contigs[0]]._contig_index
        return
    for contig in contigs:
        sel = (meta['contigName'] == contig).to_numpy()
        if not isinstance(block, GenotypeBlock):
            raise ValueError('a resident block may not span several contigs (ResidentSource splits them at load time)')
        sub = GenotypeBlock(np.ascontiguousarray(block.data[sel]), block.format, block.missing)
        yield meta[sel], sub, Y_state[contig]._contig_index


def _ordered_columns(out: pd.DataFrame, verbose_output) -> pd.DataFrame:
    """Exact reference column order: pass-through columns, then results (verbose columns last)."""
    passthrough = [c for c in out.columns if c not in STAT_NAMES + VERBOSE_NAMES + ('phenotype', )]
    return out[passthrough + list(STAT_NAMES) + ['phenotype'] + (list(VERBOSE_NAMES) if verbose_output else [])]


def _run_work(engine, work, n_work_blocks, phenotype_names, verbose_output, np_dt, shard_ranges=None) -> pd.DataFrame:
    """Runs (meta, block, contig index) items through the device lanes -- the H2D copy of block i+1 overlaps the kernels
    of block i -- and assembles the result frames in input order.  ``work(first, last)`` yields the items of the source
    blocks [first, last).  With several GPUs and a known block count every GPU takes a contiguous range of blocks
    (SURVEY.md section 8e); otherwise the blocks go round robin.  Either way each block is exactly the block one GPU
    would have run, so the rows are the same bit for bit."""
    parts = []
    if isinstance(engine, DeviceGroup) and len(engine) > 1 and n_work_blocks is not None:
        n_dev = len(engine)
        metas = [[] for _ in range(n_dev)]

        ranges = shard_ranges(engine.devices) if shard_ranges else [contiguous_blocks(n_work_blocks, d, n_dev) for d in range(n_dev)]

        def feed(d):
            for meta, block, ci in work(*ranges[d]):
                metas[d].append(meta)
                yield block, ci

        done = [[] for _ in range(n_dev)]
        for d, seq, res in engine.run_shards([feed(d) for d in range(n_dev)]):
            done[d].append(_assemble(metas[d][seq], res, phenotype_names, verbose_output, np_dt))
            metas[d][seq] = None
        for d in range(n_dev):
            parts += done[d]
    else:
        metas = []

        def feed_all():
            for meta, block, ci in work(0, None):
                metas.append(meta)
                yield block, ci

        for res in engine.run_pipelined(feed_all()):
            parts.append(_assemble(metas.pop(0), res, phenotype_names, verbose_output, np_dt))
    if not parts:
        return pd.DataFrame(columns=_output_columns([], verbose_output))
    out = pd.concat(parts) if len(parts) > 1 else parts[0]
    return _ordered_columns(out, verbose_output)


def _run_frames(blocks, values_column, Y_state, phenotype_names, verbose_output, np_dt) -> pd.DataFrame:
    """pandas blocks (one frame or a stream of frames): what map_func does at lin_reg.py:277-282, pipelined."""
    n_blocks = len(blocks) if isinstance(blocks, (list, tuple)) else None

    def work(first, last):
        import itertools
        for pdf in itertools.islice(iter(blocks), first, last):
            pdf = gwas_fx._prepare_genotype_pdf(pdf, values_column, np_dt)
            values = pdf[_VALUES_COLUMN_NAME]
            head = values.iloc[0] if len(values) else None
            block = head if isinstance(head, GenotypeBlock) else as_block(values.array if hasattr(values, 'array') else values,
                                                                          np_dt)
            yield from _split_by_contig(pdf.drop(columns=[_VALUES_COLUMN_NAME]), block, Y_state)

    return _run_work(_engine_of(Y_state), work, n_blocks, phenotype_names, verbose_output, np_dt)


def _run_source(source, Y_state, phenotype_names, verbose_output, np_dt) -> pd.DataFrame:
    """Packed blocks (e.g. PLINK 2-bit) straight from a GenotypeSource."""

    def work(first, last):
        for meta, block in source.blocks(first, last):
            yield from _split_by_contig(meta, block, Y_state)

    engine = _engine_of(Y_state)
    ranges = getattr(source, 'shard_ranges', None)
    if ranges is not None and not isinstance(engine, DeviceGroup):
        ranges(getattr(engine, 'devices', [engine.device]))  # a resident source must live on the GPU that runs it
    return _run_work(engine, work, source.n_blocks(), phenotype_names, verbose_output, np_dt, ranges)


def _is_arrow_input(obj) -> bool:
    mod = type(obj).__module__ or ''
    return mod.startswith('pyarrow') and type(obj).__name__ in ('Table', 'RecordBatch', 'RecordBatchReader',
                                                                'RecordBatchFileReader')


def _run_arrow(data, values_column, Y_state, phenotype_names, verbose_output, np_dt):
    """Arrow-native intake: record batches whose values column is a
    list<double|float|int8> go to the device straight from the Arrow child buffer -- no pandas
    objects, no np.column_stack (lin_reg.py:226-227), no per-row Python lists on the way out
    (lin_reg.py:248-252).  Returns a pyarrow Table with the reference's column order."""
    import pyarrow as pa
    if not isinstance(values_column, str):
        raise ValueError('values_column must be a column name for Arrow input')
    if values_column == gwas_fx._GENOTYPES_COLUMN_NAME:
        raise ValueError(f'The values column should not be called "{gwas_fx._GENOTYPES_COLUMN_NAME}"')
    if isinstance(data, pa.Table):
        batches = data.to_batches()
        schema = data.schema
    elif isinstance(data, pa.RecordBatch):
        batches, schema = [data], data.schema
    else:
        batches, schema = data, data.schema
    if values_column not in schema.names:
        raise ValueError(f'genotype_df has no column named "{values_column}"')
    is_loco = isinstance(Y_state, dict)
    engine = _engine_of(Y_state)
    keep_cols = [n for n in schema.names if n not in (values_column, gwas_fx._GENOTYPES_COLUMN_NAME)]
    P = len(phenotype_names)
    names = np.asarray(list(phenotype_names), dtype=object)
    pending = []

    def work():
        for batch in batches:
            if batch.num_rows == 0:
                continue
            X = as_block(batch.column(schema.get_field_index(values_column)), np_dt)
            meta = batch.select(keep_cols)
            if is_loco:
                contig = batch.column(schema.get_field_index('contigName')).to_numpy(zero_copy_only=False)
                for c in pd.unique(contig):
                    sel = np.flatnonzero(contig == c)
                    sub = GenotypeBlock(np.ascontiguousarray(X.data[sel]), X.format, X.missing)
                    pending.append(meta.take(pa.array(sel)))
                    yield sub, Y_state[c]._contig_index
            else:
                pending.append(meta)
                yield X, 0

    out_batches = []
    for res in engine.run_pipelined(work()):
        meta = pending.pop(0)
        G = meta.num_rows
        idx = pa.array(np.tile(np.arange(G, dtype=np.int64), P))
        cols = [meta.column(i).take(idx) for i in range(meta.num_columns)]
        fields = [meta.schema.field(i) for i in range(meta.num_columns)]
        pa_dt = pa.float32() if np_dt == np.float32 else pa.float64()
        for k in STAT_NAMES:
            cols.append(pa.array(np.ravel(res[k]).astype(np_dt, copy=False)))
            fields.append(pa.field(k, pa_dt))
        cols.append(pa.array(np.repeat(names, G), type=pa.string()))
        fields.append(pa.field('phenotype', pa.string()))
        if verbose_output:
            cols.append(pa.array(np.ravel(res['n']).astype(np.int32)))
            fields.append(pa.field('n', pa.int32()))
            for k in ('sum_x', 'y_transpose_x'):
                cols.append(pa.array(np.ravel(res[k]).astype(np_dt, copy=False)))
                fields.append(pa.field(k, pa_dt))
        out_batches.append(pa.RecordBatch.from_arrays(cols, schema=pa.schema(fields)))
    if not out_batches:
        pa_dt = pa.float32() if np_dt == np.float32 else pa.float64()
        fields = [schema.field(n) for n in keep_cols] + [pa.field(k, pa_dt) for k in STAT_NAMES]
        fields.append(pa.field('phenotype', pa.string()))
        if verbose_output:
            fields += [pa.field('n', pa.int32()), pa.field('sum_x', pa_dt), pa.field('y_transpose_x', pa_dt)]
        return pa.schema(fields).empty_table()
    return pa.Table.from_batches(out_batches)


def _spark_map(genotype_df, np_dt, values_column, map_func, verbose_output):
    """Spark wiring of the reference (functions.py:56-68, lin_reg.py:259-284).  Only reachable when
    pyspark is installed; each Python worker loads libglow_b200.so and drives its own GPU."""
    from pyspark.sql import functions as fx
    from pyspark.sql.types import (ArrayType, DoubleType, FloatType, IntegerType, StringType, StructField,
                                   StructType)
    sql_type = FloatType() if np_dt == np.float32 else DoubleType()
    if isinstance(values_column, str):
        out = genotype_df.withColumn(_VALUES_COLUMN_NAME,
                                     fx.col(values_column).cast(ArrayType(sql_type))).drop(values_column)
    else:
        out = genotype_df.withColumn(_VALUES_COLUMN_NAME, values_column.cast(ArrayType(sql_type)))
    if gwas_fx._GENOTYPES_COLUMN_NAME in [f.name for f in genotype_df.schema]:
        out = out.drop(gwas_fx._GENOTYPES_COLUMN_NAME)
    result_fields = [
        StructField('effect', sql_type),
        StructField('stderror', sql_type),
        StructField('tvalue', sql_type),
        StructField('pvalue', sql_type),
        StructField('phenotype', StringType())
    ]
    if verbose_output:
        result_fields += [
            StructField('n', IntegerType()),
            StructField('sum_x', sql_type),
            StructField('y_transpose_x', sql_type)
        ]
    fields = [f for f in out.schema.fields if f.name != _VALUES_COLUMN_NAME] + result_fields
    return out.mapInPandas(map_func, StructType(fields))
