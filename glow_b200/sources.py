"""Genotype block sources that feed the device without materialising float64 arrays on the host.

The reference reaches the block kernel through Spark rows of ``array<double>`` (functions.py:56-68);
its PLINK datasource (core/.../plink/PlinkFileFormat.scala) decodes the 2-bit .bed payload in the JVM
first.  Here a source hands the packed bytes straight to the device, where the decode
(PlinkRowToInternalRowConverter.scala:40-47,56-66 + genotype_states + optional mean_substitute) is
fused into the GEMM operand load.
"""
import os
from typing import Iterator, Optional, Sequence, Tuple

import numpy as np
import pandas as pd

from .engine import GenotypeBlock

PLINK_MAGIC = bytes([0x6c, 0x1b, 0x01])  # PlinkFileFormat.scala:159 (variant-major)


class GenotypeSource:
    """Iterable of (metadata frame, GenotypeBlock); the metadata columns pass through to the output."""

    def blocks(self) -> Iterator[Tuple[pd.DataFrame, GenotypeBlock]]:
        raise NotImplementedError


class ArraySource(GenotypeSource):
    """In-memory variant-major matrix ([G, N] float64/float32/int8, or packed uint8 with format='plink2')."""

    def __init__(self, data: np.ndarray, meta: Optional[pd.DataFrame] = None, format: Optional[str] = None,
                 missing: str = 'minus_one', block_variants: int = 8192):
        self.data = data
        if format is None:
            format = {np.dtype(np.float64): 'f64', np.dtype(np.float32): 'f32', np.dtype(np.int8): 'i8'}[data.dtype]
        self.format = format
        self.missing = missing
        self.meta = meta if meta is not None else pd.DataFrame(index=pd.RangeIndex(data.shape[0]))
        self.block_variants = int(block_variants)

    def blocks(self):
        G = self.data.shape[0]
        for g0 in range(0, G, self.block_variants):
            g1 = min(G, g0 + self.block_variants)
            yield self.meta.iloc[g0:g1], GenotypeBlock(np.ascontiguousarray(self.data[g0:g1]), self.format,
                                                       self.missing)


class PlinkBedSource(GenotypeSource):
    """Variant-major PLINK .bed payload (+ optional .bim metadata).

    ``mean_substitute=True`` gives ``mean_substitute(genotype_states(genotypes))`` semantics,
    ``False`` plain ``genotype_states`` (missing calls enter as -1)."""

    def __init__(self, bed, n_samples: int, meta: Optional[pd.DataFrame] = None, mean_substitute: bool = True,
                 block_variants: int = 8192):
        if isinstance(bed, (str, os.PathLike)):
            raw = np.fromfile(bed, dtype=np.uint8)
        else:
            raw = np.frombuffer(bed, dtype=np.uint8) if isinstance(bed, (bytes, bytearray, memoryview)) else np.asarray(
                bed, dtype=np.uint8)
        if raw.size < 3 or bytes(raw[:3]) != PLINK_MAGIC:
            raise ValueError('not a variant-major PLINK .bed payload (bad magic bytes)')
        self.n_samples = int(n_samples)
        self.row_bytes = (self.n_samples + 3) // 4  # PlinkFileFormat.scala:282-284
        body = raw[3:]
        if body.size % self.row_bytes:
            raise ValueError('.bed size is not a whole number of variant blocks for this sample count')
        self.packed = body.reshape(-1, self.row_bytes)
        G = self.packed.shape[0]
        self.meta = meta if meta is not None else pd.DataFrame({'variant_idx': np.arange(G)})
        if len(self.meta) != G:
            raise ValueError('metadata rows do not match the number of variants in the .bed payload')
        self.missing = 'mean' if mean_substitute else 'minus_one'
        self.block_variants = int(block_variants)

    @staticmethod
    def read_bim(path) -> pd.DataFrame:
        """.bim -> Glow's variant columns (contigName, names, position, start(0-based), alleles)."""
        bim = pd.read_csv(path, sep=r'\s+', header=None,
                          names=['contigName', 'name', 'position', 'pos1', 'alternate', 'reference'],
                          dtype={'contigName': str})
        return pd.DataFrame({
            'contigName': bim['contigName'],
            'names': bim['name'],
            'start': bim['pos1'] - 1,
            'end': bim['pos1'] - 1 + bim['reference'].astype(str).str.len(),
            'referenceAllele': bim['reference'],
            'alternateAllele': bim['alternate'],
        })

    def blocks(self):
        G = self.packed.shape[0]
        for g0 in range(0, G, self.block_variants):
            g1 = min(G, g0 + self.block_variants)
            yield self.meta.iloc[g0:g1], GenotypeBlock(np.ascontiguousarray(self.packed[g0:g1]), 'plink2',
                                                       self.missing)
