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
                 block_variants: int = 8192, read_threads: Optional[int] = None):
        self.n_samples = int(n_samples)
        self.row_bytes = (self.n_samples + 3) // 4  # PlinkFileFormat.scala:282-284
        self._path = None
        if isinstance(bed, (str, os.PathLike)):
            # a file is STREAMED: blocks() reads each block straight into a pinned staging buffer (a ring of three, so
            # that two blocks can be in flight on the device lanes), never the whole payload into pageable memory
            self._path = os.fspath(bed)
            size = os.path.getsize(self._path)
            with open(self._path, 'rb') as f:
                magic = f.read(3)
            if size < 3 or magic != PLINK_MAGIC:
                raise ValueError('not a variant-major PLINK .bed payload (bad magic bytes)')
            if (size - 3) % self.row_bytes:
                raise ValueError('.bed size is not a whole number of variant blocks for this sample count')
            G = (size - 3) // self.row_bytes
            self.packed = np.memmap(self._path, dtype=np.uint8, mode='r', offset=3, shape=(G, self.row_bytes)) if G else \
                np.empty((0, self.row_bytes), dtype=np.uint8)
        else:
            raw = np.frombuffer(bed, dtype=np.uint8) if isinstance(bed, (bytes, bytearray, memoryview)) else np.asarray(
                bed, dtype=np.uint8)
            if raw.size < 3 or bytes(raw[:3]) != PLINK_MAGIC:
                raise ValueError('not a variant-major PLINK .bed payload (bad magic bytes)')
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
        self.read_threads = max(1, int(read_threads)) if read_threads else min(16, os.cpu_count() or 4)

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

    _ring_cache = {}  # (block_variants, row_bytes, n_buffers) -> pinned buffers: page-locking 1 GB costs ~0.4 s

    def _staging_ring(self, n_buffers: int):
        """Pinned staging buffers, kept for the next call with the same geometry (plain numpy arrays when no CUDA runtime
        is present, e.g. in CPU-only tests)."""
        shape = (self.block_variants, self.row_bytes)
        key = shape + (n_buffers, )
        ring = PlinkBedSource._ring_cache.get(key)
        if ring is None:
            try:
                from . import _capi
                ring = [_capi.pinned_empty(shape, np.uint8) for _ in range(n_buffers)]
            except Exception:  # noqa: BLE001 -- no library / no device: the data path still works, just unpinned
                ring = [np.empty(shape, dtype=np.uint8) for _ in range(n_buffers)]
            PlinkBedSource._ring_cache.clear()  # one geometry at a time: the buffers are large
            PlinkBedSource._ring_cache[key] = ring
        return ring

    def blocks(self):
        G = self.packed.shape[0]
        if self._path is None or G == 0:
            for g0 in range(0, G, self.block_variants):
                g1 = min(G, g0 + self.block_variants)
                yield self.meta.iloc[g0:g1], GenotypeBlock(np.ascontiguousarray(self.packed[g0:g1]), 'plink2',
                                                           self.missing)
            return
        # streamed from the file: block k lands in staging buffer k % 3; the consumer (engine.run_pipelined, two
        # lanes) has finished with block k - 3 before it asks for block k, so the buffer is free again
        from concurrent.futures import ThreadPoolExecutor
        ring = self._staging_ring(3)
        fd = os.open(self._path, os.O_RDONLY)
        try:
            with ThreadPoolExecutor(self.read_threads) as pool:
                for k, g0 in enumerate(range(0, G, self.block_variants)):
                    g1 = min(G, g0 + self.block_variants)
                    buf = ring[k % 3][:g1 - g0]
                    flat = buf.reshape(-1)
                    nbytes = flat.size
                    base = 3 + g0 * self.row_bytes
                    step = -(-nbytes // self.read_threads)

                    def read(i, flat=flat, base=base, nbytes=nbytes, step=step):
                        lo, hi = i * step, min(nbytes, (i + 1) * step)
                        done = lo
                        while done < hi:  # preadv releases the GIL: the slices are read in parallel
                            got = os.preadv(fd, [memoryview(flat[done:hi])], base + done)
                            if got <= 0:
                                raise IOError('unexpected end of .bed file')
                            done += got

                    list(pool.map(read, range(self.read_threads)))
                    yield self.meta.iloc[g0:g1], GenotypeBlock(buf, 'plink2', self.missing)
        finally:
            os.close(fd)
