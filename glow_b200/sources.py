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

from .engine import GenotypeBlock, ResidentBlock, contiguous_blocks

PLINK_MAGIC = bytes([0x6c, 0x1b, 0x01])  # PlinkFileFormat.scala:159 (variant-major)


class GenotypeSource:
    """Iterable of (metadata frame, GenotypeBlock); the metadata columns pass through to the output."""

    def blocks(self, first_block: int = 0, last_block: Optional[int] = None) -> Iterator[Tuple[pd.DataFrame, GenotypeBlock]]:
        """Blocks [first_block, last_block) in order (all of them by default)."""
        raise NotImplementedError

    def n_blocks(self) -> Optional[int]:
        """Number of blocks when it is known up front (lets a multi-GPU run shard contiguous ranges), else None."""
        return None


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

    def n_blocks(self):
        return -(-self.data.shape[0] // self.block_variants)

    def blocks(self, first_block=0, last_block=None):
        G = self.data.shape[0]
        last_block = self.n_blocks() if last_block is None else last_block
        for k in range(first_block, last_block):
            g0 = k * self.block_variants
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

    def n_blocks(self):
        return -(-self.packed.shape[0] // self.block_variants)

    def _staging_ring(self, n_buffers: int):
        """Pinned staging buffers of THIS source, kept for its next pass (page-locking 1 GB costs ~0.4 s); plain numpy
        arrays when no CUDA runtime is present (CPU-only tests).  Returns (buffers, pinned?)."""
        shape = (self.block_variants, self.row_bytes)
        free = self.__dict__.setdefault('_rings', [])
        for i, (key, ring, pinned) in enumerate(free):
            if key == shape + (n_buffers, ):
                free.pop(i)
                return ring, pinned
        try:
            from . import _capi
            return [_capi.pinned_empty(shape, np.uint8) for _ in range(n_buffers)], True
        except Exception:  # noqa: BLE001 -- no library / no device: the data path still works, just unpinned
            return [np.empty(shape, dtype=np.uint8) for _ in range(n_buffers)], False

    def _release_ring(self, n_buffers, ring, pinned):
        self.__dict__.setdefault('_rings', []).append(((self.block_variants, self.row_bytes, n_buffers), ring, pinned))

    def blocks(self, first_block=0, last_block=None, ring_depth: int = 3):
        """``ring_depth``: staging buffers of a streamed file = blocks the consumer may hold at once + 1 (n_lanes + 1)."""
        G = self.packed.shape[0]
        last_block = self.n_blocks() if last_block is None else last_block
        if self._path is None or G == 0:
            for k in range(first_block, last_block):
                g0 = k * self.block_variants
                g1 = min(G, g0 + self.block_variants)
                yield self.meta.iloc[g0:g1], GenotypeBlock(np.ascontiguousarray(self.packed[g0:g1]), 'plink2',
                                                           self.missing)
            return
        # streamed from the file: block k lands in staging buffer k % depth; the consumer (run_pipelined / run_shards with
        # depth - 1 lanes) has finished with block k - depth before it asks for block k, so the buffer is free again.
        # Every live iterator has its own ring (two shards of one file can be streamed at the same time).
        from concurrent.futures import ThreadPoolExecutor
        depth = max(2, int(ring_depth))
        ring, pinned = self._staging_ring(depth)
        fd = os.open(self._path, os.O_RDONLY)
        try:
            with ThreadPoolExecutor(self.read_threads) as pool:
                for i, k in enumerate(range(first_block, last_block)):
                    g0 = k * self.block_variants
                    g1 = min(G, g0 + self.block_variants)
                    buf = ring[i % depth][:g1 - g0]
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
                    yield self.meta.iloc[g0:g1], GenotypeBlock(buf, 'plink2', self.missing, pinned=pinned)
        finally:
            os.close(fd)
            self._release_ring(depth, ring, pinned)


class ResidentSource(GenotypeSource):
    """Genotype blocks loaded ONCE into HBM and kept there: every later ``linear_regression(resident, ...)`` call (another
    phenotype set, another covariate model) reads them at HBM speed instead of paying the PCIe copy of the packed
    matrix again.  With several devices the source's blocks are placed in contiguous ranges -- the placement the sharded
    run uses when called with the same device list (``linear_regression(..., devices=...)``).  Blocks that span several
    contigs are split at load time so that LOCO runs need no gather.  The blocks are uploaded as they are read: nothing
    but one block at a time is held on the host."""

    def __init__(self, source: GenotypeSource, devices: Sequence[int] = (0, ), split_contigs: bool = True):
        """``split_contigs=False`` keeps the source's blocks whole (rows then come out in exactly the source's block
        order; only for runs without LOCO offsets)."""
        import ctypes as C
        from . import _capi
        self._lib = _capi.load()
        self.devices = [int(d) for d in devices]
        n_src = source.n_blocks()
        if n_src is None:
            raise ValueError('ResidentSource needs a source that knows its number of blocks')
        n_dev = len(self.devices)
        bounds = [contiguous_blocks(n_src, d, n_dev) for d in range(n_dev)]
        self._items, self._allocs, self._ranges = [], [], []
        self.bytes_resident = 0
        k = 0
        src_iter = iter(source.blocks())
        for d, (a, b) in enumerate(bounds):
            first = len(self._items)
            dev = self.devices[d]
            while k < b:
                meta, block = next(src_iter)
                k += 1
                if split_contigs and 'contigName' in meta.columns and meta['contigName'].nunique() > 1:
                    subs = []
                    for contig in pd.unique(meta['contigName']):
                        sel = (meta['contigName'] == contig).to_numpy()
                        subs.append((meta[sel], np.ascontiguousarray(block.data[sel])))
                else:
                    subs = [(meta, np.ascontiguousarray(block.data))]
                for sub_meta, data in subs:
                    p = C.c_void_p()
                    _capi.check(self._lib.glr_device_alloc(dev, data.nbytes, C.byref(p)))
                    self._allocs.append((dev, p.value))
                    if data.nbytes:
                        _capi.check(self._lib.glr_device_upload(dev, p, data.ctypes.data, data.nbytes))
                    self.bytes_resident += data.nbytes
                    self._items.append((sub_meta, ResidentBlock(p.value or 0, data.strides[0] if data.shape[0] else 0,
                                                                int(data.shape[0]), block.format, block.missing, dev)))
            self._ranges.append((first, len(self._items)))

    def n_blocks(self):
        return len(self._items)

    def shard_ranges(self, devices: Sequence[int]):
        """Item ranges per device of a sharded run; it must use the devices the blocks were placed on."""
        if [int(d) for d in devices] != self.devices:
            raise ValueError(f'the blocks are resident on GPUs {self.devices}, the run was asked for GPUs {list(devices)}')
        return list(self._ranges)

    def blocks(self, first_block=0, last_block=None):
        last_block = len(self._items) if last_block is None else last_block
        for k in range(first_block, last_block):
            yield self._items[k]

    def close(self):
        for dev, addr in getattr(self, '_allocs', []):
            try:
                self._lib.glr_device_free(dev, addr)
            except Exception:  # noqa: BLE001
                pass
        self._allocs = []
        self._items = []

    def __del__(self):
        self.close()
