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
        depth = max(2, int(ring_depth)) + 1  # + 1: the NEXT block is read while the consumer works on this one
        ring, pinned = self._staging_ring(depth)
        fd = os.open(self._path, os.O_RDONLY)
        try:
            with ThreadPoolExecutor(self.read_threads) as pool:

                def start(i, k):
                    g0 = k * self.block_variants
                    g1 = min(G, g0 + self.block_variants)
                    buf = ring[i % depth][:g1 - g0]
                    flat = buf.reshape(-1)
                    nbytes = flat.size
                    base = 3 + g0 * self.row_bytes
                    step = -(-nbytes // self.read_threads)

                    def read(j):
                        lo, hi = j * step, min(nbytes, (j + 1) * step)
                        done = lo
                        while done < hi:  # preadv releases the GIL: the slices are read in parallel
                            got = os.preadv(fd, [memoryview(flat[done:hi])], base + done)
                            if got <= 0:
                                raise IOError('unexpected end of .bed file')
                            done += got

                    return g0, g1, buf, [pool.submit(read, j) for j in range(self.read_threads)]

                ks = list(range(first_block, last_block))
                nxt = start(0, ks[0]) if ks else None
                for i, k in enumerate(ks):
                    g0, g1, buf, futs = nxt
                    nxt = start(i + 1, ks[i + 1]) if i + 1 < len(ks) else None
                    for f in futs:
                        f.result()
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


class BgenSource(GenotypeSource):
    """BGEN v1.2 (layout 2, zlib or uncompressed) file -> hard-call blocks on the device.

    The reference reads BGEN through its Spark datasource (core/.../bgen/BgenFileIterator.scala) into probability
    arrays, turns them into calls with ``hardCallThreshold`` (0.9, BgenSchemaInferrer.scala:42) and the regression
    sees ``genotype_states`` of those calls.  Here the host only parses the variant records and inflates the probability
    blocks; the stored 8/16/32-bit probabilities go to the GPU as they are, where ``glr_bgen_to_plink2`` turns them into
    2-bit rows in HBM -- exact small integers, so a BGEN file takes the int8 tensor-core path of the regression.
    Supported: unphased, diploid, biallelic blocks (what Glow's ``genotype_states`` path is defined for)."""

    def __init__(self, path, hard_call_threshold: float = 0.9, mean_substitute: bool = True, block_variants: int = 2048,
                 device: int = 0, split_contigs: bool = True):
        """``split_contigs``: hand a block out in runs of one contig (needed for LOCO offsets; a run is a row range of the
        device buffer).  With False the rows of the result follow the file's blocks exactly."""
        import struct
        self.split_contigs = bool(split_contigs)
        self._path = os.fspath(path)
        self.threshold = float(hard_call_threshold)
        self.missing = 'mean' if mean_substitute else 'minus_one'
        self.block_variants = int(block_variants)
        self.device = int(device)
        with open(self._path, 'rb') as f:
            head = f.read(16)
            offset, hlen, M, N = struct.unpack('<IIII', head)
            rest = f.read(hlen - 12)
            if rest[:4] not in (b'bgen', b'\0\0\0\0'):
                raise ValueError('not a BGEN file (bad magic)')
            flags, = struct.unpack('<I', rest[-4:])
            self._compression, layout = flags & 3, (flags >> 2) & 15
            if layout != 2 or self._compression not in (0, 1):
                raise ValueError(f'only BGEN v1.2 layout 2 with zlib or no compression is supported (layout {layout}, '
                                 f'compression {self._compression})')
            self.n_samples, self.n_variants = int(N), int(M)
            # one pass over the variant records: identifiers and the byte range of every probability block
            f.seek(4 + offset)
            contigs, rsids, starts, ends, refs, alts, spans = [], [], [], [], [], [], []
            for _ in range(M):
                lid, = struct.unpack('<H', f.read(2))
                f.seek(lid, 1)
                lrs, = struct.unpack('<H', f.read(2))
                rsids.append(f.read(lrs).decode())
                lchr, = struct.unpack('<H', f.read(2))
                contigs.append(f.read(lchr).decode())
                vpos, nall = struct.unpack('<IH', f.read(6))
                alleles = []
                for _a in range(nall):
                    la, = struct.unpack('<I', f.read(4))
                    alleles.append(f.read(la).decode())
                if nall != 2:
                    raise ValueError('only biallelic BGEN variants are supported')
                clen, = struct.unpack('<I', f.read(4))
                spans.append((f.tell(), clen))
                f.seek(clen, 1)
                starts.append(vpos - 1)
                ends.append(vpos - 1 + len(alleles[0]))
                refs.append(alleles[0])
                alts.append(alleles[1])
        self._spans = spans
        self.meta = pd.DataFrame({'contigName': contigs, 'names': rsids, 'start': np.asarray(starts, dtype=np.int64),
                                  'end': np.asarray(ends, dtype=np.int64), 'referenceAllele': refs, 'alternateAllele': alts})
        self.row_bytes = (self.n_samples + 3) // 4
        self._ring = []  # device buffers [(addr, bytes)], one per block in flight

    def n_blocks(self):
        return -(-self.n_variants // self.block_variants)

    def _read_block(self, g0, g1):
        """-> (probs [g, 2 N] uint8/16/32, ploidy [g, N] uint8, bits)"""
        import struct
        import zlib
        N = self.n_samples
        probs, ploidy, bits_all = [], [], None
        with open(self._path, 'rb') as f:
            for g in range(g0, g1):
                pos, clen = self._spans[g]
                f.seek(pos)
                raw = f.read(clen)
                block = zlib.decompress(raw[4:]) if self._compression == 1 else raw
                n, k, pmin, pmax = struct.unpack_from('<IHBB', block, 0)
                if n != N or k != 2 or pmin != 2 or pmax != 2:
                    raise ValueError('only diploid biallelic BGEN probability blocks are supported')
                phased, bits = struct.unpack_from('<BB', block, 8 + n)
                if phased != 0 or bits not in (8, 16, 32):
                    raise ValueError(f'unsupported BGEN block (phased={phased}, bits={bits})')
                if bits_all is None:
                    bits_all = bits
                elif bits != bits_all:
                    raise ValueError('mixed probability precisions inside one block of variants')
                ploidy.append(np.frombuffer(block, dtype=np.uint8, count=n, offset=8))
                probs.append(np.frombuffer(block, dtype={8: np.uint8, 16: '<u2', 32: '<u4'}[bits], count=2 * n, offset=10 + n))
        return np.ascontiguousarray(np.stack(probs)), np.ascontiguousarray(np.stack(ploidy)), bits_all

    def blocks(self, first_block=0, last_block=None, ring_depth: int = 3):
        import ctypes as C
        from . import _capi
        lib = _capi.load()
        last_block = self.n_blocks() if last_block is None else last_block
        nbytes = self.block_variants * self.row_bytes
        ring = []
        try:
            for _ in range(max(2, ring_depth)):
                p = C.c_void_p()
                _capi.check(lib.glr_device_alloc(self.device, nbytes, C.byref(p)))
                ring.append(p.value)
            for i, k in enumerate(range(first_block, last_block)):
                g0, g1 = k * self.block_variants, min(self.n_variants, (k + 1) * self.block_variants)
                probs, ploidy, bits = self._read_block(g0, g1)
                addr = ring[i % len(ring)]
                _capi.check(lib.glr_bgen_to_plink2(self.device, probs.ctypes.data, ploidy.ctypes.data, g1 - g0, self.n_samples, bits,
                                                   self.threshold, addr, self.row_bytes))
                meta = self.meta.iloc[g0:g1]
                # a block is handed out in runs of one contig (LOCO runs take a different Y per contig; a run is just a
                # row range of the device buffer)
                names = meta['contigName'].to_numpy()
                cuts = [0] + [j for j in range(1, len(names)) if self.split_contigs and names[j] != names[j - 1]] + [len(names)]
                for r0, r1 in zip(cuts, cuts[1:]):
                    yield meta.iloc[r0:r1], ResidentBlock(addr + r0 * self.row_bytes, self.row_bytes, r1 - r0, 'plink2',
                                                          self.missing, self.device)
        finally:
            for addr in ring:
                lib.glr_device_free(self.device, addr)
