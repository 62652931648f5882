"""Device engine: owns one glr_state handle (one GPU) and runs genotype blocks through the CUDA path.

This is the only place that talks to the C ABI for regression blocks.  There is no CPU fallback:
constructing a DeviceState without the built library or without an sm_100 device raises.
"""
import ctypes as C
from dataclasses import dataclass
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from . import _capi

FORMATS = {'f64': _capi.FMT_F64, 'f32': _capi.FMT_F32, 'i8': _capi.FMT_I8, 'plink2': _capi.FMT_PLINK2}
_ROW_BYTES = {
    _capi.FMT_F64: lambda n: 8 * n,
    _capi.FMT_F32: lambda n: 4 * n,
    _capi.FMT_I8: lambda n: n,
    _capi.FMT_PLINK2: lambda n: (n + 3) // 4,
}
STAT_NAMES = ('effect', 'stderror', 'tvalue', 'pvalue')
VERBOSE_NAMES = ('n', 'sum_x', 'y_transpose_x')


@dataclass
class GenotypeBlock:
    """One block of variants in a device-consumable encoding (variant-major rows)."""
    data: np.ndarray  # 2-D, C-contiguous rows: float64/float32/int8 [G, Ng] or uint8 [G, ceil(Ng/4)]
    format: str  # 'f64' | 'f32' | 'i8' | 'plink2'
    missing: str = 'minus_one'  # 'minus_one' (genotype_states) | 'mean' (mean_substitute)
    pinned: bool = False  # `data` lives in page-locked memory (glr_host_alloc): it is copied to the device as is


@dataclass
class ResidentBlock:
    """A block that already lives in the HBM of one GPU (glow_b200.sources.ResidentSource)."""
    addr: int          # device address of the first row
    stride: int        # bytes between variants
    n_variants: int
    format: str
    missing: str
    device: int        # CUDA ordinal that owns the memory


def _arrow_list_to_matrix(col):
    """pyarrow (Chunked)Array of list<T> / large_list<T> / fixed_size_list<T> with equal lengths ->
    [G, N] numpy view of the flat child buffer (zero copy for primitive, null-free data)."""
    import pyarrow as pa
    if isinstance(col, pa.ChunkedArray):
        col = col.combine_chunks() if col.num_chunks != 1 else col.chunk(0)
    if col.null_count:
        raise ValueError('the values column contains null arrays')
    t = col.type
    G = len(col)
    if pa.types.is_fixed_size_list(t):
        n = t.list_size
    elif pa.types.is_list(t) or pa.types.is_large_list(t):
        lens = np.diff(col.offsets.to_numpy(zero_copy_only=False))
        n = int(lens[0]) if G else 0
        if G and not (lens == n).all():
            raise ValueError('all genotype arrays of a block must have the same length')
    else:
        raise ValueError(f'the values column must be an Arrow list type, got {t}')
    flat = col.flatten()
    arr = flat.to_numpy(zero_copy_only=False)  # nulls inside an array become NaN (one copy)
    return arr.reshape(G, n)


def as_block(values, dt=np.float64) -> GenotypeBlock:
    """Turns the `_glow_regression_values` column of a pandas block (a sequence of per-variant
    arrays, lin_reg.py:226) into one contiguous variant-major matrix.  No transpose: the device
    consumes variant-major rows directly (replaces np.column_stack, lin_reg.py:227)."""
    if isinstance(values, GenotypeBlock):
        return values
    arr = values if isinstance(values, np.ndarray) and values.ndim == 2 else None
    if arr is None and type(values).__module__.startswith('pyarrow'):
        arr = _arrow_list_to_matrix(values)
    elif arr is None and hasattr(values, 'dtype') and type(values.dtype).__name__ == 'ArrowDtype':
        import pyarrow as pa
        arr = _arrow_list_to_matrix(pa.chunked_array(values.array._pa_array) if hasattr(values, 'array') else
                                    values._pa_array)
    if arr is None:
        seq = list(values)
        if len(seq) == 0:
            arr = np.empty((0, 0), dtype=dt)
        else:
            arr = np.stack([np.asarray(v) for v in seq])
    if arr.dtype == np.int8:
        return GenotypeBlock(np.ascontiguousarray(arr), 'i8')
    want = np.dtype(dt)
    arr = np.ascontiguousarray(arr, dtype=want)
    return GenotypeBlock(arr, 'f32' if want == np.float32 else 'f64')


class DeviceState:
    """Replica of the regression state on one GPU (Q, per-contig Y, masks, scales, dof)."""

    def __init__(self,
                 Q: np.ndarray,
                 Y: np.ndarray,
                 Y_mask: np.ndarray,
                 YdotY: np.ndarray,
                 Y_scale: np.ndarray,
                 dof: int,
                 Y_raw: Optional[np.ndarray] = None,
                 keep_idx: Optional[np.ndarray] = None,
                 n_geno_samples: Optional[int] = None,
                 device: int = 0,
                 n_lanes: int = 2,
                 memspace: int = _capi.MEM_HOST,
                 dims: Optional[Tuple[int, int, int, int]] = None,
                 options: Optional[dict] = None,
                 _handle=None):
        """Host arrays (float64, C order): Q [N,K]; Y [C,N,P] or [N,P]; Y_mask [N,P]; YdotY [C,P] or [P];
        Y_scale [P].  With memspace=MEM_DEVICE the array arguments are integer device addresses and
        ``dims`` = (N, K, P, C) must be given.  ``options``: kernel-selection knobs (glr_options fields, e.g.
        ``{'split8': 'force', 'split8_pair': 'off'}``); the default leaves every choice to the library."""
        self._lib = _capi.load()
        self._handle = C.c_void_p()
        self._pending: Dict[int, tuple] = {}
        self._stage: Dict[int, dict] = {}
        if memspace == _capi.MEM_HOST:
            Q = np.ascontiguousarray(Q, dtype=np.float64)
            Y = np.ascontiguousarray(Y, dtype=np.float64)
            if Y.ndim == 2:
                Y = Y[None]
            Y_mask = np.ascontiguousarray(Y_mask, dtype=np.float64)
            YdotY = np.ascontiguousarray(np.atleast_2d(YdotY), dtype=np.float64)
            Y_scale = np.ascontiguousarray(Y_scale, dtype=np.float64)
            n_contigs, N, P = Y.shape
            K = Q.shape[1] if Q.ndim == 2 else 0
            if Q.size and Q.shape[0] != N:
                raise ValueError('Q and Y must have the same number of rows')
            if Y_mask.shape != (N, P) or YdotY.shape != (n_contigs, P) or Y_scale.shape != (P, ):
                raise ValueError('inconsistent state array shapes')
            if Y_raw is not None:
                Y_raw = np.ascontiguousarray(Y_raw, dtype=np.float64)
            addr = lambda a: None if a is None or a.size == 0 else a.ctypes.data
            ptrs = dict(Q=addr(Q), Y=addr(Y), Y_mask=addr(Y_mask), YdotY=addr(YdotY), Y_scale=addr(Y_scale),
                        Y_raw=addr(Y_raw))
            self._keep = (Q, Y, Y_mask, YdotY, Y_scale, Y_raw)
        else:
            N, K, P, n_contigs = dims
            ptrs = dict(Q=Q, Y=Y, Y_mask=Y_mask, YdotY=YdotY, Y_scale=Y_scale, Y_raw=Y_raw)
        if keep_idx is not None:
            keep_idx = np.ascontiguousarray(keep_idx, dtype=np.int64)
            if keep_idx.shape != (N, ):
                raise ValueError('keep_idx must have one entry per phenotype row')
        self.N, self.K, self.P, self.n_contigs = int(N), int(K), int(P), int(n_contigs)
        self.Ng = int(n_geno_samples) if n_geno_samples is not None else self.N
        self.verbose = Y_raw is not None
        self.device = device
        self.n_lanes = n_lanes
        d = _capi.StateDesc(abi_version=_capi.GLR_ABI_VERSION,
                            device=device,
                            n_samples=self.N,
                            n_geno_samples=self.Ng,
                            n_covariates=self.K,
                            n_phenotypes=self.P,
                            n_contigs=self.n_contigs,
                            verbose=int(self.verbose),
                            Q=ptrs['Q'],
                            Y=ptrs['Y'],
                            Y_mask=ptrs['Y_mask'],
                            YdotY=ptrs['YdotY'],
                            Y_scale=ptrs['Y_scale'],
                            Y_raw=ptrs['Y_raw'],
                            keep_idx=None if keep_idx is None else keep_idx.ctypes.data,
                            dof=int(dof),
                            max_block_variants=0,
                            n_lanes=n_lanes,
                            memspace=memspace,
                            opt=_capi.make_options(options))
        self._desc = d
        if _handle is None:
            _capi.check(self._lib.glr_state_create(C.byref(d), C.byref(self._handle)))
        elif _handle:  # replica created by glr_comm_broadcast_state (DeviceGroup)
            self._handle = C.c_void_p(_handle)

    @classmethod
    def _replica(cls, template: 'DeviceState', handle: int, device: int) -> 'DeviceState':
        """Wraps a state handle that glr_comm_broadcast_state created on another device."""
        r = cls.__new__(cls)
        r._lib = template._lib
        r._handle = C.c_void_p(handle)
        r._pending, r._stage = {}, {}
        for k in ('N', 'K', 'P', 'n_contigs', 'Ng', 'verbose', 'n_lanes'):
            setattr(r, k, getattr(template, k))
        r.device = device
        r._desc = None
        return r

    # ------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, '_handle', None) is not None and self._handle.value:
            self._lib.glr_state_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def unique_masks(self) -> int:
        return int(self._lib.glr_state_unique_masks(self._handle))

    def _block_desc(self, fmt: int, addr: int, stride: int, G: int, contig: int, missing: str, memspace: int):
        return _capi.BlockDesc(format=fmt,
                               memspace=memspace,
                               genotypes=addr,
                               row_stride_bytes=stride,
                               n_variants=G,
                               contig=contig,
                               missing_policy=_capi.MISSING_MEAN if missing == 'mean' else _capi.MISSING_AS_MINUS_ONE)

    def _check_block(self, block: GenotypeBlock):
        fmt = FORMATS[block.format]
        data = block.data
        if data.ndim != 2 or (data.size and not data.flags['C_CONTIGUOUS']):
            raise ValueError('genotype block must be a C-contiguous 2-D array')
        need = _ROW_BYTES[fmt](self.Ng)
        have = data.shape[1] * data.dtype.itemsize
        if data.shape[0] and have != need:
            raise ValueError(f'genotype rows have {have} bytes, expected {need} for {self.Ng} samples '
                             f'in format {block.format}')
        return fmt, data

    # Pinned staging (per lane, grown on demand).  A device-to-host copy into pageable memory returns only when it has
    # finished, which would serialise the lanes; results therefore land in pinned buffers and wait() copies them out.
    # Pageable genotype blocks are staged the same way so that their H2D copy is asynchronous as well.
    _STAGE_INPUT_LIMIT = 2 << 30

    def _pinned(self, lane: int, key: str, nbytes: int) -> np.ndarray:
        st = self._stage.setdefault(lane, {})
        buf = st.get(key)
        if buf is None or buf.size < nbytes:
            st[key] = buf = _capi.pinned_empty((-(-max(nbytes + nbytes // 4, 4096) // 64) * 64, ), np.uint8)
        return buf

    def submit(self, lane: int, block: GenotypeBlock, contig: int = 0, out: Optional[Dict[str, np.ndarray]] = None,
               pinned_input: bool = False):
        """Queues one block on a lane: H2D copy, kernels and D2H copy are all asynchronous.  ``out`` (optional) must be
        pinned arrays; without it the results land in the lane's pinned staging and wait() returns copies.
        ``pinned_input``: the block's memory is already page-locked (sources that stage their own blocks)."""
        resident = isinstance(block, ResidentBlock)
        if resident:
            if block.device != self.device:
                raise ValueError(f'resident block lives on GPU {block.device}, this state on GPU {self.device}')
            fmt, data, G = FORMATS[block.format], None, int(block.n_variants)
        else:
            fmt, data = self._check_block(block)
            G = int(data.shape[0])
        names = STAT_NAMES + (VERBOSE_NAMES if self.verbose else ())
        staged_out = out is None
        if staged_out:
            n = self.P * G
            flat = self._pinned(lane, 'out', 8 * n * len(names))[:8 * n * len(names)].view(np.float64)
            out = {k: flat[i * n:(i + 1) * n].reshape(self.P, G) for i, k in enumerate(names)}
        if resident:
            bd = self._block_desc(fmt, block.addr if G else None, block.stride, G, contig, block.missing, _capi.MEM_DEVICE)
        else:
            if G and not pinned_input and not getattr(block, 'pinned', False) and data.nbytes <= self._STAGE_INPUT_LIMIT:
                stage = self._pinned(lane, 'in', data.nbytes)[:data.nbytes].view(data.dtype).reshape(data.shape)
                np.copyto(stage, data)
                data = stage
            bd = self._block_desc(fmt, data.ctypes.data if G else None, data.strides[0] if G else 0, G, contig,
                                  block.missing, _capi.MEM_HOST)
        bo = _capi.BlockOut(memspace=_capi.MEM_HOST, **{k: out[k].ctypes.data if G else None for k in names})
        _capi.check(self._lib.glr_block_submit(self._handle, lane, C.byref(bd), C.byref(bo)))
        self._pending[lane] = (out, data, staged_out)  # keep host buffers alive until wait()
        return out

    def wait(self, lane: int) -> Dict[str, np.ndarray]:
        _capi.check(self._lib.glr_block_wait(self._handle, lane))
        out, _, staged = self._pending.pop(lane)
        return {k: v.copy() for k, v in out.items()} if staged else out

    def run(self, block: GenotypeBlock, contig: int = 0) -> Dict[str, np.ndarray]:
        """One block, synchronously: the device counterpart of lin_reg.py:226-249.
        Returns P x G_b float64 arrays."""
        self.submit(0, block, contig)
        return self.wait(0)

    def run_device(self, fmt: str, geno_addr: int, stride: int, G: int, out_addrs: Dict[str, int], contig: int = 0,
                   missing: str = 'mean', lane: int = 0, wait: bool = True):
        """Block and results already resident in HBM (device addresses)."""
        names = STAT_NAMES + (VERBOSE_NAMES if self.verbose else ())
        bd = self._block_desc(FORMATS[fmt], geno_addr, stride, G, contig, missing, _capi.MEM_DEVICE)
        bo = _capi.BlockOut(memspace=_capi.MEM_DEVICE, **{k: out_addrs[k] for k in names})
        _capi.check(self._lib.glr_block_submit(self._handle, lane, C.byref(bd), C.byref(bo)))
        if wait:
            _capi.check(self._lib.glr_block_wait(self._handle, lane))

    def wait_device(self, lane: int = 0):
        _capi.check(self._lib.glr_block_wait(self._handle, lane))

    def timing(self, lane: int = 0) -> Dict[str, float]:
        t = _capi.Timing()
        _capi.check(self._lib.glr_block_timing(self._handle, lane, C.byref(t)))
        return {k: getattr(t, k) for k, _ in _capi.Timing._fields_}

    def lane_stream(self, lane: int = 0) -> int:
        s = C.c_void_p()
        _capi.check(self._lib.glr_lane_stream(self._handle, lane, C.byref(s)))
        return s.value or 0

    def run_pipelined(self, blocks: Iterable[Tuple[GenotypeBlock, int]]) -> Iterator[Dict[str, np.ndarray]]:
        """Streams (block, contig) pairs through the lanes so that the H2D copy of block i+1
        overlaps the kernels of block i; yields results in submission order."""
        inflight: List[int] = []
        lane = 0
        for block, contig in blocks:
            if len(inflight) == self.n_lanes:
                yield self.wait(inflight.pop(0))
            self.submit(lane, block, contig)
            inflight.append(lane)
            lane = (lane + 1) % self.n_lanes
        while inflight:
            yield self.wait(inflight.pop(0))


def contiguous_blocks(n_blocks: int, rank: int, world: int) -> Tuple[int, int]:
    """Blocks [start, stop) of shard `rank` of `world`: contiguous ranges, so concatenating the shards in rank order
    reproduces the single-GPU order, and whole blocks, so every block is the one a single GPU would have run."""
    return (n_blocks * rank) // world, (n_blocks * (rank + 1)) // world


class DeviceGroup:
    """State replicated on several GPUs of this process: built and packed on the first device,
    every packed array sent to the others by one grouped NCCL broadcast (glr_comm_broadcast_state), then variant blocks
    are spread over (device, lane) slots and the results gathered in submission order.  The counterpart of the
    reference's Spark partitioning of variant rows with the state captured in the closure (lin_reg.py:277-284)."""

    def __init__(self, devices: Sequence[int], **state_kwargs):
        devices = [int(d) for d in devices]
        if not devices or len(set(devices)) != len(devices):
            raise ValueError('devices must be a non-empty list of distinct CUDA ordinals')
        self._lib = _capi.load()
        self.devices = devices
        self._comm = C.c_void_p()
        if len(devices) == 1:
            self.states = [DeviceState(device=devices[0], **state_kwargs)]
            self.transport = 'single'
            return
        arr = (C.c_int32 * len(devices))(*devices)
        _capi.check(self._lib.glr_comm_init(arr, len(devices), C.byref(self._comm)))
        first = DeviceState(device=devices[0], _handle=0, **state_kwargs)  # validates + builds the descriptor only
        handles = (C.c_void_p * len(devices))()
        rc = self._lib.glr_comm_broadcast_state(self._comm, C.byref(first._desc), handles)
        if rc:
            self._lib.glr_comm_destroy(self._comm)
            self._comm = C.c_void_p()
            _capi.check(rc)
        first._handle = C.c_void_p(handles[0])
        self.states = [first] + [DeviceState._replica(first, handles[i], devices[i]) for i in range(1, len(devices))]
        self.transport = self._lib.glr_comm_transport(self._comm).decode()

    def __len__(self):
        return len(self.states)

    @property
    def unique_masks(self) -> int:
        return self.states[0].unique_masks

    @property
    def verbose(self) -> bool:
        return self.states[0].verbose

    def close(self):
        for s in getattr(self, 'states', []):
            s.close()
        self.states = []
        if getattr(self, '_comm', None) is not None and self._comm.value:
            self._lib.glr_comm_destroy(self._comm)
            self._comm = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def slots(self) -> List[Tuple[int, int]]:
        """(device index, lane) pairs, device-minor: consecutive blocks go to different GPUs."""
        n_lanes = self.states[0].n_lanes
        return [(d, l) for l in range(n_lanes) for d in range(len(self.states))]

    def run_pipelined(self, blocks: Iterable[Tuple[GenotypeBlock, int]], assign=None) -> Iterator[Dict[str, np.ndarray]]:
        """Streams (block, contig) pairs over every (device, lane) slot and yields the results in SUBMISSION order (the
        ordered gather): the output is identical, bit for bit, to running the same blocks on one GPU.
        ``assign(k)`` -> device index of block k (default: round robin)."""
        n_dev, n_lanes = len(self.states), self.states[0].n_lanes
        free = [list(range(n_lanes)) for _ in range(n_dev)]
        inflight: List[Tuple[int, int]] = []
        for k, (block, contig) in enumerate(blocks):
            d = assign(k) if assign is not None else k % n_dev
            while not free[d]:  # results come back in order: drain from the front until this device has a lane
                dd, ll = inflight.pop(0)
                yield self.states[dd].wait(ll)
                free[dd].append(ll)
            lane = free[d].pop(0)
            self.states[d].submit(lane, block, contig)
            inflight.append((d, lane))
        while inflight:
            dd, ll = inflight.pop(0)
            yield self.states[dd].wait(ll)


    def run_shards(self, shard_iters: Sequence[Iterator[Tuple[GenotypeBlock, int]]]):
        """Contiguous sharding: ``shard_iters[d]`` yields the (block, contig) pairs of device d's variant range.  All
        devices are fed concurrently (each keeps its lanes full); yields ``(d, seq, result)`` as blocks complete, in
        order within a device.  Concatenating by (d, seq) gives the single-GPU order."""
        n_dev, n_lanes = len(self.states), self.states[0].n_lanes
        if len(shard_iters) != n_dev:
            raise ValueError('one shard iterator per device')
        iters = [iter(it) for it in shard_iters]
        free = [list(range(n_lanes)) for _ in range(n_dev)]
        inflight: List[List[Tuple[int, int, int]]] = [[] for _ in range(n_dev)]  # (lane, seq, submit counter)
        done = [False] * n_dev
        seq = [0] * n_dev
        counter = 0
        while True:
            progressed = False
            for d in range(n_dev):
                if done[d] or not free[d]:
                    continue
                try:
                    block, contig = next(iters[d])
                except StopIteration:
                    done[d] = True
                    continue
                lane = free[d].pop(0)
                self.states[d].submit(lane, block, contig)
                inflight[d].append((lane, seq[d], counter))
                seq[d] += 1
                counter += 1
                progressed = True
            if progressed:
                continue
            oldest = [(q[0][2], d) for d, q in enumerate(inflight) if q]
            if not oldest:
                return
            _, d = min(oldest)
            lane, sq, _ = inflight[d].pop(0)
            res = self.states[d].wait(lane)
            free[d].append(lane)
            yield d, sq, res


def t_two_sided_pvalues(t: np.ndarray, dof: float, device: int = 0) -> np.ndarray:
    """2 * t.sf(|t|, dof) on the device (lin_reg.py:245)."""
    lib = _capi.load()
    t = np.ascontiguousarray(t, dtype=np.float64)
    out = np.empty_like(t)
    _capi.check(lib.glr_t_two_sided_pvalues(device, t.ctypes.data, t.size, float(dof), out.ctypes.data))
    return out


def decode_block(block: GenotypeBlock, n_samples: int, device: int = 0) -> np.ndarray:
    """Device decode of a packed block to float64 [G, N] (A0 semantics)."""
    lib = _capi.load()
    data = np.ascontiguousarray(block.data)
    G = data.shape[0]
    out = np.empty((G, n_samples), dtype=np.float64)
    bd = _capi.BlockDesc(format=FORMATS[block.format],
                         memspace=_capi.MEM_HOST,
                         genotypes=data.ctypes.data if G else None,
                         row_stride_bytes=data.strides[0] if G else 0,
                         n_variants=G,
                         contig=0,
                         missing_policy=_capi.MISSING_MEAN if block.missing == 'mean' else _capi.MISSING_AS_MINUS_ONE)
    _capi.check(lib.glr_decode_block(device, C.byref(bd), n_samples, out.ctypes.data))
    return out
