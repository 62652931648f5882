"""ctypes binding of include/glow_b200.h.  There is no CPU fallback: if the CUDA library is missing
or no sm_100 device is visible, every compute entry point raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('GLOW_B200_LIB') or os.path.join(_HERE, 'lib', 'libglow_b200.so')  # (override: A/B runs of differently built libraries)

GLR_ABI_VERSION = 3
FMT_F64, FMT_F32, FMT_I8, FMT_PLINK2 = 0, 1, 2, 3
MEM_HOST, MEM_DEVICE = 0, 1
MISSING_AS_MINUS_ONE, MISSING_MEAN = 0, 1
ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_STATE = -1, -2, -3, -4

AUTO, OFF, FORCE = 0, 1, 2  # glr_tristate
PATH_SPLIT8, PATH_SPLIT8_PAIR, PATH_SPLIT8_NSPLIT, PATH_PREPASS, PATH_DMMA = 1, 2, 4, 8, 16
PATH_MASK_SPARSE, PATH_MASK_INT8, PATH_MASK_DMMA, PATH_OPTIMISTIC, PATH_REPLAYED, PATH_MASK_GENERIC = 32, 64, 128, 256, 512, 1024
PATH_Q_DIGITS_LO = 2048

_dp = C.POINTER(C.c_double)


class Options(C.Structure):
    """glr_options: kernel-selection knobs, all zero = defaults."""
    _fields_ = [('split8', C.c_int32), ('split8_pair', C.c_int32), ('mask_sparse', C.c_int32), ('mask_int8', C.c_int32),
                ('keep_intercept', C.c_int32), ('no_tail', C.c_int32), ('sample_split', C.c_int32),
                ('optimistic_calls', C.c_int32), ('q_digits', C.c_int32)]


_TRI = {None: AUTO, 'auto': AUTO, 'off': OFF, 'never': OFF, 'force': FORCE, 'always': FORCE}


def make_options(opts=None) -> 'Options':
    """dict -> glr_options.  Tri-state knobs take 'auto' | 'off' | 'force'."""
    o = Options()
    for k, v in (opts or {}).items():
        if k not in dict(Options._fields_):
            raise ValueError(f'unknown option {k!r}')
        if k in ('keep_intercept', 'no_tail', 'sample_split', 'q_digits'):
            setattr(o, k, int(v))
        else:
            if v not in _TRI:
                raise ValueError(f'option {k!r}: expected auto | off | force, got {v!r}')
            setattr(o, k, _TRI[v])
    return o


class StateDesc(C.Structure):
    _fields_ = [('abi_version', C.c_int32), ('device', C.c_int32), ('n_samples', C.c_int64),
                ('n_geno_samples', C.c_int64), ('n_covariates', C.c_int32), ('n_phenotypes', C.c_int32),
                ('n_contigs', C.c_int32), ('verbose', C.c_int32), ('Q', C.c_void_p), ('Y', C.c_void_p),
                ('Y_mask', C.c_void_p), ('YdotY', C.c_void_p), ('Y_scale', C.c_void_p), ('Y_raw', C.c_void_p),
                ('keep_idx', C.c_void_p), ('dof', C.c_int64), ('max_block_variants', C.c_int32),
                ('n_lanes', C.c_int32), ('memspace', C.c_int32), ('reserved', C.c_int32), ('opt', Options)]


class BlockDesc(C.Structure):
    _fields_ = [('format', C.c_int32), ('memspace', C.c_int32), ('genotypes', C.c_void_p),
                ('row_stride_bytes', C.c_int64), ('n_variants', C.c_int32), ('contig', C.c_int32),
                ('missing_policy', C.c_int32), ('reserved', C.c_int32)]


class BlockOut(C.Structure):
    _fields_ = [('memspace', C.c_int32), ('reserved', C.c_int32), ('effect', C.c_void_p), ('stderror', C.c_void_p),
                ('tvalue', C.c_void_p), ('pvalue', C.c_void_p), ('n', C.c_void_p), ('sum_x', C.c_void_p),
                ('y_transpose_x', C.c_void_p)]


class LogRegDesc(C.Structure):
    _fields_ = [('abi_version', C.c_int32), ('device', C.c_int32), ('n_samples', C.c_int64), ('n_geno_samples', C.c_int64),
                ('n_covariates', C.c_int32), ('n_phenotypes', C.c_int32), ('n_contigs', C.c_int32), ('reserved', C.c_int32),
                ('C', C.c_void_p), ('gamma', C.c_void_p), ('Y_res', C.c_void_p), ('inv_CtGammaC', C.c_void_p),
                ('keep_idx', C.c_void_p), ('Q', C.c_void_p)]


class Timing(C.Structure):
    _fields_ = [('total_ms', C.c_float), ('prepass_ms', C.c_float), ('gram_ms', C.c_float), ('masked_ms', C.c_float),
                ('epilogue_ms', C.c_float), ('kernel_launches', C.c_int32), ('path', C.c_int32)]


# every symbol include/glow_b200.h declares: (restype, argtypes)
SYMBOLS = {
    'glr_abi_version': (C.c_int, []),
    'glr_last_error': (C.c_char_p, []),
    'glr_device_count': (C.c_int, []),
    'glr_state_create': (C.c_int, [C.POINTER(StateDesc), C.POINTER(C.c_void_p)]),
    'glr_state_destroy': (C.c_int, [C.c_void_p]),
    'glr_state_unique_masks': (C.c_int, [C.c_void_p]),
    'glr_run_block': (C.c_int, [C.c_void_p, C.POINTER(BlockDesc), C.POINTER(BlockOut)]),
    'glr_block_submit': (C.c_int, [C.c_void_p, C.c_int, C.POINTER(BlockDesc), C.POINTER(BlockOut)]),
    'glr_block_wait': (C.c_int, [C.c_void_p, C.c_int]),
    'glr_block_timing': (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Timing)]),
    'glr_lane_stream': (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    'glr_host_alloc': (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    'glr_host_free': (C.c_int, [C.c_void_p]),
    'glr_t_two_sided_pvalues': (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_double, C.c_void_p]),
    'glr_decode_block': (C.c_int, [C.c_int, C.POINTER(BlockDesc), C.c_int64, C.c_void_p]),
    'glr_tensor_i8_issue_rate': (C.c_int, [C.c_int, C.c_double, C.POINTER(C.c_double)]),
    'glr_logreg_state_create': (C.c_int, [C.POINTER(LogRegDesc), C.POINTER(C.c_void_p)]),
    'glr_logreg_state_destroy': (C.c_int, [C.c_void_p]),
    'glr_logreg_run_block': (C.c_int, [C.c_void_p, C.POINTER(BlockDesc), C.c_void_p, C.c_void_p]),
    'glr_device_alloc': (C.c_int, [C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]),
    'glr_device_free': (C.c_int, [C.c_int, C.c_void_p]),
    'glr_device_upload': (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]),
    'glr_bgen_to_plink2': (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_double, C.c_void_p,
                                     C.c_int64]),
    'glr_loco_predict': (C.c_int, [C.c_int, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    'glr_comm_init': (C.c_int, [C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_void_p)]),
    'glr_comm_destroy': (C.c_int, [C.c_void_p]),
    'glr_comm_size': (C.c_int, [C.c_void_p]),
    'glr_comm_transport': (C.c_char_p, [C.c_void_p]),
    'glr_comm_broadcast_state': (C.c_int, [C.c_void_p, C.POINTER(StateDesc), C.POINTER(C.c_void_p)]),
}

_lib = None


class GlowB200Error(RuntimeError):
    pass


def load():
    """Loads the C-ABI library.  Raises if it has not been built (python -m glow_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GlowB200Error(f'{LIB_PATH} is missing: build it with `python -m glow_b200.build` '
                            '(glow_b200 has no CPU fallback)')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.glr_abi_version() != GLR_ABI_VERSION:
        raise GlowB200Error('libglow_b200.so ABI version mismatch: rebuild the library')
    _lib = lib
    return lib


def check(rc):
    if rc == 0:
        return
    msg = load().glr_last_error().decode('utf-8', 'replace')
    if rc == ERR_INVALID:
        raise ValueError(msg)
    raise GlowB200Error(f'glow_b200 error {rc}: {msg}')


def ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def pinned_empty(shape, dtype):
    """numpy array backed by pinned host memory (freed when the array is garbage collected)."""
    lib = load()
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    check(lib.glr_host_alloc(max(n, 1), C.byref(p)))
    buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    class _Owner:
        def __init__(self, addr):
            self.addr = addr

        def __del__(self):
            try:
                lib.glr_host_free(C.c_void_p(self.addr))
            except Exception:
                pass

    holder = _Owner(p.value)
    # keep the owner alive as long as any view of the buffer lives
    buf._owner = holder
    return arr
