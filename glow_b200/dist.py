"""Multi-GPU plumbing: variants are sharded across ranks (one process per GPU), the regression state
is replicated with ONE broadcast, and there is no per-block collective.

The reference parallelises the same way -- Spark partitions of variant rows, state shipped to every
worker by closure capture (lin_reg.py:277-284).  torch.distributed is only the transport: NCCL over
NVLink on the GPUs, gloo in the CPU tests.
"""
from typing import Dict, Optional, Tuple

import numpy as np

STATE_FIELDS = ('Q', 'Y', 'Y_mask', 'YdotY', 'Y_scale', 'Y_raw')


def variant_shard(n_variants: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [start, stop) of rank `rank`: concatenating the ranks' results in rank
    order reproduces the single-GPU variant order exactly."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError('bad rank / world size')
    return (n_variants * rank) // world, (n_variants * (rank + 1)) // world


def pack_state(Q, Y, Y_mask, YdotY, Y_scale, dof, Y_raw=None) -> Tuple[np.ndarray, np.ndarray]:
    """Flattens the state into one float64 payload + an int64 header (N, K, P, C, verbose, dof)."""
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    if Y.ndim == 2:
        Y = Y[None]
    C, N, P = Y.shape
    K = Q.shape[1] if Q.ndim == 2 else 0
    YdotY = np.ascontiguousarray(np.atleast_2d(YdotY), dtype=np.float64)
    parts = [Q.ravel(), Y.ravel(), np.ascontiguousarray(Y_mask, dtype=np.float64).ravel(), YdotY.ravel(),
             np.ascontiguousarray(Y_scale, dtype=np.float64).ravel()]
    if Y_raw is not None:
        parts.append(np.ascontiguousarray(Y_raw, dtype=np.float64).ravel())
    header = np.array([N, K, P, C, int(Y_raw is not None), int(dof)], dtype=np.int64)
    return np.concatenate(parts), header


def state_layout(header) -> Dict[str, Tuple[int, Tuple[int, ...]]]:
    """Field -> (offset in doubles, shape) inside the flat payload."""
    N, K, P, C, verbose, _ = (int(x) for x in header)
    shapes = [('Q', (N, K)), ('Y', (C, N, P)), ('Y_mask', (N, P)), ('YdotY', (C, P)), ('Y_scale', (P, ))]
    if verbose:
        shapes.append(('Y_raw', (N, P)))
    out, off = {}, 0
    for name, shp in shapes:
        out[name] = (off, shp)
        off += int(np.prod(shp))
    out['_total'] = (off, ())
    return out


def unpack_state(flat: np.ndarray, header) -> Dict[str, Optional[np.ndarray]]:
    lay = state_layout(header)
    res = {k: None for k in STATE_FIELDS}
    for name, (off, shp) in lay.items():
        if name != '_total':
            res[name] = flat[off:off + int(np.prod(shp))].reshape(shp)
    res['dof'] = int(header[5])
    return res


def broadcast_state(state: Optional[dict], src: int = 0, device=None):
    """One broadcast of the payload (plus a 6-integer header) from rank `src`.

    `state` (on rank src) holds Q, Y, Y_mask, YdotY, Y_scale, dof[, Y_raw]; other ranks pass None.
    Returns (flat torch tensor on `device`, header ndarray).  On CUDA the tensor stays on the
    device and DeviceState is built from its addresses (memspace=DEVICE): no host round trip."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_initialized() else 0
    dev = device if device is not None else 'cpu'
    hdr = torch.zeros(6, dtype=torch.int64, device=dev)
    flat_np = None
    if rank == src:
        flat_np, header = pack_state(state['Q'], state['Y'], state['Y_mask'], state['YdotY'], state['Y_scale'],
                                     state['dof'], state.get('Y_raw'))
        hdr.copy_(torch.from_numpy(header))
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(hdr, src=src)
    header = hdr.cpu().numpy()
    total = state_layout(header)['_total'][0]
    flat = torch.empty(total, dtype=torch.float64, device=dev)
    if rank == src:
        flat.copy_(torch.from_numpy(flat_np))
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat, src=src)  # the single data-path collective of the whole run
    return flat, header


def device_state_from_flat(flat, header, device_index: int, n_lanes: int = 2, keep_idx=None, n_geno_samples=None,
                           options=None):
    """Builds this rank's DeviceState straight from the broadcast buffer (CUDA tensor)."""
    from . import _capi
    from .engine import DeviceState
    lay = state_layout(header)
    N, K, P, C, verbose, dof = (int(x) for x in header)
    base = flat.data_ptr()
    addr = lambda name: base + 8 * lay[name][0] if name in lay else None
    return DeviceState(Q=addr('Q'), Y=addr('Y'), Y_mask=addr('Y_mask'), YdotY=addr('YdotY'), Y_scale=addr('Y_scale'),
                       Y_raw=addr('Y_raw') if verbose else None, dof=dof, keep_idx=keep_idx,
                       n_geno_samples=n_geno_samples, device=device_index, n_lanes=n_lanes,
                       memspace=_capi.MEM_DEVICE, dims=(N, K, P, C), options=options)
