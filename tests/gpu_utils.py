import numpy as np


def has_gpu():
    try:
        from glow_b200 import _capi
        return _capi.load().glr_device_count() > 0
    except Exception:
        return False


def pack_plink(states: np.ndarray) -> np.ndarray:
    """states int [G, N] with values {0,1,2,-1} -> .bed blocks uint8 [G, ceil(N/4)]
    (code 00 -> 2 alt, 01 -> missing, 10 -> 1, 11 -> 0; sample s in byte s//4 at bits 2*(s%4))."""
    G, N = states.shape
    code = np.empty((G, N), dtype=np.uint8)
    code[states == 2] = 0
    code[states == -1] = 1
    code[states == 1] = 2
    code[states == 0] = 3
    nb = (N + 3) // 4
    padded = np.zeros((G, nb * 4), dtype=np.uint8)  # PLINK pads the last byte with 0 bits
    padded[:, :N] = code
    return (padded[:, 0::4] | (padded[:, 1::4] << 2) | (padded[:, 2::4] << 4) | (padded[:, 3::4] << 6)).astype(np.uint8)
