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


def gen_packed_gpu(n_variants, n_samples, seed, missing=0.01, device=0, chunk=256):
    """PLINK 2-bit block generated on the GPU with torch (plumbing only): per-variant MAF ~ U(0.01, 0.5), Hardy-Weinberg
    frequencies, `missing` missing calls (code 01).  Returns the packed block as a host uint8 array [G, ceil(N/4)]."""
    import torch
    dev = torch.device('cuda', device)
    row_bytes = (n_samples + 3) // 4
    out = torch.empty((n_variants, row_bytes), dtype=torch.uint8, device=dev)
    npad = row_bytes * 4
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    for c0 in range(0, n_variants, chunk):
        c1 = min(n_variants, c0 + chunk)
        f = torch.rand((c1 - c0, 1), generator=g, device=dev) * 0.49 + 0.01
        u = torch.rand((c1 - c0, npad), generator=g, device=dev)
        p0 = (1 - f)**2
        p1 = p0 + 2 * f * (1 - f)
        code = torch.where(u < p0, 3, torch.where(u < p1, 2, 0)).to(torch.uint8)
        if missing > 0:
            m = torch.rand((c1 - c0, npad), generator=g, device=dev) < missing
            code = torch.where(m, torch.ones_like(code), code)
        code[:, n_samples:] = 0
        code = code.view(c1 - c0, row_bytes, 4)
        out[c0:c1] = code[:, :, 0] | (code[:, :, 1] << 2) | (code[:, :, 2] << 4) | (code[:, :, 3] << 6)
    host = out.cpu().numpy()
    del out
    torch.cuda.empty_cache()
    return host


def unpack_states(packed_rows, n_samples):
    """packed uint8 [g, ceil(N/4)] -> int8 states {0,1,2,-1} [g, N] (same mapping as pack_plink)."""
    lut = np.array([2, -1, 1, 0], dtype=np.int8)
    codes = np.empty((packed_rows.shape[0], packed_rows.shape[1] * 4), dtype=np.uint8)
    for j in range(4):
        codes[:, j::4] = (packed_rows >> (2 * j)) & 3
    return lut[codes[:, :n_samples]]


def write_bgen(path, probs, missing=None, compress=True, bits=8, contig='1'):
    """Minimal BGEN v1.2 layout-2 writer for tests: probs integer [G, N, 2] = stored P(hom ref), P(het) (unphased, diploid,
    biallelic); missing: bool [G, N] or None."""
    import struct
    import zlib
    G, N, _ = probs.shape
    dt = {8: np.uint8, 16: '<u2', 32: '<u4'}[bits]
    body = b''
    for g in range(G):
        ident = f'id{g}'.encode()
        rs = f'rs{g}'.encode()
        ch = contig.encode() if isinstance(contig, str) else str(contig[g]).encode()
        rec = struct.pack('<H', len(ident)) + ident + struct.pack('<H', len(rs)) + rs + struct.pack('<H', len(ch)) + ch
        rec += struct.pack('<IH', 1000 + g, 2) + struct.pack('<I', 1) + b'A' + struct.pack('<I', 1) + b'G'
        ploidy = np.full(N, 2, dtype=np.uint8)
        if missing is not None:
            ploidy[missing[g]] |= 0x80
        block = struct.pack('<IHBB', N, 2, 2, 2) + ploidy.tobytes() + struct.pack('<BB', 0, bits) + \
            np.ascontiguousarray(probs[g], dtype=dt).tobytes()
        if compress:
            z = zlib.compress(block)
            rec += struct.pack('<I', len(z) + 4) + struct.pack('<I', len(block)) + z
        else:
            rec += struct.pack('<I', len(block)) + block
        body += rec
    flags = (1 if compress else 0) | (2 << 2)
    header = struct.pack('<III', 20, G, N) + b'bgen' + struct.pack('<I', flags)
    with open(path, 'wb') as f:
        f.write(struct.pack('<I', 20) + header + body)
