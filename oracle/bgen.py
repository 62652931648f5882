"""TEST INFRASTRUCTURE ONLY -- minimal BGEN v1.2 (layout 2, zlib) reader for the regenie fixtures.

Restates what the parity tests need of the reference's BGEN datasource
(core/src/main/scala/io/projectglow/bgen/BgenFileIterator.scala): unphased, diploid, biallelic
probability blocks -> hard calls with Glow's default threshold 0.9
(core/.../common/datasourceOptions.scala `hardCallThreshold`, BgenSchemaInferrer.scala:42) ->
genotype_states (number of alternate alleles, -1 when no call clears the threshold or the sample
is missing).  Used only by oracle/make_golden.py to build tests/golden/regenie_*.npz.
"""
import struct
import zlib

import numpy as np


def read_bgen_states(path, hard_call_threshold=0.9):
    """Returns (contigs, positions(1-based), rsids, states int8 [M, N])."""
    with open(path, 'rb') as f:
        data = f.read()
    offset, = struct.unpack_from('<I', data, 0)
    hlen, M, N = struct.unpack_from('<III', data, 4)
    assert data[16:20] in (b'bgen', b'\0\0\0\0')
    flags, = struct.unpack_from('<I', data, 4 + hlen - 4)
    compression = flags & 3
    layout = (flags >> 2) & 15
    assert layout == 2 and compression in (0, 1), (layout, compression)
    pos = 4 + offset
    contigs, positions, rsids, rows = [], [], [], []
    for _ in range(M):
        lid, = struct.unpack_from('<H', data, pos); pos += 2 + lid
        lrs, = struct.unpack_from('<H', data, pos); pos += 2
        rsid = data[pos:pos + lrs].decode(); pos += lrs
        lchr, = struct.unpack_from('<H', data, pos); pos += 2
        chrom = data[pos:pos + lchr].decode(); pos += lchr
        vpos, nalleles = struct.unpack_from('<IH', data, pos); pos += 6
        for _a in range(nalleles):
            la, = struct.unpack_from('<I', data, pos); pos += 4 + la
        clen, = struct.unpack_from('<I', data, pos); pos += 4
        if compression == 1:
            dlen, = struct.unpack_from('<I', data, pos)
            block = zlib.decompress(data[pos + 4:pos + clen])
            assert len(block) == dlen
        else:
            block = data[pos:pos + clen]
        pos += clen
        n, k, pmin, pmax = struct.unpack_from('<IHBB', block, 0)
        assert n == N and k == 2 and pmin == 2 and pmax == 2, 'only diploid biallelic blocks are supported'
        ploidy = np.frombuffer(block, dtype=np.uint8, count=n, offset=8)
        missing = (ploidy & 0x80) != 0
        phased, bits = struct.unpack_from('<BB', block, 8 + n)
        assert phased == 0 and bits in (8, 16, 32), (phased, bits)
        dt = {8: np.uint8, 16: '<u2', 32: '<u4'}[bits]
        probs = np.frombuffer(block, dtype=dt, count=2 * n, offset=10 + n).astype(np.float64) / (2.0**bits - 1)
        p_aa, p_ab = probs[0::2], probs[1::2]
        p_bb = 1.0 - p_aa - p_ab
        p = np.stack([p_aa, p_ab, p_bb], axis=1)
        best = p.argmax(axis=1)
        called = (p.max(axis=1) >= hard_call_threshold) & ~missing
        rows.append(np.where(called, best, -1).astype(np.int8))  # genotype index = number of alt alleles
        contigs.append(chrom)
        positions.append(vpos)
        rsids.append(rsid)
    return contigs, np.array(positions), rsids, np.stack(rows)
