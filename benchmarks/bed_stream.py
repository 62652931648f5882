"""End-to-end throughput of the PUBLIC API on a PLINK .bed FILE: glow_b200.gwas.linear_regression(PlinkBedSource(path)).
The file is synthetic (UK-Biobank shape, written to tmpfs so that the storage device is not what is measured); the timed
region covers everything a user waits for: streaming the file into pinned staging buffers, H2D, kernels, D2H and the
pandas result frame.

  python benchmarks/bed_stream.py [n_variants] [block_variants]
"""
import json
import os
import sys
import time

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import glow_b200
    from bench import gen_packed_device
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    bv = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    N, K = 500_000, 16
    dev = torch.device('cuda', 0)
    packed = gen_packed_device(torch, G, N, 5, dev).cpu().numpy()
    path = '/dev/shm/glow_b200_stream.bed' if os.path.isdir('/dev/shm') else '/tmp/glow_b200_stream.bed'
    with open(path, 'wb') as f:
        f.write(bytes([0x6c, 0x1b, 0x01]))
        f.write(packed.tobytes())
    del packed
    rg = np.random.default_rng(3)
    cov = pd.DataFrame(rg.standard_normal((N, K)))
    phe = pd.DataFrame(rg.standard_normal((N, 1)), columns=['trait'])
    meta = pd.DataFrame({'idx': np.arange(G)})
    out = None
    times = []
    for rep in range(3):
        t0 = time.perf_counter()
        out = glow_b200.gwas.linear_regression(glow_b200.PlinkBedSource(path, N, meta, block_variants=bv), phe, cov)
        times.append(time.perf_counter() - t0)
    os.unlink(path)
    best = min(times[1:])
    print(json.dumps({'n_variants': G, 'n_samples': N, 'block_variants': bv, 'seconds': times,
                      'tests_per_s': G / best, 'file_gb_per_s': G * ((N + 3) // 4) / best / 1e9,
                      'finite': float(np.isfinite(out['pvalue']).mean())}))


if __name__ == '__main__':
    main()
