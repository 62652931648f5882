#!/usr/bin/env python
"""Device-timed throughput of the block loop on the other BASELINE.json configs (bench.py covers
configs[2]).  Inputs are generated on the device; results stay in HBM.  Prints one JSON line per
config with stage times and the achieved TFLOP/s (executed flops) and genotype GB/s.

  python benchmarks/config_sweep.py [cfg2 cfg4 cfg5 cfg3]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def prep_state(rg, N, K_cov, P, missing=0.0, n_contigs=1):
    C = np.hstack([np.ones((N, 1)), rg.standard_normal((N, K_cov))])
    Q = np.linalg.qr(C)[0]
    Y = rg.standard_normal((N, P))
    mask = np.ones((N, P))
    if missing:
        mask = (rg.random((N, P), dtype=np.float32) >= missing).astype(np.float64)
        Y *= mask
    Y -= Y.mean(axis=0)
    Y = (Y - Q @ (Q.T @ Y)) * mask
    scale = np.sqrt(np.sum(Y**2, axis=0) / (mask.sum(axis=0) - Q.shape[1]))
    Y /= scale[None, :]
    Ys, ydys = [], []
    for c in range(n_contigs):
        Yc = Y if n_contigs == 1 else (Y - 0.1 * rg.standard_normal((N, P))) * mask
        Ys.append(Yc)
        ydys.append(np.sum(Yc * Yc, axis=0))
    return Q, np.stack(Ys), mask, np.stack(ydys), scale, N - C.shape[1] - 1


def run(name, N, K_cov, P, G, fmt, missing=0.0, n_contigs=1, steps=3, warmup=2):
    import torch
    from bench import gen_packed_device
    from glow_b200.engine import DeviceState
    dev = torch.device('cuda', 0)
    rg = np.random.default_rng(11)
    Q, Y, mask, ydy, scale, dof = prep_state(rg, N, K_cov, P, missing, n_contigs)
    st = DeviceState(Q=Q, Y=Y, Y_mask=mask, YdotY=ydy, Y_scale=scale, dof=dof)
    del Y, mask
    if fmt == 'plink2':
        X = gen_packed_device(torch, G, N, 5, dev)
        stride = X.shape[1]
    else:
        X = torch.empty((G, N), dtype=torch.float64, device=dev)
        for g0 in range(0, G, 4096):
            g1 = min(G, g0 + 4096)
            f = torch.rand((g1 - g0, 1), device=dev) * 0.49 + 0.01
            u = torch.rand((g1 - g0, N), device=dev)
            X[g0:g1] = (u > (1 - f)**2).double() + (u > (1 - f)**2 + 2 * f * (1 - f)).double() + 0.1 * torch.rand(
                (g1 - g0, N), device=dev).double()
        stride = N * 8
    outs = {k: torch.empty((P, G), dtype=torch.float64, device=dev) for k in ('effect', 'stderror', 'tvalue', 'pvalue')}
    addrs = {k: v.data_ptr() for k, v in outs.items()}
    torch.cuda.synchronize()
    acc = {}
    for i in range(warmup + steps):
        st.run_device(fmt, X.data_ptr(), stride, G, addrs, contig=0, missing='mean')
        if i >= warmup:
            t = st.timing(0)
            for k, v in t.items():
                acc[k] = acc.get(k, 0.0) + v / steps
    K = K_cov + 1
    U = st.unique_masks
    flops_gram = 2.0 * N * (K + P) * G
    flops_masked = float(N) * G * (2.0 * ((K + 3) // 4) * 4 + 2 + 2.0 * U) if U else 0.0
    geno_bytes = G * stride
    line = dict(config=name, N=N, K=K, P=P, G=G, format=fmt, unique_masks=U, tests=P * G,
                tests_per_s=P * G / (acc['total_ms'] / 1e3), stage_ms={k: round(v, 4) for k, v in acc.items()},
                gram_tflops=flops_gram / (acc['gram_ms'] / 1e3) / 1e12,
                masked_tflops=(flops_masked / (acc['masked_ms'] / 1e3) / 1e12) if U else None,
                genotype_gbs_in_gram=geno_bytes / (acc['gram_ms'] / 1e3) / 1e9,
                result_gbs_in_epilogue=4 * 8 * P * G / (acc['epilogue_ms'] / 1e3) / 1e9,
                finite=float(torch.isfinite(outs['pvalue']).double().mean().item()))
    print(json.dumps(line), flush=True)
    st.close()


import os as _os
CONFIGS = {
    # configs[1]: dense fp64 dosages, 10k x 100k x 10 phenotypes x 10 covariates (whole config in one block)
    'cfg2': dict(N=10_000, K_cov=10, P=10, G=100_000, fmt='f64'),
    # configs[2]: bench.py's workload, here for one-stop comparison
    'cfg3': dict(N=500_000, K_cov=16, P=1, G=int(_os.environ.get('SWEEP_G', 37_888)), fmt='plink2'),
    # configs[3]: 50k samples x 2 000 phenotypes with 10 % missing each (distinct masks), 2-bit; G scaled to one block
    'cfg4': dict(N=50_000, K_cov=10, P=2_000, G=8_192, fmt='plink2', missing=0.1),
    # configs[4]: 200k samples x 20 phenotypes, LOCO (2 of 22 contig states built here), 2-bit
    'cfg5': dict(N=200_000, K_cov=16, P=20, G=37_888, fmt='plink2', n_contigs=2),
    # not a BASELINE config: the UK-Biobank shape with missing phenotype values (the usual case in practice)
    'cfg3m': dict(N=500_000, K_cov=16, P=1, G=37_888, fmt='plink2', missing=0.02),
    'cfg5m': dict(N=200_000, K_cov=16, P=20, G=37_888, fmt='plink2', missing=0.02),
}

if __name__ == '__main__':
    which = sys.argv[1:] or ['cfg2', 'cfg4', 'cfg5']
    for w in which:
        run(w, **CONFIGS[w])
