#!/usr/bin/env python
"""A few blocks of the logistic score test at the configs[2] shape, for ncu:
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^(split8|square|logreg|split8_finalize) --csv \\
      --log-file gpurun_out/logreg_launches.csv python benchmarks/profile_logreg.py
Launch order per block: split8_kernel on [gamma o C | Y_res] (fused counts), square_kernel (mu -> mu^2), split8_kernel
with the squared-call table on the gamma column, logreg_epilogue_kernel."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from glow_b200.engine import ResidentBlock
    from glow_b200.gwas import log_reg as lr
    dev = torch.device('cuda', 0)
    rg = np.random.default_rng(17)
    N, K, G = 500_000, 17, int(os.environ.get('PROFILE_G', 37_888))
    Cm = np.column_stack([np.ones(N), rg.standard_normal((N, K - 1))])
    y = (rg.random(N) < 1.0 / (1.0 + np.exp(-(0.3 * Cm[:, 1] - 0.5)))).astype(np.float64)
    y_res, gamma, inv = lr._prepare_one_phenotype(Cm, y, None)
    state = lr.LogRegDeviceState(Cm, [lr.LogRegState(inv[None], gamma[:, None], y_res[:, None], None)], device=0)
    X = bench.gen_packed_device(torch, G, N, 9, dev)
    blk = ResidentBlock(X.data_ptr(), X.shape[1], G, 'plink2', 'mean', 0)
    for _ in range(int(os.environ.get('PROFILE_REPS', 3))):
        t0 = time.perf_counter()
        res = state.run(blk)
        print('block', G, 'variants:', round((time.perf_counter() - t0) * 1e3, 3), 'ms; finite', float(np.isfinite(res['pvalue']).mean()))
    state.close()


if __name__ == '__main__':
    main()
