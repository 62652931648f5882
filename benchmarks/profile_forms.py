#!/usr/bin/env python
"""A few launches of each form of the first-pass kernel at the configs[2] shape, for ncu:
  ncu --set full --clock-control none --import-source on -k regex:^split8_kernel -o gpurun_out/r02_split8 python benchmarks/profile_forms.py
Launch order (after one warm-up each): full form (1 % missing calls), optimistic form (fully called), 10 000-variant
block (sample range split)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from glow_b200.engine import DeviceState
    dev = torch.device('cuda', 0)
    N, G = 500_000, int(os.environ.get('PROFILE_G', 37_888))
    Q, Y, mask, ydy, scale, dof = bench.host_state(N)
    st = DeviceState(Q=Q, Y=Y, Y_mask=mask, YdotY=ydy, Y_scale=scale, dof=dof, options={'optimistic_calls': 'off'})
    st_opt = DeviceState(Q=Q, Y=Y, Y_mask=mask, YdotY=ydy, Y_scale=scale, dof=dof, options={'optimistic_calls': 'force'})
    packed = bench.gen_packed_device(torch, G, N, 5, dev)
    bench.MISSING_RATE = 0.0
    full = bench.gen_packed_device(torch, G, N, 6, dev)
    outs = {k: torch.empty((1, G), dtype=torch.float64, device=dev) for k in ('effect', 'stderror', 'tvalue', 'pvalue')}
    addrs = {k: v.data_ptr() for k, v in outs.items()}
    rb = packed.shape[1]
    reps = int(os.environ.get('PROFILE_REPS', 2))
    for _ in range(reps):
        st.run_device('plink2', packed.data_ptr(), rb, G, addrs, missing='mean')
    print('full form', st.timing(0))
    for _ in range(reps):
        st_opt.run_device('plink2', full.data_ptr(), rb, G, addrs, missing='mean')
    print('optimistic form', st_opt.timing(0))
    for _ in range(reps):
        st.run_device('plink2', packed.data_ptr(), rb, 10_000, addrs, missing='mean')
    print('10k block', st.timing(0))


if __name__ == '__main__':
    main()
