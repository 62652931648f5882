"""Randomised parity sweep (not part of the test suite: run by hand on a GPU box).  Small random shapes around the tile /
column-block / stage boundaries of the kernels, every forced path, against the oracle at the north_star tolerances.

  python benchmarks/fuzz_parity.py [n_cases] [seed]
"""
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    from glow_b200.engine import DeviceState, GenotypeBlock
    from gpu_utils import pack_plink
    from helpers import assert_stats_close
    from oracle import linreg_oracle as orc
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rg = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    edges_n = [129, 130, 255, 256, 257, 383, 384, 385, 511, 512, 513, 640, 1000, 1025, 2049, 4097, 6150]
    edges_g = [1, 7, 8, 17, 18, 19, 35, 36, 37, 127, 128, 129, 255, 256, 257, 300]
    modes = [dict(), dict(GLR_SPLIT8='2', GLR_S8_PAIR='2'), dict(GLR_SPLIT8='2', GLR_S8_PAIR='0'), dict(GLR_SPLIT8='0'),
             dict(GLR_MASK8='2', GLR_SPLIT8='2'), dict(GLR_MASK_SPARSE='2'), dict(GLR_MASK_SPARSE='0', GLR_MASK8='0'),
             # round 2: sample-range split of the int8 kernel, optimistic form (replayed when the block has missing calls)
             dict(GLR_SPLIT8='2', GLR_SPLIT='3'), dict(GLR_SPLIT8='2', GLR_S8_PAIR='0', GLR_SPLIT='2'),
             dict(GLR_SPLIT8='2', GLR_OPTIMISTIC='2'), dict(GLR_SPLIT8='2', GLR_S8_PAIR='0', GLR_SPLIT='5', GLR_OPTIMISTIC='2'),
             dict(GLR_SPLIT8='2', GLR_Q_DIGITS='7'), dict(GLR_SPLIT8='2', GLR_Q_DIGITS='3', GLR_SPLIT='2'),
             dict(GLR_SPLIT8='2', GLR_Q_DIGITS='6', GLR_OPTIMISTIC='2')]
    knobs = ('GLR_SPLIT8', 'GLR_S8_PAIR', 'GLR_MASK8', 'GLR_MASK_SPARSE', 'GLR_SPLIT', 'GLR_OPTIMISTIC', 'GLR_Q_DIGITS')
    done = 0
    for case in range(n_cases):
        N = int(rg.choice(edges_n))
        G = int(rg.choice(edges_g))
        K = int(rg.choice([0, 1, 2, 5, 16, 17, 18, 20, 33, 70]))
        P = int(rg.choice([1, 2, 3, 17, 18, 19, 36, 40]))
        if N - K - 2 < 5:
            continue
        add_icpt = bool(rg.random() < 0.8) or K == 0
        miss_ph = float(rg.choice([0.0, 0.0, 0.03, 0.3]))
        miss_gt = float(rg.choice([0.0, 0.02, 0.2]))
        fmt = str(rg.choice(['plink2', 'plink2', 'i8', 'f64']))
        policy = str(rg.choice(['mean', 'minus_one']))
        states = rg.integers(0, 3, (G, N))
        states[:, 0] = np.arange(G) % 3
        states[:, -1] = (np.arange(G) + 1) % 3
        if miss_gt:
            states[rg.random((G, N)) < miss_gt] = -1
            states[:, :3] = [0, 1, 2]
        Y = rg.standard_normal((N, P))
        if miss_ph:
            Y[rg.random((N, P)) < miss_ph] = np.nan
            Y[:K + 8] = rg.standard_normal((K + 8, P))  # enough observed rows per phenotype
        phe = pd.DataFrame(Y)
        cov = pd.DataFrame(rg.standard_normal((N, K))) if K else None
        st = orc.prepare_state(phe, cov, add_intercept_col=add_icpt)
        X = states.astype(np.float64)
        Xo = orc.mean_substitute(X) if policy == 'mean' else X
        ref = orc.block_stats(Xo, st)
        if fmt == 'plink2':
            blk = GenotypeBlock(pack_plink(states), 'plink2', policy)
        elif fmt == 'i8':
            blk = GenotypeBlock(states.astype(np.int8), 'i8', policy)
        else:
            blk = GenotypeBlock(Xo, 'f64')
        for mode in modes:
            for k in knobs:
                os.environ.pop(k, None)
            os.environ.update(mode)
            ys = [st.y_state]
            try:
                eng = DeviceState(Q=st.Q, Y=np.stack([y.Y for y in ys]), Y_mask=st.Y_mask, YdotY=np.stack([y.YdotY for y in ys]),
                                  Y_scale=st.Y_scale, dof=st.dof)
            except ValueError as exc:
                if 'covariates' in str(exc):
                    continue
                raise
            got = eng.run(blk)
            eng.close()
            label = f'case {case} N={N} G={G} K={K} P={P} icpt={add_icpt} miss_ph={miss_ph} miss_gt={miss_gt} {fmt} {policy} {mode}'
            assert_stats_close(got, ref, label=label)
        done += 1
    print(f'fuzz ok: {done} cases x {len(modes)} modes')


if __name__ == '__main__':
    main()
