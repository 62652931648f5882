"""Run-to-run determinism of a block (debugging aid): runs the same float64 block several times and reports which
(phenotype row, variant) results differ from the first run.  python benchmarks/determinism_check.py [G] [reps]"""
import sys, os, numpy as np, pandas as pd
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from glow_b200.engine import DeviceState, GenotypeBlock
from oracle import linreg_oracle as orc
rg = np.random.default_rng(1)
N, P, K = int(os.environ.get("DET_N", 10_000)), int(os.environ.get("DET_P", 10)), 10
G = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cov = pd.DataFrame(rg.standard_normal((N, K)))
X = rg.integers(0, 3, (G, N)).astype(np.float64) + 0.1 * rg.random((G, N), dtype=np.float32)
phe = pd.DataFrame(rg.standard_normal((N, P)))
st = orc.prepare_state(phe, cov)
FMT = os.environ.get('DET_FMT', 'f64')
if FMT == 'f32':
    X = X.astype(np.float32)
elif FMT == 'plink2':
    from gpu_utils import pack_plink
    X = pack_plink(np.floor(X).astype(np.int64))
PIN = bool(int(os.environ.get('DET_PIN', '0')))
if PIN:
    from glow_b200 import _capi
    Xp = _capi.pinned_empty(X.shape, X.dtype)
    Xp[:] = X
    X = Xp
for opts in ({'sample_split': 1, 'split8': 'off'}, {'sample_split': 3, 'split8': 'off'}):
    eng = DeviceState(Q=st.Q, Y=st.y_state.Y[None], Y_mask=st.Y_mask, YdotY=st.y_state.YdotY[None], Y_scale=st.Y_scale, dof=st.dof, options=opts)
    runs = [eng.run(GenotypeBlock(X, FMT, 'mean', pinned=PIN)) for _ in range(reps)]
    for rep in range(1, reps):
        for k in ('effect', 'stderror'):
            d = np.argwhere(runs[rep][k] != runs[0][k])
            if len(d):
                rows = sorted(set(d[:, 0].tolist()))
                vs = d[:, 1]
                print(opts, 'rep', rep, k, len(d), 'rows', rows, 'variants', int(vs.min()), '..', int(vs.max()), 'n_distinct', len(set(vs.tolist())),
                      'mod128', sorted(set((vs % 128).tolist()))[:20])
    print(opts, 'done')
    eng.close()
