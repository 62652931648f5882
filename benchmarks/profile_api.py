import cProfile, pstats, os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, pandas as pd, torch
import glow_b200, bench
dev = torch.device('cuda', 0)
N, G = 500_000, 32768
packed = bench.gen_packed_device(torch, G, N, 5, dev).cpu().numpy()
path = '/dev/shm/x.bed'
with open(path, 'wb') as f:
    f.write(bytes([0x6c, 0x1b, 0x01])); f.write(packed.tobytes())
del packed
rg = np.random.default_rng(3)
cov = pd.DataFrame(rg.standard_normal((N, 16))); phe = pd.DataFrame(rg.standard_normal((N, 1)), columns=['t'])
meta = pd.DataFrame({'idx': np.arange(G)})
src = glow_b200.PlinkBedSource(path, N, meta, block_variants=8192)
for i in range(2):
    t0 = time.perf_counter(); out = glow_b200.gwas.linear_regression(src, phe, cov); print('call', i, time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable()
out = glow_b200.gwas.linear_regression(src, phe, cov)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
os.unlink(path)
