#!/usr/bin/env python
"""Benchmark of the linear-regression GWAS block loop on B200 (BASELINE.json metric:
variant x phenotype tests/sec, device-timed).

Workload (config.workload): BASELINE.json configs[2], the configuration the metric is quoted on --
UK-Biobank shape, N = 500 000 samples, 16 covariates + intercept (K = 17), P = 1 phenotype,
PLINK 2-bit packed genotypes (125 000 B per variant, 1 % missing calls, mean-imputed inside the
decode).  One batch = 37 888 variants per GPU (4.7 GB packed, far beyond L2); one STEP = `--batches-per-step`
batches back to back (default 16: 20 steps are a 0.9 s timed region, long enough for the board to settle at its
power-capped clock).  Variants are independent, so N GPUs each take their own batches (weak scaling, no data-path
collective; the state is replicated with one NCCL broadcast before the timed region).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

`value`  : tests/s with the packed batch resident in HBM, timed with CUDA events on the stream
           the kernels are launched on, max over ranks.
`e2e`    : tests/s through the C-ABI block call with PINNED HOST buffers -- H2D copy of the packed
           batch and D2H copy of the four result arrays inside the timed region, sub-blocks
           pipelined over the library's lanes (`--e2e-batches-per-step` batches per step).  `e2e.h2d_ceiling`
           is what bare cudaMemcpyAsync copies of the same buffers reach on this box at this GPU count: the
           number the e2e path is bounded by.
`e2e_resident`: the same call with the batch already resident in HBM (glow_b200.ResidentSource) and
           the results copied to pinned host memory: the steady state of many phenotype sets against one
           loaded genotype shard.
`e2e_api`: the public API on a .bed FILE in tmpfs: glow_b200.gwas.linear_regression(PlinkBedSource(path)).
`roofline`: the dominant kernel (split8_kernel, int8 tensor cores) against the tcgen05.mma issue rate measured in
           this run; `traffic` comes from the committed ncu capture named in `traffic_source`.
`configs`: device-timed tests/s and the dominant kernel's roofline fraction for the other BASELINE.json configs
           (N = 1 only).
`cpu_baseline` / --impl reference: the oracle port of the reference's _linear_regression_inner
           (numpy + BLAS on all host cores) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SAMPLES = 500_000
N_COV = 16
N_PHENO = 1
MISSING_RATE = 0.01
SEED = 20261019


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batches-per-step', type=int, default=16)
    ap.add_argument('--e2e-batches-per-step', type=int, default=4)
    ap.add_argument('--skip-configs', action='store_true', help='do not measure the other BASELINE configs')
    ap.add_argument('--skip-e2e-api', action='store_true')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--variants-per-step', type=int, default=148 * 256)
    ap.add_argument('--samples', type=int, default=N_SAMPLES)
    ap.add_argument('--e2e-subblocks', type=int, default=32)
    ap.add_argument('--skip-cpu-baseline', action='store_true')
    ap.add_argument('--skip-e2e', action='store_true')
    ap.add_argument('--skip-dmma-arm', action='store_true', help='primary path only (ncu launch lists of the step)')
    ap.add_argument('--skip-peak', action='store_true', help='do not run the cuBLAS DGEMM peak measurement (ncu runs)')
    ap.add_argument('--cpu-block-variants', type=int, default=128)
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------------------
def host_state(n, seed=SEED):
    """Covariates N(0,1), one N(0,1) phenotype; driver-side prep exactly as lin_reg.py:142-159."""
    rg = np.random.default_rng(seed)
    C = np.hstack([np.ones((n, 1)), rg.standard_normal((n, N_COV))])
    Q = np.linalg.qr(C)[0]
    Y = rg.standard_normal((n, N_PHENO))
    mask = np.ones_like(Y)
    Y -= Y.mean(axis=0)
    Y -= Q @ (Q.T @ Y)
    scale = np.sqrt(np.sum(Y**2, axis=0) / (mask.sum(axis=0) - Q.shape[1]))
    Y /= scale[None, :]
    ydy = np.sum(Y * Y, axis=0)
    dof = n - C.shape[1] - 1
    return Q, Y, mask, ydy, scale, dof


def gen_packed_device(torch, n_variants, n_samples, seed, device, chunk=256):
    """PLINK 2-bit block generated on the device: per-variant MAF ~ U(0.01, 0.5), Hardy-Weinberg
    genotype frequencies, MISSING_RATE missing calls (code 01).  Seeded per (seed, chunk) so a
    shard is reproducible independently of the GPU count."""
    row_bytes = (n_samples + 3) // 4
    out = torch.empty((n_variants, row_bytes), dtype=torch.uint8, device=device)
    npad = row_bytes * 4
    for c0 in range(0, n_variants, chunk):
        c1 = min(n_variants, c0 + chunk)
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1_000_003 + c0)
        f = torch.rand((c1 - c0, 1), generator=g, device=device) * 0.49 + 0.01
        u = torch.rand((c1 - c0, npad), generator=g, device=device)
        p0 = (1 - f)**2
        p1 = p0 + 2 * f * (1 - f)
        # code: 11 = hom ref (0), 10 = het (1), 00 = hom alt (2), 01 = missing
        code = torch.where(u < p0, 3, torch.where(u < p1, 2, 0)).to(torch.uint8)
        m = torch.rand((c1 - c0, npad), generator=g, device=device) < MISSING_RATE
        code = torch.where(m, torch.ones_like(code), code)
        code[:, n_samples:] = 0
        code = code.view(c1 - c0, row_bytes, 4)
        out[c0:c1] = code[:, :, 0] | (code[:, :, 1] << 2) | (code[:, :, 2] << 4) | (code[:, :, 3] << 6)
        del u, m, code
    return out


def pack_plink_host(states):
    """int8 states {0,1,2,-1} [G, N] -> .bed blocks uint8 [G, ceil(N/4)] (00 = 2, 01 = missing, 10 = 1, 11 = 0)."""
    G, n = states.shape
    lut = np.array([3, 2, 0, 1], dtype=np.uint8)  # state 0,1,2,-1 -> code
    nb = (n + 3) // 4
    code = np.zeros((G, nb * 4), dtype=np.uint8)
    code[:, :n] = lut[states.astype(np.int64) & 3]
    return code[:, 0::4] | (code[:, 1::4] << 2) | (code[:, 2::4] << 4) | (code[:, 3::4] << 6)


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                r = subprocess.run(
                    ['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits'],
                    capture_output=True, text=True, timeout=5)
                if r.returncode == 0 and r.stdout.strip():
                    self.rows.append([x.strip() for x in r.stdout.strip().split(',')])
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit())
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {
            'sm_mhz': sm[len(sm) // 2] if sm else None,
            'sm_max_mhz': float(self.rows[0][1]) if self.rows else None,
            'power_w_max': max((float(r[2]) for r in self.rows if r[2].replace('.', '').isdigit()), default=None),
            'reasons': sorted(reasons),
            'samples': len(self.rows)
        }


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference block kernel on the host cores
# --------------------------------------------------------------------------------------------------
def use_all_host_cores():
    """BLAS on every host core, also under torchrun (which exports OMP_NUM_THREADS=1)."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    return os.cpu_count()


def cpu_reference_arm(packed_host, n_samples, Q, Y, mask, ydy, scale, dof, block_variants, n_blocks, warm=1):
    """Times oracle/linreg_oracle.block_stats (= _linear_regression_inner, lin_reg.py:226-249, incl.
    the np.column_stack transpose) on `n_blocks` blocks of `block_variants` variants; BLAS uses every
    host core.  Decode + mean substitution (JVM side in the reference) are NOT timed."""
    from oracle import linreg_oracle as orc
    use_all_host_cores()
    st = orc.OracleState(Q, mask, scale, dof, ['y0'], orc.OracleYState(Y, ydy), None, None)
    times = []
    for b in range(warm + n_blocks):
        blk = packed_host[b * block_variants:(b + 1) * block_variants]
        X = orc.mean_substitute(orc.decode_plink_states(blk, n_samples))
        t0 = time.perf_counter()
        orc.block_stats(X, st)
        dt = time.perf_counter() - t0
        if b >= warm:
            times.append(dt)
    return block_variants * N_PHENO / (sum(times) / len(times)), times


def run_reference_impl(args):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n = args.samples
    use_all_host_cores()
    Q, Y, mask, ydy, scale, dof = host_state(n)
    bv = args.cpu_block_variants
    rg = np.random.default_rng(SEED)
    total_blocks = args.warmup + args.steps
    from oracle import linreg_oracle as orc
    st = orc.OracleState(Q, mask, scale, dof, ['y0'], orc.OracleYState(Y, ydy), None, None)
    times = []
    for b in range(total_blocks):
        f = rg.uniform(0.01, 0.5, (bv, 1))
        states = rg.binomial(2, f, (bv, n)).astype(np.int8)
        states[rg.random(states.shape, dtype=np.float32) < MISSING_RATE] = -1
        packed = pack_plink_host(states)
        X = orc.mean_substitute(orc.decode_plink_states(packed, n))
        del states
        t0 = time.perf_counter()
        orc.block_stats(X, st)
        if b >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    value = bv * N_PHENO / (ms / 1e3)
    cores = os.cpu_count()
    line = {
        'impl': 'reference',
        'metric': 'variant_x_phenotype_tests_per_sec',
        'value': value,
        'unit': 'tests/s',
        'n_gpus': args.gpus,
        'steps': args.steps,
        'warmup': args.warmup,
        'ms_per_step': ms,
        'higher_is_better': True,
        'scaling': 'weak',
        'vs_baseline': None,
        'dtype': 'f64',
        'data': 'synthetic',
        'config': dict(workload_config(args, bv), batches_per_step=1, variants_per_step_per_gpu=bv),
        'cpu_baseline': {
            'value': value,
            'unit': 'tests/s',
            'cores': cores,
            'kind': 'port',
            'sample': f'{args.steps} blocks of {bv} variants x {n} samples through the oracle port of '
                      '_linear_regression_inner (numpy/OpenBLAS, all host cores); JVM decode, Arrow transfer and '
                      'Spark scheduling of the real reference are not included (no JVM in this image)'
        },
        'e2e': {'value': value, 'unit': 'tests/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0
    }
    print(json.dumps(line), flush=True)


def workload_config(args, variants_per_batch):
    return {
        'workload': 'BASELINE.json configs[2]: UK-Biobank shape, PLINK 2-bit genotypes, variant-sharded',
        'n_samples': args.samples,
        'n_covariates': N_COV,
        'n_phenotypes': N_PHENO,
        'variants_per_batch_per_gpu': variants_per_batch,
        'batches_per_step': args.batches_per_step,
        'variants_per_step_per_gpu': variants_per_batch * args.batches_per_step,
        'missing_call_rate': MISSING_RATE,
        'imputation': 'mean_substitute fused into the decode',
        'cache': 'inputs larger than L2 (packed batch 4.7 GB >> 126 MB); no L2 flush',
        'parallelism': f'variant-sharded x{args.gpus}, state replicated by one NCCL broadcast'
    }


def load_traffic():
    """DRAM bytes per launch of the dominant kernels from the committed ncu --set full captures (profiles/)."""
    try:
        return json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
    except Exception:  # noqa: BLE001
        return {}


# --------------------------------------------------------------------------------------------------
# the other BASELINE.json configs (device-timed, N = 1): tests/s + roofline fraction of the dominant kernel
# --------------------------------------------------------------------------------------------------
def _prep_state(rg, N, K_cov, P, missing=0.0, n_contigs=1):
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
    for _ in range(n_contigs):
        Yc = Y if n_contigs == 1 else (Y - 0.1 * rg.standard_normal((N, P))) * mask
        Ys.append(Yc)
        ydys.append(np.sum(Yc * Yc, axis=0))
    return Q, np.stack(Ys), mask, np.stack(ydys), scale, N - C.shape[1] - 1


def digit_rows(n_q, n_y, path, unique_masks):
    """Digit rows the int8 contraction of a block needed: 7 per phenotype-side column; per covariate column 7, or -- when
    the block ran on the reduced-digit image (glr_options.q_digits = automatic) -- 6 if a masked pass follows, else 5."""
    from glow_b200 import _capi
    qd = 7
    if path & _capi.PATH_Q_DIGITS_LO:
        qd = int(os.environ.get('GLR_Q_DIGITS', '0')) or (6 if unique_masks > 0 else 5)
    return qd * n_q + 7 * n_y


def run_config(torch, dev, name, label, N, K_cov, P, G, fmt, hbm_peak, i8_peak, missing=0.0, n_contigs=1, call_missing=MISSING_RATE,
               min_seconds=0.3):
    from glow_b200 import _capi
    from glow_b200.engine import DeviceState
    rg = np.random.default_rng(11)
    Q, Y, mask, ydy, scale, dof = _prep_state(rg, N, K_cov, P, missing, n_contigs)
    st = DeviceState(Q=Q, Y=Y, Y_mask=mask, YdotY=ydy, Y_scale=scale, dof=dof, device=dev.index)
    del Y, mask
    if fmt in ('plink2', 'i8'):
        global MISSING_RATE
        keep, MISSING_RATE = MISSING_RATE, call_missing
        X = gen_packed_device(torch, G, N, 5, dev)
        MISSING_RATE = keep
        stride = X.shape[1]
        if fmt == 'i8':  # genotype_states as int8: unpack the 2-bit block on the device (00 -> 2, 01 -> -1, 10 -> 1, 11 -> 0)
            lut = torch.tensor([2, -1, 1, 0], dtype=torch.int8, device=dev)
            S8 = torch.empty((G, N), dtype=torch.int8, device=dev)
            for g0 in range(0, G, 2048):
                blk = X[g0:g0 + 2048].to(torch.int64)
                codes = torch.stack([(blk >> (2 * j)) & 3 for j in range(4)], dim=2).reshape(blk.shape[0], -1)[:, :N]
                S8[g0:g0 + 2048] = lut[codes]
                del blk, codes
            X, stride = S8, N
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
    for _ in range(3):
        st.run_device(fmt, X.data_ptr(), stride, G, addrs, contig=0, missing='mean')
    acc, reps, t0 = {}, 0, time.perf_counter()
    while reps < 3 or time.perf_counter() - t0 < min_seconds:
        st.run_device(fmt, X.data_ptr(), stride, G, addrs, contig=0, missing='mean')
        t = st.timing(0)
        for k, v in t.items():
            acc[k] = acc.get(k, 0.0) + v
        path = t['path']
        reps += 1
    for k in acc:
        acc[k] /= reps
    K = K_cov + 1
    U = st.unique_masks
    cols = (K - 1) + P
    line = dict(config=label, N=N, K=K, P=P, variants_per_block=G, format=fmt, unique_masks=U, missing_call_rate=call_missing if fmt == 'plink2' else 0.0,
                tests_per_s=P * G / (acc['total_ms'] / 1e3), block_ms=acc['total_ms'], blocks_timed=reps,
                stage_ms={k: round(acc[k], 4) for k in ('prepass_ms', 'gram_ms', 'masked_ms', 'epilogue_ms')}, path_flags=int(path),
                genotype_gbs=float(G) * stride / (acc['gram_ms'] / 1e3) / 1e9,
                finite=float(torch.isfinite(outs['pvalue']).double().mean().item()))
    gram_s = acc['gram_ms'] / 1e3
    if path & _capi.PATH_SPLIT8:
        operands = 1 if path & _capi.PATH_OPTIMISTIC else 2
        ops = 2.0 * N * digit_rows(K - 1, cols - (K - 1), path, U) * operands * G
        line['roofline'] = dict(kernel='split8_kernel' + (' (optimistic form: no missing-call operand)' if operands == 1 else ''),
                                bound='tensor', achieved=ops / gram_s / 1e12, peak=i8_peak, unit='TFLOP/s',
                                frac=ops / gram_s / 1e12 / i8_peak if i8_peak == i8_peak else None,
                                op_type='int8 tensor-core operations, algorithmic digit rows',
                                digit_rows=digit_rows(K - 1, cols - (K - 1), path, U), share_of_block=acc['gram_ms'] / acc['total_ms'],
                                fp64_equivalent_tflops=2.0 * N * cols * G / gram_s / 1e12)
        if U and acc['masked_ms'] > 0:
            line['masked_pass'] = dict(ms=acc['masked_ms'], form='int8' if path & _capi.PATH_MASK_INT8 else
                                       ('sparse' if path & _capi.PATH_MASK_SPARSE else 'dmma'))
    else:
        nbytes = float(G) * stride + 4 * 8 * P * G
        line['roofline'] = dict(kernel='gram_kernel (FP64 DMMA, dosage rows streamed once)', bound='hbm', achieved=nbytes / gram_s / 1e9,
                                peak=hbm_peak, unit='GB/s', frac=nbytes / gram_s / 1e9 / hbm_peak, share_of_block=acc['gram_ms'] / acc['total_ms'],
                                executed_tflops=2.0 * N * (K - 1 + P) * G / gram_s / 1e12)
    st.close()
    del X, outs
    torch.cuda.empty_cache()
    return line


def run_other_configs(torch, dev, hbm_peak, i8_peak):
    out = []
    cfgs = [
        ('cfg1', 'BASELINE.json configs[1]: dense fp64 dosages 10k samples x 100k variants x 10 phenotypes x 10 covariates',
         dict(N=10_000, K_cov=10, P=10, G=100_000, fmt='f64')),
        ('cfg3', 'BASELINE.json configs[3]: 50k samples x 2 000 phenotypes, 10 % missing each (2 000 distinct masks), 2-bit; one block of 8 192 variants',
         dict(N=50_000, K_cov=10, P=2_000, G=8_192, fmt='plink2', missing=0.1)),
        ('cfg4', 'BASELINE.json configs[4]: 200k samples x 20 phenotypes, LOCO offsets (2 of the 22 contig states built), 2-bit',
         dict(N=200_000, K_cov=16, P=20, G=37_888, fmt='plink2', n_contigs=2)),
        ('cfg2_fully_called', 'configs[2] shape with fully called genotypes (imputed hard calls): the optimistic form without the missing-call operand',
         dict(N=500_000, K_cov=16, P=1, G=37_888, fmt='plink2', call_missing=0.0)),
        ('cfg2_missing_phenotype', 'configs[2] shape with 2 % of the phenotype values unobserved (masked second pass)',
         dict(N=500_000, K_cov=16, P=1, G=37_888, fmt='plink2', missing=0.02)),
        ('cfg2_int8_states', 'configs[2] shape with genotype_states as int8 arrays (1 byte per genotype) on the int8 tensor path',
         dict(N=500_000, K_cov=16, P=1, G=37_888, fmt='i8')),
        ('cfg2_spark_batch', 'configs[2] shape in blocks of 10 000 variants (Spark default Arrow batch, lin_reg.py:277-282): sample range split across CTAs',
         dict(N=500_000, K_cov=16, P=1, G=10_000, fmt='plink2')),
    ]
    for name, label, kw in cfgs:
        try:
            out.append(run_config(torch, dev, name, label, hbm_peak=hbm_peak, i8_peak=i8_peak, **kw))
        except Exception as exc:  # noqa: BLE001
            out.append(dict(config=label, error=f'{type(exc).__name__}: {exc}'))
    return out


def run_next_rows(torch, dev):
    """SURVEY.md section 8f rows built beyond the linear block loop, each with a measured rate: N3 logistic score test
    (device block loop), N2 BGEN probabilities -> 2-bit rows, N4 GloWGR LOCO apply step."""
    import ctypes as C
    from glow_b200 import _capi
    from glow_b200.engine import ResidentBlock
    from glow_b200.gwas import log_reg as lr
    out = []
    lib = _capi.load()
    # ---- N3: logistic score test at the configs[2] shape, genotypes resident in HBM ----
    try:
        rg = np.random.default_rng(17)
        N, K, G = N_SAMPLES, N_COV + 1, 37_888
        Cm = np.column_stack([np.ones(N), rg.standard_normal((N, K - 1))])
        y = (rg.random(N) < 1.0 / (1.0 + np.exp(-(0.3 * Cm[:, 1] - 0.5)))).astype(np.float64)
        y_res, gamma, inv = lr._prepare_one_phenotype(Cm, y, None)
        state = lr.LogRegDeviceState(Cm, [lr.LogRegState(inv[None], gamma[:, None], y_res[:, None], None)], device=dev.index)
        X = gen_packed_device(torch, G, N, 9, dev)
        blk = ResidentBlock(X.data_ptr(), X.shape[1], G, 'plink2', 'mean', dev.index)
        for _ in range(2):
            res = state.run(blk)
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or time.perf_counter() - t0 < 0.5:
            res = state.run(blk)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        state.close()
        out.append(dict(config='N3 logistic_regression(correction="none") block loop, configs[2] shape (500k samples, 16 covariates + '
                               'intercept, 1 binary phenotype, 2-bit, 1 % missing calls), block resident in HBM',
                        tests_per_s=G / dt, block_ms=dt * 1e3, variants_per_block=G, blocks_timed=reps,
                        finite=float(np.isfinite(res['pvalue']).mean()),
                        kernels='split8_kernel twice (int8 tensor cores: [gamma o C | Y_res]^T x with fused counts, then gamma^T x^2 '
                                'with the squared-call table) + logreg_epilogue_kernel; wall clock of glr_logreg_run_block incl. '
                                'the copy of chisq / pvalue to the host',
                        fp64_equivalent_tflops=2.0 * N * (K + 2) * G / dt / 1e12,
                        genotype_gbs_two_reads=2.0 * G * X.shape[1] / dt / 1e9))
        del X
    except Exception as exc:  # noqa: BLE001
        out.append(dict(config='N3 logistic score test', error=f'{type(exc).__name__}: {exc}'))
    # ---- N2: BGEN layout-2 probabilities (8 bit) -> PLINK 2-bit rows on the device, host in / device out ----
    try:
        rg = np.random.default_rng(18)
        G, N = 2_048, 100_000
        probs = torch.empty((G, 2 * N), dtype=torch.uint8, pin_memory=True).numpy()   # pinned: pageable memory gave 1.5-10 GB/s
        probs[:] = rg.integers(0, 256, (G, 2 * N), dtype=np.uint8)
        ploidy = torch.empty((G, N), dtype=torch.uint8, pin_memory=True).numpy()
        ploidy[:] = 2
        dst = torch.empty((G, (N + 3) // 4), dtype=torch.uint8, device=dev)
        call = lambda: _capi.check(lib.glr_bgen_to_plink2(dev.index, probs.ctypes.data, ploidy.ctypes.data, G, N, 8, 0.9,
                                                          C.c_void_p(dst.data_ptr()), dst.shape[1]))  # noqa: E731
        call()
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or time.perf_counter() - t0 < 0.3:
            call()
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        out.append(dict(config='N2 glr_bgen_to_plink2: 8-bit BGEN probability pairs of 2 048 variants x 100 000 samples (pinned host '
                               'memory) -> hard calls -> 2-bit rows in HBM',
                        input_gbs=(probs.nbytes + ploidy.nbytes) / dt / 1e9, genotypes_per_s=G * N / dt, block_ms=dt * 1e3,
                        note='bound by the host-to-device copy of 3 bytes per genotype'))
        del probs, ploidy, dst
    except Exception as exc:  # noqa: BLE001
        out.append(dict(config='N2 BGEN conversion', error=f'{type(exc).__name__}: {exc}'))
    # ---- N4: GloWGR LOCO apply step ----
    try:
        rg = np.random.default_rng(19)
        N, H, NB, NC = 50_000, 1_000, 10, 22
        Xr = torch.empty((N, H), dtype=torch.float64, pin_memory=True).numpy()
        Xr[:] = rg.standard_normal((N, H))
        mu, sig = Xr.mean(0), Xr.std(0)
        B = rg.standard_normal((NB, H)) / H
        hc = rg.integers(0, NC, H).astype(np.int32)
        sb = (np.arange(N) * NB // N).astype(np.int32)
        res = np.empty((NC, N))
        call = lambda: _capi.check(lib.glr_loco_predict(dev.index, N, H, NB, NC, Xr.ctypes.data, mu.ctypes.data, sig.ctypes.data,
                                                        B.ctypes.data, hc.ctypes.data, sb.ctypes.data, None, 0, None,
                                                        res.ctypes.data))  # noqa: E731
        call()
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or time.perf_counter() - t0 < 0.3:
            call()
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        ref = ((Xr[:64] - mu) / sig * B[sb[:64]])
        want = np.stack([ref[:, hc != c].sum(1) for c in range(NC)])
        out.append(dict(config='N4 glr_loco_predict: 50 000 samples x 1 000 reduced headers, 10 sample blocks, 22 chromosomes, pinned '
                               'host in / host out',
                        input_gbs=Xr.nbytes / dt / 1e9, offsets_per_s=NC * N / dt, call_ms=dt * 1e3,
                        max_abs_err_vs_numpy=float(np.abs(res[:, :64] - want).max())))
    except Exception as exc:  # noqa: BLE001
        out.append(dict(config='N4 LOCO apply', error=f'{type(exc).__name__}: {exc}'))
    return out


def run_e2e_api(torch, dev, n_variants=65_536, block_variants=8_192):
    """Public API on a .bed FILE in tmpfs: everything a user waits for (file -> pinned staging -> H2D -> kernels -> D2H ->
    pandas frame), state creation included."""
    import pandas as pd
    import glow_b200
    N = N_SAMPLES
    packed = gen_packed_device(torch, n_variants, N, 5, dev).cpu().numpy()
    path = '/dev/shm/glow_b200_stream.bed' if os.path.isdir('/dev/shm') else '/tmp/glow_b200_stream.bed'
    with open(path, 'wb') as f:
        f.write(bytes([0x6c, 0x1b, 0x01]))
        f.write(packed.tobytes())
    del packed
    rg = np.random.default_rng(3)
    cov = pd.DataFrame(rg.standard_normal((N, N_COV)))
    phe = pd.DataFrame(rg.standard_normal((N, 1)), columns=['trait'])
    meta = pd.DataFrame({'idx': np.arange(n_variants)})
    times = []
    try:
        src = glow_b200.PlinkBedSource(path, N, meta, block_variants=block_variants)
        for _ in range(3):
            t0 = time.perf_counter()
            out = glow_b200.gwas.linear_regression(src, phe, cov, device=dev.index)
            times.append(time.perf_counter() - t0)
    finally:
        os.unlink(path)
    best = min(times[1:])
    return {'value': n_variants / best, 'unit': 'tests/s', 'seconds': times, 'n_variants': n_variants,
            'file_gb_per_s': n_variants * ((N + 3) // 4) / best / 1e9, 'rows': int(len(out)),
            'call': 'glow_b200.gwas.linear_regression(PlinkBedSource(<.bed in tmpfs>), phenotype_df, covariate_df): includes the '
                    'host-side state preparation (QR of 500k x 17), state packing, the file read and the pandas result frame'}


# --------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference_impl(args)
        return

    import torch
    import torch.distributed as dist
    from glow_b200 import _capi
    from glow_b200.engine import DeviceState, GenotypeBlock, ResidentBlock

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    lib = _capi.load()
    if lib.glr_device_count() <= 0:
        raise RuntimeError('no sm_100 device: glow_b200 has no CPU fallback')

    n, K, P = args.samples, N_COV + 1, N_PHENO
    G = args.variants_per_step
    B = max(1, args.batches_per_step)

    # ---- state: rank 0 prepares it on the host, ONE NCCL broadcast replicates it ----
    from glow_b200 import dist as gdist
    host = None
    st_in = None
    if rank == 0:
        host = host_state(n)
        Q, Y, mask, ydy, scale, dof = host
        st_in = dict(Q=Q, Y=Y, Y_mask=mask, YdotY=ydy, Y_scale=scale, dof=dof)
    flat, header = gdist.broadcast_state(st_in, src=0, device=dev)
    state = gdist.device_state_from_flat(flat, header, local_rank, n_lanes=2)

    # ---- synthetic batch, resident in HBM ----
    packed = gen_packed_device(torch, G, n, SEED + 17 * rank, dev)
    row_bytes = packed.shape[1]
    outs = {k: torch.empty((P, G), dtype=torch.float64, device=dev) for k in ('effect', 'stderror', 'tvalue', 'pvalue')}
    out_addrs = {k: v.data_ptr() for k, v in outs.items()}
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def measure(st, sample_clocks, batches):
        """W warm-up steps, then K timed steps of `batches` batches each on the lane stream of `st` (CUDA events around the
        whole timed region, max over ranks)."""
        lane_stream = torch.cuda.ExternalStream(st.lane_stream(0), device=dev)

        def step_device():
            for _ in range(batches):
                st.run_device('plink2', packed.data_ptr(), row_bytes, G, out_addrs, missing='mean', lane=0, wait=True)

        for _ in range(args.warmup):
            step_device()
        barrier()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stage = {'prepass_ms': 0.0, 'gram_ms': 0.0, 'masked_ms': 0.0, 'epilogue_ms': 0.0, 'total_ms': 0.0}
        launches = 0
        ev0.record(lane_stream)
        for _ in range(args.steps):
            for _ in range(batches):
                st.run_device('plink2', packed.data_ptr(), row_bytes, G, out_addrs, missing='mean', lane=0, wait=True)
                t = st.timing(0)
                for k in stage:
                    stage[k] += t[k]
                launches += t['kernel_launches']
        ev1.record(lane_stream)
        barrier()
        clocks = sampler.stop() if sampler else None
        ms_step = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
        for k in stage:
            stage[k] /= args.steps * batches   # per batch
        return ms_step, stage, launches, clocks, st.timing(0)['path']

    # primary arm: the library's default path (2-bit genotypes -> exact int8 digit-split GEMM on tcgen05)
    ms_step, stage, launches, clocks, path = measure(state, True, B)
    value = world * G * B * P / (ms_step / 1e3)
    t_primary = outs['tvalue'].clone()

    # secondary arm in the same run: the FP64 tensor-core (DMMA) path forced for the same batch (2 batches per step)
    ms_dmma, stage_dmma, paths_agree = None, None, None
    if not args.skip_dmma_arm:
        from glow_b200.dist import state_layout
        state_dmma = gdist.device_state_from_flat(flat, header, local_rank, n_lanes=1, options={'split8': 'off'})
        ms_dmma, stage_dmma, _, _, _ = measure(state_dmma, False, 1)
        paths_agree = bool(torch.allclose(outs['tvalue'], t_primary, rtol=1e-9, atol=1e-12, equal_nan=True))
        state_dmma.close()
        state.run_device('plink2', packed.data_ptr(), row_bytes, G, out_addrs, missing='mean', lane=0, wait=True)

    # sanity: results are real numbers for (nearly) all variants
    pv = outs['pvalue']
    frac_finite = float(torch.isfinite(pv).double().mean().item())

    # ---- e2e: pinned host buffers through the C-ABI block call, lanes pipelined ----
    e2e, e2e_res = None, None
    names = ('effect', 'stderror', 'tvalue', 'pvalue')
    if not args.skip_e2e:
        Be = max(1, args.e2e_batches_per_step)
        nsub = max(1, args.e2e_subblocks)
        gs = G // nsub
        host_packed = torch.empty((G, row_bytes), dtype=torch.uint8, pin_memory=True)
        host_packed.copy_(packed)
        hp = host_packed.numpy()
        host_out = [{k: torch.empty((P, gs), dtype=torch.float64, pin_memory=True).numpy() for k in names}
                    for _ in range(nsub)]
        torch.cuda.synchronize()

        def step_e2e():
            inflight = []
            for _ in range(Be):
                for i in range(nsub):
                    lane = i % state.n_lanes
                    if len(inflight) == state.n_lanes:
                        state.wait(inflight.pop(0))
                    state.submit(lane, GenotypeBlock(hp[i * gs:(i + 1) * gs], 'plink2', 'mean', pinned=True), 0, out=host_out[i])
                    inflight.append(lane)
            while inflight:
                state.wait(inflight.pop(0))

        for _ in range(max(1, args.warmup - 1)):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e_val = world * gs * nsub * Be * P * args.steps / dt
        # the host path (smaller sub-blocks) must agree with the resident path
        same = bool(np.allclose(host_out[0]['tvalue'], outs['tvalue'][:, :gs].cpu().numpy(), rtol=1e-9, atol=1e-12,
                                equal_nan=True))
        # what the platform gives: the same pinned sub-blocks copied with bare cudaMemcpyAsync on two streams, all ranks at once
        barrier()
        scratch = torch.empty((2, gs, row_bytes), dtype=torch.uint8, device=dev)
        streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
        reps = max(2, min(args.steps, 6)) * Be
        t0 = time.perf_counter()
        for _ in range(reps):
            for i in range(nsub):
                with torch.cuda.stream(streams[i % 2]):
                    scratch[i % 2].copy_(host_packed[i * gs:(i + 1) * gs], non_blocking=True)
        torch.cuda.synchronize()
        dt_c = max_over_ranks(time.perf_counter() - t0)
        ceiling_gbs = world * reps * gs * nsub * row_bytes / dt_c / 1e9
        del scratch
        h2d = int(gs * nsub * row_bytes) * Be
        e2e = {
            'value': e2e_val,
            'unit': 'tests/s',
            'h2d_bytes_per_step': h2d,
            'd2h_bytes_per_step': int(4 * 8 * P * gs * nsub) * Be,
            'batches_per_step': Be,
            'seconds_timed': dt,
            'subblocks': nsub,
            'lanes': state.n_lanes,
            'timer': 'host wall clock around submit..wait of every sub-block (copies inside), max over ranks',
            'matches_device_path': same,
            'h2d_gbs': world * h2d * args.steps / dt / 1e9,
            'h2d_ceiling': {'gbs': ceiling_gbs, 'frac': (world * h2d * args.steps / dt / 1e9) / ceiling_gbs,
                            'how': 'bare cudaMemcpyAsync of the same pinned sub-blocks on two streams, every rank at once, no kernels'},
            'note': 'bound by the host-to-device copy of 125 kB per test: compare h2d_gbs with h2d_ceiling.gbs (whole job)'
        }
        # resident genotypes: the batch stays in HBM (glow_b200.ResidentSource), only the results cross PCIe
        rblk = ResidentBlock(packed.data_ptr(), row_bytes, G, 'plink2', 'mean', local_rank)
        pin_out = {k: torch.empty((P, G), dtype=torch.float64, pin_memory=True).numpy() for k in names}
        for _ in range(2):
            state.submit(0, rblk, 0, out=pin_out)
            state.wait(0)
        barrier()
        t0 = time.perf_counter()
        nres = args.steps * Be
        for _ in range(nres):
            state.submit(0, rblk, 0, out=pin_out)
            state.wait(0)
        dt_r = max_over_ranks(time.perf_counter() - t0)
        e2e_res = {'value': world * G * P * nres / dt_r, 'unit': 'tests/s', 'h2d_bytes_per_step': 0,
                   'd2h_bytes_per_step': int(4 * 8 * P * G) * Be, 'seconds_timed': dt_r,
                   'note': 'genotype shard loaded into HBM once (glow_b200.ResidentSource), every later phenotype set runs against '
                           'it: C-ABI call with a device-resident block, results copied to pinned host memory inside the timed region'}
        del host_packed

    # ---- rooflines ----
    roof = None
    dmma = None
    cpu = None
    configs = None
    next_rows = None
    e2e_api = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        hbm_peak = peaks.get('hbm_gbs', 6650.0)
        hbm_src = 'MEASURED_PEAKS.json hbm_gbs' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s (B200_PROFILING.md)'
        traffic = load_traffic()

        def best_of(fn, n=7, skip=2):
            best = 1e9
            for i in range(n):
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record()
                fn()
                s1.record()
                torch.cuda.synchronize()
                if i >= skip:
                    best = min(best, s0.elapsed_time(s1))
            return best

        fp64_peak = float('nan')
        int8_peak, int8_src = float('nan'), 'not measured'
        if not args.skip_peak:
            # FP64 tensor peak: cuBLAS DGEMM 8192^3 (MEASURED_PEAKS.json has no fp64 entry)
            a = torch.randn((8192, 8192), dtype=torch.float64, device=dev)
            b = torch.randn((8192, 8192), dtype=torch.float64, device=dev)
            fp64_peak = 2 * 8192**3 / (best_of(lambda: torch.matmul(a, b)) / 1e3) / 1e12
            del a, b
            try:
                a8 = torch.randint(-127, 127, (8192, 8192), dtype=torch.int8, device=dev)
                b8 = torch.randint(-127, 127, (8192, 8192), dtype=torch.int8, device=dev)
                int8_peak = 2 * 8192**3 / (best_of(lambda: torch._int_mm(a8, b8)) / 1e3) / 1e12
                int8_src = 'cuBLASLt int8 GEMM 8192^3 via torch._int_mm, best of 5, measured in this run'
                del a8, b8
            except Exception as exc:  # noqa: BLE001
                int8_src = f'torch._int_mm unavailable: {type(exc).__name__}'
        # what the int8 tensor pipe does on this box right now: back-to-back tcgen05.mma with resident operands on every
        # SM, ~6 ms (long enough for the clocks to settle), through the library's own probe
        issue_rate, issue_src = 4340.0, 'benchmarks/micro/umma_i8_ts.cu (0.6 ms burst, measured earlier on this pool)'
        if not args.skip_peak:
            try:
                import ctypes as C
                r = C.c_double()
                _capi.check(_capi.load().glr_tensor_i8_issue_rate(local_rank, 6.0, C.byref(r)))
                issue_rate, issue_src = float(r.value), 'glr_tensor_i8_issue_rate: 6 ms of back-to-back tcgen05.mma kind::i8, measured in this run'
            except Exception as exc:  # noqa: BLE001
                issue_src += f' (live probe failed: {type(exc).__name__})'
        bf16 = peaks.get('bf16_tflops')
        # Executed formulation: S = [Q(:,1:) | Ytilde]^T X.  The intercept column of Q is constant, so its product
        # with x is q0 * sum(x) and comes from the genotype counts: its flops are NOT executed and not counted.
        cols = (K - 1) + P
        flops_f64_equiv = 2.0 * n * cols * G                      # FP64 flops of the contraction as formulated
        flops_ref = float(n) * (4 * K + 4 * P + 2) * G            # reference formulation, SURVEY.md section 8(d)
        gram_s = stage['gram_ms'] / 1e3
        # int8 digit-split kernel: 7 digit rows per phenotype column, 5-7 per covariate column (what the block actually
        # ran with), two int8 operands (genotype, missing indicator)
        rows = digit_rows(K - 1, P, path, state.unique_masks)
        ops_int8 = 2.0 * n * rows * 2 * G
        n_cb = -(-cols // 18)                                     # column blocks of up to 18 columns
        nc_rows = -(-((rows if n_cb == 1 else 7 * -(-cols // n_cb)) + 1) // 16) * 16  # rows per block incl. the ones = MMA N
        ops_int8_issued = 2.0 * (-(-n // 128) * 128) * nc_rows * n_cb * 2 * (-(-G // 128) * 128)
        achieved = ops_int8 / gram_s / 1e12
        default_shape = n == N_SAMPLES and G == 148 * 256
        roof = {
            'kernel': 'split8_kernel (tcgen05.mma.cta_group::2 kind::i8 on CTA pairs, A = decoded genotypes of 2 x 128 '
                      'variants in TMEM, B = digit rows split over the pair\'s shared memory, M256 N128 K32, accumulators in '
                      'TMEM; radix-256 digit split of the FP64 operand (7 digits per phenotype column, 5-7 per covariate '
                      'column with a certified error bound per test), '
                      '2-bit decode + genotype counts fused, FP64 recombination in the epilogue)',
            'bound': 'tensor',
            'achieved': achieved,
            'peak': issue_rate,
            'unit': 'TFLOP/s',
            'op_type': 'int8 tensor-core operations (2 per multiply-accumulate), algorithmic = digit rows actually needed; half of '
                       'them multiply the missing-call indicator (1 % dense): see configs.cfg2_fully_called for the form without it',
            'frac': achieved / issue_rate,
            'traffic': traffic.get('split8_kernel', {}).get('dram_bytes_per_launch') if default_shape else None,
            'traffic_source': traffic.get('split8_kernel', {}).get('source', 'no capture committed') +
                              ' (ncu --set full capture of one launch at this shape: a constant of the build, not measured in this run)',
            'peak_source': issue_src,
            'peak_alternatives': {
                'library_int8_gemm_tops': int8_peak,
                'library_int8_gemm_source': int8_src,
                'frac_of_library_int8_gemm': achieved / int8_peak,
                'twice_MEASURED_PEAKS_bf16_burst_tops': 2 * bf16 if bf16 else None,
                'frac_of_twice_bf16_burst': achieved / (2 * bf16) if bf16 else None,
            },
            'note': 'peak = the int8 tensor pipe fed back to back with constant resident operands on every SM (it holds '
                    '~1.9 GHz); MEASURED_PEAKS.json has no int8 entry -- 2 x its bf16 burst figure is given beside it. Under a real '
                    'load the board sits at its power cap and the SM clock drops (see clocks)',
            'ops_per_launch_algorithmic': ops_int8,
            'digit_rows': rows,
            'mma_n': nc_rows,
            'issued_tops_incl_padding': ops_int8_issued / gram_s / 1e12,
            'fp64_equivalent_tflops': flops_f64_equiv / gram_s / 1e12,
            'reference_equivalent_tflops': flops_ref / gram_s / 1e12,
            'algorithmic_bytes_per_launch': float(G) * row_bytes + float(n) * rows,
            'avg_launch_ms': stage['gram_ms'],
            'launches_timed': args.steps * B,
            'share_of_step': stage['gram_ms'] / max(stage['total_ms'], 1e-9),
            'stages': {
                'genotype_read_inside_kernel': {
                    'bound': 'hbm',
                    'achieved_gbs': G * row_bytes / max(stage['gram_ms'], 1e-9) / 1e6,
                    'peak_gbs': hbm_peak,
                    'peak_source': hbm_src,
                    'note': 'decode, counts and mean imputation are fused into the contraction kernel: no pre-pass'
                },
                'epilogue': {
                    'bound': 'hbm',
                    'achieved_gbs': (4 * 8 * P * G + 8 * (K + P + 1) * G) / max(stage['epilogue_ms'], 1e-9) / 1e6,
                    'peak_gbs': hbm_peak,
                    'ms': stage['epilogue_ms']
                }
            }
        }
        dmma_cols = (cols // 8) * 8 if 1 <= cols % 8 <= 3 and cols // 8 >= 1 else ((cols + 7) // 8) * 8
        d_s = stage_dmma['gram_ms'] / 1e3 if stage_dmma else float('nan')
        dmma = None if stage_dmma is None else {
            'note': 'same batch through the FP64 tensor-core path (options split8=off): what float / dosage input uses',
            'value': world * G * P / (ms_dmma / 1e3),
            'unit': 'tests/s',
            'ms_per_batch': ms_dmma,
            'stage_ms': stage_dmma,
            'agrees_with_primary_path_rel_1e-9': paths_agree,
            'roofline': {
                'kernel': 'gram_kernel<PLINK2, MT=2, NT=2, TAIL=1, WARPS=16> (mma.sync.m8n8k4.f64 DMMA for the 16 covariate '
                          'columns, FP64 FMA for the phenotype column, fused 2-bit decode)',
                'bound': 'tensor',
                'achieved': flops_f64_equiv / d_s / 1e12,
                'peak': fp64_peak,
                'unit': 'TFLOP/s',
                'frac': flops_f64_equiv / d_s / 1e12 / fp64_peak,
                'traffic': traffic.get('gram_kernel_plink2', {}).get('dram_bytes_per_launch') if default_shape else None,
                'traffic_source': traffic.get('gram_kernel_plink2', {}).get('source', 'no capture committed'),
                'algorithmic_bytes_per_launch': float(G) * row_bytes + 8.0 * n * cols,
                'peak_source': 'cuBLAS DGEMM 8192^3 best of 5, measured in this run (MEASURED_PEAKS.json has no fp64 entry)',
                'issued_tflops_incl_padding': 2.0 * n * max(dmma_cols, cols) * G / d_s / 1e12,
                'reference_equivalent_tflops': flops_ref / d_s / 1e12,
                'avg_launch_ms': stage_dmma['gram_ms'],
                'prepass_gbs': G * row_bytes / max(stage_dmma['prepass_ms'], 1e-9) / 1e6,
                'hbm_peak_gbs': hbm_peak
            }
        }
        if not args.skip_cpu_baseline:
            bv = args.cpu_block_variants
            nb = 3
            sample = packed[:bv * (nb + 1)].cpu().numpy()
            if host is None:
                host = host_state(n)
            Q, Y, mask, ydy, scale, dof_h = host
            v, times = cpu_reference_arm(sample, n, Q, Y, mask, ydy, scale, dof_h, bv, nb)
            cpu = {
                'value': v,
                'unit': 'tests/s',
                'cores': os.cpu_count(),
                'kind': 'port',
                'sample': f'{nb} blocks of {bv} variants x {n} samples (1 warm-up block discarded) through the oracle '
                          'port of _linear_regression_inner incl. np.column_stack; numpy/OpenBLAS on all host cores; '
                          'decode + mean substitution not timed',
                'block_seconds': times
            }
    # the other configs and the public-API run use the whole GPU: only at N = 1, after the main measurement
    if world == 1:
        del packed, outs
        state.close()
        del flat
        torch.cuda.empty_cache()
        if not args.skip_configs:
            configs = run_other_configs(torch, dev, hbm_peak, issue_rate)
            next_rows = run_next_rows(torch, dev)
        if not args.skip_e2e_api:
            try:
                e2e_api = run_e2e_api(torch, dev)
            except Exception as exc:  # noqa: BLE001
                e2e_api = {'error': f'{type(exc).__name__}: {exc}'}

    if rank == 0:
        line = {
            'metric': 'variant_x_phenotype_tests_per_sec',
            'value': value,
            'unit': 'tests/s',
            'n_gpus': world,
            'steps': args.steps,
            'warmup': args.warmup,
            'ms_per_step': ms_step,
            'ms_per_batch': ms_step / B,
            'seconds_timed': ms_step * args.steps / 1e3,
            'higher_is_better': True,
            'scaling': 'weak',
            'vs_baseline': None,
            'dtype': 'f64',
            'dtype_detail': 'results and the DMMA path are FP64; the default first pass computes the same FP64 contraction '
                            'exactly as int8 x int8 -> int32 tensor-core products of a 55-bit fixed-point operand '
                            '(7 radix-256 digits), recombined in FP64',
            'data': 'synthetic',
            'config': workload_config(args, G),
            'clocks': clocks,
            'e2e': e2e,
            'e2e_resident': e2e_res,
            'e2e_api': e2e_api,
            'gpu_launches': launches,
            'path_flags': int(path),
            'roofline': roof,
            'dmma_f64_path': dmma,
            'configs': configs,
            'next_rows': next_rows,
            'cpu_baseline': cpu,
            'stage_ms_per_batch': stage,
            'finite_fraction': frac_finite
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
