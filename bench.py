#!/usr/bin/env python
"""Benchmark of the linear-regression GWAS block loop on B200 (BASELINE.json metric:
variant x phenotype tests/sec, device-timed).

Workload (config.workload): BASELINE.json configs[2], the configuration the metric is quoted on --
UK-Biobank shape, N = 500 000 samples, 16 covariates + intercept (K = 17), P = 1 phenotype,
PLINK 2-bit packed genotypes (125 000 B per variant, 1 % missing calls, mean-imputed inside the
decode).  One step = one batch of G_step variants per GPU (4.7 GB packed at the default 37 888);
variants are independent, so N GPUs each take their own batch (weak scaling, no data-path
collective; the state is replicated with one NCCL broadcast before the timed region).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

`value`  : tests/s with the packed batch resident in HBM, timed with CUDA events on the stream
           the kernels are launched on, max over ranks.
`e2e`    : tests/s through the C-ABI block call with PINNED HOST buffers -- H2D copy of the packed
           batch and D2H copy of the four result arrays inside the timed region, sub-blocks
           pipelined over the library's lanes.
`roofline`: the DMMA contraction kernel against the FP64 tensor peak measured in this run with a
           cuBLAS DGEMM (MEASURED_PEAKS.json has no fp64 entry; HBM peak is taken from it).
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
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
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
        'config': workload_config(args, bv),
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


def workload_config(args, variants_per_step):
    return {
        'workload': 'BASELINE.json configs[2]: UK-Biobank shape, PLINK 2-bit genotypes, variant-sharded',
        'n_samples': args.samples,
        'n_covariates': N_COV,
        'n_phenotypes': N_PHENO,
        'variants_per_step_per_gpu': variants_per_step,
        'missing_call_rate': MISSING_RATE,
        'imputation': 'mean_substitute fused into the decode',
        'cache': 'inputs larger than L2 (packed batch >> 126 MB); no L2 flush',
        'parallelism': f'variant-sharded x{args.gpus}, state replicated by one NCCL broadcast'
    }


# --------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference_impl(args)
        return

    import torch
    import torch.distributed as dist
    from glow_b200 import _capi
    from glow_b200.engine import DeviceState, GenotypeBlock

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

    def measure(st, sample_clocks):
        """W warm-up steps, then K timed steps on the lane stream of `st` (CUDA events, max over ranks)."""
        lane_stream = torch.cuda.ExternalStream(st.lane_stream(0), device=dev)

        def step_device():
            st.run_device('plink2', packed.data_ptr(), row_bytes, G, out_addrs, missing='mean', lane=0, wait=False)

        for _ in range(args.warmup):
            step_device()
            st.wait_device(0)
        barrier()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stage = {'prepass_ms': 0.0, 'gram_ms': 0.0, 'masked_ms': 0.0, 'epilogue_ms': 0.0, 'total_ms': 0.0}
        launches = 0
        ev0.record(lane_stream)
        for _ in range(args.steps):
            step_device()
            st.wait_device(0)
            t = st.timing(0)
            for k in stage:
                stage[k] += t[k]
            launches += t['kernel_launches']
        ev1.record(lane_stream)
        barrier()
        clocks = sampler.stop() if sampler else None
        ms_total = ev0.elapsed_time(ev1)
        t_max = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
        ms_step = float(t_max.item()) / args.steps
        for k in stage:
            stage[k] /= args.steps
        return ms_step, stage, launches, clocks

    # primary arm: the library's default path (2-bit genotypes -> exact int8 digit-split GEMM on tcgen05)
    ms_step, stage, launches, clocks = measure(state, True)
    value = world * G * P / (ms_step / 1e3)
    t_primary = outs['tvalue'].clone()

    # secondary arm in the same run: the FP64 tensor-core (DMMA) path forced for the same batch
    ms_dmma, stage_dmma, paths_agree = None, None, None
    if not args.skip_dmma_arm:
        os.environ['GLR_SPLIT8'] = '0'
        state_dmma = gdist.device_state_from_flat(flat, header, local_rank, n_lanes=1)
        os.environ.pop('GLR_SPLIT8', None)
        ms_dmma, stage_dmma, _, _ = measure(state_dmma, False)
        paths_agree = bool(torch.allclose(outs['tvalue'], t_primary, rtol=1e-9, atol=1e-12, equal_nan=True))
        state_dmma.close()
        state.run_device('plink2', packed.data_ptr(), row_bytes, G, out_addrs, missing='mean', lane=0, wait=True)

    # sanity: results are real numbers for (nearly) all variants
    pv = outs['pvalue']
    frac_finite = float(torch.isfinite(pv).double().mean().item())

    # ---- e2e: pinned host buffers through the C-ABI block call, lanes pipelined ----
    e2e = None
    if not args.skip_e2e:
        nsub = max(1, args.e2e_subblocks)
        gs = G // nsub
        host_packed = torch.empty((G, row_bytes), dtype=torch.uint8, pin_memory=True)
        host_packed.copy_(packed)
        hp = host_packed.numpy()
        names = ('effect', 'stderror', 'tvalue', 'pvalue')
        host_out = [{k: torch.empty((P, gs), dtype=torch.float64, pin_memory=True).numpy() for k in names}
                    for _ in range(nsub)]
        torch.cuda.synchronize()

        def step_e2e():
            inflight = []
            for i in range(nsub):
                lane = i % state.n_lanes
                if len(inflight) == state.n_lanes:
                    state.wait(inflight.pop(0))
                state.submit(lane, GenotypeBlock(hp[i * gs:(i + 1) * gs], 'plink2', 'mean'), 0, out=host_out[i])
                inflight.append(lane)
            while inflight:
                state.wait(inflight.pop(0))

        for _ in range(max(1, args.warmup - 1)):
            step_e2e()
        barrier()
        e2e_launches = 0
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        e2e_val = world * gs * nsub * P * args.steps / float(t_e.item())
        # the device-resident and host paths must agree bit for bit
        # the host path (smaller sub-blocks, split reduction order) must agree with the resident path
        same = bool(np.allclose(host_out[0]['tvalue'], outs['tvalue'][:, :gs].cpu().numpy(), rtol=1e-9, atol=1e-12,
                                equal_nan=True))
        e2e = {
            'value': e2e_val,
            'unit': 'tests/s',
            'h2d_bytes_per_step': int(gs * nsub * row_bytes),
            'd2h_bytes_per_step': int(4 * 8 * P * gs * nsub),
            'subblocks': nsub,
            'lanes': state.n_lanes,
            'timer': 'host wall clock around submit..wait of every sub-block (copies inside), max over ranks',
            'matches_device_path': same
        }
        del host_packed

    # ---- rooflines ----
    roof = None
    dmma = None
    cpu = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        hbm_peak = peaks.get('hbm_gbs', 6650.0)
        hbm_src = 'MEASURED_PEAKS.json hbm_gbs' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s (B200_PROFILING.md)'

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
            # INT8 tensor peak: cuBLASLt int8 GEMM 8192^3 through torch._int_mm; if that is not available the
            # rate of back-to-back tcgen05.mma kind::i8 (benchmarks/micro/umma_i8_rate.cu, 4.28 Pop/s) is used
            try:
                a8 = torch.randint(-127, 127, (8192, 8192), dtype=torch.int8, device=dev)
                b8 = torch.randint(-127, 127, (8192, 8192), dtype=torch.int8, device=dev)
                int8_peak = 2 * 8192**3 / (best_of(lambda: torch._int_mm(a8, b8)) / 1e3) / 1e12
                int8_src = 'cuBLASLt int8 GEMM 8192^3 via torch._int_mm, best of 5, measured in this run'
                del a8, b8
            except Exception as exc:  # noqa: BLE001
                int8_peak = 4280.0
                int8_src = f'tcgen05.mma kind::i8 issue-rate microbenchmark (torch._int_mm unavailable: {type(exc).__name__})'
        # what the int8 tensor pipe does on this box right now: back-to-back tcgen05.mma with resident operands on every
        # SM, ~6 ms (long enough for the clocks to settle), through the library's own probe
        issue_rate, issue_src = 4340.0, 'benchmarks/micro/umma_i8_ts.cu (0.6 ms burst, measured earlier on this pool)'
        if not args.skip_peak:
            try:
                import ctypes as C
                from glow_b200 import _capi
                r = C.c_double()
                _capi.check(_capi.load().glr_tensor_i8_issue_rate(local_rank, 6.0, C.byref(r)))
                issue_rate, issue_src = float(r.value), 'glr_tensor_i8_issue_rate: 6 ms of back-to-back tcgen05.mma kind::i8, measured in this run'
            except Exception as exc:  # noqa: BLE001
                issue_src += f' (live probe failed: {type(exc).__name__})'
        # Executed formulation: S = [Q(:,1:) | Ytilde]^T X.  The intercept column of Q is constant, so its product
        # with x is q0 * sum(x) and comes from the genotype counts: its flops are NOT executed and not counted.
        cols = (K - 1) + P
        flops_f64_equiv = 2.0 * n * cols * G                      # FP64 flops of the contraction as formulated
        flops_ref = float(n) * (4 * K + 4 * P + 2) * G            # reference formulation, SURVEY.md section 8(d)
        gram_s = stage['gram_ms'] / 1e3
        # int8 digit-split kernel: 7 digit rows per column, two int8 operands (genotype, missing indicator)
        ops_int8 = 2.0 * n * (7 * cols) * 2 * G
        n_cb = -(-cols // 18)                                     # column blocks of up to 18 columns
        nc_rows = -(-(7 * -(-cols // n_cb) + 1) // 16) * 16       # digit rows per block incl. the row of ones = MMA N
        ops_int8_issued = 2.0 * (-(-n // 128) * 128) * nc_rows * n_cb * 2 * (-(-G // 128) * 128)
        achieved = ops_int8 / gram_s / 1e12
        roof = {
            'kernel': 'split8_kernel (tcgen05.mma.cta_group::2 kind::i8 on CTA pairs, A = decoded genotypes of 2 x 128 '
                      'variants in TMEM, B = digit rows split over the pair\'s shared memory, M256 N128 K32, accumulators in '
                      'TMEM; exact 7-digit radix-256 split of the FP64 operand, '
                      '2-bit decode + genotype counts fused, FP64 recombination in the epilogue)',
            'bound': 'tensor',
            'achieved': achieved,
            'peak': issue_rate,
            'unit': 'TFLOP/s',
            'op_type': 'int8 tensor-core operations (2 per multiply-accumulate), algorithmic = digit rows actually needed',
            'frac': achieved / issue_rate,
            # dram bytes of one launch at the default shape, ncu --set full, profiles/r01_split8_full.txt
            'traffic': 4.876e9 if (n == N_SAMPLES and G == 148 * 256) else None,
            'peak_source': issue_src,
            # back-to-back tcgen05.mma kind::i8 from one thread per SM, operands resident: 64 clk per M128 N128 K32; a
            # library GEMM and this kernel both run below it
            'library_int8_gemm_tops': int8_peak,
            'library_int8_gemm_source': int8_src,
            'frac_of_library_int8_gemm': achieved / int8_peak,
            'note': 'peak = the int8 tensor pipe fed back to back with constant resident operands on every SM (it holds '
                    '~1.9 GHz); a library int8 GEMM on random data (as MEASURED_PEAKS.json measures bf16: 1625 TFLOP/s burst, '
                    'i.e. 3250 TOP/s for int8 on the same pipe) runs BELOW this kernel: under a real load the board sits at '
                    'its power cap and the SM clock drops (ncu: 1.59 GHz with the tensor pipe 97.5 % active for this kernel)',
            'ops_per_launch_algorithmic': ops_int8,
            'issued_tops_incl_padding': ops_int8_issued / gram_s / 1e12,
            'fp64_equivalent_tflops': flops_f64_equiv / gram_s / 1e12,
            'reference_equivalent_tflops': flops_ref / gram_s / 1e12,
            'algorithmic_bytes_per_launch': float(G) * row_bytes + 7.0 * n * cols,
            'avg_launch_ms': stage['gram_ms'],
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
            'note': 'same batch through the FP64 tensor-core path (GLR_SPLIT8=0): what float / dosage input, verbose '
                    'states and small blocks use',
            'value': world * G * P / (ms_dmma / 1e3),
            'unit': 'tests/s',
            'ms_per_step': ms_dmma,
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
                # dram bytes of one launch at the default shape, ncu --set full, profiles/r01_gram_full.txt
                'traffic': 4.818e9 if (n == N_SAMPLES and G == 148 * 256) else None,
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

    if rank == 0:
        line = {
            'metric': 'variant_x_phenotype_tests_per_sec',
            'value': value,
            'unit': 'tests/s',
            'n_gpus': world,
            'steps': args.steps,
            'warmup': args.warmup,
            'ms_per_step': ms_step,
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
            'gpu_launches': launches,
            'roofline': roof,
            'dmma_f64_path': dmma,
            'cpu_baseline': cpu,
            'stage_ms': stage,
            'finite_fraction': frac_finite
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
