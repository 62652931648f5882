// Pass 2 on the integer tensor cores: D[u][v] = sum_s mask_u[s] * r_v[s]^2,  r_v = x_v - Q (Q^T x_v)
// (lin_reg.py:241, the masked sum of squares of the residualised genotypes) for states with MANY distinct
// missingness patterns, where the FP64 DMMA form (masked_kernel, 37 TFLOP/s peak) dominates the block time.
//
// The masks are 0/1, so the contraction is exact in integers once r^2 is written in fixed point:
//   1. rsq_kernel<PASS 0>: r^2 for every (variant, sample) with FP64 FMAs (coefficients Q^T x from pass 1), per-variant
//      maximum (atomicMax on the bit pattern of a non-negative double);
//   2. rsq_kernel<PASS 1>: the same r^2 (bitwise), scaled by 2^54 / max, rounded to a 55-bit integer and split into 7
//      balanced radix-256 int8 digits, written as the K-major, 128-byte-swizzled A-operand image
//      [tile of 18 variants = 126 digit rows][128-sample stage][128 rows][128 B];
//   3. mask_gemm_kernel: int8 GEMM  Z = digits x masks^T  with tcgen05.mma.cta_group::2 (CTA pairs, M = 256 digit
//      rows = 36 variants, N = up to 256 masks, K = 32 samples per instruction), both operands streamed by bulk
//      copies (TMA engine) from pre-swizzled images, int32 accumulators in TMEM (|.| <= 128 N: exact);
//   4. recombine_kernel: D[u][v] = max_v 2^-54 sum_i 256^i Z[(v, i)][u] in FP64.
// The only rounding beyond the FP64 evaluation of r^2 itself (which the DMMA form shares) is the 2^-55 max_v
// quantisation of each term of a sum of N non-negative terms.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "../../include/glow_b200.h"
#include "common.cuh"
#include "contract.cuh"
#include "kernels.h"

namespace glr {

constexpr int kM8K = 128;        // samples per stage
constexpr int kM8Digits = 7;
constexpr int kM8TileV = 18;     // variants per 128-row digit tile (126 rows used)
constexpr int kM8Rows = 128;
constexpr int kM8Ring = 6;
constexpr int kM8Threads = 256;  // GEMM: warp 0 producer, warp 1 issuer, warp 2 TMEM, warps 4-7 epilogue

// ------------------------------------------------------------------------------------------------
// r^2 sweeps.  Block = 8 warps x 32 lanes; lane = variant (256 consecutive variants per block), the block walks one
// 128-sample stage.  Q rows of the stage sit in shared memory and are read as warp-wide broadcasts; the lane keeps
// its variant's coefficients in registers.
// ------------------------------------------------------------------------------------------------
constexpr int kRsqThreads = 256;
constexpr int kRsqMaxK = 64;

template <int FMT>
__device__ __forceinline__ double m8_elem(const uint8_t* row, int64_t s, double miss) {
  if (FMT == GLR_FMT_F64) return reinterpret_cast<const double*>(row)[s];
  if (FMT == GLR_FMT_F32) return (double)reinterpret_cast<const float*>(row)[s];
  if (FMT == GLR_FMT_I8) {
    const int v = (int)reinterpret_cast<const int8_t*>(row)[s];
    return v == -1 ? miss : (double)v;
  }
  const uint32_t code = (row[s >> 2] >> (2 * (s & 3))) & 3u;
  return XLoader<GLR_FMT_PLINK2, 1>::dec(code, miss);
}

struct RsqParams {
  BlockView x;         // x.v1 = missing substitute per variant
  const double* qg;    // [n_geno][K] rows of Q in genotype-sample order (zero rows for dropped samples)
  const double* A;     // coefficients Q^T X: [K][ldS]
  int64_t ldS;
  int32_t K;
  int32_t v0;          // first variant of the sub-range the images cover
  int32_t nv;          // variants in the sub-range
  int64_t n_stages;
  unsigned long long* rmax_bits;  // [G] bit pattern of max_s r^2
  int8_t* z8;          // digit image of the sub-range
};

// KT: compile-time bound of K (the coefficient vector stays in registers)
template <int FMT, int PASS, int KT>
__global__ void __launch_bounds__(kRsqThreads) rsq_kernel(const RsqParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  double* qs = reinterpret_cast<double*>(smem_raw);  // [128][K]
  const int K = p.K;
  const int64_t stage = blockIdx.x;
  const int64_t s0 = stage * kM8K;
  const int64_t n = p.x.n_geno;
  for (int i = threadIdx.x; i < kM8K * K; i += kRsqThreads) {
    const int64_t s = s0 + i / K;
    qs[i] = s < n ? p.qg[s * K + (i % K)] : 0.0;
  }
  __syncthreads();
  const int vl = blockIdx.y * kRsqThreads + threadIdx.x;  // variant inside the sub-range
  if (vl >= p.nv) return;
  const int v = p.v0 + vl;
  const uint8_t* row = p.x.base + (int64_t)v * p.x.stride;
  const double miss = (FMT == GLR_FMT_I8 || FMT == GLR_FMT_PLINK2) ? p.x.v1[v] : 0.0;
  double a[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) a[k] = k < K ? p.A[(int64_t)k * p.ldS + v] : 0.0;

  double inv = 0.0;
  if (PASS == 1) {
    const double mx = __longlong_as_double((long long)p.rmax_bits[v]);
    inv = mx > 0.0 ? 18014398509481984.0 /* 2^54 */ / mx : 0.0;
  }
  double mx_local = 0.0;
  const int tile = vl / kM8TileV, row0 = (vl - tile * kM8TileV) * kM8Digits;
  int8_t* img = PASS == 1 ? p.z8 + ((int64_t)tile * p.n_stages + stage) * (kM8Rows * kM8K) : nullptr;

  for (int c32 = 0; c32 < kM8K / 32; ++c32) {  // 32 samples = two 16-byte chunks of every digit row
    uint32_t dig[kM8Digits][8];
#pragma unroll
    for (int q4 = 0; q4 < 8; ++q4) {
      unsigned long long qq[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int sl = c32 * 32 + q4 * 4 + j;
        const int64_t s = s0 + sl;
        double r = s < n ? m8_elem<FMT>(row, s, miss) : 0.0;
        const double* qrow = qs + sl * K;
#pragma unroll
        for (int k = 0; k < KT; ++k)
          if (k < K) r = fma(-qrow[k], a[k], r);
        const double r2 = r * r;
        if (PASS == 0) mx_local = fmax(mx_local, r2);
        // 55-bit fixed point, then bias every byte by 128: the bytes of qq are the balanced digits + 128
        if (PASS == 1) qq[j] = (unsigned long long)__double2ll_rn(r2 * inv) + 0x0080808080808080ull;
      }
      if (PASS == 1) {
#pragma unroll
        for (int i = 0; i < kM8Digits; ++i) {
          // byte i of the four values -> one word (sample order = natural), minus the bias
          const uint32_t sel = i < 4 ? (uint32_t)(i | ((4 + i) << 4)) : (uint32_t)((i - 4) | (i << 4));
          const uint32_t lo01 = i < 4 ? __byte_perm((uint32_t)qq[0], (uint32_t)qq[1], sel)
                                      : __byte_perm((uint32_t)(qq[0] >> 32), (uint32_t)(qq[1] >> 32), sel);
          const uint32_t lo23 = i < 4 ? __byte_perm((uint32_t)qq[2], (uint32_t)qq[3], sel)
                                      : __byte_perm((uint32_t)(qq[2] >> 32), (uint32_t)(qq[3] >> 32), sel);
          dig[i][q4] = __byte_perm(lo01, lo23, 0x5410u) ^ 0x80808080u;
        }
      }
    }
    if (PASS == 1) {
#pragma unroll
      for (int i = 0; i < kM8Digits; ++i) {
        const int r = row0 + i;
        int8_t* dst = img + (int64_t)r * kM8K;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int chunk = (c32 * 2 + h) ^ (r & 7);  // 128-byte swizzle
          *reinterpret_cast<uint4*>(dst + chunk * 16) = make_uint4(dig[i][4 * h], dig[i][4 * h + 1], dig[i][4 * h + 2], dig[i][4 * h + 3]);
        }
      }
    }
  }
  if (PASS == 0) atomicMax(p.rmax_bits + v, (unsigned long long)__double_as_longlong(mx_local));
}

// Fast form for K <= 16: a warp handles ONE variant at a time over the block's 128-sample stage, lane = 4 consecutive
// samples.  The lane keeps the Q rows of its 4 samples in registers for the whole block (no shared memory at all), the
// variant's coefficients are warp-uniform loads, the packed genotypes of a variant are one coalesced 32-byte read per
// warp, and every digit row of the stage is written as one full 128-byte line per warp.
template <int FMT>
__device__ __forceinline__ void m8_load4(const uint8_t* row, int64_t s, int64_t n, double miss, double (&x)[4]) {
  if (FMT == GLR_FMT_PLINK2) {
    const uint32_t b = s < n ? (uint32_t)__ldg(row + (s >> 2)) : 0xFFu;
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = s + j < n ? XLoader<GLR_FMT_PLINK2, 1>::dec((b >> (2 * j)) & 3u, miss) : 0.0;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = s + j < n ? m8_elem<FMT>(row, s + j, miss) : 0.0;
  }
}

template <int FMT, int PASS, int KT>
__global__ void __launch_bounds__(kRsqThreads) rsq4_kernel(const RsqParams p) {
  // per-block operands of the 256 variants, loaded once (coalesced) so that the variant loop sees shared-memory latency
  __shared__ double As[KT][kRsqThreads];           // coefficients
  __shared__ double ms[kRsqThreads];               // missing substitute (PASS 0/1), then 2^54 / max (PASS 1) in invs
  __shared__ double invs[kRsqThreads];
  __shared__ __align__(16) uint8_t gb[FMT == GLR_FMT_PLINK2 ? kRsqThreads : 1][32];  // packed genotypes of the stage
  const int K = p.K;
  const int64_t stage = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t s = stage * kM8K + 4 * lane;  // this lane's first sample
  const int64_t n = p.x.n_geno;
  const int vl0 = blockIdx.y * kRsqThreads;
  const int nvb = min(kRsqThreads, p.nv - vl0);
  {
    const int t = threadIdx.x;
    if (t < nvb) {
      const int v = p.v0 + vl0 + t;
#pragma unroll
      for (int k = 0; k < KT; ++k) As[k][t] = k < K ? __ldg(p.A + (int64_t)k * p.ldS + v) : 0.0;
      ms[t] = (FMT == GLR_FMT_I8 || FMT == GLR_FMT_PLINK2) ? __ldg(p.x.v1 + v) : 0.0;
      if (PASS == 1) {
        const double mx = __longlong_as_double((long long)__ldg(p.rmax_bits + v));
        invs[t] = mx > 0.0 ? 18014398509481984.0 /* 2^54 */ / mx : 0.0;
      }
      if (FMT == GLR_FMT_PLINK2) {
        const uint8_t* row = p.x.base + (int64_t)v * p.x.stride;
        const int64_t b0 = stage * 32, row_bytes = (n + 3) >> 2;
#pragma unroll
        for (int i = 0; i < 32; ++i) gb[t][i] = b0 + i < row_bytes ? __ldg(row + b0 + i) : (uint8_t)0xFF;
      }
    }
  }
  double q[4][KT];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < KT; ++k) q[j][k] = (k < K && s + j < n) ? __ldg(p.qg + (s + j) * K + k) : 0.0;
  __syncthreads();

  for (int t = warp; t < nvb; t += kRsqThreads / 32) {
    const int vl = vl0 + t;
    const int v = p.v0 + vl;
    const double miss = ms[t];
    double x[4];
    if (FMT == GLR_FMT_PLINK2) {
      const uint32_t b = gb[t][lane];
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = s + j < n ? XLoader<GLR_FMT_PLINK2, 1>::dec((b >> (2 * j)) & 3u, miss) : 0.0;
    } else {
      m8_load4<FMT>(p.x.base + (int64_t)v * p.x.stride, s, n, miss, x);
    }
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      if (k < K) {
        const double ak = As[k][t];  // warp-uniform (broadcast)
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = fma(-q[j][k], ak, x[j]);
      }
    }
    if (PASS == 0) {
      double m = fmax(fmax(x[0] * x[0], x[1] * x[1]), fmax(x[2] * x[2], x[3] * x[3]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (lane == 0) atomicMax(p.rmax_bits + v, (unsigned long long)__double_as_longlong(m));
    } else {
      const double inv = invs[t];
      unsigned long long qq[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) qq[j] = (unsigned long long)__double2ll_rn((x[j] * x[j]) * inv) + 0x0080808080808080ull;
      const int tile = vl / kM8TileV, row0 = (vl - tile * kM8TileV) * kM8Digits;
      int8_t* img = p.z8 + ((int64_t)tile * p.n_stages + stage) * (kM8Rows * kM8K);
#pragma unroll
      for (int i = 0; i < kM8Digits; ++i) {
        const uint32_t sel = i < 4 ? (uint32_t)(i | ((4 + i) << 4)) : (uint32_t)((i - 4) | (i << 4));
        const uint32_t lo01 = i < 4 ? __byte_perm((uint32_t)qq[0], (uint32_t)qq[1], sel)
                                    : __byte_perm((uint32_t)(qq[0] >> 32), (uint32_t)(qq[1] >> 32), sel);
        const uint32_t lo23 = i < 4 ? __byte_perm((uint32_t)qq[2], (uint32_t)qq[3], sel)
                                    : __byte_perm((uint32_t)(qq[2] >> 32), (uint32_t)(qq[3] >> 32), sel);
        const int r = row0 + i;
        // lane's 4 bytes sit at k = 4 lane: chunk (lane >> 2) ^ (r & 7), offset 4 (lane & 3)
        *reinterpret_cast<uint32_t*>(img + (int64_t)r * kM8K + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2))) =
            __byte_perm(lo01, lo23, 0x5410u) ^ 0x80808080u;
      }
    }
  }
}

template <int FMT, int KT>
static void launch_rsq4_kt(const RsqParams& p, cudaStream_t st) {
  const dim3 grid((unsigned)p.n_stages, (unsigned)ceil_div(p.nv, kRsqThreads));
  rsq4_kernel<FMT, 0, KT><<<grid, kRsqThreads, 0, st>>>(p);
  rsq4_kernel<FMT, 1, KT><<<grid, kRsqThreads, 0, st>>>(p);
}

template <int FMT, int KT>
static void launch_rsq_kt(const RsqParams& p, cudaStream_t st) {
  const dim3 grid((unsigned)p.n_stages, (unsigned)ceil_div(p.nv, kRsqThreads));
  const size_t smem = (size_t)kM8K * p.K * sizeof(double);
  cudaFuncSetAttribute(rsq_kernel<FMT, 0, KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(rsq_kernel<FMT, 1, KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  rsq_kernel<FMT, 0, KT><<<grid, kRsqThreads, smem, st>>>(p);
  rsq_kernel<FMT, 1, KT><<<grid, kRsqThreads, smem, st>>>(p);
}
template <int FMT>
static void launch_rsq_fmt(const RsqParams& p, cudaStream_t st) {
  if (p.K <= 8) launch_rsq4_kt<FMT, 8>(p, st);
  else if (p.K <= 12) launch_rsq4_kt<FMT, 12>(p, st);
  else if (p.K <= 16) launch_rsq4_kt<FMT, 16>(p, st);
  else if (p.K <= 24) launch_rsq_kt<FMT, 24>(p, st);
  else launch_rsq_kt<FMT, kRsqMaxK>(p, st);
}

// ------------------------------------------------------------------------------------------------
// int8 GEMM on CTA pairs
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t m8_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ void m8_mma_ss2(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void m8_commit2(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void m8_arrive_cta(uint64_t* bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];\n" ::"r"(remote) : "memory");
}
__device__ __forceinline__ bool m8_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void m8_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void m8_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
      "[%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

struct MaskGemmParams {
  const int8_t* z8;   // digit image: [tile][stage][128 rows][128 B]
  const int8_t* m8;   // mask image:  [mask block][stage][NB rows][128 B]
  int32_t n_tiles;    // digit tiles of the sub-range
  int32_t n_mb;       // mask blocks
  int32_t NB;         // mask rows per block = MMA N (multiple of 16, <= 256)
  int64_t n_stages;
  int32_t* Z;         // [tile][mask block][128 rows][NB] int32
};

struct MaskGemmSmem {
  uint64_t full[kM8Ring];      // this CTA's bulk copies of a stage have landed
  uint64_t peer_full[kM8Ring]; // leader: the peer CTA's copies have landed
  uint64_t done[kM8Ring];      // the MMAs of a stage have completed (multicast to both CTAs)
  uint64_t tmem_full;
  uint64_t tmem_empty;
  uint32_t tmem_base;
  uint32_t pad;
};

__global__ void __launch_bounds__(kM8Threads, 1) mask_gemm_kernel(const MaskGemmParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t NB = (uint32_t)p.NB;
  constexpr uint32_t a_bytes = kM8Rows * kM8K;            // this CTA's digit tile of a stage
  const uint32_t b_stage_bytes = NB * kM8K;               // a whole mask stage in the image
  const uint32_t b_bytes = b_stage_bytes / 2;             // the half this CTA holds
  const uint32_t stage_bytes = a_bytes + b_bytes;
  MaskGemmSmem* sm = reinterpret_cast<MaskGemmSmem*>(smem_raw + (size_t)kM8Ring * (a_bytes + 16384));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(rank));
  if (tid == 0) {
    for (int s = 0; s < kM8Ring; ++s) {
      mbar_init(&sm->full[s], 1);
      mbar_init(&sm->peer_full[s], 1);
      mbar_init(&sm->done[s], 1);
    }
    mbar_init(&sm->tmem_full, 1);
    mbar_init(&sm->tmem_empty, 8);  // 4 epilogue warps of each CTA
    mbar_fence_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&sm->tmem_base)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  m8_cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = sm->tmem_base;

  // work item = (pair of digit tiles, mask block); mask block fastest so that a wave shares few digit tiles
  const int n_tp = (p.n_tiles + 1) / 2;
  const int n_items = n_tp * p.n_mb;
  const int first = (int)(blockIdx.x >> 1), stride = (int)(gridDim.x >> 1);
  const int n_st = (int)p.n_stages;

  if (warp == 0) {
    // ---- producer: this CTA's digit tile and its half of the mask rows ----
    if (lane == 0) {
      uint32_t g = 0;
      for (int item = first; item < n_items; item += stride) {
        const int tp = item / p.n_mb, mb = item - tp * p.n_mb;
        const int tile = min(tp * 2 + (int)rank, p.n_tiles - 1);
        const int8_t* srcA = p.z8 + (int64_t)tile * n_st * a_bytes;
        const int8_t* srcB = p.m8 + (int64_t)mb * n_st * b_stage_bytes + rank * b_bytes;
        for (int k = 0; k < n_st; ++k, ++g) {
          const uint32_t s = g % kM8Ring;
          if (g >= (uint32_t)kM8Ring) mbar_wait_backoff(&sm->done[s], ((g / kM8Ring) & 1u) ^ 1u);
          unsigned char* dst = smem_raw + (size_t)s * (a_bytes + 16384);
          mbar_expect_tx(&sm->full[s], stage_bytes);
          bulk_g2s(dst, srcA + (int64_t)k * a_bytes, a_bytes, &sm->full[s]);
          bulk_g2s(dst + a_bytes, srcB + (int64_t)k * b_stage_bytes, b_bytes, &sm->full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (rank != 0) {
      // peer: forward "my operands of stage g have landed" to the leader
      uint32_t g = 0;
      for (int item = first; item < n_items; item += stride)
        for (int k = 0; k < n_st; ++k, ++g) {
          const uint32_t s = g % kM8Ring;
          mbar_wait(&sm->full[s], (g / kM8Ring) & 1u);
          if (lane == 0) m8_arrive_cta(&sm->peer_full[s], 0);
          __syncwarp();
        }
    } else {
      // leader: converged warp, one elected lane issues
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((NB >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t ring_addr = smem_u32(smem_raw);
      uint32_t g = 0;
      int icount = 0;
      for (int item = first; item < n_items; item += stride, ++icount) {
        if (icount > 0) mbar_wait(&sm->tmem_empty, (uint32_t)((icount - 1) & 1));
        for (int k = 0; k < n_st; ++k, ++g) {
          const uint32_t s = g % kM8Ring, ph = (g / kM8Ring) & 1u;
          mbar_wait(&sm->full[s], ph);
          mbar_wait(&sm->peer_full[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const uint32_t base = ring_addr + s * (a_bytes + 16384);
          const uint64_t da0 = m8_desc(base), db0 = m8_desc(base + a_bytes);
          if (m8_elect_one()) {
#pragma unroll
            for (int kk = 0; kk < kM8K / 32; ++kk)
              m8_mma_ss2(tmem, da0 + (uint64_t)(kk * 2), db0 + (uint64_t)(kk * 2), idesc, (k > 0 || kk > 0) ? 1u : 0u);
            m8_commit2(&sm->done[s]);
            if (k == n_st - 1) m8_commit2(&sm->tmem_full);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    // ---- epilogue: TMEM lane = digit row; the int32 sums go to the scratch the recombination reads ----
    const int q = warp & 3;
    const int rowl = q * 32 + lane;
    int icount = 0;
    for (int item = first; item < n_items; item += stride, ++icount) {
      const int tp = item / p.n_mb, mb = item - tp * p.n_mb;
      const int tile = tp * 2 + (int)rank;
      mbar_wait(&sm->tmem_full, (uint32_t)(icount & 1));
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      int32_t* dst = p.Z + (((int64_t)min(tile, p.n_tiles - 1) * p.n_mb + mb) * kM8Rows + rowl) * NB;
      for (uint32_t c0 = 0; c0 < NB; c0 += 32) {
        uint32_t z[32];
        m8_tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c0, z);
        if (tile < p.n_tiles) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            if (c0 + i < NB) *reinterpret_cast<uint4*>(dst + c0 + i) = make_uint4(z[i], z[i + 1], z[i + 2], z[i + 3]);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) m8_arrive_cta(&sm->tmem_empty, 0);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  m8_cluster_sync();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(256));
}

// D[u][v0 + vl] = max 2^-54 sum_i 256^i Z[(vl, i)][u]
__global__ void recombine_kernel(const int32_t* Z, const unsigned long long* rmax_bits, int v0, int nv, int U, int n_mb, int NB,
                                 double* D, int64_t ldS) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)nv * U) return;
  const int u = (int)(idx % U), vl = (int)(idx / U);
  const int tile = vl / kM8TileV, row0 = (vl - tile * kM8TileV) * kM8Digits;
  const int mb = u / NB, col = u - mb * NB;
  const int32_t* z = Z + (((int64_t)tile * n_mb + mb) * kM8Rows + row0) * NB + col;
  double acc = 0.0;
#pragma unroll
  for (int i = kM8Digits - 1; i >= 0; --i) acc = fma(acc, 256.0, (double)z[(int64_t)i * NB]);
  const double mx = __longlong_as_double((long long)rmax_bits[v0 + vl]);
  D[(int64_t)u * ldS + v0 + vl] = acc * (mx * 5.5511151231257827e-17 /* 2^-54 */);
}

// ------------------------------------------------------------------------------------------------
// mask image: unique masks, row-major [N x U] doubles -> [mask block][stage][NB rows][128 B] int8, swizzled
// ------------------------------------------------------------------------------------------------
__global__ void pack_mask8_kernel(const double* mu, int64_t N, int U, const int64_t* geno_to_row, int64_t Ng, int NB, int64_t n_stages,
                                  int8_t* dst) {
  const int64_t total = (int64_t)U * n_stages * kM8K;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int u = (int)(idx % U);
  const int64_t sg = idx / U;
  int8_t m = 0;
  if (sg < Ng) {
    const int64_t row = geno_to_row ? geno_to_row[sg] : sg;
    if (row >= 0 && row < N) m = mu[row * U + u] != 0.0 ? 1 : 0;
  }
  const int mb = u / NB, r = u - mb * NB;
  const int64_t stage = sg / kM8K;
  const int k = (int)(sg % kM8K);
  dst[((int64_t)mb * n_stages + stage) * NB * kM8K + (int64_t)r * kM8K + (((k >> 4) ^ (r & 7)) << 4) + (k & 15)] = m;
}

__global__ void gather_q_kernel(const double* Q, int64_t N, int K, const int64_t* geno_to_row, int64_t Ng, double* qg) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Ng * K) return;
  const int64_t sg = idx / K;
  const int64_t row = geno_to_row ? geno_to_row[sg] : sg;
  qg[idx] = (row >= 0 && row < N) ? Q[row * K + (idx % K)] : 0.0;
}

void mask8_geometry(int U, int* n_mb, int* NB) {
  *NB = U <= 256 ? (int)round_up(U, 16) : 256;
  *n_mb = (U + *NB - 1) / *NB;
}

void launch_pack_mask8(const double* mu, int64_t N, int U, const int64_t* geno_to_row, int64_t Ng, int NB, int64_t n_stages,
                       int8_t* dst, cudaStream_t st) {
  const int64_t total = (int64_t)U * n_stages * kM8K;
  pack_mask8_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(mu, N, U, geno_to_row, Ng, NB, n_stages, dst);
}

void launch_gather_q(const double* Q, int64_t N, int K, const int64_t* geno_to_row, int64_t Ng, double* qg, cudaStream_t st) {
  if (K <= 0) return;
  gather_q_kernel<<<(unsigned)ceil_div(Ng * K, 256), 256, 0, st>>>(Q, N, K, geno_to_row, Ng, qg);
}

size_t mask8_tile_bytes(int64_t n_stages) { return (size_t)n_stages * kM8Rows * kM8K; }
size_t mask8_z_bytes(int n_tiles, int n_mb, int NB) { return (size_t)n_tiles * n_mb * kM8Rows * NB * sizeof(int32_t); }
int mask8_tile_variants() { return kM8TileV; }

// one sub-range of variants [v0, v0 + nv): r^2 max, digit image, GEMM, recombination into D[U][ldS]
void launch_mask8(const Mask8Params& p, int sm_count, cudaStream_t st) {
  if (p.nv <= 0) return;
  RsqParams rp;
  rp.x = p.x;
  rp.qg = p.qg;
  rp.A = p.A;
  rp.ldS = p.ldS;
  rp.K = p.K;
  rp.v0 = p.v0;
  rp.nv = p.nv;
  rp.n_stages = p.n_stages;
  rp.rmax_bits = p.rmax_bits;
  rp.z8 = p.z8;
  cudaMemsetAsync(p.rmax_bits + p.v0, 0, (size_t)p.nv * sizeof(unsigned long long), st);
  switch (p.x.format) {
    case GLR_FMT_F64: launch_rsq_fmt<GLR_FMT_F64>(rp, st); break;
    case GLR_FMT_F32: launch_rsq_fmt<GLR_FMT_F32>(rp, st); break;
    case GLR_FMT_I8: launch_rsq_fmt<GLR_FMT_I8>(rp, st); break;
    default: launch_rsq_fmt<GLR_FMT_PLINK2>(rp, st); break;
  }
  MaskGemmParams gp;
  gp.z8 = p.z8;
  gp.m8 = p.m8;
  gp.n_tiles = (p.nv + kM8TileV - 1) / kM8TileV;
  gp.n_mb = p.n_mb;
  gp.NB = p.NB;
  gp.n_stages = p.n_stages;
  gp.Z = p.Z;
  const int n_items = (gp.n_tiles + 1) / 2 * gp.n_mb;
  const size_t smem = (size_t)kM8Ring * (kM8Rows * kM8K + 16384) + sizeof(MaskGemmSmem) + 64;
  cudaFuncSetAttribute(mask_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * std::min(n_items, sm_count / 2));
  cfg.blockDim = dim3(kM8Threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, mask_gemm_kernel, gp);
  const int64_t total = (int64_t)p.nv * p.U;
  recombine_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(p.Z, p.rmax_bits, p.v0, p.nv, p.U, p.n_mb, p.NB, p.D, p.ldS);
}

}  // namespace glr
