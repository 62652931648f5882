// Exact FP64-accurate contraction S = W^T X on the 5th-generation tensor cores (tcgen05.mma kind::i8,
// accumulators in TMEM) for PLINK 2-bit genotypes.
//
// Blackwell has no tcgen05 FP64 kind and DMMA tops out at 37 TFLOP/s, but 2-bit genotypes are small
// integers, so the product can be computed EXACTLY with integer tensor-core MMAs (Ozaki-style
// error-free splitting of the FP64 operand):
//   * every column of W = [Q(:,1:) | Ytilde] is scaled by its max |w| and rounded to a 55-bit signed
//     integer, then written as 8 balanced radix-128 digits d_i in [-64, 63]  (int8);
//   * a genotype row is x = xi + mu * e with xi in {0,1,2} (missing -> 0) and e the missing-call
//     indicator, both int8;
//   * Zx_i = d_i^T xi and Ze_i = d_i^T e are int32 dot products (|.| <= 64 * 2 * N < 2^31 for
//     N < 1.6e7): exact on the tensor cores;
//   * S[c][v] = scale_c * 2^-54 * sum_i 128^i (Zx_i + mu_v Ze_i), assembled in FP64 in the epilogue.
// The only rounding is the 2^-54 (relative to the column maximum) quantisation of W -- finer than
// the FP64 mantissa for the large entries that dominate the sums -- and the final FP64 Horner sum.
//
// CTA roles (384 threads, 1 CTA/SM, persistent over 128-variant tiles):
//   warp 0   : lane 0 streams the packed digit operand (swizzled image, one bulk copy per stage)
//   warp 1   : lane 0 issues tcgen05.mma; MMA completion (tcgen05.commit) frees the stage
//   warp 2   : TMEM allocation / deallocation
//   warps 4-11: decode the 2-bit rows of the tile into the two int8 A operands (xi, e) directly in the
//               128-byte-swizzled K-major layout the MMA reads; warps 4-7 also run the epilogue
//               (tcgen05.ld of their 32 TMEM lanes = 32 variants, digit recombination, FP64 store).
#include <algorithm>

#include "../../include/glow_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace glr {

constexpr int kS8Stages = 4;
constexpr int kS8K = 128;           // samples per stage = bytes per operand row per stage
constexpr int kS8TileV = 128;       // variants per tile = MMA M
constexpr int kS8Threads = 384;
constexpr int kS8DecodeWarps = 8;
constexpr int kS8FirstDecodeWarp = 4;
constexpr int kS8Slices = 8;

__device__ __forceinline__ uint64_t s8_desc(uint32_t saddr) {
  // K-major operand, SWIZZLE_128B: start >> 4 | LBO (unused) = 1 | SBO = 1024 B | version 1 | layout 2
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ void s8_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void s8_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// 4 two-bit codes (one byte) -> PRMT selector with one code per nibble
__device__ __forceinline__ uint32_t s8_selector(uint32_t b) {
  uint32_t t = (b | (b << 4)) & 0x0F0Fu;
  return (t | (t << 2)) & 0x3333u;
}

struct Split8Smem {
  uint64_t full_w[kS8Stages];
  uint64_t full_x[kS8Stages];
  uint64_t empty[kS8Stages];
  uint64_t tmem_full;
  uint64_t tmem_empty;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kS8Threads, 1) split8_kernel(const Split8Params p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1024-byte aligned operand stages first, bookkeeping after
  const int NB = p.NB;
  const uint32_t w_bytes = (uint32_t)NB * kS8K;
  const uint32_t stage_bytes = w_bytes + 2u * kS8TileV * kS8K;
  unsigned char* stages = smem_raw;
  Split8Smem* sm = reinterpret_cast<Split8Smem*>(smem_raw + (size_t)kS8Stages * stage_bytes);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kS8Stages; ++s) {
      mbar_init(&sm->full_w[s], 1);
      mbar_init(&sm->full_x[s], kS8DecodeWarps);
      mbar_init(&sm->empty[s], 1);
    }
    mbar_init(&sm->tmem_full, 1);
    mbar_init(&sm->tmem_empty, 4);
    mbar_fence_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&sm->tmem_base)),
                 "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = sm->tmem_base;
  const uint32_t tmem_x = tmem, tmem_e = tmem + 256;

  const int n_tiles = (p.x.n_variants + kS8TileV - 1) / kS8TileV;
  const int64_t n_st = p.n_stages;  // 128-sample stages covering the genotype samples

  if (warp == 0) {
    // ---- digit-operand producer ----
    if (lane == 0) {
      int64_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int64_t k = 0; k < n_st; ++k, ++it) {
          const int s = (int)(it % kS8Stages);
          if (it >= kS8Stages) mbar_wait_backoff(&sm->empty[s], (uint32_t)((it / kS8Stages) & 1) ^ 1u);
          mbar_expect_tx(&sm->full_w[s], w_bytes);
          bulk_g2s(stages + (size_t)s * stage_bytes, p.w8 + k * w_bytes, w_bytes, &sm->full_w[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(kS8TileV >> 4) << 24);
    int64_t it = 0;
    int tcount = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
      // the epilogue of the previous tile must have drained the accumulators
      if (tcount > 0) mbar_wait(&sm->tmem_empty, (uint32_t)((tcount - 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      for (int64_t k = 0; k < n_st; ++k, ++it) {
        const int s = (int)(it % kS8Stages);
        const uint32_t ph = (uint32_t)((it / kS8Stages) & 1);
        mbar_wait(&sm->full_w[s], ph);
        mbar_wait(&sm->full_x[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        if (lane == 0) {
          const uint32_t w_addr = smem_u32(stages + (size_t)s * stage_bytes);
          const uint32_t x_addr = w_addr + w_bytes;
          const uint32_t e_addr = x_addr + kS8TileV * kS8K;
#pragma unroll
          for (int kk = 0; kk < kS8K / 32; ++kk) {
            const uint32_t acc = (k > 0 || kk > 0) ? 1u : 0u;
            const uint64_t dw = s8_desc(w_addr + kk * 32);
            s8_mma(tmem_x, s8_desc(x_addr + kk * 32), dw, idesc, acc);
            s8_mma(tmem_e, s8_desc(e_addr + kk * 32), dw, idesc, acc);
          }
          s8_commit(&sm->empty[s]);                       // frees the stage when these MMAs retire
          if (k == n_st - 1) s8_commit(&sm->tmem_full);  // accumulators of the tile are complete
        }
        __syncwarp();
      }
    }
  } else if (warp >= kS8FirstDecodeWarp) {
    // ---- decode (all 8 warps) + epilogue (first 4 warps) ----
    const int dt = tid - kS8FirstDecodeWarp * 32;  // 0..255
    const int vloc = dt >> 1, half = dt & 1;       // variant in tile, which 16 bytes of the stage's 32
    const int64_t row_bytes = (p.x.n_geno + 3) >> 2;
    int64_t it = 0;
    int tcount = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
      const int v = min(tile * kS8TileV + vloc, p.x.n_variants - 1);
      const uint8_t* row = p.x.base + (int64_t)v * p.x.stride;
      uint32_t nxt[5];
      uint32_t sh_nxt;
      auto load = [&](int64_t k) {
        const int64_t b0 = k * 32 + half * 16;
        if (b0 + 19 <= row_bytes) {
          const uint8_t* q8 = row + b0;
          const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(q8) & 3);
          const uint32_t* q = reinterpret_cast<const uint32_t*>(q8 - a);
#pragma unroll
          for (int i = 0; i < 4; ++i) nxt[i] = __ldg(q + i);
          nxt[4] = a ? __ldg(q + 4) : 0u;
          sh_nxt = a * 8;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint32_t w = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const int64_t off = b0 + i * 4 + b;
              const uint32_t byte = off < row_bytes ? (uint32_t)__ldg(row + off) : 0xFFu;  // code 11 = 0
              w |= byte << (8 * b);
            }
            nxt[i] = w;
          }
          nxt[4] = 0u;
          sh_nxt = 0;
        }
      };
      load(0);
      for (int64_t k = 0; k < n_st; ++k, ++it) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = __funnelshift_r(nxt[i], nxt[i + 1], sh_nxt);
        if (k + 1 < n_st) load(k + 1);
        const int s = (int)(it % kS8Stages);
        if (it >= kS8Stages) mbar_wait(&sm->empty[s], (uint32_t)((it / kS8Stages) & 1) ^ 1u);
        unsigned char* xrow = stages + (size_t)s * stage_bytes + w_bytes + vloc * kS8K;
        unsigned char* erow = xrow + kS8TileV * kS8K;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t xs[4], es[4];
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const uint32_t sel = s8_selector((w[i] >> (8 * b)) & 0xFFu);
            xs[b] = __byte_perm(0x00010002u, 0u, sel);  // code 00 -> 2, 01 -> 0 (missing), 10 -> 1, 11 -> 0
            es[b] = __byte_perm(0x00000100u, 0u, sel);  // code 01 -> 1
          }
          const int chunk = (half * 4 + i) ^ (vloc & 7);  // 128-byte swizzle
          *reinterpret_cast<uint4*>(xrow + chunk * 16) = make_uint4(xs[0], xs[1], xs[2], xs[3]);
          *reinterpret_cast<uint4*>(erow + chunk * 16) = make_uint4(es[0], es[1], es[2], es[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic writes -> async proxy (MMA reads)
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm->full_x[s]);
      }
      // ---- epilogue: 4 warps x 32 lanes = the 128 variants of the tile ----
      if (warp < kS8FirstDecodeWarp + 4) {
        mbar_wait(&sm->tmem_full, (uint32_t)(tcount & 1));
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const int q = warp - kS8FirstDecodeWarp;
        const int64_t vg = (int64_t)tile * kS8TileV + q * 32 + lane;  // this lane's variant (may exceed n_variants)
        const double mu = p.x.v1[min(vg, (int64_t)p.x.n_variants - 1)];
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        for (int c0 = 0; c0 < NB; c0 += 16) {
          uint32_t zx[16], ze[16];
          tmem_ld16(tmem_x + lane_base + c0, zx);
          tmem_ld16(tmem_e + lane_base + c0, ze);
#pragma unroll
          for (int h = 0; h < 2; ++h) {  // two columns of W per 16 accumulator columns (8 digits each)
            const int c = c0 / 8 + h;
            if (c < p.n_cols) {
              double acc = 0.0;
#pragma unroll
              for (int i = kS8Slices - 1; i >= 0; --i)
                acc = fma(acc, 128.0, fma(mu, (double)(int32_t)ze[h * 8 + i], (double)(int32_t)zx[h * 8 + i]));
              p.S[(int64_t)c * p.ldS + vg] = acc * p.scale[c];
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm->tmem_empty);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512));
}

void launch_split8(const Split8Params& p, int sm_count, cudaStream_t st) {
  if (p.x.n_variants <= 0) return;
  const int n_tiles = (p.x.n_variants + kS8TileV - 1) / kS8TileV;
  const size_t smem = (size_t)kS8Stages * ((size_t)p.NB * kS8K + 2 * kS8TileV * kS8K) + sizeof(Split8Smem) + 64;
  cudaFuncSetAttribute(split8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  split8_kernel<<<std::min(n_tiles, sm_count), kS8Threads, smem, st>>>(p);
}

// ------------------------------------------------------------------------------------------------
// packing: W (row-major [N x ld], columns c0..c0+ncols) -> per 128-sample stage a swizzled image
// [NB rows = 8 digits x column][128 bytes], row of (column c, digit i) = 8 c + i
// ------------------------------------------------------------------------------------------------
__global__ void colmax_kernel(const double* src, int64_t N, int64_t ld, int c0, int ncols, double* out) {
  __shared__ double sh[256];
  const int c = blockIdx.x;
  double m = 0.0;
  for (int64_t i = threadIdx.x; i < N; i += 256) m = fmax(m, fabs(src[i * ld + c0 + c]));
  sh[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) out[c] = sh[0];
}

__global__ void pack_split8_kernel(const double* src, int64_t N, int64_t ld, int c0, int ncols, const double* colmax,
                                   const int64_t* geno_to_row, int64_t Ng, int8_t* dst, int dst_col0, int NB,
                                   int64_t n_stages) {
  const int64_t total = (int64_t)ncols * n_stages * kS8K;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % ncols);
  const int64_t sg = idx / ncols;  // genotype sample index (padded to whole stages)
  double v = 0.0;
  if (sg < Ng) {
    const int64_t row = geno_to_row ? geno_to_row[sg] : sg;
    if (row >= 0 && row < N) v = src[row * ld + c0 + c];
  }
  const double mx = colmax[c];
  long long q = mx > 0.0 ? __double2ll_rn(v / mx * 18014398509481984.0 /* 2^54 */) : 0ll;
  const int64_t stage = sg / kS8K;
  const int k = (int)(sg % kS8K);
  int8_t* base = dst + stage * (int64_t)NB * kS8K;
#pragma unroll
  for (int i = 0; i < kS8Slices; ++i) {
    long long d = ((q + 64) & 127) - 64;  // balanced digit in [-64, 63]
    q = (q - d) >> 7;
    const int r = (dst_col0 + c) * kS8Slices + i;
    base[(int64_t)r * kS8K + (((k >> 4) ^ (r & 7)) << 4) + (k & 15)] = (int8_t)d;
  }
}

void launch_pack_split8(const double* src, int64_t N, int64_t ld, int c0, int ncols, double* colmax_scratch,
                        const int64_t* geno_to_row, int64_t Ng, int8_t* dst, int dst_col0, int NB, int64_t n_stages,
                        cudaStream_t st) {
  if (ncols <= 0) return;
  colmax_kernel<<<ncols, 256, 0, st>>>(src, N, ld, c0, ncols, colmax_scratch);
  const int64_t total = (int64_t)ncols * n_stages * kS8K;
  pack_split8_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(src, N, ld, c0, ncols, colmax_scratch, geno_to_row,
                                                                      Ng, dst, dst_col0, NB, n_stages);
}

}  // namespace glr
