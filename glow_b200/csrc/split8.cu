// Exact FP64-accurate contraction S = W^T X on the 5th-generation tensor cores (tcgen05.mma kind::i8,
// accumulators in TMEM) for PLINK 2-bit genotypes, with the genotype counts (mean imputation, sum x,
// sum x^2) fused into the same pass over the packed rows.
//
// Blackwell has no tcgen05 FP64 kind and DMMA tops out at 37 TFLOP/s, but 2-bit genotypes are small
// integers, so the product can be computed EXACTLY with integer tensor-core MMAs (Ozaki-style
// error-free splitting of the FP64 operand):
//   * every column of W = [Q(:,1:) | Ytilde] is scaled by its max |w| and rounded to a signed 55-bit
//     integer (quantum 2^-54 of the column maximum, i.e. an operand error of about one FP64 ulp of the largest entries), then written as 7 balanced radix-256 digits
//     d_i in [-128, 127]  (int8);
//   * a genotype row is x = xi + mu * e with xi in {0,1,2} (missing -> 0) and e the missing-call
//     indicator, both int8;
//   * Zx_i = d_i^T xi and Ze_i = d_i^T e are int32 dot products (|.| <= 128 * 2 * N < 2^31 for
//     N < 8.3e6): exact on the tensor cores;
//   * S[c][v] = max|w_c| * 2^-54 * sum_i 256^i (Zx_i + mu_v Ze_i), assembled in FP64 in the epilogue.
// The only roundings are the 2^-54 quantisation of W relative to its column maximum (finer than the
// FP64 mantissa for the entries that dominate the sums; the accumulation itself is exact, unlike a
// floating-point GEMM) and the final FP64 Horner sum.
//
// MMA shape: A = digit rows (M = 128: up to 18 columns x 7 digits; two M blocks for up to 36
// columns), B = [xi of 128 variants | e of the same 128 variants] (N = 256), K = 32 samples per
// instruction, 128 samples per pipeline stage.  D[digit row][variant | 128 + variant] lives in TMEM.
//
// CTA roles (384 threads, 1 CTA/SM, persistent over 128-variant tiles):
//   warp 0    : lane 0 streams the packed digit operand (pre-swizzled image, one bulk copy per stage)
//   warps 1, 3: lane 0 of each issues tcgen05.mma for alternate stages; tcgen05.commit frees the buffers /
//               publishes the accumulators
//   warp 2    : TMEM allocation / deallocation
//   warps 4-11: decode the 2-bit rows of the tile into the int8 B operand (xi, e) directly in the
//               128-byte-swizzled K-major layout the MMA reads, and count the genotype codes on the way;
//               warps 4-7 also run the epilogue (tcgen05.ld of their 32 TMEM lanes, transpose through
//               an L2-resident scratch, digit recombination per variant, FP64 store).
// The digit operand and the decoded genotypes travel through separate rings (the first is filled
// from L2 and needs depth to cover the copy latency, the second is filled locally).
#include <algorithm>

#include "../../include/glow_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace glr {

constexpr int kS8K = 128;      // samples per stage = bytes per operand row per stage
constexpr int kS8TileV = 128;  // variants per tile
constexpr int kS8N = 256;      // MMA N = xi | e of the tile
constexpr int kS8M = 128;      // digit rows per M block
constexpr int kS8Digits = 7;
constexpr int kS8DecodeWarps = 8;
constexpr int kS8FirstDecodeWarp = 4;
constexpr int kS8Threads = (kS8FirstDecodeWarp + kS8DecodeWarps) * 32;
constexpr int kS8Words = 4;    // 32-bit words (16 samples each) per decode thread per stage: 2 threads per variant
constexpr int kS8MaxRing = 8;

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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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

// 4 two-bit codes (one byte) -> PRMT selector with one code per nibble
__device__ __forceinline__ uint32_t s8_selector(uint32_t b) {
  uint32_t t = (b | (b << 4)) & 0x0F0Fu;
  return (t | (t << 2)) & 0x3333u;
}

struct Split8Smem {
  uint64_t full_w[kS8MaxRing];
  uint64_t empty_w[kS8MaxRing];
  uint64_t full_x[kS8MaxRing];
  uint64_t empty_x[kS8MaxRing];
  uint64_t tmem_full;
  uint64_t tmem_empty;
  uint64_t tile_started;
  uint32_t tmem_base;
  uint32_t pad;
  int32_t counts[2][kS8TileV][4];  // per tile parity: n00, n01 (missing), n10 per variant
};

__global__ void __launch_bounds__(kS8Threads, 1) split8_kernel(const Split8Params p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int MB = p.MB, DW = p.n_stage_bufs, DX = p.n_x_bufs;
  const uint32_t w_bytes = (uint32_t)MB * kS8M * kS8K;
  constexpr uint32_t x_bytes = (uint32_t)kS8N * kS8K;
  unsigned char* w_ring = smem_raw;  // 1024-byte aligned operand buffers first, bookkeeping after
  unsigned char* x_ring = smem_raw + (size_t)DW * w_bytes;
  Split8Smem* sm = reinterpret_cast<Split8Smem*>(x_ring + (size_t)DX * x_bytes);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < DW; ++s) {
      mbar_init(&sm->full_w[s], 1);
      mbar_init(&sm->empty_w[s], 1);
    }
    for (int s = 0; s < DX; ++s) {
      mbar_init(&sm->full_x[s], kS8DecodeWarps);
      mbar_init(&sm->empty_x[s], 1);
    }
    mbar_init(&sm->tmem_full, 2);   // both MMA issuer threads commit their last stage of the tile
    mbar_init(&sm->tmem_empty, 4);
    mbar_init(&sm->tile_started, 1);
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

  const int n_tiles = (p.x.n_variants + kS8TileV - 1) / kS8TileV;
  const int64_t n_st = p.n_stages;  // 128-sample stages covering the genotype samples

  if (warp == 0) {
    // ---- digit-operand producer ----
    if (lane == 0) {
      // ring slot and phase are carried incrementally: a 64-bit modulo by a run-time ring depth costs
      // hundreds of cycles per stage on the issuing thread
      int s = 0;
      uint32_t ph = 0;
      bool wrapped = false;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int64_t k = 0; k < n_st; ++k) {
          if (wrapped) mbar_wait_backoff(&sm->empty_w[s], ph ^ 1u);
          mbar_expect_tx(&sm->full_w[s], w_bytes);
          bulk_g2s(w_ring + (size_t)s * w_bytes, p.w8 + k * w_bytes, w_bytes, &sm->full_w[s]);
          if (++s == DW) {
            s = 0;
            ph ^= 1u;
            wrapped = true;
          }
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ---- MMA issuers: two single threads (warp 1 and warp 3, lane 0) take alternate stages, so that one
    // is issuing while the other is committing / waiting for the next stage's operands.  Integer
    // accumulation is order independent; the only ordering needed is that the stage that OVERWRITES the
    // accumulators (first stage of a tile) is issued before the other thread's first stage of that tile. ----
    if (lane == 0) {
      const int j = warp == 1 ? 0 : 1;
      // instruction descriptor: D = S32, A = B = signed 8 bit, both K-major, N >> 3, M >> 4
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kS8N >> 3) << 17) | ((uint32_t)(kS8M >> 4) << 24);
      int sw = j % DW;  // ring slot / phase of this thread's next stage (global stage index j, j+2, ...)
      uint32_t ph = (uint32_t)(j / DW) & 1u;
      int base_par = 0;  // parity of the global index of the tile's first stage
      int tcount = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
        const int k0 = (j - base_par + 2) & 1;  // first stage of this thread inside the tile
        for (int64_t k = k0; k < n_st; k += 2) {
          if (k == 0) {
            // the epilogue of the previous tile must have drained the accumulators
            if (tcount > 0) mbar_wait(&sm->tmem_empty, (uint32_t)((tcount - 1) & 1));
          } else if (k == 1) {
            mbar_wait(&sm->tile_started, (uint32_t)(tcount & 1));  // stage 0 (overwrite) has been issued
          }
          mbar_wait(&sm->full_w[sw], ph);
          mbar_wait(&sm->full_x[sw], ph);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const uint32_t w_addr = smem_u32(w_ring + (size_t)sw * w_bytes);
          const uint32_t x_addr = smem_u32(x_ring + (size_t)sw * x_bytes);
#pragma unroll
          for (int kk = 0; kk < kS8K / 32; ++kk) {
            const uint32_t acc = (k > 0 || kk > 0) ? 1u : 0u;
            const uint64_t db = s8_desc(x_addr + kk * 32);
            for (int mb = 0; mb < MB; ++mb)
              s8_mma(tmem + mb * kS8N, s8_desc(w_addr + mb * (kS8M * kS8K) + kk * 32), db, idesc, acc);
          }
          if (k == 0) mbar_arrive(&sm->tile_started);
          s8_commit(&sm->empty_w[sw]);                  // ONE commit per stage: frees both buffers of the stage
          if (k + 2 >= n_st) s8_commit(&sm->tmem_full);  // this thread's last stage of the tile (barrier count 2)
          sw += 2;
          while (sw >= DW) {
            sw -= DW;
            ph ^= 1u;
          }
        }
        base_par = (base_par + (int)(n_st & 1)) & 1;
      }
    }
  } else if (warp >= kS8FirstDecodeWarp) {
    // ---- decode + genotype counts (8 warps), epilogue (first 4 of them) ----
    const int dt = tid - kS8FirstDecodeWarp * 32;  // 0..255
    const int vloc = dt >> 1, half = dt & 1;       // variant in tile, which 16 bytes of the stage's 32
    const int64_t n_geno = p.x.n_geno;
    const int64_t row_bytes = (n_geno + 3) >> 2;
    const int64_t full_bytes = n_geno >> 2;
    int s = 0;
    uint32_t ph = 0;
    bool wrapped = false;
    int tcount = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
      const int v = min(tile * kS8TileV + vloc, p.x.n_variants - 1);
      const uint8_t* row = p.x.base + (int64_t)v * p.x.stride;
      uint32_t nxt[kS8Words + 1];
      uint32_t sh_nxt;
      auto load = [&](int64_t k) {
        const int64_t b0 = k * 32 + half * 16;
        if (b0 + 4 * kS8Words + 3 <= full_bytes) {  // every word touched lies inside the whole bytes of the row
          const uint8_t* q8 = row + b0;
          const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(q8) & 3);
          const uint32_t* q = reinterpret_cast<const uint32_t*>(q8 - a);
#pragma unroll
          for (int i = 0; i < kS8Words; ++i) nxt[i] = __ldg(q + i);
          nxt[kS8Words] = a ? __ldg(q + kS8Words) : 0u;
          sh_nxt = a * 8;
        } else {
#pragma unroll
          for (int i = 0; i < kS8Words; ++i) {
            uint32_t w = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const int64_t off = b0 + i * 4 + b;
              uint32_t byte = 0xFFu;  // beyond the row: code 11 = genotype 0, not counted
              if (off < row_bytes) {
                byte = (uint32_t)__ldg(row + off);
                if (off == full_bytes) byte |= (0xFFu << (2 * (int)(n_geno & 3))) & 0xFFu;  // pad bits of the last byte
              }
              w |= byte << (8 * b);
            }
            nxt[i] = w;
          }
          nxt[kS8Words] = 0u;
          sh_nxt = 0;
        }
      };
      int n00 = 0, n01 = 0, n10 = 0;
      // a 128-byte line of the row feeds 4 stages: pull lines into L2 well ahead (one lane per line)
      if (half == 0) {
#pragma unroll
        for (int l = 0; l < 4; ++l)
          if ((int64_t)l * 128 < row_bytes) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(row + l * 128));
      }
      load(0);
      for (int64_t k = 0; k < n_st; ++k) {
        uint32_t w[kS8Words];
#pragma unroll
        for (int i = 0; i < kS8Words; ++i) w[i] = __funnelshift_r(nxt[i], nxt[i + 1], sh_nxt);
        if (k + 1 < n_st) load(k + 1);
        if (half == 0 && (k & 3) == 0 && (k + 16) * 32 < row_bytes)
          asm volatile("prefetch.global.L2 [%0];\n" ::"l"(row + (k + 16) * 32));
        if (wrapped) mbar_wait(&sm->empty_w[s], ph ^ 1u);
        unsigned char* xrow = x_ring + (size_t)s * x_bytes + vloc * kS8K;
        unsigned char* erow = xrow + kS8TileV * kS8K;
#pragma unroll
        for (int i = 0; i < kS8Words; ++i) {
          const uint32_t lo = w[i] & 0x55555555u, hi = (w[i] >> 1) & 0x55555555u;
          n01 += __popc(lo & ~hi);
          n10 += __popc(hi & ~lo);
          n00 += __popc(~(lo | hi) & 0x55555555u);
          uint32_t xs[4], es[4];
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const uint32_t sel = s8_selector(__byte_perm(w[i], 0u, 0x4440u + b));
            xs[b] = __byte_perm(0x00010002u, 0u, sel);  // code 00 -> 2, 01 -> 0 (missing), 10 -> 1, 11 -> 0
            es[b] = __byte_perm(0x00000100u, 0u, sel);  // code 01 -> 1
          }
          const int chunk = (half * kS8Words + i) ^ (vloc & 7);  // 128-byte swizzle
          *reinterpret_cast<uint4*>(xrow + chunk * 16) = make_uint4(xs[0], xs[1], xs[2], xs[3]);
          *reinterpret_cast<uint4*>(erow + chunk * 16) = make_uint4(es[0], es[1], es[2], es[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic writes -> async proxy (MMA reads)
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm->full_x[s]);
        if (++s == DX) {
          s = 0;
          ph ^= 1u;
          wrapped = true;
        }
      }
      // per-variant code counts: the two lanes of a variant are neighbours
      n00 += __shfl_xor_sync(0xffffffffu, n00, 1);
      n01 += __shfl_xor_sync(0xffffffffu, n01, 1);
      n10 += __shfl_xor_sync(0xffffffffu, n10, 1);
      if (half == 0) {
        int32_t* c = sm->counts[tcount & 1][vloc];
        c[0] = n00;
        c[1] = n01;
        c[2] = n10;
      }
      asm volatile("bar.sync 1, 256;\n" ::: "memory");  // all decode warps: counts of this tile are in shared memory

      // ---- epilogue: 4 warps ----
      if (warp < kS8FirstDecodeWarp + 4) {
        const int q = warp - kS8FirstDecodeWarp;
        const int et = q * 32 + lane;  // 0..127: TMEM lane first, then variant of the tile
        int32_t* scratch = p.scratch + (int64_t)blockIdx.x * MB * kS8M * kS8N;
        mbar_wait(&sm->tmem_full, (uint32_t)(tcount & 1));
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        for (int mb = 0; mb < MB; ++mb) {
          int32_t* dst = scratch + ((int64_t)mb * kS8M + et) * kS8N;  // digit row `et` of block mb
          for (int c0 = 0; c0 < kS8N; c0 += 32) {
            uint32_t z[32];
            tmem_ld32(tmem + mb * kS8N + ((uint32_t)(q * 32) << 16) + c0, z);
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<uint4*>(dst + c0 + i) = make_uint4(z[i], z[i + 1], z[i + 2], z[i + 3]);
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm->tmem_empty);      // accumulators may be overwritten
        __threadfence_block();
        asm volatile("bar.sync 2, 128;\n" ::: "memory");  // the 4 epilogue warps: scratch is complete

        // thread `et` now owns variant `et` of the tile
        const int64_t vg = (int64_t)tile * kS8TileV + et;
        const int32_t* c = sm->counts[tcount & 1][et];
        const double c00 = c[0], c01 = c[1], c10 = c[2];
        double mu = -1.0;
        if (p.fused_counts) {
          if (p.missing_policy == GLR_MISSING_MEAN && c01 < (double)n_geno)
            mu = (2.0 * c00 + c10) / ((double)n_geno - c01);  // MeanSubstitute.scala:57-84
          if (vg < p.x.n_variants) {
            p.v1_out[vg] = mu;
            p.mom_out[vg] = 4.0 * c00 + c10 + (mu * mu) * c01;
            p.mom_out[p.ldS + vg] = 2.0 * c00 + c10 + mu * c01;
          }
        } else {
          mu = p.x.v1[min(vg, (int64_t)p.x.n_variants - 1)];
        }
        const volatile int32_t* zs = scratch;
        for (int col = 0; col < p.n_cols; ++col) {
          double acc = 0.0;
#pragma unroll
          for (int i = kS8Digits - 1; i >= 0; --i) {
            const int r = col * kS8Digits + i;
            const double zx = (double)zs[(int64_t)r * kS8N + et];
            const double ze = (double)zs[(int64_t)r * kS8N + kS8TileV + et];
            acc = fma(acc, 256.0, fma(mu, ze, zx));
          }
          p.S[(int64_t)col * p.ldS + vg] = acc * p.scale[col];
        }
        asm volatile("bar.sync 2, 128;\n" ::: "memory");  // scratch may be reused by the next tile
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512));
}

size_t split8_scratch_bytes(int MB, int sm_count) { return (size_t)sm_count * MB * kS8M * kS8N * sizeof(int32_t); }

void launch_split8(Split8Params p, int sm_count, cudaStream_t st) {
  if (p.x.n_variants <= 0) return;
  const int n_tiles = (p.x.n_variants + kS8TileV - 1) / kS8TileV;
  // 192 KB of operand buffers: MB = 1 -> 4 x 16 KB digits + 4 x 32 KB genotypes; MB = 2 -> 3 x 32 KB + 3 x 32 KB
  p.n_stage_bufs = p.MB == 1 ? 4 : 3;  // the two rings advance together (one completion barrier per stage)
  p.n_x_bufs = p.n_stage_bufs;
  const size_t smem = (size_t)p.n_stage_bufs * p.MB * kS8M * kS8K + (size_t)p.n_x_bufs * kS8N * kS8K + sizeof(Split8Smem) + 64;
  cudaFuncSetAttribute(split8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  split8_kernel<<<std::min(n_tiles, sm_count), kS8Threads, smem, st>>>(p);
}

// ------------------------------------------------------------------------------------------------
// packing: W (row-major [N x ld], columns c0..c0+ncols) -> per 128-sample stage a swizzled image
// [MB*128 digit rows][128 bytes]; row of (column c, digit i) = 7 c + i
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
                                   const int64_t* geno_to_row, int64_t Ng, int8_t* dst, int dst_col0, int MB,
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
  int8_t* base = dst + stage * (int64_t)MB * kS8M * kS8K;
#pragma unroll
  for (int i = 0; i < kS8Digits; ++i) {
    const long long d = ((q + 128) & 255) - 128;  // balanced digit in [-128, 127]
    q = (q - d) >> 8;
    const int r = (dst_col0 + c) * kS8Digits + i;
    base[(int64_t)r * kS8K + (((k >> 4) ^ (r & 7)) << 4) + (k & 15)] = (int8_t)d;
  }
}

void launch_pack_split8(const double* src, int64_t N, int64_t ld, int c0, int ncols, double* colmax_scratch,
                        const int64_t* geno_to_row, int64_t Ng, int8_t* dst, int dst_col0, int MB, int64_t n_stages,
                        cudaStream_t st) {
  if (ncols <= 0) return;
  colmax_kernel<<<ncols, 256, 0, st>>>(src, N, ld, c0, ncols, colmax_scratch);
  const int64_t total = (int64_t)ncols * n_stages * kS8K;
  pack_split8_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(src, N, ld, c0, ncols, colmax_scratch, geno_to_row,
                                                                      Ng, dst, dst_col0, MB, n_stages);
}

}  // namespace glr
