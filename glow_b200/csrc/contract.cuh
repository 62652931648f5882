// FP64 tensor-core contraction kernels with the genotype decode fused into the operand load.
//
//   gram_kernel   : S = W^T X              (pass 1; W = [Q | Ytilde | verbose columns], packed)
//   masked_kernel : D = Mu^T (X - Q A)^2   (pass 2; only for phenotypes with missing samples)
//
// Replaces np.column_stack + _residualize_in_place + `Y.T @ X` + the masked einsum of
// python/glow/gwas/lin_reg.py:226-241 (and functions.py:136-143, :78-82).
//
// CTA = 8 consumer warps + 1 producer warp.  The producer streams the packed, shared operand
// (W / Q / masks) through a 4-stage shared-memory ring with bulk async copies (TMA engine,
// cp.async.bulk + mbarrier complete_tx); consumers wait on the stage's "full" barrier, read their
// A fragments with conflict-free 16-byte LDS and release the stage on its "empty" barrier.
// The genotype operand never touches shared memory: each lane loads exactly the bytes of its own
// B fragments from global/L2 (2-bit: one 16-byte row segment per 64 samples) one unit ahead of use
// and decodes them in registers.
#pragma once
#include "../../include/glow_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace glr {

constexpr int kWarps = 8;
constexpr int kThreads = (kWarps + 1) * 32;
constexpr int kStages = 4;     // masked pass / large stages
constexpr int kMaxStages = 8;  // gram pass with small stages (MT <= 4): deeper ring, warps drift further apart
constexpr int kBarBytes = 128;  // full[4] + empty[4] mbarriers, padded

// ------------------------------------------------------------------------------------------------
// Genotype loaders: one per format.  A "unit" is UG sample groups (8 samples each) for all NT
// n-tiles of the warp.  load() fills the `nxt` registers, advance() makes them current, get()
// decodes the two values this lane owns in group g of the current unit (samples 2j, 2j+1 of the
// group).  Samples at or beyond n_geno decode to a finite value (0 for float formats): the packed
// shared operand is zero there, so they contribute nothing.
// ------------------------------------------------------------------------------------------------
template <int FMT, int NT>
struct XLoader;

__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr));
  return v;
}

// PLINK 2-bit.  Decode = table lookup: every variant of the warp has a 4-entry table
// {code 00 -> 2, 01 -> missing substitute, 10 -> 1, 11 -> 0} in shared memory
// (PlinkRowToInternalRowConverter.scala:40-47 + genotype_states + mean_substitute); a lane fetches its two
// values of a sample group with two 8-byte LDS.  Lanes 0-15 / 16-31 of an LDS.64 touch 4 variants
// x 4 entries = 16 distinct 8-byte bank pairs (or broadcast the same entry): always conflict free.
template <int NT>
struct XLoader<GLR_FMT_PLINK2, NT> {
  static constexpr int UG = 8;  // one 64-sample chunk = 16 bytes per variant
  static constexpr int kLutDoublesPerWarp = NT * 8 * 4;
  const uint8_t* row[NT];
  uint32_t lut[NT];  // shared-space byte address of this lane's variant table
  uint32_t cur[NT][5], nxt[NT][5];
  uint32_t sh_cur[NT], sh_nxt[NT];
  int64_t row_bytes;
  int nib_shift;

  __device__ __forceinline__ void init(const BlockView& x, int v0, int r, int j, double* lut_warp, int lane) {
    row_bytes = (x.n_geno + 3) >> 2;
    nib_shift = 4 * j;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int v = min(v0 + nt * 8 + r, x.n_variants - 1);
      row[nt] = x.base + (int64_t)v * x.stride;
      // lane 4*r' + c writes entry c of variant r' of this n-tile
      const int c = lane & 3;
      const int vv = min(v0 + nt * 8 + (lane >> 2), x.n_variants - 1);
      lut_warp[(nt * 8 + (lane >> 2)) * 4 + c] = dec((uint32_t)c, x.v1[vv]);
      lut[nt] = smem_u32(lut_warp + (nt * 8 + r) * 4);
    }
    __syncwarp();
  }
  __device__ __forceinline__ void load(int64_t s0) {
    const int64_t b0 = s0 >> 2;
    const bool fast = b0 + 19 <= row_bytes;  // every word touched lies inside the row
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const uint8_t* p = row[nt] + b0;
      if (fast) {
        const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3);
        const uint32_t* q = reinterpret_cast<const uint32_t*>(p - a);
#pragma unroll
        for (int i = 0; i < 4; ++i) nxt[nt][i] = __ldg(q + i);
        nxt[nt][4] = a ? __ldg(q + 4) : 0u;
        sh_nxt[nt] = a * 8;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t w = 0;
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int64_t off = b0 + i * 4 + b;
            // beyond the row: 0xFF = code 11 everywhere = 0.0
            const uint32_t byte = off < row_bytes ? (uint32_t)__ldg(row[nt] + off) : 0xFFu;
            w |= byte << (8 * b);
          }
          nxt[nt][i] = w;
        }
        nxt[nt][4] = 0u;
        sh_nxt[nt] = 0;
      }
    }
  }
  __device__ __forceinline__ void advance() {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int i = 0; i < 5; ++i) cur[nt][i] = nxt[nt][i];
      sh_cur[nt] = sh_nxt[nt];
    }
  }
  static __device__ __forceinline__ double dec(uint32_t code, double miss) {
    const int hi = (code == 0u) ? 0x40000000 : ((code == 2u) ? 0x3FF00000 : 0);
    const double v = __hiloint2double(hi, 0);
    return (code == 1u) ? miss : v;
  }
  __device__ __forceinline__ void get(int g, int nt, double& xa, double& xb) const {
    const int i = g >> 1;
    const uint32_t w = __funnelshift_r(cur[nt][i], cur[nt][i + 1], sh_cur[nt]);
    const uint32_t nib = w >> (16 * (g & 1) + nib_shift);
    xa = lds_f64(lut[nt] + ((nib & 3u) << 3));
    xb = lds_f64(lut[nt] + ((nib & 12u) << 1));
  }
};

// float64 / float32 arrays
template <typename T, int NT>
struct XLoaderFloat {
  static constexpr int UG = 4;
  using T2 = typename std::conditional<sizeof(T) == 8, double2, float2>::type;
  const T* row[NT];
  T2 cur[NT][UG], nxt[NT][UG];
  int64_t n;
  int joff;
  bool vec;

  static constexpr int kLutDoublesPerWarp = 0;
  __device__ __forceinline__ void init(const BlockView& x, int v0, int r, int j, double*, int) {
    n = x.n_geno;
    joff = 2 * j;
    vec = ((reinterpret_cast<uintptr_t>(x.base) | (uintptr_t)x.stride) & (2 * sizeof(T) - 1)) == 0;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      int v = min(v0 + nt * 8 + r, x.n_variants - 1);
      row[nt] = reinterpret_cast<const T*>(x.base + (int64_t)v * x.stride);
    }
  }
  __device__ __forceinline__ void load(int64_t s0) {
    const bool fast = s0 + UG * 8 <= n;
    // pull the lines two more units ahead into L2 (no registers held): one lane per 128-byte line
    if (joff == 0) {
      const int64_t sp = s0 + 2 * UG * 8;
      constexpr int kPerLine = 128 / (int)sizeof(T);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < UG * 8; e += kPerLine)
          if (sp + e < n) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(row[nt] + sp + e));
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int g = 0; g < UG; ++g) {
        const int64_t s = s0 + 8 * g + joff;
        const T* p = row[nt] + s;
        T2 v;
        if (fast && vec) {
          v = __ldg(reinterpret_cast<const T2*>(p));
        } else if (fast) {
          v.x = __ldg(p);
          v.y = __ldg(p + 1);
        } else {
          v.x = s < n ? __ldg(p) : T(0);
          v.y = s + 1 < n ? __ldg(p + 1) : T(0);
        }
        nxt[nt][g] = v;
      }
    }
  }
  __device__ __forceinline__ void advance() {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int g = 0; g < UG; ++g) cur[nt][g] = nxt[nt][g];
  }
  __device__ __forceinline__ void get(int g, int nt, double& xa, double& xb) const {
    xa = (double)cur[nt][g].x;
    xb = (double)cur[nt][g].y;
  }
};
// float64 / float32 arrays for the first pass: same fragment order, but the prefetch goes through a per-warp ring in
// shared memory filled with cp.async (16 / 8 bytes per lane, zero-filled past the end of the row) instead of
// registers: kDepth - 1 units of 32 samples are in flight per warp with no register cost, which is what it takes
// to keep HBM busy with 9 warps per SM.  Every lane copies exactly the bytes it later reads, so completion is a
// per-thread cp.async.wait_group.  Rows that are not 16-byte (8-byte for float) aligned use the register path.
template <typename T, int NT, int DEPTH>
struct XLoaderFloatAsync {
  static constexpr int UG = 4;
  static constexpr int kDepth = DEPTH;
  using T2 = typename std::conditional<sizeof(T) == 8, double2, float2>::type;
  static constexpr int kSlotT2 = NT * UG * 32;                                   // T2 elements per unit
  static constexpr int kLutDoublesPerWarp = kDepth * kSlotT2 * (int)sizeof(T2) / 8;
  XLoaderFloat<T, NT> fb;  // register path (unaligned rows)
  const T* row[NT];
  T2 cur[NT][UG];
  int64_t n;
  int joff, lane, slot_w, slot_r;
  uint32_t ring;  // shared-memory address of this warp's ring
  bool vec, primed;

  __device__ __forceinline__ void init(const BlockView& x, int v0, int r, int j, double* warp_smem, int lane_) {
    n = x.n_geno;
    joff = 2 * j;
    lane = lane_;
    vec = ((reinterpret_cast<uintptr_t>(x.base) | (uintptr_t)x.stride) & (2 * sizeof(T) - 1)) == 0;
    ring = smem_u32(warp_smem);
    slot_w = slot_r = 0;
    primed = false;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      int v = min(v0 + nt * 8 + r, x.n_variants - 1);
      row[nt] = reinterpret_cast<const T*>(x.base + (int64_t)v * x.stride);
    }
    if (!vec) fb.init(x, v0, r, j, warp_smem, lane_);
  }
  __device__ __forceinline__ void issue(int64_t s0) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int g = 0; g < UG; ++g) {
        const int64_t s = s0 + 8 * g + joff;
        const int64_t left = n - s;
        const uint32_t bytes = left >= 2 ? 2u * (uint32_t)sizeof(T) : (left == 1 ? (uint32_t)sizeof(T) : 0u);
        const T* src = row[nt] + (left > 0 ? s : 0);
        const uint32_t dst = ring + (uint32_t)(((slot_w * NT + nt) * UG + g) * 32 + lane) * (uint32_t)sizeof(T2);
        if (sizeof(T) == 8)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
        else
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
      }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    slot_w = slot_w + 1 == kDepth ? 0 : slot_w + 1;
  }
  // load(s0): the unit starting at sample s0 must be available at the next advance()
  __device__ __forceinline__ void load(int64_t s0) {
    if (!vec) {
      fb.load(s0);
      return;
    }
    if (!primed) {
#pragma unroll
      for (int d = 0; d < kDepth - 2; ++d) issue(s0 + d * (UG * 8));
      primed = true;
    }
    issue(s0 + (kDepth - 2) * (UG * 8));  // kDepth - 1 groups are now in flight
  }
  __device__ __forceinline__ void advance() {
    if (!vec) {
      fb.advance();
      return;
    }
    asm volatile("cp.async.wait_group %0;\n" ::"n"(kDepth - 2) : "memory");  // the oldest group has landed
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int g = 0; g < UG; ++g) {
        const uint32_t a = ring + (uint32_t)(((slot_r * NT + nt) * UG + g) * 32 + lane) * (uint32_t)sizeof(T2);
        if (sizeof(T) == 8) {
          double2 v;
          asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(a));
          cur[nt][g].x = v.x;
          cur[nt][g].y = v.y;
        } else {
          float2 v;
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];\n" : "=f"(v.x), "=f"(v.y) : "r"(a));
          cur[nt][g].x = v.x;
          cur[nt][g].y = v.y;
        }
      }
    slot_r = slot_r + 1 == kDepth ? 0 : slot_r + 1;
  }
  __device__ __forceinline__ void get(int g, int nt, double& xa, double& xb) const {
    if (!vec) {
      fb.get(g, nt, xa, xb);
      return;
    }
    xa = (double)cur[nt][g].x;
    xb = (double)cur[nt][g].y;
  }
};

// loader of the first pass: the cp.async ring for float formats, the plain loaders otherwise.  The ring depth is what
// fits next to the W ring (8 warps x 4 KB per unit for f64; the W ring is 8 stages deep up to 4 m-tiles, 4 beyond).
// (16 consumer warps -- float blocks with few columns -- have half the ring per warp: 3 units of 32 samples in flight)
template <int FMT, int NT, int MT, int WARPS = kWarps>
struct GramLoader : XLoader<FMT, NT> {};
template <int NT, int MT, int WARPS>
struct GramLoader<GLR_FMT_F64, NT, MT, WARPS> : XLoaderFloatAsync<double, NT, (WARPS > 8 ? 3 : (MT <= 2 ? 4 : (MT == 4 ? 2 : 3)))> {};
template <int NT, int MT, int WARPS>
struct GramLoader<GLR_FMT_F32, NT, MT, WARPS> : XLoaderFloatAsync<float, NT, (WARPS > 8 ? 3 : (MT <= 2 ? 4 : (MT == 4 ? 2 : 3)))> {};

template <int NT>
struct XLoader<GLR_FMT_F64, NT> : XLoaderFloat<double, NT> {};
template <int NT>
struct XLoader<GLR_FMT_F32, NT> : XLoaderFloat<float, NT> {};

// int8 genotype states, -1 = missing
template <int NT>
struct XLoader<GLR_FMT_I8, NT> {
  static constexpr int UG = 8;
  const int8_t* row[NT];
  double v1[NT];
  int cur[NT][UG], nxt[NT][UG];  // two int8 packed in the low 16 bits
  int64_t n;
  int joff;

  static constexpr int kLutDoublesPerWarp = 0;
  __device__ __forceinline__ void init(const BlockView& x, int v0, int r, int j, double*, int) {
    n = x.n_geno;
    joff = 2 * j;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      int v = min(v0 + nt * 8 + r, x.n_variants - 1);
      row[nt] = reinterpret_cast<const int8_t*>(x.base + (int64_t)v * x.stride);
      v1[nt] = x.v1[v];
    }
  }
  __device__ __forceinline__ void load(int64_t s0) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int g = 0; g < UG; ++g) {
        const int64_t s = s0 + 8 * g + joff;
        const int a = s < n ? (int)__ldg(row[nt] + s) : 0;
        const int b = s + 1 < n ? (int)__ldg(row[nt] + s + 1) : 0;
        nxt[nt][g] = (a & 0xFF) | ((b & 0xFF) << 8);
      }
    }
  }
  __device__ __forceinline__ void advance() {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int g = 0; g < UG; ++g) cur[nt][g] = nxt[nt][g];
  }
  __device__ __forceinline__ void get(int g, int nt, double& xa, double& xb) const {
    const int a = (int)(int8_t)(cur[nt][g] & 0xFF);
    const int b = (int)(int8_t)((cur[nt][g] >> 8) & 0xFF);
    xa = a == -1 ? v1[nt] : (double)a;
    xb = b == -1 ? v1[nt] : (double)b;
  }
};

// ------------------------------------------------------------------------------------------------
// shared-memory ring
// ------------------------------------------------------------------------------------------------
// Release of a shared-memory stage that the consumers read with ordinary loads (generic proxy) and the producer refills
// with cp.async.bulk (async proxy): the reads must be ordered before the refill ACROSS the proxies, which the mbarrier
// hand-off alone does not do -- fence.proxy.async before the arrive.  Without it a refill occasionally overtook the last
// reads of a stage: wrong values in single rows of S for one warp's 16 variants, a few times per 10^4 warps, seen only
// with short-lived CTAs of the float64 kernels (found in round 2 by the permutation test of configs[1]).
__device__ __forceinline__ void stage_release_fence() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

struct Ring {
  uint64_t* full;
  uint64_t* empty;
  __device__ __forceinline__ void init(unsigned char* smem, int tid, int n_consumer_warps, int n_stages = kStages) {
    full = reinterpret_cast<uint64_t*>(smem);
    empty = full + kMaxStages;
    if (tid == 0) {
      for (int s = 0; s < n_stages; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], n_consumer_warps);
      }
      mbar_fence_init();
    }
    __syncthreads();
  }
};

// ------------------------------------------------------------------------------------------------
// pass 1
// ------------------------------------------------------------------------------------------------
// TAIL = 0..4 extra columns beyond the MT m-tiles, accumulated with plain FP64 FMAs on the lane's
// own two samples (then reduced over the 4 lanes that share a variant): a column count of 8*MT + 1..3
// does not pay for a whole zero-padded m-tile of DMMAs.
template <int FMT, int MT, int NT, int TAIL, int WARPS>
__global__ void __launch_bounds__((WARPS + 1) * 32, 1) gram_kernel(const GramParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int kStageDoubles = kChunkGroups * (MT * kAtomDoubles + TAIL * kGroup);
  constexpr uint32_t kStageBytes = kStageDoubles * 8;
  using Loader = GramLoader<FMT, NT, MT, WARPS>;
  constexpr int UG = Loader::UG;
  constexpr int UPC = kChunkGroups / UG;  // units per chunk

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  Ring ring;
  constexpr int kStagesG = MT <= 4 ? kMaxStages : kStages;
  ring.init(smem, tid, WARPS, kStagesG);
  double* wst = reinterpret_cast<double*>(smem + kBarBytes);

  const int64_t cps = ceil_div(p.n_chunks, p.n_split);
  const int64_t c_begin = (int64_t)blockIdx.z * cps;
  const int64_t c_end = min(p.n_chunks, c_begin + cps);
  const int64_t n_it = c_end > c_begin ? c_end - c_begin : 0;

  if (warp == WARPS) {  // producer warp: one lane drives the TMA engine
    if (lane == 0) {
      const double* src = p.wpk + ((int64_t)blockIdx.y * p.n_chunks + c_begin) * kStageDoubles;
      for (int64_t it = 0; it < n_it; ++it) {
        const int s = (int)(it % kStagesG);
        if (it >= kStagesG) mbar_wait_backoff(&ring.empty[s], (uint32_t)((it / kStagesG) & 1) ^ 1u);
        mbar_expect_tx(&ring.full[s], kStageBytes);
        bulk_g2s(wst + (int64_t)s * kStageDoubles, src + it * kStageDoubles, kStageBytes, &ring.full[s]);
      }
    }
    return;
  }

  const int r = lane >> 2, j = lane & 3;
  const int v_warp = blockIdx.x * (WARPS * NT * 8) + warp * (NT * 8);
  Loader ld;
  ld.init(p.x, v_warp, r, j, wst + (int64_t)kStagesG * kStageDoubles + warp * Loader::kLutDoublesPerWarp, lane);

  double acc[MT][NT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
  double sq[NT], sx[NT];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) sq[nt] = sx[nt] = 0.0;
  double tacc[TAIL > 0 ? TAIL : 1][NT];
#pragma unroll
  for (int t = 0; t < (TAIL > 0 ? TAIL : 1); ++t)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) tacc[t][nt] = 0.0;
  const bool want_sq = (FMT == GLR_FMT_F64 || FMT == GLR_FMT_F32) && blockIdx.y == 0;

  const int64_t t_total = n_it * UPC;
  const int64_t s_begin = c_begin * kChunk;
  if (t_total > 0) ld.load(s_begin);
  for (int64_t it = 0; it < n_it; ++it) {
    const int s = (int)(it % kStagesG);
    mbar_wait(&ring.full[s], (uint32_t)((it / kStagesG) & 1));
    const double* st = wst + (int64_t)s * kStageDoubles;
    // A fragments (and tail weights) are fetched one sample group ahead of their DMMAs
    double2 a_nxt[MT];
    double2 w_nxt[TAIL > 0 ? TAIL : 1];
    {
      const double2* ap = reinterpret_cast<const double2*>(st) + lane;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) a_nxt[mt] = ap[mt * 32];
      if (TAIL > 0) {
        const double2* tp = reinterpret_cast<const double2*>(st + kChunkGroups * MT * kAtomDoubles) + j;
#pragma unroll
        for (int t = 0; t < TAIL; ++t) w_nxt[t] = tp[t * 4];
      }
    }
#pragma unroll
    for (int u = 0; u < UPC; ++u) {
      ld.advance();
      const int64_t t = it * UPC + u + 1;
      if (t < t_total) ld.load(s_begin + t * (UG * kGroup));
      // genotype values are decoded one sample group ahead of their DMMAs as well
      double xa_n[NT], xb_n[NT];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) ld.get(0, nt, xa_n[nt], xb_n[nt]);
#pragma unroll
      for (int g = 0; g < UG; ++g) {
        const int gg = u * UG + g;
        double2 a[MT];
        double2 w[TAIL > 0 ? TAIL : 1];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) a[mt] = a_nxt[mt];
#pragma unroll
        for (int t2 = 0; t2 < (TAIL > 0 ? TAIL : 1); ++t2) w[t2] = w_nxt[t2];
        double xa[NT], xb[NT];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          xa[nt] = xa_n[nt];
          xb[nt] = xb_n[nt];
        }
        if (gg + 1 < kChunkGroups) {
          const double2* ap = reinterpret_cast<const double2*>(st + (int64_t)(gg + 1) * MT * kAtomDoubles) + lane;
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) a_nxt[mt] = ap[mt * 32];
          if (TAIL > 0) {
            const double2* tp =
                reinterpret_cast<const double2*>(st + kChunkGroups * MT * kAtomDoubles + (gg + 1) * TAIL * kGroup) + j;
#pragma unroll
            for (int t2 = 0; t2 < TAIL; ++t2) w_nxt[t2] = tp[t2 * 4];
          }
        }
        if (g + 1 < UG) {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) ld.get(g + 1, nt, xa_n[nt], xb_n[nt]);
        }
        if (FMT == GLR_FMT_F64 || FMT == GLR_FMT_F32) {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            sq[nt] = fma(xa[nt], xa[nt], fma(xb[nt], xb[nt], sq[nt]));
            sx[nt] += xa[nt] + xb[nt];
          }
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt].x, xa[nt]);
        if (TAIL > 0) {
#pragma unroll
          for (int t2 = 0; t2 < TAIL; ++t2)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) tacc[t2][nt] = fma(w[t2].x, xa[nt], fma(w[t2].y, xb[nt], tacc[t2][nt]));
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt].y, xb[nt]);
      }
    }
    stage_release_fence();
    __syncwarp();
    if (lane == 0) mbar_arrive(&ring.empty[s]);
  }

  // (no thread leaves with cp.async copies in flight: the loaders run a few units ahead of use)
  asm volatile("cp.async.wait_all;\n" ::: "memory");
  // C fragment: lane (r, j) holds rows r of each m-tile, variants 2j, 2j+1 of each n-tile
  double* Sz = p.S + (int64_t)blockIdx.z * p.s_slab;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int64_t row = ((int64_t)blockIdx.y * MT + mt) * 8 + r;
      const int64_t col = v_warp + nt * 8 + 2 * j;
      *reinterpret_cast<double2*>(Sz + row * p.ldS + col) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
    }
  if (TAIL > 0) {
#pragma unroll
    for (int t = 0; t < TAIL; ++t)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        double v = tacc[t][nt];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (j == 0) Sz[((int64_t)MT * 8 + t) * p.ldS + v_warp + nt * 8 + r] = v;
      }
  }
  if (want_sq) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      double v = sq[nt], u = sx[nt];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      u += __shfl_xor_sync(0xffffffffu, u, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      u += __shfl_xor_sync(0xffffffffu, u, 2);
      if (j == 0) {
        p.mom_part[((int64_t)blockIdx.z * 2 + 0) * p.ldS + v_warp + nt * 8 + r] = v;
        p.mom_part[((int64_t)blockIdx.z * 2 + 1) * p.ldS + v_warp + nt * 8 + r] = u;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// pass 2: residualise in registers with DMMA, square, contract with the 0/1 masks
// ------------------------------------------------------------------------------------------------
constexpr int kMaskedNT = 2;
constexpr int kMaxKT = 8;  // K <= 32 covariates in the masked pass

template <int FMT, int MT>
__global__ void __launch_bounds__(kThreads, 1) masked_kernel(const MaskedParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int NT = kMaskedNT;
  using Loader = XLoader<FMT, NT>;
  constexpr int UG = Loader::UG;
  constexpr int UPC = kChunkGroups / UG;

  const int KT2 = p.KT2;
  const int mStageDoubles = kChunkGroups * MT * kAtomDoubles;
  const int qStageDoubles = kChunkGroups * KT2 * kAtomDoubles;
  const int stageDoubles = mStageDoubles + qStageDoubles;
  const uint32_t stageBytes = (uint32_t)stageDoubles * 8;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  Ring ring;
  ring.init(smem, tid, kWarps);
  double* stage0 = reinterpret_cast<double*>(smem + kBarBytes);
  double* negA_s = stage0 + (int64_t)kStages * stageDoubles;  // [kWarps][NT][2*KT2][32]

  const int64_t cps = ceil_div(p.n_chunks, p.n_split);
  const int64_t c_begin = (int64_t)blockIdx.z * cps;
  const int64_t c_end = min(p.n_chunks, c_begin + cps);
  const int64_t n_it = c_end > c_begin ? c_end - c_begin : 0;

  if (warp == kWarps) {
    if (lane == 0) {
      const double* msrc = p.mpk + ((int64_t)blockIdx.y * p.n_chunks + c_begin) * mStageDoubles;
      const double* qsrc = p.qb + c_begin * qStageDoubles;
      for (int64_t it = 0; it < n_it; ++it) {
        const int s = (int)(it % kStages);
        if (it >= kStages) mbar_wait_backoff(&ring.empty[s], (uint32_t)((it / kStages) & 1) ^ 1u);
        mbar_expect_tx(&ring.full[s], stageBytes);
        double* dst = stage0 + (int64_t)s * stageDoubles;
        bulk_g2s(dst, msrc + it * mStageDoubles, (uint32_t)mStageDoubles * 8, &ring.full[s]);
        bulk_g2s(dst + mStageDoubles, qsrc + it * qStageDoubles, (uint32_t)qStageDoubles * 8, &ring.full[s]);
      }
    }
    return;
  }

  const int r = lane >> 2, j = lane & 3;
  const int v_warp = blockIdx.x * (kWarps * NT * 8) + warp * (NT * 8);
  Loader ld;
  ld.init(p.x, v_warp, r, j, negA_s + (int64_t)kWarps * NT * 2 * KT2 * 32 + warp * Loader::kLutDoublesPerWarp, lane);

  // A operand of the residual product: lane (r, j) holds -a[4*kk + j][variant r] for k-step kk
  double* negA = negA_s + (int64_t)warp * NT * 2 * KT2 * 32;
  for (int nt = 0; nt < NT; ++nt)
    for (int kk = 0; kk < 2 * KT2; ++kk) {
      const int k = 4 * kk + j;
      const int64_t v = v_warp + nt * 8 + r;
      negA[(nt * 2 * KT2 + kk) * 32 + lane] = k < p.K ? -p.A[(int64_t)k * p.ldS + v] : 0.0;
    }
  __syncwarp();

  double acc[MT][NT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

  const int64_t t_total = n_it * UPC;
  const int64_t s_begin = c_begin * kChunk;
  if (t_total > 0) ld.load(s_begin);
  for (int64_t it = 0; it < n_it; ++it) {
    const int s = (int)(it % kStages);
    mbar_wait(&ring.full[s], (uint32_t)((it / kStages) & 1));
    const double* mst = stage0 + (int64_t)s * stageDoubles;
    const double* qst = mst + mStageDoubles;
#pragma unroll
    for (int u = 0; u < UPC; ++u) {
      ld.advance();
      const int64_t t = it * UPC + u + 1;
      if (t < t_total) ld.load(s_begin + t * (UG * kGroup));
#pragma unroll
      for (int g = 0; g < UG; ++g) {
        const int gg = u * UG + g;
        double ra[NT], rb[NT];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) ld.get(g, nt, ra[nt], rb[nt]);
        // r = x - sum_k Q[s][k] a[k][v]: C[variant][sample] += (-a)^T[variant][k] * Q^T[k][sample]
        const double2* qp = reinterpret_cast<const double2*>(qst + (int64_t)gg * KT2 * kAtomDoubles) + lane;
        for (int k2 = 0; k2 < KT2; ++k2) {
          const double2 qv = qp[k2 * 32];
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const double a0 = negA[(nt * 2 * KT2 + 2 * k2) * 32 + lane];
            const double a1 = negA[(nt * 2 * KT2 + 2 * k2 + 1) * 32 + lane];
            dmma884(ra[nt], rb[nt], a0, qv.x);
            dmma884(ra[nt], rb[nt], a1, qv.y);
          }
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          ra[nt] *= ra[nt];
          rb[nt] *= rb[nt];
        }
        const double2* mp = reinterpret_cast<const double2*>(mst + (int64_t)gg * MT * kAtomDoubles) + lane;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const double2 m = mp[mt * 32];
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            dmma884(acc[mt][nt][0], acc[mt][nt][1], m.x, ra[nt]);
            dmma884(acc[mt][nt][0], acc[mt][nt][1], m.y, rb[nt]);
          }
        }
      }
    }
    stage_release_fence();
    __syncwarp();
    if (lane == 0) mbar_arrive(&ring.empty[s]);
  }

  const int64_t rows = (int64_t)p.n_groups * MT * 8;
  double* Dz = p.D + (int64_t)blockIdx.z * rows * p.ldS;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int64_t row = ((int64_t)blockIdx.y * MT + mt) * 8 + r;
      const int64_t col = v_warp + nt * 8 + 2 * j;
      *reinterpret_cast<double2*>(Dz + row * p.ldS + col) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
    }
}

// ------------------------------------------------------------------------------------------------
// per-format dispatch (instantiated in gram_<fmt>.cu so the formats compile in parallel)
// ------------------------------------------------------------------------------------------------
template <int FMT, int MT, int NT, int TAIL, int WARPS = kWarps>
void launch_gram_inst(const GramParams& p, cudaStream_t st) {
  const int tile = WARPS * NT * 8;
  dim3 grid((unsigned)ceil_div(p.x.n_variants, tile), (unsigned)p.n_groups, (unsigned)p.n_split);
  const size_t smem = kBarBytes + (size_t)(MT <= 4 ? kMaxStages : kStages) * kChunkGroups * (MT * kAtomDoubles + TAIL * kGroup) * 8 +
                      (size_t)WARPS * GramLoader<FMT, NT, MT, WARPS>::kLutDoublesPerWarp * 8;
  auto kern = gram_kernel<FMT, MT, NT, TAIL, WARPS>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<grid, (WARPS + 1) * 32, smem, st>>>(p);
}

// tail columns are instantiated for MT <= 4 only (TAILS = false: formats that never use them)
template <int FMT, int MT, int NT, bool TAILS, int WARPS>
void launch_gram_tail(const GramParams& p, cudaStream_t st) {
  if constexpr (TAILS && MT <= 4) {
    switch (p.tail) {
      case 1: launch_gram_inst<FMT, MT, NT, 1, WARPS>(p, st); return;
      case 2: launch_gram_inst<FMT, MT, NT, 2, WARPS>(p, st); return;
      case 3: launch_gram_inst<FMT, MT, NT, 3, WARPS>(p, st); return;
      case 4: launch_gram_inst<FMT, MT, NT, 4, WARPS>(p, st); return;
      default: break;
    }
  }
  launch_gram_inst<FMT, MT, NT, 0, WARPS>(p, st);
}

// MT_LO..MT_HI: which m-tile counts this (NT, WARPS) shape is instantiated for
template <int FMT, int NT, bool TAILS, int WARPS = kWarps, int MT_LO = 1, int MT_HI = 8>
void launch_gram_mt(const GramParams& p, cudaStream_t st) {
#define GLR_MT_CASE(M)                                                                         \
  case M:                                                                                      \
    if constexpr (M >= MT_LO && M <= MT_HI) launch_gram_tail<FMT, M, NT, TAILS, WARPS>(p, st); \
    break;
  switch (p.MT) {
    GLR_MT_CASE(1)
    GLR_MT_CASE(2)
    GLR_MT_CASE(3)
    GLR_MT_CASE(4)
    GLR_MT_CASE(5)
    GLR_MT_CASE(6)
    GLR_MT_CASE(7)
    GLR_MT_CASE(8)
    default: break;
  }
#undef GLR_MT_CASE
}

template <int FMT, int MT>
void launch_masked_inst(const MaskedParams& p, cudaStream_t st) {
  const int tile = kWarps * kMaskedNT * 8;
  dim3 grid((unsigned)ceil_div(p.x.n_variants, tile), (unsigned)p.n_groups, (unsigned)p.n_split);
  const size_t stage = (size_t)kChunkGroups * (MT + p.KT2) * kAtomDoubles * 8;
  const size_t smem = kBarBytes + kStages * stage + (size_t)kWarps * kMaskedNT * 2 * p.KT2 * 32 * 8 +
                      (size_t)kWarps * XLoader<FMT, kMaskedNT>::kLutDoublesPerWarp * 8;
  auto kern = masked_kernel<FMT, MT>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<grid, kThreads, smem, st>>>(p);
}

template <int FMT>
void launch_masked_mt(const MaskedParams& p, cudaStream_t st) {
  switch (p.MT) {
    case 1: launch_masked_inst<FMT, 1>(p, st); break;
    case 2: launch_masked_inst<FMT, 2>(p, st); break;
    case 4: launch_masked_inst<FMT, 4>(p, st); break;
    default: launch_masked_inst<FMT, 8>(p, st); break;
  }
}

// defined one per format translation unit
void launch_gram_f64(const GramParams& p, cudaStream_t st);
void launch_gram_f32(const GramParams& p, cudaStream_t st);
void launch_gram_i8(const GramParams& p, cudaStream_t st);
void launch_gram_p2(const GramParams& p, cudaStream_t st);
void launch_masked_f64(const MaskedParams& p, cudaStream_t st);
void launch_masked_f32(const MaskedParams& p, cudaStream_t st);
void launch_masked_i8(const MaskedParams& p, cudaStream_t st);
void launch_masked_p2(const MaskedParams& p, cudaStream_t st);

}  // namespace glr
