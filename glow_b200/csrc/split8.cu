// Exact FP64-accurate contraction S = W^T X on the 5th-generation tensor cores (tcgen05.mma kind::i8,
// accumulators AND the genotype operand in TMEM) for PLINK 2-bit genotypes, with the genotype counts
// (mean imputation, sum x, sum x^2) fused into the same pass over the packed rows.
//
// Blackwell has no tcgen05 FP64 kind and DMMA tops out at 37 TFLOP/s, but 2-bit genotypes are small
// integers, so the product can be computed EXACTLY with integer tensor-core MMAs (Ozaki-style
// error-free splitting of the FP64 operand):
//   * every column of W = [Q(:,1:) | Ytilde] is scaled by its max |w| and rounded to a signed 55-bit
//     integer (quantum 2^-54 of the column maximum, i.e. an operand error of about one FP64 ulp of
//     the largest entries), then written as 7 balanced radix-256 digits d_i in [-128, 127]  (int8);
//   * a genotype row is x = xi + mu * e with xi in {0,1,2} (missing -> 0) and e the missing-call
//     indicator, both int8;
//   * Zx_i = d_i^T xi and Ze_i = d_i^T e are int32 dot products (|.| <= 128 * 2 * N < 2^31 for
//     N < 8.3e6): exact on the tensor cores;
//   * S[c][v] = max|w_c| * 2^-54 * sum_i 256^i (Zx_i + mu_v Ze_i), assembled in FP64 in the epilogue.
// The only roundings are the 2^-54 quantisation of W relative to its column maximum (finer than the
// FP64 mantissa for the entries that dominate the sums; the accumulation itself is exact, unlike a
// floating-point GEMM) and the final FP64 Horner sum.
//
// MMA shape.  The GENOTYPES are the A operand and live in tensor memory: TMEM lane = variant of a
// 128-variant tile, 8 columns (32 packed int8) per K = 32 samples, written straight from the decode
// registers with tcgen05.st.  The digit rows are the B operand (N = NC <= 128 rows: up to 18 columns
// x 5-7 digits + a row of ones), K-major in shared memory with the 128-byte swizzle, streamed by bulk
// copies.  Two accumulators per tile, Dx = xi * digits^T and De = e * digits^T ([128 variants x NC],
// int32) sit in TMEM columns [0, NC) and [128, 128 + NC); the A ring uses columns [256, 512).
// Compared with keeping the genotypes in shared memory (B operand) this removes the decode's shared
// memory stores and a third of the MMA's operand reads, which is what bounded that layout; and each
// TMEM lane ends up holding every digit sum of ONE variant, so the epilogue recombines in registers.
// The row of ones yields sum xi and the number of missing calls per variant for free; the count of
// homozygous-alternate codes (needed for sum x^2) is a popcount in the decode.
// States with more than 18 columns are processed in column blocks of up to 18 (the tile's rows are decoded once
// per block; the kernel stays MMA-bound).
//
// CTA pairs.  By default two CTAs of a cluster (the two SMs of a TPC) work as a pair on 256 variants with
// tcgen05.mma.cta_group::2 (M = 256): each CTA decodes its own 128 variants into its own TMEM, but holds only HALF of
// the digit rows of a stage in its shared memory — the tensor cores exchange the halves — so the digit stream costs half
// the L2 and shared-memory traffic.  The leader (rank 0) issues the MMAs; the barriers it waits on collect arrivals from
// both CTAs (remote mbarrier arrives), and tcgen05.commit multicasts the completions to both.
//
// CTA roles (S8Shape: 608 threads = 19 warps, 1 CTA/SM, persistent over (128-variant tile, column block[, sample range])
// work items):
//   warps 0-1 : issue tcgen05.mma, stage g by warp g % 2 (converged warps, one elected lane); one tcgen05.commit per
//               stage frees the stage's buffers, a second one after the warp's last stage of the item publishes the
//               accumulators (in the non-leader CTA of a pair they forward "my half of the digit stage has landed"
//               instead)
//   warps 2-17: decode, four groups of four warps.  Warp w owns TMEM lanes 32 (w % 4) ..; group g % 4 takes stage g; a
//               thread turns 32 bytes of its variant's row (128 samples) into 32 + 32 TMEM words per stage.  One group's
//               stage is a dependent chain of ~1 200 cycles (row segment -> decode -> TMEM slot -> tcgen05.st -> arrive):
//               the groups overlap their chains.  All decode warps run the epilogue (tcgen05.ld of the variant's digit
//               sums, FP64 recombination, coalesced store of S).
//   warp 18   : TMEM allocation / deallocation; lane 0 streams the packed digit operand (pre-swizzled image, one bulk
//               copy per stage)
// Digits.  The first n_lo columns of a single-block image may carry fewer than 7 digits (Split8Params.n_lo / lo_digits);
// the caller then certifies every result (xx_full_kernel in aux.cu) and re-runs a block on the 7-digit image when a bound
// fails.  Split8Params.x_lut replaces the call values 2 / 1 by other small integers (4 / 1: the squared calls of the
// logistic score test).
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "../../include/glow_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace glr {

constexpr int kS8K = 128;      // samples per stage = bytes per digit row per stage
constexpr int kS8TileV = 128;  // variants per tile = TMEM lanes
constexpr int kS8Digits = 7;
// Warp roles of a CTA: kIssuers MMA issuer warps (stage g is issued by warp g % kIssuers), then kGroups decode groups of
// 4 warps (one warp per TMEM lane quarter; group g % kGroups decodes stage g), then the producer warp (digit operand,
// TMEM allocation).  Registers per thread follow from the warps per SM sub-partition (16 K registers each): 17-20 warps
// -> 96, 21-24 -> 80.  The form with the missing-call operand keeps 64 decoded words live across the TMEM slot wait and
// needs the 96; the optimistic form needs half of that and takes more groups instead.  Measured (37 888 x 500 k, short
// runs): 2 groups / 4 issuers 2.57 | 2.35 ms (full | optimistic form), 3 / 4: 2.37 | 2.11, 4 / 2: 2.26 | 1.74; 5 or 6
// groups in the optimistic form: within 4 % of 4.  Moving registers from the issuers to the decode warps with setmaxnreg
// (4 groups / 4 issuers, 21 warps, 88 registers for the decode) was 4 % slower than 4 / 2 at 96.
#ifndef GLR_S8_GROUPS
#define GLR_S8_GROUPS 4
#endif
#ifndef GLR_S8_ISSUERS
#define GLR_S8_ISSUERS 2
#endif
#ifndef GLR_S8_GROUPS_NOE
#define GLR_S8_GROUPS_NOE GLR_S8_GROUPS
#endif
#ifndef GLR_S8_ISSUERS_NOE
#define GLR_S8_ISSUERS_NOE GLR_S8_ISSUERS
#endif
template <bool kNoE>
struct S8Shape {
  static constexpr int kGroups = kNoE ? GLR_S8_GROUPS_NOE : GLR_S8_GROUPS;
  static constexpr int kIssuers = kNoE ? GLR_S8_ISSUERS_NOE : GLR_S8_ISSUERS;  // power of two
  static constexpr int kThreads = (kIssuers + 4 * kGroups + 1) * 32;
};
constexpr int kS8MaxGroups = 8;
constexpr int kS8MaxIssuers = 4;  // (split8_plan: every sample range has at least this many stages)
constexpr int kS8Words = 8;  // 32-bit words (16 samples each) per decode thread per stage
constexpr int kS8WRing = 8;  // digit-operand stages in shared memory (also the depth of the completion barriers)
constexpr int kS8WRingMax = 16;  // CTA pairs in the optimistic form: a stage takes half as long, so twice as many digit stages
                                 // (8 KB each) must be in flight to cover the L2 latency of the bulk copies
constexpr int kS8ARing = 4;  // genotype-operand stages in TMEM: 4 x (32 + 32) columns (full form)
constexpr int kS8ARingNoE = 8;  // optimistic form (no indicator operand): 8 x 32 columns -- its MMAs take half as long, and
                                // with 4 slots the round trip MMA done -> TMEM store -> arrive -> issue bounded the kernel
                                // (ncu: tensor pipe 52 %, decode warps stalled on the slot barrier)
constexpr int kS8ARingMax = 8;
constexpr uint32_t kS8ColDe = 128, kS8ColA = 256;

__device__ __forceinline__ uint64_t s8_desc(uint32_t saddr) {
  // K-major operand, SWIZZLE_128B: start >> 4 | LBO (unused) = 1 | SBO = 1024 B | version 1 | layout 2
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]^T
__device__ __forceinline__ void s8_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// byte permute with the table in both source registers.  (__byte_perm masks the selector with 0x7777 first; the selectors
// here are already masked, and the extra LOP3 per permute was a fifth of the decode's instructions.)
__device__ __forceinline__ uint32_t s8_prmt(uint32_t table, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %1, %2;\n" : "=r"(r) : "r"(table), "r"(sel));
  return r;
}
__device__ __forceinline__ uint32_t s8_and(uint32_t a, uint32_t m) {
  uint32_t r;
  asm("and.b32 %0, %1, %2;\n" : "=r"(r) : "r"(a), "r"(m));
  return r;
}
// mbarrier wait that names the decoded words as read-write operands: their computation cannot be scheduled after it
__device__ __forceinline__ void s8_wait_keep(uint64_t* bar, uint32_t parity, uint32_t (&a)[32], uint32_t (&b)[32]) {
#pragma unroll
  for (int i = 0; i < 32; i += 8)
    asm volatile("" : "+r"(a[i]), "+r"(a[i + 1]), "+r"(a[i + 2]), "+r"(a[i + 3]), "+r"(a[i + 4]), "+r"(a[i + 5]), "+r"(a[i + 6]),
                 "+r"(a[i + 7]), "+r"(b[i]), "+r"(b[i + 1]), "+r"(b[i + 2]), "+r"(b[i + 3]), "+r"(b[i + 4]), "+r"(b[i + 5]),
                 "+r"(b[i + 6]), "+r"(b[i + 7]));
  mbar_wait(bar, parity);
}
__device__ __forceinline__ void s8_wait_keep1(uint64_t* bar, uint32_t parity, uint32_t (&a)[32]) {
#pragma unroll
  for (int i = 0; i < 32; i += 8)
    asm volatile("" : "+r"(a[i]), "+r"(a[i + 1]), "+r"(a[i + 2]), "+r"(a[i + 3]), "+r"(a[i + 4]), "+r"(a[i + 5]), "+r"(a[i + 6]),
                 "+r"(a[i + 7]));
  mbar_wait(bar, parity);
}
// ---- CTA-pair (cta_group::2) forms: one MMA drives the tensor cores of both SMs of a TPC; each CTA holds its own 128
// variants (A, D in its TMEM) and HALF of the digit rows (B) in its shared memory ----
__device__ __forceinline__ void s8_mma_ts2(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all prior MMAs of the pair, signalled on the same barrier offset in BOTH CTAs
__device__ __forceinline__ void s8_commit2(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ uint32_t s8_cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void s8_arrive_cta(uint64_t* bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
  // default (cta-scope release) semantics, as CUTLASS's ClusterBarrier::arrive(cta_id): a cluster-scope release turns
  // into a full memory barrier that waits for the decode's prefetch loads in flight (measured: 1400 clk per arrive);
  // the data handed over lives in TMEM / async-proxy shared memory and is ordered by the tcgen05 fences
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];\n" ::"r"(remote) : "memory");
}
// wait on a barrier whose arrivals come from both CTAs of the pair
__device__ __forceinline__ void s8_wait_cluster(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void s8_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ bool s8_elect_one() {
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
__device__ __forceinline__ void s8_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}

struct Split8Smem {
  uint64_t full_w[kS8WRingMax];  // bulk copy of a digit stage has landed
  uint64_t done[kS8WRingMax];    // the MMAs of a stage have completed (stage g -> barrier g % 8, parity (g / 8) & 1)
  uint64_t full_a[kS8ARingMax];  // the 4 decode warps of a stage have written its genotype operand
  uint64_t tmem_full;
  uint64_t tmem_empty;
  uint64_t tile_started;
  uint64_t peer_w[kS8WRingMax];  // pair mode, leader: the peer CTA's half of a digit stage has landed
  uint32_t tmem_base;
  uint32_t pad;
  int32_t n00[kS8MaxGroups][kS8TileV];  // per decode group: homozygous-alternate (code 00) count per variant
};

// Work-item order.  A round of concurrently running items should touch few distinct digit blocks (128 N bytes
// each) AND few distinct genotype tiles (32 N bytes each): column blocks are taken in groups of kS8CbGroup, and
// inside a group the items run tile-major, so that a wave of 148 CTAs shares ~6 digit blocks and ~25 tiles through L2.
constexpr int kS8CbGroup = 6;
__device__ __forceinline__ void s8_item(int item, int n_tiles, int n_cb, int& tile, int& cb) {
  const int per_group = n_tiles * kS8CbGroup;
  const int grp = item / per_group, rem = item - grp * per_group;
  const int gsz = min(kS8CbGroup, n_cb - grp * kS8CbGroup);  // the last group may be smaller
  tile = rem / gsz;
  cb = grp * kS8CbGroup + (rem - tile * gsz);
}

// Sample-range split (small blocks: fewer (tile, column block) items than SMs).  Item = (tile, column block, z); split z
// covers the stages [z per, (z + 1) per), the last one up to n_stages (per is even and every range has >= 4 stages, so all
// MMA issuers and both decode groups take part in every item).  The partial digit sums are int32: they are added into a
// zeroed workspace with integer atomics (exact, any order gives the same bits) and split8_finalize_kernel recombines.
__device__ __forceinline__ void s8_item_split(int item, int n_tiles, int n_cb, int n_split, int per, int n_st, int& tile, int& cb,
                                              int& kbeg, int& kcnt) {
  int z = 0, base = item;
  if (n_split > 1) {
    base = item / n_split;
    z = item - base * n_split;
  }
  s8_item(base, n_tiles, n_cb, tile, cb);
  kbeg = z * per;
  kcnt = (z == n_split - 1) ? n_st - kbeg : per;
}

#ifdef GLR_S8_PROFILE
__device__ long long g_s8_prof[160][3][8];  // [cta][issuer 0 / issuer 1 / decode warp 4][phase] clock64 sums
#define S8_T(var) const long long var = clock64()
#define S8_ACC(role, slot, t1, t0) g_s8_prof[blockIdx.x][(role) > 2 ? 1 : (role)][slot] += (t1) - (t0)
#else
#define S8_T(var)
#define S8_ACC(role, slot, t1, t0)
#endif

// kAlign: 8 = every row starts on an 8-byte boundary (base and stride multiples of 8): the 32 bytes of a stage are four 64-bit
// loads -- the row segments of a warp's 32 variants are 32 different cache lines, so every load instruction costs the L1 data
// pipe a dozen wavefronts, and with 32-bit loads that pipe was 85 % busy (ncu, round 1), co-limiting the kernel;
// 4 = 4-byte boundary: eight 32-bit loads; 0 = any alignment: nine 32-bit loads and funnel shifts.
// kPair: CTA pairs (cluster of 2, cta_group::2): a pair works on 256 variants and shares one copy of every digit
// stage (each CTA loads and holds half of the rows), which halves the shared-memory and L2 traffic of the digit stream.
// Only the leader (rank 0) issues MMAs; barriers the leader waits on collect arrivals from both CTAs, and the MMA
// completion is multicast to both.
// kNoE: "optimistic" form for blocks expected to have no missing call: no indicator operand, no De MMAs (half the tensor
// work).  The decode still checks every code; a missing call raises *p.flags and the host re-runs the block in the full form.
// kFmt: GLR_FMT_PLINK2 (2-bit rows, 32 bytes per stage) or GLR_FMT_I8 (genotype_states as int8, 128 bytes per stage; kAlign = 16
// or 4).  The int8 form assumes values in {-1, 0, 1, 2}; it checks every byte while decoding and raises *p.flags |= 2 when
// it meets anything else, upon which the host re-runs the block on the FP64 path (any int8 value is legal there).
template <int kAlign, bool kPair, bool kNoE, int kFmt = GLR_FMT_PLINK2>
__global__ void __launch_bounds__(S8Shape<kNoE>::kThreads, 1) split8_kernel(const Split8Params p) {
  constexpr int kS8Groups = S8Shape<kNoE>::kGroups, kS8Issuers = S8Shape<kNoE>::kIssuers;
  constexpr int kS8DecodeWarps = 4 * kS8Groups, kS8FirstDecodeWarp = kS8Issuers;
  constexpr int kS8ProducerWarp = kS8FirstDecodeWarp + kS8DecodeWarps;
  static_assert(kS8Groups <= kS8MaxGroups && kS8Issuers <= kS8MaxIssuers && (kS8Issuers & (kS8Issuers - 1)) == 0, "warp roles");
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int NC = p.NC;
  const uint32_t rank = kPair ? s8_cta_rank() : 0u;
  const uint32_t w_stage_bytes = (uint32_t)NC * kS8K;                   // a whole digit stage in the packed image
  const uint32_t w_bytes = kPair ? w_stage_bytes / 2 : w_stage_bytes;   // what this CTA holds of it
  unsigned char* w_ring = smem_raw;  // 1024-byte aligned operand buffers first, bookkeeping after
  constexpr int kWR = (kNoE && kPair) ? kS8WRingMax : kS8WRing, kWRLog = (kNoE && kPair) ? 4 : 3;
  Split8Smem* sm = reinterpret_cast<Split8Smem*>(w_ring + (size_t)kWR * w_bytes);

  // genotype-operand ring in TMEM: depth, log2(depth), columns per slot
  constexpr int kAR = kNoE ? kS8ARingNoE : kS8ARing, kARLog = kNoE ? 3 : 2, kASlot = kNoE ? 32 : 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kS8WRingMax; ++s) {
      mbar_init(&sm->full_w[s], 1);
      mbar_init(&sm->done[s], 1);
    }
    for (int s = 0; s < kS8WRingMax; ++s) mbar_init(&sm->peer_w[s], 1);
    for (int s = 0; s < kS8ARingMax; ++s) mbar_init(&sm->full_a[s], (kPair ? 2 : 1) * 4);  // the 4 warps of one group
    mbar_init(&sm->tmem_full, kS8Issuers);  // every MMA issuer thread commits its last stage of the item
    mbar_init(&sm->tmem_empty, (kPair ? 2 : 1) * kS8DecodeWarps);
    mbar_init(&sm->tile_started, 1);
    mbar_fence_init();
  }
  if (warp == kS8ProducerWarp) {
    if (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&sm->tmem_base)),
                   "n"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&sm->tmem_base)),
                   "n"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  if (kPair) s8_cluster_sync(); else __syncthreads();  // barriers of BOTH CTAs are initialised before any remote arrive
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = sm->tmem_base;

  // a work unit (one CTA, or one CTA pair) takes items first_item, first_item + item_stride, ...
  const int tile_v = kPair ? 2 * kS8TileV : kS8TileV;
  const int n_tiles = (p.x.n_variants + tile_v - 1) / tile_v;
  const int n_split = p.n_split, per = p.split_stages;
  const int n_items = n_tiles * p.n_cb * n_split;  // work item = (tile [pair], column block, sample range)
  const int first_item = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int item_stride = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int n_st_all = (int)p.n_stages;  // 128-sample stages covering the genotype samples (>= 2)

  // Every role walks the same sequence of stages; g = running ("global") stage index of this CTA.  Ring slots and
  // barrier parities are functions of g (power-of-two depths: masks and shifts only).
  if (warp == kS8ProducerWarp) {
    // ---- digit-operand producer ----
    if (lane == 0) {
      uint32_t g = 0;
      for (int item = first_item; item < n_items; item += item_stride) {
        int tile, cb, kbeg, kcnt;
        s8_item_split(item, n_tiles, p.n_cb, n_split, per, n_st_all, tile, cb, kbeg, kcnt);
        // pair: rows [rank NC/2, ..)
        const int8_t* src = p.w8 + ((int64_t)cb * n_st_all + kbeg) * w_stage_bytes + rank * w_bytes;
        for (int64_t k = 0; k < kcnt; ++k, ++g) {
          const uint32_t s = g & (kWR - 1);
          if (g >= (uint32_t)kWR) mbar_wait_backoff(&sm->done[s], ((g >> kWRLog) & 1u) ^ 1u);  // stage g - ring depth consumed
          mbar_expect_tx(&sm->full_w[s], w_bytes);
          bulk_g2s(w_ring + (size_t)s * w_bytes, src + k * w_stage_bytes, w_bytes, &sm->full_w[s]);
        }
      }
    }
  } else if (warp < kS8Issuers) {
    // ---- MMA issuers: four warps take the stages round robin, so that one is issuing while the others are
    // committing / waiting for their next stage's operands.  Integer accumulation is order independent; the only
    // ordering needed is that the stage that OVERWRITES the accumulators (first stage of an item) is issued
    // before the other thread's first stage of that item. ----
    // The whole warp runs the loop CONVERGED (warp-uniform indices, addresses and barrier polls live in uniform
    // registers); only the tcgen05 instructions themselves are issued by one elected lane.
    if (kPair && rank != 0) {
      // peer CTA of a pair: no MMAs to issue; these two warps forward "my half of digit stage g has landed" to the leader
      const uint32_t j = (uint32_t)__shfl_sync(0xffffffffu, warp, 0);
      uint32_t g0 = 0;
      for (int item = first_item; item < n_items; item += item_stride) {
        int tile_, cb_, kbeg_, n_st32;
        s8_item_split(item, n_tiles, p.n_cb, n_split, per, n_st_all, tile_, cb_, kbeg_, n_st32);
        for (int k = (int)((j - g0) & (uint32_t)(kS8Issuers - 1)); k < n_st32; k += kS8Issuers) {
          const uint32_t g = g0 + (uint32_t)k;
          const uint32_t sw = g & (kWR - 1);
          mbar_wait(&sm->full_w[sw], (g >> kWRLog) & 1u);
          if (lane == 0) s8_arrive_cta(&sm->peer_w[sw], 0);
          __syncwarp();
        }
        g0 += (uint32_t)n_st32;
      }
    } else {
      const uint32_t j = (uint32_t)__shfl_sync(0xffffffffu, warp, 0);
      // instruction descriptor: D = S32, A = B = signed 8 bit, both K-major, N >> 3, M >> 4
      const uint32_t idesc =
          (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NC >> 3) << 17) | ((uint32_t)((kPair ? 2 : 1) * kS8TileV >> 4) << 24);
      const uint32_t w_ring_addr = smem_u32(w_ring);
      uint32_t g0 = 0;  // global index of the item's first stage
      int icount = 0;
      S8_T(t_begin);
      for (int item = first_item; item < n_items; item += item_stride, ++icount) {
        int tile_, cb_, kbeg_, n_st32;
        s8_item_split(item, n_tiles, p.n_cb, n_split, per, n_st_all, tile_, cb_, kbeg_, n_st32);
        for (int k = (int)((j - g0) & (uint32_t)(kS8Issuers - 1)); k < n_st32; k += kS8Issuers) {
          const uint32_t g = g0 + (uint32_t)k;
          S8_T(t0);
          if (k == 0) {
            // the epilogue of the previous item must have drained the accumulators
            if (icount > 0) {
              if (kPair) s8_wait_cluster(&sm->tmem_empty, (uint32_t)((icount - 1) & 1));
              else mbar_wait(&sm->tmem_empty, (uint32_t)((icount - 1) & 1));
            }
          } else if (k < kS8Issuers) {
            mbar_wait(&sm->tile_started, (uint32_t)(icount & 1));  // stage 0 (overwrite) has been issued
          }
          const uint32_t sw = g & (kWR - 1), sa = g & (kAR - 1);
          S8_T(t1);
          mbar_wait(&sm->full_w[sw], (g >> kWRLog) & 1u);
          if (kPair) s8_wait_cluster(&sm->peer_w[sw], (g >> kWRLog) & 1u);
          S8_T(t2);
          if (kPair) s8_wait_cluster(&sm->full_a[sa], (g >> kARLog) & 1u);
          else mbar_wait(&sm->full_a[sa], (g >> kARLog) & 1u);
          S8_T(t3);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          // descriptor of the stage's digit rows: only the start-address field changes (units of 16 bytes)
          const uint64_t db0 = s8_desc(w_ring_addr + sw * w_bytes);
          const uint32_t a_x = tmem + kS8ColA + sa * kASlot, a_e = a_x + 32;
          if (s8_elect_one()) {
#pragma unroll
            for (int kk = 0; kk < kS8K / 32; ++kk) {
              const uint32_t acc = (k > 0 || kk > 0) ? 1u : 0u;
              const uint64_t db = db0 + (uint64_t)(kk * 2);
              if (kPair) {
                s8_mma_ts2(tmem, a_x + kk * 8, db, idesc, acc);
                if (!kNoE) s8_mma_ts2(tmem + kS8ColDe, a_e + kk * 8, db, idesc, acc);
              } else {
                s8_mma_ts(tmem, a_x + kk * 8, db, idesc, acc);
                if (!kNoE) s8_mma_ts(tmem + kS8ColDe, a_e + kk * 8, db, idesc, acc);
              }
            }
            if (k == 0) mbar_arrive(&sm->tile_started);
            S8_T(t4);
            // ONE commit per stage frees both operands of the stage; after this warp's last stage of the item a
            // second one publishes the accumulators (pair mode: both signalled in both CTAs)
            if (kPair) {
              s8_commit2(&sm->done[sw]);
              if (k + kS8Issuers >= n_st32) s8_commit2(&sm->tmem_full);
            } else {
              s8_commit(&sm->done[sw]);
              if (k + kS8Issuers >= n_st32) s8_commit(&sm->tmem_full);
            }
            S8_T(t5);
            S8_ACC(j, 0, t1, t0);
            S8_ACC(j, 1, t2, t1);
            S8_ACC(j, 2, t3, t2);
            S8_ACC(j, 3, t4, t3);
            S8_ACC(j, 4, t5, t4);
            S8_ACC(j, 5, 1, 0);
          }
          __syncwarp();
        }
        g0 += (uint32_t)n_st32;
      }
#ifdef GLR_S8_PROFILE
      if (lane == 0) {
        S8_T(t_end);
        S8_ACC(j, 6, t_end, t_begin);
      }
#endif
    }
  } else if (warp >= kS8FirstDecodeWarp) {
    // ---- decode + count, epilogue ----
    const int q = warp & 3;                                 // TMEM lane quarter this warp may access
    const uint32_t grp = (warp - kS8FirstDecodeWarp) >> 2;  // takes the stages with global index = grp mod kS8Groups
    const int vloc = q * 32 + lane;                         // variant of the tile = TMEM lane
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const int64_t n_geno = p.x.n_geno;
    const int64_t row_bytes = (n_geno + 3) >> 2;
    const int64_t full_bytes = n_geno >> 2;
    uint32_t g0 = 0;
    int icount = 0;
    const uint32_t x_lut = p.x_lut ? p.x_lut : 0x00010002u;
    // table of the missing-call indicator (code 01 -> 1), a run-time value so that it stays in ONE register: as a literal
    // the compiler re-materialised it from a uniform register in front of every PRMT (32 extra moves per stage and warp)
    const uint32_t e_lut = p.e_lut;
    const bool sq = p.x_lut == 0x00010004u;  // int8 states: the same table, applied arithmetically
    uint32_t miss_seen = 0;  // kNoE: OR of the code-01 bits met
    uint32_t bad_seen = 0;   // int8 form: bits of bytes outside {-1, 0, 1, 2}
    for (int item = first_item; item < n_items; item += item_stride, ++icount) {
      int tile, cb, kbeg, kcnt;
      s8_item_split(item, n_tiles, p.n_cb, n_split, per, n_st_all, tile, cb, kbeg, kcnt);
      if (kPair) tile = tile * 2 + (int)rank;  // the pair's two 128-variant halves
      const int v = min(tile * kS8TileV + vloc, p.x.n_variants - 1);
      const uint8_t* row = p.x.base + (int64_t)v * p.x.stride;
      constexpr int kIn = kFmt == GLR_FMT_I8 ? 32 : kS8Words + 1;  // input words held one stage ahead
      uint32_t nxt[kIn];
      uint32_t sh_nxt = 0;
      auto load = [&](int64_t k) {  // k = absolute stage
        if constexpr (kFmt == GLR_FMT_I8) {
          const int64_t s0 = k * kS8K;
          if (s0 + kS8K <= n_geno) {
            if (kAlign == 16) {
              const uint4* q4 = reinterpret_cast<const uint4*>(row + s0);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const uint4 t = __ldg(q4 + i);
                nxt[4 * i] = t.x;
                nxt[4 * i + 1] = t.y;
                nxt[4 * i + 2] = t.z;
                nxt[4 * i + 3] = t.w;
              }
            } else {
              const uint32_t* qw = reinterpret_cast<const uint32_t*>(row + s0);
#pragma unroll
              for (int i = 0; i < 32; ++i) nxt[i] = __ldg(qw + i);
            }
          } else {  // last stage of the row: samples beyond n_geno read as 0
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              uint32_t w = 0;
#pragma unroll
              for (int b = 0; b < 4; ++b) {
                const int64_t off = s0 + i * 4 + b;
                if (off < n_geno) w |= (uint32_t)__ldg(row + off) << (8 * b);
              }
              nxt[i] = w;
            }
          }
          return;
        }
        const int64_t b0 = k * 32;
        if (b0 + 4 * kS8Words + 3 <= full_bytes) {  // every word touched lies inside the whole bytes of the row
          const uint8_t* q8 = row + b0;
          if (kAlign == 8) {
            const uint2* q2 = reinterpret_cast<const uint2*>(q8);
#pragma unroll
            for (int i = 0; i < kS8Words / 2; ++i) {
              const uint2 t = __ldg(q2 + i);
              nxt[2 * i] = t.x;
              nxt[2 * i + 1] = t.y;
            }
            nxt[kS8Words] = 0u;
            sh_nxt = 0;
          } else {
            const uint32_t a = kAlign ? 0u : (uint32_t)(reinterpret_cast<uintptr_t>(q8) & 3);
            const uint32_t* qw = reinterpret_cast<const uint32_t*>(q8 - a);
#pragma unroll
            for (int i = 0; i < kS8Words; ++i) nxt[i] = __ldg(qw + i);
            nxt[kS8Words] = a ? __ldg(qw + kS8Words) : 0u;
            sh_nxt = a * 8;
          }
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
      int n00 = 0;
      const int k_first = (int)((grp + (uint32_t)kS8Groups - g0 % (uint32_t)kS8Groups) % (uint32_t)kS8Groups);
      // (with two decode groups an L2 prefetch of the row lines 16 stages ahead paid off; with four, every group loads its
      // segment a whole 128-byte line -- four stages -- ahead, and the prefetch instructions only cost energy: -1.9 %)
      load(kbeg + k_first);
      for (int k = k_first; k < kcnt; k += kS8Groups) {
        const uint32_t g = g0 + (uint32_t)k;
        const int64_t ka = kbeg + k;
        const uint32_t sa = g & (kAR - 1);
        const uint32_t a_x = lane_addr + kS8ColA + sa * kASlot, a_e = a_x + 32;
        if constexpr (kFmt == GLR_FMT_I8) {
          // ---- int8 genotype states: 128 bytes of the row per stage.  The registers cannot hold a decoded stage next to
          // the prefetched one, so the TMEM slot is awaited first and the stage is decoded and stored in four 32-sample parts.
          uint32_t w[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) w[i] = nxt[i];
          if (k + kS8Groups < kcnt) load(ka + kS8Groups);
          if (g >= (uint32_t)kAR) {
            uint64_t* bar = &sm->done[(g - kAR) & (kWR - 1)];
            s8_wait_keep1(bar, ((g - kAR) >> kWRLog) & 1u, w);
          }
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t xs8[8], es8[8];
#pragma unroll
            for (int h = 0; h < 2; ++h) {  // 16 samples = 4 input words -> 4 output words in the order of the 2-bit decode
              const uint32_t* in = w + 8 * c + 4 * h;
              uint32_t t[4];
              asm("prmt.b32 %0, %1, %2, 0x6420;\n" : "=r"(t[0]) : "r"(in[0]), "r"(in[1]));  // samples 0, 2, 4, 6
              asm("prmt.b32 %0, %1, %2, 0x6420;\n" : "=r"(t[1]) : "r"(in[2]), "r"(in[3]));  // 8, 10, 12, 14
              asm("prmt.b32 %0, %1, %2, 0x7531;\n" : "=r"(t[2]) : "r"(in[0]), "r"(in[1]));  // 1, 3, 5, 7
              asm("prmt.b32 %0, %1, %2, 0x7531;\n" : "=r"(t[3]) : "r"(in[2]), "r"(in[3]));  // 9, 11, 13, 15
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t e = (t[j] >> 7) & 0x01010101u;           // -1 = 0xFF: missing
                const uint32_t u = t[j] ^ (e * 0xFFu);                   // missing -> 0, everything else unchanged
                // anything but -1, 0, 1, 2: a value above 2 (bits 2..6, or bits 0 and 1 together), a negative one other than -1
                bad_seen |= (u & 0xFCFCFCFCu) | (u & (u >> 1) & 0x01010101u) | (u & (e * 0xFFu));
                n00 += __popc(u & 0x02020202u);                          // number of alternate-homozygous calls
                uint32_t xv = u & 0x03030303u;
                if (sq) xv += xv & 0x02020202u;                          // squared calls: 0, 1, 2 -> 0, 1, 4
                xs8[4 * h + j] = xv;
                es8[4 * h + j] = e;
                if (kNoE) miss_seen |= e;
              }
            }
            tmem_st8(a_x + 8 * c, xs8);
            if constexpr (!kNoE) tmem_st8(a_e + 8 * c, es8);
          }
        } else {
        S8_T(d0);
        uint32_t w[kS8Words];
#pragma unroll
        for (int i = 0; i < kS8Words; ++i) w[i] = kAlign ? nxt[i] : __funnelshift_r(nxt[i], nxt[i + 1], sh_nxt);
        if (k + kS8Groups < kcnt) load(ka + kS8Groups);
        // decode into registers first, THEN wait for the TMEM slot: the slot frees up when the MMAs of stage g - 4
        // complete, and the hand-off back to the issuer must not include the decode arithmetic
        uint32_t xs[32], es[kNoE ? 1 : 32];
        if (p.fused_counts) {  // (uniform) the count of code 00 is only needed for the fused moments
#pragma unroll
          for (int i = 0; i < kS8Words; ++i)
            n00 += __popc(~(w[i] | (w[i] << 1)) & 0xAAAAAAAAu);  // both bits of a code clear (left shift: FMA pipe, not ALU)
        }
#pragma unroll
        for (int i = 0; i < kS8Words; ++i) {
          const uint32_t ww = w[i];
          // code 01 (low bit set, high bit clear) at the even bit positions; the odd ones are masked once, at the end
          if (kNoE) miss_seen |= ww & ~(ww >> 1);
          // PRMT selectors: nibble n of `ev` = code 2n (bit 2 is a stray bit of the neighbouring code: the table
          // is passed in both source registers, so it does not matter; bit 3 must be clear), of `od` = code 2n+1.
          // Sample order inside a 16-sample group is therefore 0,2,4,6 | 8,10,12,14 | 1,3,5,7 | 9,11,13,15;
          // the digit operand is packed with the same permutation (pack_split8_kernel).
          // (masks through opaque asm: otherwise the compiler rewrites (w & m) >> 16 into two more mask operations)
          const uint32_t ev = s8_and(ww, 0x77777777u), od = s8_and(ww >> 2, 0x77777777u);
          const uint32_t ev_hi = ev >> 16, od_hi = od >> 16;
          // code 00 -> 2, 01 -> 0 (missing), 10 -> 1, 11 -> 0;  e: code 01 -> 1
          xs[4 * i + 0] = s8_prmt(x_lut, ev);
          xs[4 * i + 1] = s8_prmt(x_lut, ev_hi);
          xs[4 * i + 2] = s8_prmt(x_lut, od);
          xs[4 * i + 3] = s8_prmt(x_lut, od_hi);
          if (!kNoE) {
            es[4 * i + 0] = s8_prmt(e_lut, ev);
            es[4 * i + 1] = s8_prmt(e_lut, ev_hi);
            es[4 * i + 2] = s8_prmt(e_lut, od);
            es[4 * i + 3] = s8_prmt(e_lut, od_hi);
          }
        }
        S8_T(d1);
        // the MMAs that read this TMEM slot last (stage g - ring depth) must have completed
        if (g >= (uint32_t)kAR) {
          uint64_t* bar = &sm->done[(g - kAR) & (kWR - 1)];
          const uint32_t par = ((g - kAR) >> kWRLog) & 1u;
          if constexpr (kNoE) s8_wait_keep1(bar, par, xs);
          else s8_wait_keep(bar, par, xs, es);
        }
        S8_T(d2);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        tmem_st16(a_x, xs);
        tmem_st16(a_x + 16, xs + 16);
        if constexpr (!kNoE) {
          tmem_st16(a_e, es);
          tmem_st16(a_e + 16, es + 16);
        }
#ifdef GLR_S8_PROFILE
        if (warp == kS8FirstDecodeWarp && lane == 0) {
          S8_T(d3);
          S8_ACC(2, 0, d1, d0);
          S8_ACC(2, 1, d2, d1);
          S8_ACC(2, 2, d3, d2);
          S8_ACC(2, 5, 1, 0);
        }
#endif
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (kPair) s8_arrive_cta(&sm->full_a[sa], 0);
          else mbar_arrive(&sm->full_a[sa]);
        }
      }
      const int64_t vg = (int64_t)tile * kS8TileV + vloc;
      const bool v_ok = vg < p.x.n_variants;
      if (n_split == 1) {
        sm->n00[grp][vloc] = n00;
        asm volatile("bar.sync 1, %0;\n" ::"n"(32 * kS8DecodeWarps) : "memory");  // all decode warps: every group's counts are in shared memory
      } else if (cb == 0 && v_ok && n00 != 0) {
        atomicAdd(p.n00_ws + vg, n00);
      }

      // ---- epilogue: thread = variant; the two groups split the columns ----
      mbar_wait(&sm->tmem_full, (uint32_t)(icount & 1));
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const int col0 = cb * p.cpb, ncols = min(p.cpb, p.n_cols - col0);
      if (n_split == 1) {
        double mu = -1.0;
        {
          uint32_t zx[8], ze[8];
          const int r1 = split8_row0(p.cpb, p.n_lo, p.lo_digits);  // the row of ones follows the last column
          tmem_ld8(lane_addr + r1, zx);
          if (!kNoE) tmem_ld8(lane_addr + kS8ColDe + r1, ze);
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
          if (p.fused_counts) {
            const double sx = (double)(int32_t)zx[0];                  // 2 n00 + n10
            const double c01 = kNoE ? 0.0 : (double)(int32_t)ze[0];    // missing calls
            int c00i = 0;
#pragma unroll
            for (int gq = 0; gq < kS8Groups; ++gq) c00i += sm->n00[gq][vloc];
            const double c00 = (double)c00i;
            if (p.missing_policy == GLR_MISSING_MEAN && c01 < (double)n_geno)
              mu = sx / ((double)n_geno - c01);  // MeanSubstitute.scala:57-84
            if (cb == 0 && grp == 0) {
              if (v_ok) {
                p.v1_out[vg] = mu;
                p.mom_out[vg] = sx + 2.0 * c00 + (mu * mu) * c01;
                p.mom_out[p.ldS + vg] = sx + mu * c01;
              }
              if (!kNoE) {  // total number of missing calls of the block (drives the optimistic form of the NEXT block)
                int cm = v_ok ? (int32_t)ze[0] : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) cm += __shfl_xor_sync(0xffffffffu, cm, o);
                if (lane == 0 && cm != 0) atomicAdd(p.flags + 1, cm);
              }
            }
          } else {
            mu = p.x.v1[min(vg, (int64_t)p.x.n_variants - 1)];
          }
        }
        for (int c = (int)grp; c < ncols; c += kS8Groups) {
          uint32_t zx[8], ze[8];
          const int r0 = split8_row0(c, p.n_lo, p.lo_digits), nd = c < p.n_lo ? p.lo_digits : kS8Digits;
          tmem_ld8(lane_addr + r0, zx);
          if (!kNoE) tmem_ld8(lane_addr + kS8ColDe + r0, ze);
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
          double acc = 0.0;
#pragma unroll
          for (int i = kS8Digits - 1; i >= 0; --i)
            if (i < nd)
              acc = fma(acc, 256.0, kNoE ? (double)(int32_t)zx[i] : fma(mu, (double)(int32_t)ze[i], (double)(int32_t)zx[i]));
          p.S[(int64_t)(col0 + c) * p.ldS + vg] = acc * p.scale[col0 + c];
        }
      } else {
        // partial digit sums of this sample range: exact integer atomics into the zeroed workspace
        // zws[cb][accumulator][digit row][ldS]; the two groups split the 8-row chunks
        int32_t* zx_ws = p.z_ws + (int64_t)cb * 2 * NC * p.ldS + vg;
        for (int ch = (int)grp; ch < (NC >> 3); ch += kS8Groups) {  // rows of absent columns hold zeros: no atomics issued
          uint32_t zx[8], ze[8];
          tmem_ld8(lane_addr + ch * 8, zx);
          if (!kNoE) tmem_ld8(lane_addr + kS8ColDe + ch * 8, ze);
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
          if (v_ok) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int64_t r = ch * 8 + i;
              if (zx[i]) atomicAdd(zx_ws + r * p.ldS, (int32_t)zx[i]);
              if (!kNoE && ze[i]) atomicAdd(zx_ws + ((int64_t)NC + r) * p.ldS, (int32_t)ze[i]);
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) {  // accumulators may be overwritten
        if (kPair) s8_arrive_cta(&sm->tmem_empty, 0);
        else mbar_arrive(&sm->tmem_empty);
      }
      if (n_split == 1) asm volatile("bar.sync 1, %0;\n" ::"n"(32 * kS8DecodeWarps) : "memory");  // n00[] may be rewritten
      g0 += (uint32_t)kcnt;
    }
    if (kNoE) {
      const uint32_t m = kFmt == GLR_FMT_I8 ? miss_seen : (miss_seen & 0x55555555u);
      if (__any_sync(0xffffffffu, m != 0) && lane == 0) atomicOr(p.flags, 1);
    }
    if (kFmt == GLR_FMT_I8) {
      if (__any_sync(0xffffffffu, bad_seen != 0) && lane == 0) atomicOr(p.flags, 2);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  if (kPair) s8_cluster_sync(); else __syncthreads();
  if (warp == kS8ProducerWarp) {
    if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512));
  }
}

void split8_geometry(int n_cols, int n_lo, int lo_digits, int* n_cb, int* cpb, int* NC) {
  *n_cb = (n_cols + 17) / 18;
  *cpb = (n_cols + *n_cb - 1) / *n_cb;
  if (*n_cb > 1) n_lo = 0;
  *NC = (int)round_up((int64_t)split8_row0(*cpb, n_lo, lo_digits) + 1, 16);
}

// Recombination of the split form: thread = (variant, column block) reads the summed int32 digit rows and does exactly
// what the fused epilogue does (same operations in the same order: bit-identical S, v1, moments).
__global__ void split8_finalize_kernel(const Split8Params p, int no_e) {
  const int64_t vg = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int cb = blockIdx.y;
  if (vg >= p.x.n_variants) return;
  const int NC = p.NC;
  const int32_t* zx = p.z_ws + (int64_t)cb * 2 * NC * p.ldS + vg;
  const int32_t* ze = zx + (int64_t)NC * p.ldS;
  const double n_geno = (double)p.x.n_geno;
  double mu = -1.0;
  if (p.fused_counts) {
    const int64_t r1 = split8_row0(p.cpb, p.n_lo, p.lo_digits);
    const double sx = (double)zx[r1 * p.ldS];
    const int32_t cm = no_e ? 0 : ze[r1 * p.ldS];
    const double c01 = (double)cm;
    const double c00 = (double)p.n00_ws[vg];
    if (p.missing_policy == GLR_MISSING_MEAN && c01 < n_geno) mu = sx / (n_geno - c01);
    if (cb == 0) {
      p.v1_out[vg] = mu;
      p.mom_out[vg] = sx + 2.0 * c00 + (mu * mu) * c01;
      p.mom_out[p.ldS + vg] = sx + mu * c01;
      if (cm != 0) atomicAdd(p.flags + 1, cm);
    }
  } else {
    mu = p.x.v1[vg];
  }
  const int col0 = cb * p.cpb, ncols = min(p.cpb, p.n_cols - col0);
  for (int c = 0; c < ncols; ++c) {
    double acc = 0.0;
#pragma unroll
    const int r0 = split8_row0(c, p.n_lo, p.lo_digits), nd = c < p.n_lo ? p.lo_digits : kS8Digits;
    for (int i = kS8Digits - 1; i >= 0; --i) {
      if (i >= nd) continue;
      const int64_t r = r0 + i;
      const double dx = (double)zx[r * p.ldS];
      acc = fma(acc, 256.0, no_e ? dx : fma(mu, (double)ze[r * p.ldS], dx));
    }
    p.S[(int64_t)(col0 + c) * p.ldS + vg] = acc * p.scale[col0 + c];
  }
}

// Plan of a launch: CTA pairs or single CTAs, and the split of the sample range that fills the GPU when the block has
// fewer (tile, column block) items than work units.  Returned makespan is in stages of one unit (cost model for api.cu).
Split8Plan split8_plan(int n_variants, int n_cb, int64_t n_stages, int sm_count, int pair_mode, int force_split) {
  Split8Plan pl;
  const int items1 = (n_variants + kS8TileV - 1) / kS8TileV * n_cb;
  const int items2 = (n_variants + 2 * kS8TileV - 1) / (2 * kS8TileV) * n_cb;
  const int max_split = (int)std::max<int64_t>(1, std::min<int64_t>(64, n_stages / 8));
  auto best = [&](int items, int units, int& ns_out, int& per_out) {
    const int kOverhead = 8;  // pipeline fill + epilogue of an item, in stages
    double best_cost = 1e300;
    for (int ns = 1; ns <= max_split; ++ns) {
      if (force_split > 0 && ns != std::min(force_split, max_split)) continue;
      // ns ranges of `per` stages (even), the last one shorter but with a stage for every issuer warp:
      // n_stages - (ns - 1) per >= kS8MaxIssuers
      const int per = ns == 1 ? (int)n_stages : (int)round_up(ceil_div(n_stages, ns), 2);
      if (ns > 1 && (int64_t)(ns - 1) * per + kS8MaxIssuers > n_stages) {
        if (force_split > 0) {  // forced but not feasible: fall back to the unsplit form
          ns_out = 1;
          per_out = (int)n_stages;
          best_cost = (double)ceil_div((int64_t)items, units) * (double)(n_stages + kOverhead);
        }
        continue;
      }
      const double waves = (double)ceil_div((int64_t)items * ns, units);
      const double cost = waves * (double)(per + kOverhead);
      if (cost < best_cost * (ns > 1 ? 0.98 : 1.0)) {  // prefer fewer ranges unless clearly better
        best_cost = cost;
        ns_out = ns;
        per_out = per;
      }
      if (items >= 8 * units && force_split <= 0) break;  // many items per unit: the tail is negligible
    }
    return best_cost;
  };
  int ns1 = 1, per1 = (int)n_stages, ns2 = 1, per2 = (int)n_stages;
  const double c1 = best(items1, sm_count, ns1, per1);
  const double c2 = best(items2, sm_count / 2, ns2, per2) * 1.0;
  // pairs halve the digit stream's L2 / shared-memory traffic: preferred whenever they are not slower
  bool pair = pair_mode == GLR_FORCE || (pair_mode != GLR_OFF && c2 <= c1 * 1.02);
  pl.pair = pair ? 1 : 0;
  pl.n_split = pair ? ns2 : ns1;
  pl.split_stages = pair ? per2 : per1;
  pl.makespan_stages = pair ? c2 : c1;
  return pl;
}

size_t split8_ws_bytes(int n_cb, int NC, int64_t ldS) { return (size_t)n_cb * 2 * NC * ldS * sizeof(int32_t); }

void launch_split8(Split8Params p, int sm_count, cudaStream_t st) {
  if (p.x.n_variants <= 0) return;
  const uintptr_t misal = reinterpret_cast<uintptr_t>(p.x.base) | (uintptr_t)p.x.stride;
  const bool i8 = p.x.format == GLR_FMT_I8;
  const int align = i8 ? ((misal & 15) == 0 ? 16 : 4) : ((misal & 7) == 0 ? 8 : ((misal & 3) == 0 ? 4 : 0));
  const int n_items1 = (p.x.n_variants + kS8TileV - 1) / kS8TileV * p.n_cb * p.n_split;
  const int n_items2 = (p.x.n_variants + 2 * kS8TileV - 1) / (2 * kS8TileV) * p.n_cb * p.n_split;
  bool pair = p.pair != 0;
  const bool no_e = p.no_e != 0;
  if (p.n_split > 1) {
    cudaMemsetAsync(p.z_ws, 0, split8_ws_bytes(p.n_cb, p.NC, p.ldS), st);
    cudaMemsetAsync(p.n00_ws, 0, (size_t)p.ldS * sizeof(int32_t), st);
  }
#ifdef GLR_S8_PROFILE
  static long long zero[160][3][8];
  cudaMemcpyToSymbol(g_s8_prof, zero, sizeof(zero));
#endif
  p.e_lut = 0x00000100u;
  using Kern = void (*)(const Split8Params);
  const unsigned n_threads = no_e ? S8Shape<true>::kThreads : S8Shape<false>::kThreads;
  auto pick = [&](bool pr) -> Kern {
#define GLR_S8_K(A, F) (pr ? (no_e ? (Kern)split8_kernel<A, true, true, F> : (Kern)split8_kernel<A, true, false, F>) \
                           : (no_e ? (Kern)split8_kernel<A, false, true, F> : (Kern)split8_kernel<A, false, false, F>))
    if (i8) return align == 16 ? GLR_S8_K(16, GLR_FMT_I8) : GLR_S8_K(4, GLR_FMT_I8);
    return align == 8 ? GLR_S8_K(8, GLR_FMT_PLINK2) : (align == 4 ? GLR_S8_K(4, GLR_FMT_PLINK2) : GLR_S8_K(0, GLR_FMT_PLINK2));
#undef GLR_S8_K
  };
  if (pair) {
    const size_t smem = (size_t)(no_e ? kS8WRingMax : kS8WRing) * (p.NC / 2) * kS8K + sizeof(Split8Smem) + 64;
    Kern kern = pick(true);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * std::min(n_items2, sm_count / 2));
    cfg.blockDim = dim3(n_threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kern, p) != cudaSuccess) {  // no cluster launch on this device / partition
      cudaGetLastError();
      pair = false;
    }
  }
  if (!pair) {
    const size_t smem = (size_t)kS8WRing * p.NC * kS8K + sizeof(Split8Smem) + 64;
    Kern kern = pick(false);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<std::min(n_items1, sm_count), n_threads, smem, st>>>(p);
  }
  if (p.n_split > 1) {
    const dim3 grid((unsigned)ceil_div(p.x.n_variants, 128), (unsigned)p.n_cb);
    split8_finalize_kernel<<<grid, 128, 0, st>>>(p, no_e ? 1 : 0);
  }
#ifdef GLR_S8_PROFILE
  static long long h[160][3][8];
  cudaStreamSynchronize(st);
  cudaMemcpyFromSymbol(h, g_s8_prof, sizeof(h));
  const int nc = pair ? 2 * std::min(n_items2, sm_count / 2) : std::min(n_items1, sm_count);
  const char* names[3] = {"issuer0", "issuer1", "decode4"};
  for (int r = 0; r < 3; ++r) {
    double a[8] = {0};
    for (int c = 0; c < nc; ++c)
      for (int i = 0; i < 8; ++i) a[i] += (double)h[c][r][i];
    const double n = a[5] > 0 ? a[5] : 1;
    fprintf(stderr, "[s8 prof] %s stages/cta %.0f  clk/stage:", names[r], a[5] / nc);
    for (int i = 0; i < 5; ++i) fprintf(stderr, " %.0f", a[i] / n);
    if (a[6] > 0) fprintf(stderr, "  | loop clk per stage of the CTA: %.0f", a[6] / (2 * n));
    fprintf(stderr, "\n");
  }
#endif
}

// ------------------------------------------------------------------------------------------------
// Issue-rate probe: every SM issues `iters` x 8 back-to-back tcgen05.mma kind::i8 (A in TMEM, B in shared memory,
// M128 N128 K32, operands resident, two accumulators alternating as in split8_kernel) and nothing else.  bench.py runs
// it next to the real kernel so that the roofline's "what the tensor pipe can do on this box, right now" is measured,
// not quoted.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) s8_issue_rate_kernel(int iters) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 2 * 128 * kS8K / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x01010101u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  {
    uint32_t v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0x01010101u;
    tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + kS8ColA, v);
    tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + kS8ColA + 16, v);
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (warp == 0) {
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t base = smem_u32(smem_raw);
    for (int it = 0; it < iters; ++it) {
      const uint64_t db0 = s8_desc(base + (uint32_t)(it & 1) * 128 * kS8K);
      if (s8_elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          s8_mma_ts(tmem, tmem + kS8ColA + kk * 8, db0 + (uint64_t)(kk * 2), idesc, 1u);
          s8_mma_ts(tmem + kS8ColDe, tmem + kS8ColA + kk * 8, db0 + (uint64_t)(kk * 2), idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (s8_elect_one()) s8_commit(&bar);
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512));
}

// returns the int8 tensor-core rate in TOP/s (2 ops per multiply-accumulate) of a run of about `ms_target` milliseconds
double measure_s8_issue_rate(int sm_count, double ms_target, cudaStream_t st) {
  const size_t smem = (size_t)2 * 128 * kS8K + 1024;
  cudaFuncSetAttribute(s8_issue_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  s8_issue_rate_kernel<<<sm_count, 128, smem, st>>>(64);  // warm-up
  // 8 MMAs of 64 cycles per iteration at roughly 1.8 GHz
  const int iters = std::max(256, (int)(ms_target * 1e-3 * 1.8e9 / (8 * 64)));
  cudaEventRecord(e0, st);
  s8_issue_rate_kernel<<<sm_count, 128, smem, st>>>(iters);
  cudaEventRecord(e1, st);
  double tops = -1.0;
  if (cudaEventSynchronize(e1) == cudaSuccess && cudaGetLastError() == cudaSuccess) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    tops = 2.0 * (double)sm_count * iters * 8.0 * 128.0 * 128.0 * 32.0 / (ms * 1e-3) / 1e12;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return tops;
}

// ------------------------------------------------------------------------------------------------
// packing: W (row-major [N x ld], columns c0..c0+ncols) -> per column block and 128-sample stage a
// swizzled image [NC digit rows][128 bytes]; row of (local column c, digit i) = split8_row0(c) + i (7 c + i when every
// column has 7 digits), row split8_row0(cpb) = ones.  A column with d digits is quantised to 8 d - 1 bits.
// Inside each 16-sample group the bytes follow the sample order the decode produces (see split8_kernel).
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
                                   const int64_t* geno_to_row, int64_t Ng, int8_t* dst, int dst_col0, int cpb, int NC,
                                   int64_t n_stages, int n_lo, int lo_digits) {
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
  const int gc = dst_col0 + c, cb = gc / cpb, lc = gc - cb * cpb;
  const int nd = lc < n_lo ? lo_digits : kS8Digits;
  long long q = mx > 0.0 ? __double2ll_rn(v / mx * (double)(1ll << (8 * nd - 2))) : 0ll;
  const int64_t stage = sg / kS8K;
  const int k = (int)(sg % kS8K);
  const int r16 = k & 15;  // sample inside its 16-sample group -> byte position (register j, byte b) of the decode
  const int pos = (((r16 >> 3) | ((r16 & 1) << 1)) << 2) | ((r16 >> 1) & 3);
  int8_t* base = dst + ((int64_t)cb * n_stages + stage) * NC * kS8K;
  const int r0 = split8_row0(lc, n_lo, lo_digits);
  for (int i = 0; i < nd; ++i) {
    const long long d = ((q + 128) & 255) - 128;  // balanced digit in [-128, 127]
    q = (q - d) >> 8;
    const int r = r0 + i;
    base[(int64_t)r * kS8K + (((k >> 4) ^ (r & 7)) << 4) + pos] = (int8_t)d;
  }
  if (lc == 0) {  // the first column of a block also writes the block's row of ones
    const int r = split8_row0(cpb, n_lo, lo_digits);
    base[(int64_t)r * kS8K + (((k >> 4) ^ (r & 7)) << 4) + pos] = 1;
  }
}

void launch_pack_split8(const double* src, int64_t N, int64_t ld, int c0, int ncols, double* colmax_scratch,
                        const int64_t* geno_to_row, int64_t Ng, int8_t* dst, int dst_col0, int cpb, int NC, int64_t n_stages,
                        int n_lo, int lo_digits, cudaStream_t st) {
  if (ncols <= 0) return;
  colmax_kernel<<<ncols, 256, 0, st>>>(src, N, ld, c0, ncols, colmax_scratch);
  const int64_t total = (int64_t)ncols * n_stages * kS8K;
  pack_split8_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(src, N, ld, c0, ncols, colmax_scratch, geno_to_row,
                                                                      Ng, dst, dst_col0, cpb, NC, n_stages, n_lo, lo_digits);
}

}  // namespace glr
