// Shared device helpers: FP64 tensor-core MMA, mbarrier + bulk-copy (TMA engine) wrappers, layouts.
// sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace glr {

// ------------------------------------------------------------------------------------------------
// Operand layout ("pair-atom").  The only FP64 tensor instruction on sm_100a is DMMA.8x8x4
// (mma.sync.aligned.m8n8k4.f64).  Lane T = 4*r + j of a warp supplies
//     A[row r][col j]   (8 x 4, row = output row m, col = reduction index)
//     B[row j][col r]   (4 x 8, row = reduction index, col = output column n)
// and receives C[row r][cols 2j, 2j+1].
// The reduction index is the SAMPLE.  Samples are consumed in groups of 8 = two k-steps; within a
// group, lane column j owns samples {2j, 2j+1}: k-step "a" reduces over samples {0,2,4,6}, k-step
// "b" over {1,3,5,7}.  Every operand is packed to that order once, so that
//   * a lane reads both of its A values for a group with one conflict-free 16-byte LDS,
//   * a lane reads both of its genotype values with one 16-byte load (f64) or one nibble (2-bit),
//   * the C fragment of a [variant x sample] product is already a B fragment pair (masked pass).
// A packed "atom" = 8 samples x 8 columns = 64 doubles = 512 B, element index (lane*2 + h):
//     atom[(4*r + j)*2 + h] = M[s0 + 2*j + h][c0 + r]
// ------------------------------------------------------------------------------------------------
constexpr int kGroup = 8;          // samples per pair-atom
constexpr int kAtomDoubles = 64;   // 8 samples x 8 columns
constexpr int kChunkGroups = 8;    // pair-atoms (sample groups) per pipeline stage
constexpr int kChunk = kGroup * kChunkGroups;  // 64 samples per stage
constexpr int kMaxMT = 8;          // m-tiles (of 8 columns) per column group

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---- mbarrier / bulk async copy (cp.async.bulk = the TMA engine in 1-D mode) --------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// producer-side wait: the slot frees up a whole stage (microseconds) later, so back off between polls
// instead of burning issue slots of the SM sub-partition the producer warp shares with consumers
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(64);
  }
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes % 16 == 0, 16-B aligned).
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__host__ __device__ __forceinline__ int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
__host__ __device__ __forceinline__ int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }

}  // namespace glr
