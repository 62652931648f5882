// Bring-up test for tcgen05.mma kind::i8 (s8 x s8 -> s32) on sm_100a with hand-built descriptors.
// D[128 x N] = A[128 x K] * B[N x K]^T, both operands K-major in shared memory, 128-byte swizzle.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_i8 umma_i8.cu && ./umma_i8
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 144, KB = 128;  // one stage = 128 bytes of K per row

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  // K-major, SWIZZLE_128B: start address >> 4, LBO (ignored) = 1, SBO = 1024 B >> 4, version 1, layout type 2
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) umma_test(const int8_t* A, const int8_t* B, int K, int32_t* D) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  unsigned char* sA = smem;             // 128 rows x 128 B
  unsigned char* sB = smem + M * KB;    // 144 rows x 128 B
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = tmem_base_s;

  // instruction descriptor: c = S32 (2), a = b = signed 8 bit (1), K-major both, N >> 3, M >> 4
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);

  uint32_t phase = 0;
  for (int k0 = 0; k0 < K; k0 += KB) {
    // stage operands with the 128-byte swizzle: 16-byte chunk c of row r lands at chunk (c ^ (r & 7))
    for (int i = tid; i < (M + N) * 8; i += 128) {
      const int r = i >> 3, c = i & 7;
      const int8_t* src = r < M ? A + (size_t)r * K + k0 + c * 16 : B + (size_t)(r - M) * K + k0 + c * 16;
      unsigned char* dst = (r < M ? sA + r * KB : sB + (r - M) * KB) + ((c ^ (r & 7)) << 4);
      *reinterpret_cast<int4*>(dst) = *reinterpret_cast<const int4*>(src);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy writes -> async proxy (UMMA)
    __syncthreads();
    if (warp == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;\n");
      if (lane == 0) {
        for (int k = 0; k < KB / 32; ++k) {
          const uint64_t da = make_desc(smem_u32(sA) + k * 32);
          const uint64_t db = make_desc(smem_u32(sB) + k * 32);
          const uint32_t acc = (k0 > 0 || k > 0) ? 1u : 0u;
          asm volatile(
              "{\n"
              ".reg .pred p;\n"
              "setp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
              "}\n" ::"r"(tmem),
              "l"(da), "l"(db), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)));
      }
      __syncwarp();
    }
    // everyone waits for the MMAs of this stage before the operands are overwritten
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra D_%=;\n"
        "bra W_%=;\n"
        "D_%=:\n"
        "}\n" ::"r"(smem_u32(&bar)),
        "r"(phase)
        : "memory");
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;\n");
  }

  // epilogue: warp w reads TMEM lanes 32w..32w+31 (= rows of D), 16 columns at a time
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int i = 0; i < 16; ++i) D[(size_t)(warp * 32 + lane) * N + c0 + i] = (int32_t)v[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(256));
}

int main() {
  const int K = 1024;
  std::vector<int8_t> A((size_t)M * K), B((size_t)N * K);
  srand(1);
  for (auto& a : A) a = (int8_t)(rand() % 5 - 2);
  for (auto& b : B) b = (int8_t)(rand() % 255 - 127);
  int8_t *dA, *dB;
  int32_t* dD;
  cudaMalloc(&dA, A.size());
  cudaMalloc(&dB, B.size());
  cudaMalloc(&dD, (size_t)M * N * 4);
  cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
  const size_t smem = (size_t)(M + N) * KB + 1024;
  cudaFuncSetAttribute(umma_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  umma_test<<<1, 128, smem>>>(dA, dB, K, dD);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<int32_t> D((size_t)M * N);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  long bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      int32_t ref = 0;
      for (int k = 0; k < K; ++k) ref += (int32_t)A[(size_t)m * K + k] * (int32_t)B[(size_t)n * K + k];
      if (ref != D[(size_t)m * N + n]) {
        if (bad < 5) printf("mismatch m=%d n=%d got %d ref %d\n", m, n, D[(size_t)m * N + n], ref);
        ++bad;
      }
    }
  printf("mismatches: %ld of %d\n", bad, M * N);
  return bad != 0;
}
