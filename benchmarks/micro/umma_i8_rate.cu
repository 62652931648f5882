// Throughput of tcgen05.mma kind::i8 (and kind::f8f6f4 for comparison) with operands resident in
// shared memory: one CTA per SM issues `iters` x 4 MMAs (M=128, N=256, K=32) back to back.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_i8_rate umma_i8_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}

template <int KIND, int N>
__global__ void __launch_bounds__(128, 1) rate(int iters, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (128 + N) * 128 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = tmem_base_s;
  // i8: c = S32 (2), a = b = signed (1).  f8f6f4: c = F32 (1), a = b = E4M3 (0).
  const uint32_t idesc = KIND == 0 ? ((2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24))
                                   : ((1u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24));
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    if (lane == 0) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = make_desc(smem_u32(smem) + k * 32);
          const uint64_t db = make_desc(smem_u32(smem + 128 * 128) + k * 32);
          if (KIND == 0)
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                         "l"(da), "l"(db), "r"(idesc), "r"(1u));
          else
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                         "l"(da), "l"(db), "r"(idesc), "r"(1u));
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)));
    }
    __syncwarp();
  }
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(&bar)), "r"(0u) : "memory");
  if (tid == 0) {
    t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(256));
}

template <int KIND, int N>
void run(const char* name) {
  const int iters = 2000;
  long long* d;
  cudaMalloc(&d, 148 * 8);
  const size_t smem = (size_t)(128 + N) * 128 + 1024;
  cudaFuncSetAttribute(rate<KIND, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  rate<KIND, N><<<148, 128, smem>>>(10, d);
  cudaEventRecord(e0);
  rate<KIND, N><<<148, 128, smem>>>(iters, d);
  cudaEventRecord(e1);
  cudaError_t e = cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long c;
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  const double macs = 148.0 * iters * 4 * 128.0 * N * 32;
  printf("%-22s N=%3d  %s  %.3f ms  %.0f clk/MMA  %.2f Pop/s\n", name, N, cudaGetErrorString(e), ms, (double)c / (iters * 4.0),
         2 * macs / ms / 1e12);
  cudaFree(d);
}

int main() {
  run<0, 256>("kind::i8");
  run<0, 128>("kind::i8");
  run<1, 256>("kind::f8f6f4 (e4m3)");
  run<1, 128>("kind::f8f6f4 (e4m3)");
  return 0;
}
