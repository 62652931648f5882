// Issue rate of tcgen05.mma.cta_group::2 kind::i8 (A in TMEM, B split across the CTA pair's shared memory).
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_i8_pair umma_i8_pair.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) pair_rate(int iters, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(rank));
  for (int i = tid; i < (128 + N) * 128 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (16u << 24);  // M = 256
  const uint32_t a_col = 2 * N <= 256 ? 256 : 480;
  {
    uint32_t v = 0x01010101u;
    for (int c = 0; c < 32; c += 8)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};\n" ::"r"(tmem + ((uint32_t)(warp * 32) << 16) +
                                                                                                   a_col + (2 * N <= 256 ? c : (c & 8))),
                   "r"(v));
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  if (rank == 0 && warp == 0) {
    if (lane == 0) {
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = make_desc(smem_u32(smem + N * 64) + k * 32);  // SS form: A = 128 rows per CTA after the B half
          const uint64_t db = make_desc(smem_u32(smem) + k * 32);          // B: N/2 rows per CTA
          const uint32_t a = tmem + a_col + (2 * N <= 256 ? k * 8 : (k & 1) * 8);
          if (TS) {
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem),
                         "r"(a), "l"(db), "r"(idesc), "r"(1u));
            if (2 * N <= 256)
              asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem + N),
                           "r"(a), "l"(db), "r"(idesc), "r"(1u));
          } else {
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                         "l"(da), "l"(db), "r"(idesc), "r"(1u));
            if (2 * N <= 256)
              asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem + N),
                           "l"(da), "l"(db), "r"(idesc), "r"(1u));
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(&bar)),
                   "h"((uint16_t)3));
    }
    __syncwarp();
  }
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(&bar)), "r"(0u) : "memory");
  if (tid == 0 && rank == 0) cycles[blockIdx.x / 2] = clock64();
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512));
}

template <int N, bool TS>
void run() {
  const int iters = 2000;
  long long* d;
  cudaMalloc(&d, 148 * 8);
  const size_t smem = (size_t)(128 + N) * 128 + 1024;
  cudaFuncSetAttribute(pair_rate<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaLaunchKernelEx(&cfg, pair_rate<N, TS>, 16, d);
  cudaEventRecord(e0);
  cudaLaunchKernelEx(&cfg, pair_rate<N, TS>, iters, d);
  cudaEventRecord(e1);
  cudaError_t e = cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const int per_it = (2 * N <= 256) ? 8 : 4;
  const double macs = 74.0 * iters * per_it * 256.0 * N * 32;
  printf("pair %s N=%3d  %s  %.3f ms  %.1f ns/MMA  %.2f Pop/s\n", TS ? "TS" : "SS", N, cudaGetErrorString(e), ms, ms * 1e6 / (iters * per_it),
         2 * macs / ms / 1e12);
  cudaFree(d);
}

int main() {
  run<128, true>();
  run<256, true>();
  run<64, true>();
  run<128, false>();
  run<256, false>();
  return 0;
}
