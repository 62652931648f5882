// tcgen05.mma kind::i8 with the A operand in TENSOR MEMORY (written by tcgen05.st from registers) and
// B K-major in shared memory (128-byte swizzle): layout bring-up + issue rate.
//   D[128 x N] (s32, TMEM) = A[128 x 32] (s8, TMEM: lane = row, 8 columns of 4 packed bytes) * B[N x 32]^T
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_i8_ts umma_i8_ts.cu && ./umma_i8_ts
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d),
      "r"(a), "l"(db), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)));
}
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(
                   smem_u32(bar)),
               "r"(ph)
               : "memory");
}
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
}

constexpr int M = 128;

// ---------------- correctness ----------------
template <int N>
__global__ void __launch_bounds__(128, 1) ts_check(const int8_t* A, const int8_t* B, int K, int32_t* D) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
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
  const uint32_t a_col = 256;  // A staging columns
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  uint32_t phase = 0;
  for (int k0 = 0; k0 < K; k0 += 128) {
    // B stage: N rows x 128 bytes, swizzled
    for (int i = tid; i < N * 8; i += 128) {
      const int r = i >> 3, c = i & 7;
      *reinterpret_cast<int4*>(smem + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const int4*>(B + (size_t)r * K + k0 + c * 16);
    }
    // A stage: thread = row; 128 bytes = 32 columns, 8 per K32 block
    for (int kb = 0; kb < 4; ++kb) {
      uint32_t v[8];
      const uint32_t* src = reinterpret_cast<const uint32_t*>(A + (size_t)tid * K + k0 + kb * 32);
      for (int i = 0; i < 8; ++i) v[i] = src[i];
      st8(tmem + ((uint32_t)(warp * 32) << 16) + a_col + kb * 8, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;\n");
      for (int kb = 0; kb < 4; ++kb)
        mma_ts(tmem, tmem + a_col + kb * 8, make_desc(smem_u32(smem) + kb * 32), idesc, (k0 > 0 || kb > 0) ? 1u : 0u);
      commit(&bar);
    }
    wait(&bar, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;\n");
  }
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int i = 0; i < 16; ++i) D[(size_t)tid * N + c0 + i] = (int32_t)v[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512));
}

template <int N>
int check() {
  const int K = 1024;
  std::vector<int8_t> A((size_t)M * K), B((size_t)N * K);
  srand(1);
  for (auto& a : A) a = (int8_t)(rand() % 3);
  for (auto& b : B) b = (int8_t)(rand() % 255 - 127);
  int8_t *dA, *dB;
  int32_t* dD;
  cudaMalloc(&dA, A.size());
  cudaMalloc(&dB, B.size());
  cudaMalloc(&dD, (size_t)M * N * 4);
  cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
  const size_t smem = (size_t)N * 128 + 1024;
  cudaFuncSetAttribute(ts_check<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ts_check<N><<<1, 128, smem>>>(dA, dB, K, dD);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<int32_t> D((size_t)M * N);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  long bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      int32_t ref = 0;
      for (int k = 0; k < K; ++k) ref += (int32_t)A[(size_t)m * K + k] * (int32_t)B[(size_t)n * K + k];
      if (ref != D[(size_t)m * N + n]) {
        if (bad < 4) printf("  mismatch m=%d n=%d got %d ref %d\n", m, n, D[(size_t)m * N + n], ref);
        ++bad;
      }
    }
  printf("TS check N=%d: %s, mismatches %ld of %d\n", N, cudaGetErrorString(e), bad, M * N);
  cudaFree(dA), cudaFree(dB), cudaFree(dD);
  return bad != 0;
}

// ---------------- rate ----------------
// warp 0 lane 0 issues MMAs (two accumulators alternate, as the xi / e pair would); warps 4..11 optionally hammer
// tcgen05.st into other TMEM columns, and optionally also 16-byte shared-memory stores, to see what disturbs the pipe.
template <int N, int NISS>
__global__ void __launch_bounds__(384, 1) ts_rate(int iters, int st_mode, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 2 * N * 128 / 4; i += 384) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(NISS));
    asm volatile("fence.mbarrier_init.release.cluster;\n");
    stop = 0;
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
  const uint32_t a_col = 2 * N <= 256 ? 256 : 480;
  if (warp >= 4) {
    uint32_t v[8];
    for (int i = 0; i < 8; ++i) v[i] = 0x01010101u;
    st8(tmem + ((uint32_t)((warp & 3) * 32) << 16) + a_col, v);
    st8(tmem + ((uint32_t)((warp & 3) * 32) << 16) + a_col + 8, v);
    st8(tmem + ((uint32_t)((warp & 3) * 32) << 16) + a_col + 16, v);
    st8(tmem + ((uint32_t)((warp & 3) * 32) << 16) + a_col + 24, v);
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  if (warp < NISS) {
    if (lane == 0) {
      const long long t0 = clock64();
      for (int it = warp; it < iters; it += NISS) {
        for (int k = 0; k < 4; ++k) {
          const uint64_t db = make_desc(smem_u32(smem + (it & 1) * N * 128) + k * 32);
          const uint32_t a = tmem + a_col + (2 * N <= 256 ? k * 8 : (k & 1) * 8);
          mma_ts(tmem, a, db, idesc, 1u);
          if (2 * N <= 512 - 32) mma_ts(tmem + N, a, db, idesc, 1u);
        }
      }
      commit(&bar);
      wait(&bar, 0);
      if (warp == 0) {
        cycles[blockIdx.x] = clock64() - t0;
        stop = 1;
      }
    }
    __syncwarp();
  } else if (warp >= 4 && st_mode) {
    uint32_t v[8];
    for (int i = 0; i < 8; ++i) v[i] = 0x01010101u;
    const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (2 * N <= 256 ? 384 : 496) + (warp >> 2 & 1) * 8;
    while (!stop) {
      if (st_mode & 1) {
        st8(ta, v);
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
      }
      if (st_mode & 2) {
        *reinterpret_cast<uint4*>(smem + 2 * N * 128 + tid * 16) = make_uint4(v[0], v[1], v[2], v[3]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512));
}

template <int N, int NISS>
void rate(int st_mode, int iters = 2000) {
  long long* d;
  cudaMalloc(&d, 148 * 8);
  const size_t smem = (size_t)2 * N * 128 + 384 * 16 + 1024;
  cudaFuncSetAttribute(ts_rate<N, NISS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  ts_rate<N, NISS><<<148, 384, smem>>>(16, st_mode, d);
  cudaEventRecord(e0);
  ts_rate<N, NISS><<<148, 384, smem>>>(iters, st_mode, d);
  cudaEventRecord(e1);
  cudaError_t e = cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long c;
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  const int per_it = (2 * N <= 480) ? 8 : 4;
  const double macs = 148.0 * iters * per_it * 128.0 * N * 32;
  printf("TS rate N=%3d issuers=%d st_mode=%d iters=%d  %s  %.3f ms  %.1f clk/MMA  %.2f Pop/s  SM clock %.0f MHz\n", N, NISS, st_mode,
         iters, cudaGetErrorString(e), ms, (double)c / (iters * (double)per_it), 2 * macs / ms / 1e12, (double)c / ms / 1e3);
  cudaFree(d);
}

int main() {
  int bad = check<128>();
  bad |= check<256>();
  bad |= check<144>();
  for (int m = 0; m < 4; ++m) rate<128, 1>(m);
  for (int m = 0; m < 4; ++m) rate<128, 2>(m);
  rate<256, 1>(0);
  rate<256, 1>(1);
  rate<64, 1>(0);
  rate<128, 2>(0, 20000);  // ~6 ms: long enough for the power management to react
  rate<128, 2>(3, 100000);
  rate<112, 1>(0);
  rate<96, 1>(0);
  return bad;
}
