// Microbenchmark: FP64 FMA vs DMMA.8x8x4 throughput on sm_100a, alone and mixed in one warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu && ./fp64_pipes
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NDMMA, int NDFMA>
__global__ void __launch_bounds__(512) mix(double* out, int iters, double a, double b) {
  double c[8][2], f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; f[i] = i + threadIdx.x; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NDMMA; ++i) dmma(c[i % 8][0], c[i % 8][1], a, b);
#pragma unroll
    for (int i = 0; i < NDFMA; ++i) f[i % 8] = fma(f[i % 8], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// fine interleave: one dfma (or dependent pair) after every dmma, 4 independent fma chains
template <int PAIR>
__global__ void __launch_bounds__(512) interleaved(double* out, int iters, double a, double b) {
  double c[8][2], f[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
#pragma unroll
  for (int i = 0; i < 4; ++i) f[i] = i + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dmma(c[i][0], c[i][1], a, b);
      if (PAIR) { if (i % 2 == 0) f[(i / 2) % 4] = fma(a, f[(i / 2) % 4], fma(b, a, f[(i / 2) % 4])); }
      else f[i % 4] = fma(f[i % 4], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < 4; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int PAIR>
void run_i(const char* name, int warps) {
  double* out;
  cudaMalloc(&out, 148 * 1024 * 8);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  interleaved<PAIR><<<148, warps * 32>>>(out, 100, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  interleaved<PAIR><<<148, warps * 32>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma_dmma = 148.0 * warps * iters * 8 * 256.0, fma_dfma = 148.0 * warps * iters * 8 * 32.0;
  printf("%-28s warps=%2d  %8.3f ms  %6.2f TFLOP/s  clk/iter/warp=%.1f\n", name, warps, ms,
         2 * (fma_dmma + fma_dfma) / ms / 1e9, ms * 1e-3 * 1.965e9 / iters / (warps / 4.0));
  cudaFree(out);
}

template <int NDMMA, int NDFMA>
void run(const char* name, int warps) {
  double* out;
  cudaMalloc(&out, 148 * 1024 * 8);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  mix<NDMMA, NDFMA><<<148, warps * 32>>>(out, 100, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  mix<NDMMA, NDFMA><<<148, warps * 32>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma_dmma = 148.0 * warps * iters * NDMMA * 256.0, fma_dfma = 148.0 * warps * iters * NDFMA * 32.0;
  double clk_per_iter_per_smsp = ms * 1e-3 * 1.965e9 / iters / (warps / 4.0);
  printf("%-28s warps=%2d  %8.3f ms  %6.2f TFLOP/s  (dmma %5.2f + dfma %5.2f)  clk/iter/warp=%.1f\n", name, warps, ms,
         2 * (fma_dmma + fma_dfma) / ms / 1e9, 2 * fma_dmma / ms / 1e9, 2 * fma_dfma / ms / 1e9, clk_per_iter_per_smsp);
  cudaFree(out);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<8, 0>("dmma only (8/iter)", w);
    run<0, 8>("dfma only (8/iter)", w);
    run<0, 64>("dfma only (64/iter)", w);
    run<8, 8>("8 dmma + 8 dfma", w);
    run<8, 4>("8 dmma + 4 dfma", w);
    run<16, 16>("16 dmma + 16 dfma", w);
    run_i<0>("interleaved 8x(dmma,dfma)", w);
    run_i<1>("interleaved dep. pairs", w);
  }
  return 0;
}
