// Instantiates the contraction kernels for GLR_FMT_F64 (one translation unit per format so they build in parallel).
#include "contract.cuh"

namespace glr {
void launch_gram_f64(const GramParams& p, cudaStream_t st) {
  // (Tried for dosage blocks with few columns, configs[1]: 12 and 16 consumer warps, to hide the FP64 pipe latencies that
  // 2 warps per SM sub-partition leave exposed -- the kernel needs ~168 registers and spills at the 128 / 96 those shapes
  // allow: 3.3 ms instead of 2.2 ms; 16 warps x 1 n-tile fits 96 registers but reads every A fragment twice: 2.42 ms.)
  launch_gram_mt<GLR_FMT_F64, 2, true>(p, st);
}
void launch_masked_f64(const MaskedParams& p, cudaStream_t st) { launch_masked_mt<GLR_FMT_F64>(p, st); }
}  // namespace glr
