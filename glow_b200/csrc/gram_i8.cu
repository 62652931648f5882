// Instantiates the contraction kernels for GLR_FMT_I8 (one translation unit per format so they build in parallel).
#include "contract.cuh"

namespace glr {
void launch_gram_i8(const GramParams& p, cudaStream_t st) {
  launch_gram_mt<GLR_FMT_I8, 2, true>(p, st);
}
void launch_masked_i8(const MaskedParams& p, cudaStream_t st) { launch_masked_mt<GLR_FMT_I8>(p, st); }
}  // namespace glr
