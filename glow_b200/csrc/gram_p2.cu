// Instantiates the contraction kernels for GLR_FMT_PLINK2 (one translation unit per format so they build in parallel).
#include "contract.cuh"

namespace glr {
void launch_gram_p2(const GramParams& p, cudaStream_t st) {
  if (p.nt == 4) {
    launch_gram_mt<GLR_FMT_PLINK2, 4>(p, st);
    return;
  }
  launch_gram_mt<GLR_FMT_PLINK2, 2>(p, st);
}
void launch_masked_p2(const MaskedParams& p, cudaStream_t st) { launch_masked_mt<GLR_FMT_PLINK2>(p, st); }
}  // namespace glr
