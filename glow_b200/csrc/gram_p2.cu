// Instantiates the contraction kernels for GLR_FMT_PLINK2 (one translation unit per format so they build in parallel).
#include "contract.cuh"

namespace glr {
void launch_gram_p2(const GramParams& p, cudaStream_t st) {
  // 16 consumer warps x 2 n-tiles: 4 warps per SM sub-partition hide the LDS / decode latency
  // behind the DMMA pipe better than 8 warps x 4 n-tiles (measured 24.2 vs 25.4 ms per 37 888
  // variants x 500 k samples); both tile 256 variants per CTA.
  if (p.warps == 16) {
    launch_gram_mt<GLR_FMT_PLINK2, 2, true, 16>(p, st);
    return;
  }
  launch_gram_mt<GLR_FMT_PLINK2, 2, true>(p, st);
}
void launch_masked_p2(const MaskedParams& p, cudaStream_t st) { launch_masked_mt<GLR_FMT_PLINK2>(p, st); }
}  // namespace glr
