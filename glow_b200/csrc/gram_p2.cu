// Instantiates the contraction kernels for GLR_FMT_PLINK2 (one translation unit per format so they build in parallel).
#include "contract.cuh"

namespace glr {
void launch_gram_p2(const GramParams& p, cudaStream_t st) {
  // up to 4 m-tiles: 16 consumer warps x 2 n-tiles (256 variants per CTA; 4 warps per SM sub-partition
  // keep the DMMA pipe fed; the 96-register budget of a 544-thread CTA is enough);
  // 5..8 m-tiles: 8 warps x 2 n-tiles (128 variants per CTA, 168 registers for the accumulators)
  if (p.warps == 16) {
    launch_gram_mt<GLR_FMT_PLINK2, 2, true, 16, 1, 4>(p, st);
    return;
  }
  launch_gram_mt<GLR_FMT_PLINK2, 2, true, 8, 5, 8>(p, st);
}
void launch_masked_p2(const MaskedParams& p, cudaStream_t st) { launch_masked_mt<GLR_FMT_PLINK2>(p, st); }
}  // namespace glr
