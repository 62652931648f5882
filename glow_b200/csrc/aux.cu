// Memory-bound and once-per-call kernels: decode pre-pass, split reduction, statistics epilogue,
// operand packing.  The FP64 tensor-core contractions live in contract.cuh.
#include <algorithm>

#include "contract.cuh"

namespace glr {

// ------------------------------------------------------------------------------------------------
// element decode shared by the non-GEMM kernels (A0 semantics)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double decode_elem(const BlockView& x, const uint8_t* row, int64_t s, double miss) {
  switch (x.format) {
    case GLR_FMT_F64: return reinterpret_cast<const double*>(row)[s];
    case GLR_FMT_F32: return (double)reinterpret_cast<const float*>(row)[s];
    case GLR_FMT_I8: {
      const int v = (int)reinterpret_cast<const int8_t*>(row)[s];
      return v == -1 ? miss : (double)v;
    }
    default: {
      const uint32_t code = (row[s >> 2] >> (2 * (s & 3))) & 3u;
      return XLoader<GLR_FMT_PLINK2, 1>::dec(code, miss);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// pre-pass (PLINK2 / I8): one CTA per variant; integer counts -> missing substitute + sum x^2.
// HBM-bound: reads the block once (N/4 bytes per variant for 2-bit).
// ------------------------------------------------------------------------------------------------
constexpr int kPreThreads = 256;

__device__ __forceinline__ long long block_sum_ll(long long v, long long* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  long long t = 0;
  for (int w = 0; w < kPreThreads / 32; ++w) t += sh[w];
  return t;
}

__global__ void __launch_bounds__(kPreThreads) prepass_kernel(BlockView x, int policy, double* v1, double* mom,
                                                              int64_t ld) {
  __shared__ long long sh[kPreThreads / 32];
  const int g = blockIdx.x;
  const uint8_t* row = x.base + (int64_t)g * x.stride;
  const int64_t n = x.n_geno;
  long long c_miss = 0, s1 = 0, s2 = 0;  // missing count, sum x, sum x^2 over non-missing
  if (x.format == GLR_FMT_PLINK2) {
    // codes: 00 -> 2, 01 -> missing, 10 -> 1, 11 -> 0
    const int64_t full_bytes = n >> 2;
    const uint8_t* p = row;
    const uint8_t* pe = row + full_bytes;
    const uint8_t* a0 = reinterpret_cast<const uint8_t*>((reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t)15);
    if (a0 > pe) a0 = pe;
    const uint8_t* a1 = a0 + ((pe - a0) & ~(int64_t)15);
    long long c00 = 0, c01 = 0, c10 = 0;
    // 16-byte aligned middle
    const uint4* q = reinterpret_cast<const uint4*>(a0);
    const int64_t nq = (a1 - a0) >> 4;
    for (int64_t i = threadIdx.x; i < nq; i += kPreThreads) {
      const uint4 w4 = __ldg(q + i);
      const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t lo = ws[k] & 0x55555555u, hi = (ws[k] >> 1) & 0x55555555u;
        c01 += __popc(lo & ~hi);
        c10 += __popc(hi & ~lo);
        c00 += __popc(~(lo | hi) & 0x55555555u);
      }
    }
    // unaligned head / tail bytes and the final partial byte
    const int64_t head = a0 - p, tail = pe - a1;
    for (int64_t i = threadIdx.x; i < head + tail + 1; i += kPreThreads) {
      const uint8_t* bp;
      int valid = 4;
      if (i < head) {
        bp = p + i;
      } else if (i < head + tail) {
        bp = a1 + (i - head);
      } else {
        bp = pe;
        valid = (int)(n & 3);
      }
      if (valid) {
        const uint32_t b = *bp;
        for (int k = 0; k < valid; ++k) {
          const uint32_t code = (b >> (2 * k)) & 3u;
          c00 += code == 0u;
          c01 += code == 1u;
          c10 += code == 2u;
        }
      }
    }
    c_miss = c01;
    s1 = 2 * c00 + c10;
    s2 = 4 * c00 + c10;
  } else {  // I8
    const int8_t* r8 = reinterpret_cast<const int8_t*>(row);
    for (int64_t s = threadIdx.x; s < n; s += kPreThreads) {
      const int v = r8[s];
      if (v == -1) {
        ++c_miss;
      } else {
        s1 += v;
        s2 += v * v;
      }
    }
  }
  c_miss = block_sum_ll(c_miss, sh);
  s1 = block_sum_ll(s1, sh);
  s2 = block_sum_ll(s2, sh);
  if (threadIdx.x == 0) {
    double m = -1.0;
    if (policy == GLR_MISSING_MEAN && c_miss < n) m = (double)s1 / (double)(n - c_miss);  // MeanSubstitute.scala:57-84
    v1[g] = m;
    mom[g] = (double)s2 + (m * m) * (double)c_miss;
    mom[ld + g] = (double)s1 + m * (double)c_miss;
  }
}

void launch_prepass(const BlockView& x, int missing_policy, double* v1, double* mom, int64_t ld, cudaStream_t st) {
  if (x.n_variants <= 0) return;
  prepass_kernel<<<x.n_variants, kPreThreads, 0, st>>>(x, missing_policy, v1, mom, ld);
}

// ------------------------------------------------------------------------------------------------
__global__ void reduce_splits_kernel(const double* part, double* out, int64_t total, int n_split) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  double s = part[i];
  for (int z = 1; z < n_split; ++z) s += part[(int64_t)z * total + i];
  out[i] = s;
}
void launch_reduce_splits(const double* part, double* out, int64_t rows, int64_t ld, int n_split, cudaStream_t st) {
  const int64_t total = rows * ld;
  if (total <= 0) return;
  reduce_splits_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(part, out, total, n_split);
}

// ------------------------------------------------------------------------------------------------
// epilogue: lin_reg.py:241-249 per (phenotype, variant); variants on x (coalesced), phenotypes
// strided over y.  HBM-bound on the 4 x 8 B result writes.
// ------------------------------------------------------------------------------------------------
constexpr int kEpiThreads = 256;

// xx_full[v] = sum x^2 - |Q^T x|^2  (= |x - Q Q^T x|^2, the fully observed XdotX)
//
// Digit certificate.  A column q_c stored with d digits is the exact column q_c + e_c with |e_c[s]| <= 0.51 u_c,
// u_c = max|q_c| 2^-(8d-2) (round to nearest plus the rounding of the division), and the integer contraction is exact, so
// the computed a_c = q_c.x + e_c.x with |e_c.x| <= 0.51 u_c sum|x| =: t_c, and
//   |xx_computed - xx_exact| = |sum_c (a_c^2 - (a_c - e_c.x)^2)| <= sum_c (2 |a_c| + t_c) t_c.
// sum|x| <= max(sum x, sum x^2): the entries are >= 0 (mean substitution) or in {-1, 0, 1, 2} (|x| <= x^2).
// A phenotype with missing samples takes XdotX = xx - D (complement form) or XdotX = D, D = sum over a sample subset of
// r_s^2 with r = x - Q a: the error of a moves r_s by sum_c q_c[s] (e_c.x), so (Cauchy-Schwarz, |q_c| = 1)
//   |D_computed - D_exact| <= (2 sqrt(D) + T) T,  T = sum_c t_c.
// Both bounds are written next to xx; the epilogue compares them with every test's XdotX.
__global__ void __launch_bounds__(kEpiThreads) xx_full_kernel(const double* S, const double* sxx, int K, int G,
                                                               int64_t ldS, double* xx, const double* lo_scale,
                                                               const double* sum_x_row, int n_lo, int lo_row0) {
  const int v = blockIdx.x * kEpiThreads + threadIdx.x;
  if (v >= G) return;
  double a2 = 0.0;
  for (int k = 0; k < K; ++k) {
    const double a = S[(int64_t)k * ldS + v];
    a2 = fma(a, a, a2);
  }
  xx[v] = sxx[v] - a2;
  if (lo_scale) {
    const double sabs = fmax(sum_x_row[v], sxx[v]);
    double cert = 0.0, tsum = 0.0;
    for (int c = 0; c < n_lo; ++c) {
      const double t = 0.51 * lo_scale[c] * sabs;
      cert += (2.0 * fabs(S[(int64_t)(lo_row0 + c) * ldS + v]) + t) * t;
      tsum += t;
    }
    xx[ldS + v] = cert;
    xx[2 * ldS + v] = tsum;
  }
}

// one thread per (phenotype, variant) test: the p-value loops are latency bound, so parallelism
// across tests (not work per thread) is what hides them
__global__ void __launch_bounds__(kEpiThreads) epilogue_kernel(const EpilogueParams p, const double* xx_full) {
  const int v = blockIdx.x * kEpiThreads + threadIdx.x;
  const int ph = blockIdx.y;
  if (v >= p.G) return;
  const int mid = p.mask_id[ph];
  const double XdotX = mid < 0 ? xx_full[v] : (p.d_is_complement ? xx_full[v] - p.D[(int64_t)mid * p.ldS + v] : p.D[(int64_t)mid * p.ldS + v]);
  const double XdotY = p.S[(int64_t)(p.K + ph) * p.ldS + v];
  // (a variant that is numerically inside the span of the covariates has no meaningful statistics in any arithmetic)
  if (p.lo_scale) {
    double cert = (mid < 0 || p.d_is_complement) ? xx_full[p.ldS + v] : 0.0;
    if (mid >= 0) {
      const double T = xx_full[2 * p.ldS + v];
      cert += (2.0 * sqrt(fabs(p.D[(int64_t)mid * p.ldS + v])) + T) * T;
    }
    if (cert > kDigitCertTol * XdotX && XdotX > 1e-9 * p.sxx[v]) atomicOr(p.flags, 4);
  }
  const double recip = 1.0 / XdotX;                                             // lin_reg.py:241
  const double beta = XdotY * recip;                                            // :242
  const double se = sqrt((p.YdotY[ph] * recip - beta * beta) / p.dof);          // :243
  const double T = beta / se;                                                   // :244
  const double pv = t_two_sided_pvalue(T, p.pc);                                // :245
  const int64_t o = (int64_t)ph * p.G + v;
  p.effect[o] = beta * p.Y_scale[ph];                                           // :247-248
  p.stderror[o] = se * p.Y_scale[ph];                                           // :249
  p.tvalue[o] = T;
  p.pvalue[o] = pv;
  if (p.verbose) {                                                              // :234-237
    p.n[o] = p.n_obs[ph];
    p.sum_x[o] = p.S[(int64_t)(p.K + p.P + ph) * p.ldS + v];
    p.ytx[o] = p.S[(int64_t)(p.K + 2 * p.P + ph) * p.ldS + v];
  }
}
void launch_epilogue(const EpilogueParams& p, double* xx_scratch, cudaStream_t st) {
  if (p.G <= 0 || p.P <= 0) return;
  const unsigned gx = (unsigned)ceil_div(p.G, kEpiThreads);
  xx_full_kernel<<<gx, kEpiThreads, 0, st>>>(p.S, p.sxx, p.K, p.G, p.ldS, xx_scratch, p.lo_scale, p.sum_x_row, p.n_lo,
                                             p.lo_row0);
  for (int p0 = 0; p0 < p.P; p0 += 65535) {  // gridDim.y limit
    EpilogueParams q = p;
    const int np = p.P - p0 < 65535 ? p.P - p0 : 65535;
    q.P = p.P;
    q.mask_id += p0;
    q.YdotY += p0;
    q.Y_scale += p0;
    q.n_obs += p0;
    q.effect += (int64_t)p0 * p.G;
    q.stderror += (int64_t)p0 * p.G;
    q.tvalue += (int64_t)p0 * p.G;
    q.pvalue += (int64_t)p0 * p.G;
    if (p.verbose) {
      q.n += (int64_t)p0 * p.G;
      q.sum_x += (int64_t)p0 * p.G;
      q.ytx += (int64_t)p0 * p.G;
    }
    q.S = p.S + (int64_t)p0 * p.ldS;  // rows K+ph, K+P+ph, K+2P+ph all shift by p0
    epilogue_kernel<<<dim3(gx, (unsigned)np), kEpiThreads, 0, st>>>(q, xx_scratch);
  }
}

__global__ void pvalues_kernel(const double* t, int64_t n, PvalConst pc, double* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = t_two_sided_pvalue(t[i], pc);
}
void launch_pvalues(const double* t, int64_t n, PvalConst pc, double* p, cudaStream_t st) {
  if (n <= 0) return;
  pvalues_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, st>>>(t, n, pc, p);
}

__global__ void decode_kernel(BlockView x, double* out) {
  const int g = blockIdx.y;
  const uint8_t* row = x.base + (int64_t)g * x.stride;
  const double miss = (x.format == GLR_FMT_I8 || x.format == GLR_FMT_PLINK2) ? x.v1[g] : 0.0;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < x.n_geno; s += (int64_t)gridDim.x * blockDim.x)
    out[(int64_t)g * x.n_geno + s] = decode_elem(x, row, s, miss);
}
void launch_decode(const BlockView& x, double* out, cudaStream_t st) {
  if (x.n_variants <= 0 || x.n_geno <= 0) return;
  dim3 grid((unsigned)std::min<int64_t>(ceil_div(x.n_geno, 256), (int64_t)64), (unsigned)x.n_variants);
  decode_kernel<<<grid, 256, 0, st>>>(x, out);
}

__global__ void dropped_moments_kernel(BlockView x, const int64_t* drop_idx, int64_t n_drop, double* mom, int64_t ld) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= x.n_variants) return;
  const uint8_t* row = x.base + (int64_t)g * x.stride;
  const double miss = (x.format == GLR_FMT_I8 || x.format == GLR_FMT_PLINK2) ? x.v1[g] : 0.0;
  double s2 = 0.0, s1 = 0.0;
  for (int64_t i = 0; i < n_drop; ++i) {
    const double v = decode_elem(x, row, drop_idx[i], miss);
    s2 = fma(v, v, s2);
    s1 += v;
  }
  mom[g] -= s2;
  mom[ld + g] -= s1;
}
void launch_dropped_moments(const BlockView& x, const int64_t* drop_idx, int64_t n_drop, double* mom, int64_t ld,
                            cudaStream_t st) {
  if (x.n_variants <= 0 || n_drop <= 0) return;
  dropped_moments_kernel<<<(unsigned)ceil_div(x.n_variants, 128), 128, 0, st>>>(x, drop_idx, n_drop, mom, ld);
}

// Float blocks with dropped samples: the reference removes those rows before any arithmetic (np.delete, lin_reg.py:228-229),
// so whatever they hold (NaN included) must not reach a sum.  The block is staged in library memory and the dropped
// entries are set to zero there; W has zero rows for them and their moments are then zero as well.
__global__ void zero_dropped_kernel(uint8_t* base, int64_t stride, int G, int elem_bytes, const int64_t* drop_idx, int64_t n_drop) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)G * n_drop) return;
  const int64_t g = idx / n_drop, i = idx - g * n_drop;
  uint8_t* row = base + g * stride;
  if (elem_bytes == 8) reinterpret_cast<double*>(row)[drop_idx[i]] = 0.0;
  else reinterpret_cast<float*>(row)[drop_idx[i]] = 0.0f;
}
void launch_zero_dropped(const uint8_t* base, int64_t stride, int G, int format, const int64_t* drop_idx, int64_t n_drop,
                         cudaStream_t st) {
  if (G <= 0 || n_drop <= 0) return;
  const int64_t total = (int64_t)G * n_drop;
  zero_dropped_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(const_cast<uint8_t*>(base), stride, G,
                                                                       format == GLR_FMT_F64 ? 8 : 4, drop_idx, n_drop);
}

// ------------------------------------------------------------------------------------------------
// BGEN v1.2 layout-2 probability blocks (unphased, diploid, biallelic) -> Glow hard calls -> PLINK 2-bit rows on the
// device.  Semantics of the reference's BGEN datasource + hard_calls + genotype_states
// (core/.../bgen/BgenFileIterator.scala, BgenSchemaInferrer.scala:42 hardCallThreshold = 0.9,
// core/.../sql/expressions/VariantUtilExprs.scala:37-59): p = stored integer / (2^bits - 1); the call is the most
// probable genotype if its probability reaches the threshold, else missing; a sample flagged missing (top bit of its
// ploidy byte) is missing; state = number of alternate alleles.  Thread = 4 samples of a variant = one output byte
// (state 0 -> code 11, 1 -> 10, 2 -> 00, missing -> 01: PlinkRowToInternalRowConverter.scala:40-47 inverted).
template <typename T>
__global__ void bgen_to_plink2_kernel(const T* probs, const uint8_t* ploidy, int G, int64_t N, double denom, double threshold,
                                      uint8_t* out, int64_t out_stride) {
  const int64_t row_bytes = (N + 3) >> 2;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)G * row_bytes) return;
  const int64_t g = idx / row_bytes, bpos = idx - g * row_bytes;
  const T* pr = probs + g * 2 * N;
  const uint8_t* pl = ploidy + g * N;
  uint32_t byte = 0;
  for (int k = 0; k < 4; ++k) {
    const int64_t s = bpos * 4 + k;
    uint32_t code = 0;  // padding bits of the last byte stay 0, as in a .bed file
    if (s < N) {
      const double p_aa = (double)pr[2 * s] / denom, p_ab = (double)pr[2 * s + 1] / denom;
      const double p_bb = 1.0 - p_aa - p_ab;
      int best = 0;
      double pmax = p_aa;
      if (p_ab > pmax) {
        pmax = p_ab;
        best = 1;
      }
      if (p_bb > pmax) {
        pmax = p_bb;
        best = 2;
      }
      const bool called = pmax >= threshold && !(pl[s] & 0x80);
      code = !called ? 1u : (best == 0 ? 3u : (best == 1 ? 2u : 0u));
    }
    byte |= code << (2 * k);
  }
  out[g * out_stride + bpos] = (uint8_t)byte;
}
void launch_bgen_to_plink2(const void* probs, const uint8_t* ploidy, int G, int64_t N, int bits, double threshold, uint8_t* out,
                           int64_t out_stride, cudaStream_t st) {
  if (G <= 0) return;
  const int64_t total = (int64_t)G * ((N + 3) >> 2);
  const unsigned grid = (unsigned)ceil_div(total, 256);
  const double denom = std::ldexp(1.0, bits) - 1.0;
  if (bits == 8) bgen_to_plink2_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t*)probs, ploidy, G, N, denom, threshold, out, out_stride);
  else if (bits == 16) bgen_to_plink2_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t*)probs, ploidy, G, N, denom, threshold, out, out_stride);
  else bgen_to_plink2_kernel<uint32_t><<<grid, 256, 0, st>>>((const uint32_t*)probs, ploidy, G, N, denom, threshold, out, out_stride);
}

// ------------------------------------------------------------------------------------------------
// GloWGR LOCO apply step: RidgeRegression.transform_loco (python/glow/wgr/
// ridge_regression.py:199-234) applies the fitted model with the headers of one chromosome removed, once per
// chromosome; per (sample block, label) the apply step is apply_model (python/glow/wgr/ridge_udfs.py:259-347):
// X = [covariates | (X_raw - mu) / sig] (assemble_block, model_functions.py:118-158), prediction = X @ B.
// Warp = one sample.  The headers are visited chromosome by chromosome (perm = header indices grouped by chromosome,
// seg = group boundaries); the per-chromosome partial sums are then combined so that offsets[c] = covariate term +
// sum of the partials of every OTHER chromosome.
__global__ void __launch_bounds__(256) loco_predict_kernel(const double* X, int64_t N, int H, const double* mu, const double* sig,
                                                           const double* B, const int32_t* perm, const int32_t* seg, int n_chr,
                                                           const int32_t* sample_block, const double* cov, int Kc,
                                                           const double* Bcov, double* out) {
  extern __shared__ double part_s[];  // [warps][n_chr]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (s >= N) return;
  const int sb = sample_block[s];
  const double* xr = X + s * H;
  const double* br = B + (int64_t)sb * H;
  double* part = part_s + warp * n_chr;
  for (int c = 0; c < n_chr; ++c) {
    double acc = 0.0;
    for (int i = seg[c] + lane; i < seg[c + 1]; i += 32) {
      const int h = perm[i];
      acc = fma((xr[h] - mu[h]) / sig[h], br[h], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) part[c] = acc;
  }
  double cterm = 0.0;
  for (int k = lane; k < Kc; k += 32) cterm = fma(cov[s * Kc + k], Bcov[(int64_t)sb * Kc + k], cterm);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cterm += __shfl_xor_sync(0xffffffffu, cterm, o);
  __syncwarp();
  for (int c = lane; c < n_chr; c += 32) {
    double acc = cterm;
    for (int c2 = 0; c2 < n_chr; ++c2)
      if (c2 != c) acc += part[c2];
    out[(int64_t)c * N + s] = acc;
  }
}
void launch_loco_predict(const double* X, int64_t N, int H, const double* mu, const double* sig, const double* B, const int32_t* perm,
                         const int32_t* seg, int n_chr, const int32_t* sample_block, const double* cov, int Kc, const double* Bcov,
                         double* out, cudaStream_t st) {
  if (N <= 0 || n_chr <= 0) return;
  const int warps = 8;
  loco_predict_kernel<<<(unsigned)ceil_div(N, warps), warps * 32, (size_t)warps * n_chr * sizeof(double), st>>>(
      X, N, H, mu, sig, B, perm, seg, n_chr, sample_block, cov, Kc, Bcov, out);
}

__global__ void intercept_row_kernel(const double* sum_x, double q0, double* row0, int G) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < G) row0[v] = q0 * sum_x[v];
}
void launch_intercept_row(const double* sum_x, double q0, double* row0, int G, cudaStream_t st) {
  if (G <= 0) return;
  intercept_row_kernel<<<(unsigned)ceil_div(G, 256), 256, 0, st>>>(sum_x, q0, row0, G);
}

// ------------------------------------------------------------------------------------------------
// state packing (runs once per linear_regression call)
// ------------------------------------------------------------------------------------------------
// T[k][p] = sum_s Q[s][k] Y[s][p]; one CTA per (p, k), fixed-order tree reduction (deterministic)
__global__ void __launch_bounds__(256) qty_kernel(const double* Q, const double* Y, int64_t N, int K, int P, double* T) {
  __shared__ double sh[256];
  const int ph = blockIdx.x, k = blockIdx.y;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < N; i += 256) s = fma(Q[i * K + k], Y[i * P + ph], s);
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) T[(int64_t)k * P + ph] = sh[0];
}
void launch_qty(const double* Q, const double* Y, int64_t N, int K, int P, double* T, cudaStream_t st) {
  if (K <= 0 || P <= 0) return;
  qty_kernel<<<dim3((unsigned)P, (unsigned)K), 256, 0, st>>>(Q, Y, N, K, P, T);
}
__global__ void residualize_kernel(const double* Q, const double* T, double* Y, int64_t N, int K, int P) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * P) return;
  const int64_t s = i / P;
  const int ph = (int)(i % P);
  double acc = 0.0;
  for (int k = 0; k < K; ++k) acc = fma(Q[s * K + k], T[(int64_t)k * P + ph], acc);
  Y[i] -= acc;
}
void launch_residualize(const double* Q, const double* T, double* Y, int64_t N, int K, int P, cudaStream_t st) {
  if (N * P <= 0) return;
  residualize_kernel<<<(unsigned)ceil_div(N * P, 256), 256, 0, st>>>(Q, T, Y, N, K, P);
}

// dst layout: [group][chunk][8 sample groups][MT][64]; element (lane*2 + h) of an atom =
// M[s0 + 2*(lane&3) + h][col0 + (lane>>2)]
__global__ void pack_a_kernel(const double* src, int64_t N, int64_t ld, int c0, int ncols, const int64_t* geno_to_row,
                              int64_t Ng, double* dst, int dst_col0, int MT, int tail, int64_t n_chunks) {
  const int64_t total = (int64_t)ncols * n_chunks * kChunk;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % ncols);       // source column (fastest: coalesced reads of a row)
  const int64_t sg = i / ncols;         // genotype-sample index, padded
  double v = 0.0;
  if (sg < Ng) {
    const int64_t row = geno_to_row ? geno_to_row[sg] : sg;
    if (row >= 0 && row < N) v = src[row * ld + c0 + c];
  }
  const int dcol = dst_col0 + c;
  const int64_t chunk = sg / kChunk;
  const int g = (int)((sg % kChunk) / kGroup);
  const int w = (int)(sg % kGroup);
  const int64_t stage_doubles = (int64_t)kChunkGroups * (MT * kAtomDoubles + tail * kGroup);
  if (tail > 0 && dcol >= MT * 8) {  // tail column (single column group)
    const int t = dcol - MT * 8;
    dst[chunk * stage_doubles + (int64_t)kChunkGroups * MT * kAtomDoubles + (g * tail + t) * kGroup + w] = v;
    return;
  }
  const int group = dcol / (MT * 8), mt = (dcol / 8) % MT, r = dcol % 8;
  const int j = w >> 1, h = w & 1;
  dst[(group * n_chunks + chunk) * stage_doubles + ((int64_t)g * MT + mt) * kAtomDoubles + (4 * r + j) * 2 + h] = v;
}
void launch_pack_a(const double* src, int64_t N, int64_t ld, int c0, int ncols, const int64_t* geno_to_row,
                   int64_t Ng, double* dst, int dst_col0, int MT, int tail, int64_t n_chunks, cudaStream_t st) {
  const int64_t total = (int64_t)ncols * n_chunks * kChunk;
  if (total <= 0) return;
  pack_a_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(src, N, ld, c0, ncols, geno_to_row, Ng, dst, dst_col0,
                                                                 MT, tail, n_chunks);
}

// dst layout: [chunk][8 sample groups][KT2][64]; element (lane*2 + h) = Q[s0 + (lane>>2)][4*(2*k2 + h) + (lane&3)]
__global__ void pack_qb_kernel(const double* Q, int64_t N, int K, const int64_t* geno_to_row, int64_t Ng, double* dst,
                               int KT2, int64_t n_chunks) {
  const int kpad = KT2 * 8;
  const int64_t total = (int64_t)kpad * n_chunks * kChunk;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int k = (int)(i % kpad);
  const int64_t sg = i / kpad;
  double v = 0.0;
  if (sg < Ng && k < K) {
    const int64_t row = geno_to_row ? geno_to_row[sg] : sg;
    if (row >= 0 && row < N) v = Q[row * K + k];
  }
  const int64_t chunk = sg / kChunk;
  const int g = (int)((sg % kChunk) / kGroup);
  const int n = (int)(sg % kGroup);          // sample within group = output column of the residual product
  const int kk = k >> 2, jj = k & 3;         // k-step, lane column
  const int k2 = kk >> 1, h = kk & 1;
  const int64_t atom = (chunk * kChunkGroups + g) * KT2 + k2;
  dst[atom * kAtomDoubles + (4 * n + jj) * 2 + h] = v;
}
void launch_pack_qb(const double* Q, int64_t N, int K, const int64_t* geno_to_row, int64_t Ng, double* dst, int KT2,
                    int64_t n_chunks, cudaStream_t st) {
  const int64_t total = (int64_t)KT2 * 8 * n_chunks * kChunk;
  if (total <= 0) return;
  pack_qb_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(Q, N, K, geno_to_row, Ng, dst, KT2, n_chunks);
}

// ------------------------------------------------------------------------------------------------
int gram_tile_variants(int nt, int warps) { return warps * nt * 8; }

void launch_gram(const GramParams& p, cudaStream_t st) {
  if (p.x.n_variants <= 0) return;
  switch (p.x.format) {
    case GLR_FMT_F64: launch_gram_f64(p, st); break;
    case GLR_FMT_F32: launch_gram_f32(p, st); break;
    case GLR_FMT_I8: launch_gram_i8(p, st); break;
    default: launch_gram_p2(p, st); break;
  }
}
void launch_masked(const MaskedParams& p, cudaStream_t st) {
  if (p.x.n_variants <= 0) return;
  switch (p.x.format) {
    case GLR_FMT_F64: launch_masked_f64(p, st); break;
    case GLR_FMT_F32: launch_masked_f32(p, st); break;
    case GLR_FMT_I8: launch_masked_i8(p, st); break;
    default: launch_masked_p2(p, st); break;
  }
}

}  // namespace glr
