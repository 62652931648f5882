// Pass 2 for states with FEW missing phenotype values (the usual case: a handful of phenotypes, each with a few percent
// of the samples unobserved):  D_u[v] = sum_s mask_u[s] r_v[s]^2  =  |r_v|^2  -  sum_{s : mask_u[s] = 0} r_v[s]^2.
// The first term is the fully observed XdotX the epilogue already has (sum x^2 - |Q^T x|^2); this kernel evaluates the
// COMPLEMENT: r_v[s] = x_v[s] - q_s . (Q^T x_v) only at the unobserved samples of each distinct mask, on the FP64
// tensor cores.  Work is 2 K flops per (variant, unobserved sample) instead of a dense N x U contraction: at the
// UK-Biobank shape with 2 % of one phenotype missing that is 10 000 samples per variant instead of 500 000.
//
// Block = 8 warps, each with NV blocks of 8 variants (256 variants per block for NV = 4); grid.y splits a mask's sorted
// list of unobserved samples, grid.z walks the masks.  The Q rows of 128 list entries are staged in shared memory; the
// genotypes of the entries are gathered from the packed rows (the list is sorted, so a lane walks forward through its
// rows and re-uses the sectors it has pulled into L1).
#include <algorithm>
#include <cstdlib>

#include "../../include/glow_b200.h"
#include "common.cuh"
#include "contract.cuh"
#include "kernels.h"

namespace glr {

constexpr int kSpThreads = 256;
constexpr int kSpTile = 128;

template <int FMT>
__device__ __forceinline__ double sp_elem(const uint8_t* row, int64_t s, double miss) {
  if (FMT == GLR_FMT_F64) return __ldg(reinterpret_cast<const double*>(row) + s);
  if (FMT == GLR_FMT_F32) return (double)__ldg(reinterpret_cast<const float*>(row) + s);
  if (FMT == GLR_FMT_I8) {
    const int v = (int)__ldg(reinterpret_cast<const int8_t*>(row) + s);
    return v == -1 ? miss : (double)v;
  }
  const uint32_t code = ((uint32_t)__ldg(row + (s >> 2)) >> (2 * (s & 3))) & 3u;
  return XLoader<GLR_FMT_PLINK2, 1>::dec(code, miss);
}

// The residuals of 8 variants x 8 list entries are one m8n8k4 accumulator tile:
//   C[variant r][entry c] = x  -  sum_k a[variant r][k] Q[entry c][k]     (ceil(K/4) DMMAs, C initialised with x)
// A operand = -a (per lane KT4 values per variant block, kept in registers for the whole kernel), B operand = the Q rows of
// the 8 entries (one LDS.64 per k-step from the tile in shared memory, re-used by the warp's NV variant blocks), and
// the lane then squares its two residuals into a per-variant-row accumulator.  (A first version with one DFMA chain per
// (variant, entry) kept the FP64 pipe 14 % busy: one DMMA per 64 x 4 multiply-adds instead of one DFMA per 32 is 5x
// fewer instructions for the same FP64 datapath time.)
// Lane (r, j) = 4 r + j of the m8n8k4 fragment layout: A[r][j], B[j][r], C[r][2j..2j+1].
template <int FMT, int KT4, int NV>
__global__ void __launch_bounds__(kSpThreads) sparse_dmma_kernel(const SparseMaskedParams p) {
  extern __shared__ __align__(16) unsigned char sp_smem[];
  const int K = p.K;
  const int Ks = 4 * KT4 + 2;  // row stride of the Q tile in doubles: k-steps of 4, +2 keeps the LDS.64 at 2 wavefronts
  double* qs = reinterpret_cast<double*>(sp_smem);                        // [128][Ks], zero beyond K and beyond cnt
  int32_t* sidx = reinterpret_cast<int32_t*>(qs + (size_t)kSpTile * Ks);  // [128], -1 beyond cnt
  const int u = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, r = lane >> 2, j = lane & 3;
  const int64_t l0 = p.off[u], l1 = p.off[u + 1];
  const int64_t per = ceil_div(l1 - l0, (int64_t)gridDim.y);
  const int64_t j0 = l0 + (int64_t)blockIdx.y * per, j1 = min(l1, j0 + per);

  // this warp's NV blocks of 8 variants; lane row r
  const int v_warp = blockIdx.x * (kSpThreads / 32 * NV * 8) + warp * (NV * 8);
  const uint8_t* row[NV];
  double miss[NV], na[NV][KT4], acc[NV];
#pragma unroll
  for (int nv = 0; nv < NV; ++nv) {
    const int v = min(v_warp + nv * 8 + r, p.x.n_variants - 1);
    row[nv] = p.x.base + (int64_t)v * p.x.stride;
    miss[nv] = (FMT == GLR_FMT_I8 || FMT == GLR_FMT_PLINK2) ? p.x.v1[v] : 0.0;
    acc[nv] = 0.0;
#pragma unroll
    for (int kk = 0; kk < KT4; ++kk) {
      const int k = 4 * kk + j;
      na[nv][kk] = k < K ? -p.A[(int64_t)k * p.ldS + v] : 0.0;
    }
  }

  for (int64_t t0 = j0; t0 < j1; t0 += kSpTile) {
    const int cnt = (int)min((int64_t)kSpTile, j1 - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < kSpTile * Ks; i += kSpThreads) {
      const int e = i / Ks, k = i - e * Ks;
      qs[i] = (e < cnt && k < K) ? p.qm[(t0 + e) * ((K + 1) & ~1) + k] : 0.0;
    }
    if ((int)threadIdx.x < kSpTile) sidx[threadIdx.x] = (int)threadIdx.x < cnt ? p.idx[t0 + threadIdx.x] : -1;
    __syncthreads();
    const int n_groups = (cnt + 7) >> 3;
    for (int g = 0; g < n_groups; ++g) {
      const int e0 = g * 8;
      // C tiles start as x at (variant row r, entries e0 + 2j, e0 + 2j + 1); padding entries (-1) enter as 0
      const int s_a = sidx[e0 + 2 * j], s_b = sidx[e0 + 2 * j + 1];
      double c0[NV], c1[NV];
#pragma unroll
      for (int nv = 0; nv < NV; ++nv) {
        c0[nv] = s_a >= 0 ? sp_elem<FMT>(row[nv], (int64_t)s_a, miss[nv]) : 0.0;
        c1[nv] = s_b >= 0 ? sp_elem<FMT>(row[nv], (int64_t)s_b, miss[nv]) : 0.0;
      }
      const double* qrow = qs + (size_t)(e0 + r) * Ks + j;  // B[j][r] = Q[entry e0 + r][4 kk + j]
#pragma unroll
      for (int kk = 0; kk < KT4; ++kk) {
        if (4 * kk < K) {
          const double b = qrow[4 * kk];
#pragma unroll
          for (int nv = 0; nv < NV; ++nv) dmma884(c0[nv], c1[nv], na[nv][kk], b);
        }
      }
#pragma unroll
      for (int nv = 0; nv < NV; ++nv) acc[nv] = fma(c0[nv], c0[nv], fma(c1[nv], c1[nv], acc[nv]));
    }
  }
  // the 4 lanes of a variant row hold the sums over their entry columns
#pragma unroll
  for (int nv = 0; nv < NV; ++nv) {
    double t = acc[nv];
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    const int v = v_warp + nv * 8 + r;
    if (j == 0 && v < p.x.n_variants) p.comp[((int64_t)blockIdx.y * p.U + u) * p.ldS + v] = t;
  }
}

// Generic form for any number of covariates (K > 64, where the coefficients no longer fit the DMMA kernel's registers):
// thread = variant, plain FP64 FMAs.  The coefficient reads a[k][v] are coalesced across the warp (consecutive variants),
// the Q row of the list entry is a broadcast read.  Same output as sparse_dmma_kernel.
template <int FMT>
__global__ void __launch_bounds__(kSpThreads) sparse_generic_kernel(const SparseMaskedParams p) {
  const int K = p.K, Kp = (K + 1) & ~1;
  const int u = blockIdx.z;
  const int64_t l0 = p.off[u], l1 = p.off[u + 1];
  const int64_t per = ceil_div(l1 - l0, (int64_t)gridDim.y);
  const int64_t j0 = l0 + (int64_t)blockIdx.y * per, j1 = min(l1, j0 + per);
  const int v = blockIdx.x * kSpThreads + threadIdx.x;
  if (v >= p.x.n_variants) return;
  const uint8_t* row = p.x.base + (int64_t)v * p.x.stride;
  const double miss = (FMT == GLR_FMT_I8 || FMT == GLR_FMT_PLINK2) ? p.x.v1[v] : 0.0;
  const double* a = p.A + v;
  double acc = 0.0;
  for (int64_t e = j0; e < j1; ++e) {
    double r = sp_elem<FMT>(row, (int64_t)p.idx[e], miss);
    const double* q = p.qm + e * Kp;
    for (int k = 0; k < K; ++k) r = fma(-__ldg(a + (int64_t)k * p.ldS), __ldg(q + k), r);
    acc = fma(r, r, acc);
  }
  p.comp[((int64_t)blockIdx.y * p.U + u) * p.ldS + v] = acc;
}

template <int FMT, int KT4, int NV>
static void launch_spd(const SparseMaskedParams& p, cudaStream_t st) {
  const int Ks = 4 * KT4 + 2;
  const size_t smem = (size_t)kSpTile * Ks * sizeof(double) + kSpTile * sizeof(int32_t);
  cudaFuncSetAttribute(sparse_dmma_kernel<FMT, KT4, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int v_block = kSpThreads / 32 * NV * 8;
  const dim3 grid((unsigned)ceil_div(p.x.n_variants, v_block), (unsigned)p.n_parts, (unsigned)p.U);
  sparse_dmma_kernel<FMT, KT4, NV><<<grid, kSpThreads, smem, st>>>(p);
}
template <int FMT>
static void launch_spd_fmt(const SparseMaskedParams& p, cudaStream_t st) {
  if (p.K <= 12) launch_spd<FMT, 3, 4>(p, st);
  else if (p.K <= 20) launch_spd<FMT, 5, 4>(p, st);
  else if (p.K <= 32) launch_spd<FMT, 8, 4>(p, st);
  else if (p.K <= 64) launch_spd<FMT, 16, 2>(p, st);
  else {
    const dim3 grid((unsigned)ceil_div(p.x.n_variants, kSpThreads), (unsigned)p.n_parts, (unsigned)p.U);
    sparse_generic_kernel<FMT><<<grid, kSpThreads, 0, st>>>(p);
  }
}

void launch_sparse_masked(const SparseMaskedParams& p, cudaStream_t st) {
  if (p.x.n_variants <= 0 || p.U <= 0) return;
  switch (p.x.format) {
    case GLR_FMT_F64: launch_spd_fmt<GLR_FMT_F64>(p, st); break;
    case GLR_FMT_F32: launch_spd_fmt<GLR_FMT_F32>(p, st); break;
    case GLR_FMT_I8: launch_spd_fmt<GLR_FMT_I8>(p, st); break;
    default: launch_spd_fmt<GLR_FMT_PLINK2>(p, st); break;
  }
}

// qm[j][k] = Q[rows[j]][k] (k < K), zero padded to Kp columns
__global__ void gather_rows_kernel(const double* Q, int K, int Kp, const int64_t* rows, int64_t n, double* qm) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * Kp) return;
  const int64_t j = idx / Kp;
  const int k = (int)(idx - j * Kp);
  qm[idx] = k < K ? Q[rows[j] * K + k] : 0.0;
}
void launch_gather_rows(const double* Q, int K, int Kp, const int64_t* rows, int64_t n, double* qm, cudaStream_t st) {
  if (n <= 0) return;
  gather_rows_kernel<<<(unsigned)ceil_div(n * Kp, 256), 256, 0, st>>>(Q, K, Kp, rows, n, qm);
}

}  // namespace glr
