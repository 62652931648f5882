// Host-callable launchers of the CUDA kernels (internal; the public surface is include/glow_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pvalue.cuh"

namespace glr {

// Device-side view of one genotype block (A0/A1: replaces np.column_stack at lin_reg.py:226-227
// by reading the variant-major rows directly as the GEMM operand).
struct BlockView {
  const uint8_t* base;
  int64_t stride;     // bytes between variants
  int32_t n_variants; // G_b
  int32_t format;     // glr_format
  int64_t n_geno;     // samples per variant (genotype-array length)
  const double* v1;   // [G_b] value substituted for a missing call (I8 / PLINK2), else unused
};

// ---- pre-pass: per-variant missing substitute and sum of squares (I8 / PLINK2) ------------------
// v1[g]     = mean of non-missing (policy MEAN; -1 if all missing) or -1
// mom[0][g] = sum_s x^2, mom[1][g] = sum_s x, missing replaced by v1[g]   (mom is [2][ld])
void launch_prepass(const BlockView& x, int missing_policy, double* v1, double* mom, int64_t ld, cudaStream_t st);

// ---- pass 1: S = W^T X on FP64 tensor cores ----------------------------------------------------
struct GramParams {
  BlockView x;
  const double* wpk;   // packed W of this contig: [group][chunk][8 groups][MT][64]
  int32_t MT;          // m-tiles per column group (1,2,3,4,6,8)
  int32_t tail;        // 0..3 extra columns accumulated with FP64 FMAs (single column group, MT <= 4)
  int32_t n_groups;    // column groups
  int64_t n_chunks;    // 64-sample chunks (W is zero padded to a whole chunk)
  int32_t n_split;     // split of the sample range across CTAs (deterministic 2-stage reduce)
  int32_t nt;          // n-tiles (8 variants) per warp: 2 or 4
  int32_t warps;       // consumer warps per CTA: 8, or 16 (PLINK2 with nt == 2 only)
  double* S;           // first row this kernel writes; split z starts at S + z * s_slab
  int64_t s_slab;      // elements between the per-split slabs of S
  double* mom_part;    // [n_split][2][ldS]: sum x^2, sum x (F64 / F32 only, written by column group 0)
  int64_t ldS;         // >= G_b rounded up to the CTA tile
};
int gram_tile_variants(int nt, int warps = 8);  // variants per CTA
void launch_gram(const GramParams& p, cudaStream_t st);

// ---- pass 1, integer tensor-core variant (PLINK 2-bit only): exact digit-split GEMM on tcgen05 ----
struct Split8Params {
  BlockView x;           // PLINK2 block; x.v1 = missing substitute per variant (used when !fused_counts)
  const int8_t* w8;      // packed digits of this contig: [column block][stage][NC rows][128 B], 128-byte swizzled
  const double* scale;   // [n_cols] column scale = max|w_c| * 2^-(8 digits - 2)
  int32_t n_cols;        // columns of W in the GEMM (7 digit rows each, see n_lo)
  int32_t n_cb;          // column blocks (1 for <= 18 columns, 2 for <= 36)
  int32_t cpb;           // columns per block; the digit row after the last column of every block is the row of ones
  int32_t n_lo;          // the first n_lo columns carry lo_digits digits instead of 7 (single-block images only; the
  int32_t lo_digits;     // caller certifies the results, see digit_certificate in aux.cu)
  int32_t NC;            // digit rows per block = MMA N (multiple of 16, <= 128)
  int32_t fused_counts;  // 1: the kernel derives v1 / sum x^2 / sum x from the codes it decodes (no pre-pass)
  int32_t missing_policy;
  int64_t n_stages;      // 128-sample stages (>= 2)
  double* S;             // first GEMM row of the S workspace, [n_cols][ldS]
  int64_t ldS;
  double* v1_out;        // [ldS]      (fused_counts)
  double* mom_out;       // [2][ldS]   (fused_counts): sum x^2, sum x
  int32_t pair;          // 1: CTA pairs (cta_group::2)
  int32_t no_e;          // 1: optimistic form without the missing-call operand (raises flags[0] if the block has one)
  int32_t n_split;       // split of the sample range across work items (1 = none)
  int32_t split_stages;  // stages per range (even; the last range takes the remainder)
  int32_t* z_ws;         // n_split > 1: zero-initialised by the launcher, [n_cb][2][NC][ldS] int32 partial digit sums
  int32_t* n00_ws;       // n_split > 1: [ldS] int32 code-00 counts
  uint32_t x_lut;        // 2-bit code -> genotype value, one byte per code (0 = the default 0x00010002: 00 -> 2, 10 -> 1);
                         // 0x00010004 contracts the SQUARED calls (logistic score test: sum gamma x^2); the int8-states form knows these two tables
  uint32_t e_lut;        // 2-bit code -> missing-call indicator (0x00000100), set by launch_split8
  int32_t* flags;        // [2] device ints: [0] |= 1 when the optimistic form met a missing call,
                         //                  [1] += number of missing calls of the block (full form, fused_counts)
};
struct Split8Plan {
  int pair = 0, n_split = 1, split_stages = 0;
  double makespan_stages = 0.0;  // stages one work unit runs for
};
Split8Plan split8_plan(int n_variants, int n_cb, int64_t n_stages, int sm_count, int pair_mode, int force_split);
size_t split8_ws_bytes(int n_cb, int NC, int64_t ldS);
void launch_split8(Split8Params p, int sm_count, cudaStream_t st);
// n_lo / lo_digits: the reduced-digit image (n_lo = 0 for the plain one); it exists only when one block holds all columns
void split8_geometry(int n_cols, int n_lo, int lo_digits, int* n_cb, int* cpb, int* NC);
// digit row of local column lc inside its block, and of the row of ones (= split8_row0(cpb, ...))
__host__ __device__ inline int split8_row0(int lc, int n_lo, int lo_digits) {
  return 7 * lc - (7 - lo_digits) * (lc < n_lo ? lc : n_lo);
}
double measure_s8_issue_rate(int sm_count, double ms_target, cudaStream_t st);  // TOP/s, < 0 on failure
void launch_pack_split8(const double* src, int64_t N, int64_t ld, int c0, int ncols, double* colmax_scratch,
                        const int64_t* geno_to_row, int64_t Ng, int8_t* dst, int dst_col0, int cpb, int NC,
                        int64_t n_stages, int n_lo, int lo_digits, cudaStream_t st);

// ---- pass 2 (only when some phenotype has missing samples): D = Mu^T (X - Q A)^2 ----------------
struct MaskedParams {
  BlockView x;
  const double* qb;    // packed Q, B-operand layout: [chunk][8 groups][KT2][64]  (KT2 = pairs of k-steps)
  const double* mpk;   // packed unique masks: [group][chunk][8][MT][64]
  const double* A;     // coefficients Q^T X: rows 0..K-1 of the reduced S, [K][ldS]
  int32_t K;
  int32_t KT2;         // ceil(ceil(K/4)/2)
  int32_t MT;
  int32_t n_groups;
  int64_t n_chunks;
  int32_t n_split;
  double* D;           // [n_split][n_groups*MT*8][ldS]
  int64_t ldS;
};
void launch_masked(const MaskedParams& p, cudaStream_t st);

// ---- pass 2, integer tensor-core variant (many distinct masks): see mask8.cu ----
struct Mask8Params {
  BlockView x;           // x.v1 = missing substitute per variant
  const double* qg;      // [n_geno][K] rows of Q in genotype-sample order
  const double* A;       // coefficients Q^T X: [K][ldS]
  int64_t ldS;
  int32_t K;
  int32_t v0, nv;        // sub-range of variants
  int64_t n_stages;      // 128-sample stages
  unsigned long long* rmax_bits;  // [G] scratch: max_s r^2 per variant
  int8_t* z8;            // digit image of the sub-range: [tile][stage][128][128]
  const int8_t* m8;      // mask image: [mask block][stage][NB][128]
  int32_t n_mb, NB, U;
  int32_t* Z;            // int32 scratch [tile][mask block][128][NB]
  double* D;             // out: [U][ldS]
};
void mask8_geometry(int U, int* n_mb, int* NB);
size_t mask8_tile_bytes(int64_t n_stages);
size_t mask8_z_bytes(int n_tiles, int n_mb, int NB);
int mask8_tile_variants();
void launch_pack_mask8(const double* mu, int64_t N, int U, const int64_t* geno_to_row, int64_t Ng, int NB, int64_t n_stages,
                       int8_t* dst, cudaStream_t st);
void launch_gather_q(const double* Q, int64_t N, int K, const int64_t* geno_to_row, int64_t Ng, double* qg, cudaStream_t st);
void launch_mask8(const Mask8Params& p, int sm_count, cudaStream_t st);

// ---- pass 2, sparse complement form (few unobserved phenotype values): see masked_sparse.cu ----
struct SparseMaskedParams {
  BlockView x;           // x.v1 = missing substitute per variant
  const double* A;       // coefficients Q^T X: [K][ldS]
  int64_t ldS;
  int32_t K;
  const int32_t* idx;    // unobserved genotype-sample indices of all distinct masks, concatenated (sorted per mask)
  const double* qm;      // [total][Kp] Q rows of those samples, Kp = K rounded up to even
  const int64_t* off;    // [U + 1] list boundaries
  int32_t U;
  int32_t n_parts;       // split of every list across grid.y
  double* comp;          // out: [n_parts][U][ldS] partial sums of r^2 over the unobserved samples
};
void launch_sparse_masked(const SparseMaskedParams& p, cudaStream_t st);
void launch_gather_rows(const double* Q, int K, int Kp, const int64_t* rows, int64_t n, double* qm, cudaStream_t st);

// sums the n_split partial slabs in fixed order: out[r][v] = sum_z part[z][r][v]
void launch_reduce_splits(const double* part, double* out, int64_t rows, int64_t ld, int n_split, cudaStream_t st);

// ---- epilogue: A5-A7 (lin_reg.py:241-249) --------------------------------------------------------
struct EpilogueParams {
  const double* S;       // reduced [Mpad][ldS]; rows 0..K-1 = Q^T X, K..K+P-1 = Ytilde^T X,
                         // (verbose) K+P..K+2P-1 = mask^T X, K+2P..K+3P-1 = Yraw^T X
  const double* sxx;     // [ldS]
  const double* D;       // reduced [Upad][ldS] or nullptr
  int32_t d_is_complement; // 1: D holds the sum over the UNOBSERVED samples; XdotX = xx_full - D
  const int32_t* mask_id;// [P]: -1 = fully observed phenotype, else row of D
  const double* YdotY;   // [P] of this contig
  const double* Y_scale; // [P]
  const double* n_obs;   // [P] non-missing count per phenotype (verbose)
  int64_t ldS;
  int32_t K, P, G;
  int32_t verbose;
  double dof;
  PvalConst pc;
  double *effect, *stderror, *tvalue, *pvalue;  // [P][G]
  double *n, *sum_x, *ytx;                      // verbose, [P][G]
  // digit certificate (reduced-digit image of split8.cu): rows lo_row0 .. lo_row0 + n_lo - 1 of S were computed from
  // columns quantised to lo_scale[c]; a test whose certified error of XdotX exceeds kDigitCertTol * XdotX raises flags |= 4
  const double* lo_scale = nullptr;  // [n_lo] quantum of each reduced column, nullptr = nothing to certify
  const double* sum_x_row = nullptr; // [ldS] sum of x per variant
  int32_t n_lo = 0, lo_row0 = 0;
  int32_t* flags = nullptr;
};
constexpr double kDigitCertTol = 1e-11;  // 100 x inside the 1e-9 parity tolerance of the statistics
// xx_scratch: [3][ldS] doubles of workspace (fully observed XdotX per variant; the two certified error bounds)
void launch_epilogue(const EpilogueParams& p, double* xx_scratch, cudaStream_t st);

void launch_pvalues(const double* t, int64_t n, PvalConst pc, double* p, cudaStream_t st);
void launch_decode(const BlockView& x, double* out, cudaStream_t st);  // out [G_b][n_geno]
// sums over the listed (dropped) samples of x^2 and x, subtracted from mom[0] / mom[1] when rows are dropped
void launch_dropped_moments(const BlockView& x, const int64_t* drop_idx, int64_t n_drop, double* mom, int64_t ld,
                            cudaStream_t st);
// zeroes the listed (dropped) samples of a STAGED float block (base is library-owned memory)
void launch_zero_dropped(const uint8_t* base, int64_t stride, int G, int format, const int64_t* drop_idx, int64_t n_drop,
                         cudaStream_t st);
// BGEN layout-2 probability pairs (bits = 8 | 16 | 32 per probability) + ploidy bytes -> PLINK 2-bit rows (device buffers)
void launch_bgen_to_plink2(const void* probs, const uint8_t* ploidy, int G, int64_t N, int bits, double threshold, uint8_t* out,
                           int64_t out_stride, cudaStream_t st);
// GloWGR LOCO apply step: out[c][s] = cov[s] . Bcov[block(s)] + sum over headers h of chromosomes != c of
// (X[s][h] - mu[h]) / sig[h] * B[block(s)][h]; perm/seg = header indices grouped by chromosome (device arrays)
void launch_loco_predict(const double* X, int64_t N, int H, const double* mu, const double* sig, const double* B, const int32_t* perm,
                         const int32_t* seg, int n_chr, const int32_t* sample_block, const double* cov, int Kc, const double* Bcov,
                         double* out, cudaStream_t st);
// S row of an intercept column that is not in the GEMM: row0[v] = q0 * sum_x[v]
void launch_intercept_row(const double* sum_x, double q0, double* row0, int G, cudaStream_t st);

// ---- state packing (once per call of linear_regression) -----------------------------------------
struct PackDims {
  int64_t N, Ng;
  int32_t K, P;
};
// T = Q^T Y (K x P)
void launch_qty(const double* Q, const double* Y, int64_t N, int K, int P, double* T, cudaStream_t st);
// Ytilde = Y - Q T
void launch_residualize(const double* Q, const double* T, double* Y, int64_t N, int K, int P, cudaStream_t st);
// Packs columns [c0, c0+ncols) of a row-major [N x ld] matrix into pair-atom A layout at column
// offset `dst_col0` of a packed operand with `MT` m-tiles per group.
// `tail` extra columns (dst columns >= 8*MT, single group) are stored per stage after the atoms as
// [8 sample groups][tail][8 samples] in natural sample order.
void launch_pack_a(const double* src, int64_t N, int64_t ld, int c0, int ncols, const int64_t* geno_to_row,
                   int64_t Ng, double* dst, int dst_col0, int MT, int tail, int64_t n_chunks, cudaStream_t st);
// Packs Q into the masked-pass B layout.
void launch_pack_qb(const double* Q, int64_t N, int K, const int64_t* geno_to_row, int64_t Ng, double* dst, int KT2,
                    int64_t n_chunks, cudaStream_t st);

}  // namespace glr
