// C ABI of the logistic-regression score test: the block kernel
// _logistic_regression_inner of python/glow/gwas/log_reg.py:278-350 with correction = 'none'.
//
// Reference formulation, per phenotype p (log_reg.py:286, 316-321), gamma = yhat (1 - yhat) (0 at unobserved samples):
//   X_hat = C (C^T Gamma C)^-1 C^T Gamma X,   X_res = X - X_hat
//   chisq = (sum_s X_res Y_res)^2 / (sum_s X_res^2 gamma),   pvalue = chi2.sf(chisq, 1)
// With b = C^T Gamma x, q = (C^T Gamma C)^-1 b and t = C^T Y_res (a per-phenotype constant, ~0 at the fitted null model
// but kept: the reference keeps it):
//   sum_s X_res Y_res   = Y_res^T x - t . q
//   sum_s X_res^2 gamma = x^T Gamma x - 2 b . q + q^T (C^T Gamma C) q
// (the last two terms are evaluated as written, with the stored C^T Gamma C, not collapsed to b . q: the reference's
// inverse is a numerical one).  So a block needs two contractions and an O(K^2) epilogue per test:
//   pass 1  S = [gamma_p o C | Y_res_p]_p^T X        (K + 1 columns per phenotype)   -> gram_kernel (FP64 DMMA)
//   pass 2  D = gamma^T X^2                          (one weight column per phenotype) -> masked_kernel with K = 0:
//           its "masks" are arbitrary FP64 weights and r = x when there is no coefficient to subtract
//   epilogue: logreg_epilogue_kernel
// The decode of packed genotypes (PLINK 2-bit / int8 states, optional mean substitution) is the same fused operand
// load as in the linear path.
// PLINK 2-bit blocks and 4-byte aligned int8 genotype states (more than 384 samples) take both contractions on the int8 tensor cores instead
// (split8.cu, the linear path's kernel): pass 1 on the digit image of W with the genotype counts and the mean substitute
// fused; pass 2 on the digit image of the gamma columns with the SQUARED-call table (00 -> 4, 10 -> 1) and mu^2 as the
// missing substitute, because x^2 = xi^2 + mu^2 e when xi e = 0.  Exact integer accumulation, FP64 recombination:
// 4.9 ms instead of 79 ms per 37 888 x 500 000 block.  GLR_LOGREG_SPLIT8=0 (read at state creation) keeps the DMMA path.
//
// correction = 'approx-firth' first residualises the genotypes LINEARLY against the covariate basis, X <- X - Q Q^T X
// (log_reg.py:312-313), and runs the score test on that.  With a = Q^T x (K more columns of the same contraction) the
// quantities of x_lin = x - Q a follow from those of x:
//   b_lin = b - (C^T Gamma Q) a,   v_lin = v - (Q^T Y_res) . a,   D_lin = D - 2 a . (R^-T b) + a^T (Q^T Gamma Q) a
// (Q^T Gamma x = R^-T C^T Gamma x because C = Q R), so no residual matrix is formed on the device either.  The Firth
// refits of the flagged tests stay on the host, as in the reference.
#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "../../include/glow_b200.h"
#include "api_internal.h"
#include "common.cuh"
#include "kernels.h"

using namespace glr;

#define CUL(expr)                                                                                   \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) return api_fail(GLR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

namespace {
struct Buf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    if (cudaMalloc(&p, bytes + bytes / 8 + 256) != cudaSuccess) return api_fail(GLR_ERR_CUDA, "cudaMalloc failed (logistic workspace)");
    cap = bytes + bytes / 8 + 256;
    return 0;
  }
  template <typename T>
  T* as() const { return reinterpret_cast<T*>(p); }
};
int mt_masked(int n_mtiles) {
  for (int o : {1, 2, 4, 8})
    if (n_mtiles <= o) return o;
  return 8;
}
}  // namespace

struct glr_logreg_state {
  int device = 0;
  int64_t N = 0, Ng = 0, n_chunks = 0;
  int K = 0, P = 0, C = 1;
  int Mcols = 0, MT = 1, n_groups = 1;   // pass 1 operand
  int MT2 = 1, n_groups2 = 1;            // pass 2 operand (gamma)
  int64_t w_contig = 0, g_contig = 0;    // doubles per contig in wpk / gpk
  double* wpk = nullptr;                 // [C][n_groups][n_chunks][8][MT][64]
  double* gpk = nullptr;                 // [C][n_groups2][n_chunks][8][MT2][64]
  double* qb0 = nullptr;                 // zero "Q" operand of the K = 0 residual step
  double* inv = nullptr;                 // [C][P][K][K]
  double* ctgc = nullptr;                // [C][P][K][K]   C^T Gamma C
  double* tvec = nullptr;                // [C][P][K]      C^T Y_res
  int lin = 0;                           // 1: linear residualisation first (approx-firth): Q columns follow in W
  double* m1 = nullptr;                  // [C][P][K][K]   C^T Gamma Q
  double* m2 = nullptr;                  // [C][P][K][K]   Q^T Gamma Q
  double* t2 = nullptr;                  // [C][P][K]      Q^T Y_res
  double* rinvt = nullptr;               // [K][K]         R^-T, R = Q^T C
  int64_t* drop_idx = nullptr;
  int64_t n_drop = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  Buf geno, out, s_part, s_red, mom_part, mom, v1, d_part, d_red;
  // int8 tensor-core form of both contractions for PLINK 2-bit blocks (split8.cu): digit images of W and of gamma
  int s8 = 0;
  int s8_ncb = 0, s8_cpb = 0, s8_NC = 0, s8g_ncb = 0, s8g_cpb = 0, s8g_NC = 0;
  int64_t s8_stages = 0, s8_contig = 0, s8g_contig = 0;
  int8_t *w8 = nullptr, *g8 = nullptr;         // [C][column block][stage][NC][128]
  double *s8_scale = nullptr, *s8g_scale = nullptr;  // [C][Mcols], [C][P]
  Buf s8_ws, s8_n00, s8_flags, v1sq;
};

__global__ void square_kernel(const double* in, double* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] * in[i];
}
__global__ void scale_exp_kernel(double* v, int n, int e) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = ldexp(v[i], e);
}

struct LogregEpi {
  const double *S, *D;
  int64_t ldS;
  int K, P, G, lin;
  const double *inv, *ctgc, *tvec;     // of this contig
  const double *m1, *m2, *t2, *rinvt;  // lin only
  double *chisq, *pvalue;
};

// thread = (phenotype, variant)
__global__ void logreg_epilogue_kernel(const LogregEpi e) {
  constexpr int kLocalK = 64;
  const int K = e.K, G = e.G;
  const int64_t ldS = e.ldS;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)e.P * G) return;
  const int p = (int)(idx / G);
  const int g = (int)(idx - (int64_t)p * G);
  const double* bs = e.S + (int64_t)p * (K + 1) * ldS + g;  // b[k] = bs[k * ldS]
  double v = bs[(int64_t)K * ldS];                           // Y_res^T x
  double Dv = e.D[(int64_t)p * ldS + g];                     // x^T Gamma x
  const double* ip = e.inv + (int64_t)p * K * K;
  const double* ap = e.ctgc + (int64_t)p * K * K;
  const double* tp = e.tvec + (int64_t)p * K;
  double tq = 0.0, bq = 0.0, qaq = 0.0;
  if (K <= kLocalK) {
    double b[kLocalK], q[kLocalK];
    for (int c = 0; c < K; ++c) b[c] = bs[(int64_t)c * ldS];
    if (e.lin) {
      // linear residualisation against Q folded into the contractions (see the header of this file)
      const double* as = e.S + (int64_t)e.P * (K + 1) * ldS + g;  // a[k] = Q^T x
      const double* m1 = e.m1 + (int64_t)p * K * K;
      const double* m2 = e.m2 + (int64_t)p * K * K;
      const double* t2 = e.t2 + (int64_t)p * K;
      double a[kLocalK];
      for (int c = 0; c < K; ++c) a[c] = as[(int64_t)c * ldS];
      double cross = 0.0, quad = 0.0, ta = 0.0;
      for (int c = 0; c < K; ++c) {
        double rb = 0.0, m2a = 0.0;
        for (int d = 0; d < K; ++d) {
          rb = fma(e.rinvt[c * K + d], b[d], rb);  // (R^-T b)[c] = (Q^T Gamma x)[c]
          m2a = fma(m2[c * K + d], a[d], m2a);
        }
        cross = fma(a[c], rb, cross);
        quad = fma(a[c], m2a, quad);
        ta = fma(t2[c], a[c], ta);
      }
      Dv = Dv - 2.0 * cross + quad;
      v -= ta;
      for (int c = 0; c < K; ++c) {
        double m1a = 0.0;
        for (int d = 0; d < K; ++d) m1a = fma(m1[c * K + d], a[d], m1a);
        q[c] = b[c] - m1a;
      }
      for (int c = 0; c < K; ++c) b[c] = q[c];
    }
    for (int c = 0; c < K; ++c) {
      double qc = 0.0;
      for (int d = 0; d < K; ++d) qc = fma(ip[c * K + d], b[d], qc);
      q[c] = qc;
      tq = fma(tp[c], qc, tq);
      bq = fma(b[c], qc, bq);
    }
    for (int c = 0; c < K; ++c) {
      double aq = 0.0;
      for (int d = 0; d < K; ++d) aq = fma(ap[c * K + d], q[d], aq);
      qaq = fma(q[c], aq, qaq);
    }
  } else {  // many covariates (score test only): q is recomputed instead of stored, O(K^3) per test
    for (int c = 0; c < K; ++c) {
      double qc = 0.0;
      for (int d = 0; d < K; ++d) qc = fma(ip[c * K + d], bs[(int64_t)d * ldS], qc);
      tq = fma(tp[c], qc, tq);
      bq = fma(bs[(int64_t)c * ldS], qc, bq);
      double aq = 0.0;
      for (int d = 0; d < K; ++d) {
        double qd = 0.0;
        for (int f = 0; f < K; ++f) qd = fma(ip[d * K + f], bs[(int64_t)f * ldS], qd);
        aq = fma(ap[c * K + d], qd, aq);
      }
      qaq = fma(qc, aq, qaq);
    }
  }
  const double num = (v - tq) * (v - tq);
  const double den = Dv - 2.0 * bq + qaq;
  const double x2 = num / den;
  e.chisq[idx] = x2;
  // scipy.stats.chi2.sf(x, 1) = erfc(sqrt(x / 2)); sf of a negative statistic is 1
  e.pvalue[idx] = x2 != x2 ? x2 : (x2 < 0.0 ? 1.0 : erfc(sqrt(0.5 * x2)));
}

extern "C" {

int glr_logreg_state_destroy(glr_logreg_state* st) {
  if (!st) return 0;
  cudaSetDevice(st->device);
  if (st->stream) cudaStreamSynchronize(st->stream);
  for (void* p : {(void*)st->wpk, (void*)st->gpk, (void*)st->qb0, (void*)st->inv, (void*)st->ctgc, (void*)st->tvec, (void*)st->drop_idx,
                  (void*)st->m1, (void*)st->m2, (void*)st->t2, (void*)st->rinvt, (void*)st->w8, (void*)st->g8,
                  (void*)st->s8_scale, (void*)st->s8g_scale})
    if (p) cudaFree(p);
  for (Buf* b : {&st->geno, &st->out, &st->s_part, &st->s_red, &st->mom_part, &st->mom, &st->v1, &st->d_part, &st->d_red,
                 &st->s8_ws, &st->s8_n00, &st->s8_flags, &st->v1sq})
    if (b->p) cudaFree(b->p);
  if (st->stream) cudaStreamDestroy(st->stream);
  delete st;
  return 0;
}

int glr_logreg_state_create(const glr_logreg_desc* d, glr_logreg_state** out) {
  if (!d || !out) return api_fail(GLR_ERR_INVALID, "null argument");
  *out = nullptr;
  if (d->abi_version != GLR_ABI_VERSION) return api_fail(GLR_ERR_INVALID, "ABI version mismatch");
  if (d->n_samples <= 0 || d->n_covariates <= 0 || d->n_phenotypes <= 0 || d->n_contigs <= 0)
    return api_fail(GLR_ERR_INVALID, "bad dimensions");
  if (d->n_geno_samples < d->n_samples) return api_fail(GLR_ERR_INVALID, "n_geno_samples < n_samples");
  if (!d->C || !d->gamma || !d->Y_res || !d->inv_CtGammaC) return api_fail(GLR_ERR_INVALID, "missing state array");
  if (d->n_geno_samples != d->n_samples && !d->keep_idx) return api_fail(GLR_ERR_INVALID, "keep_idx required when genotype rows are dropped");
  if (int rc = api_check_device(d->device)) return rc;
  glr_logreg_state* st = new glr_logreg_state();
  struct Guard {
    glr_logreg_state* s;
    ~Guard() { if (s) glr_logreg_state_destroy(s); }
  } guard{st};
  st->device = d->device;
  const int64_t N = st->N = d->n_samples, Ng = st->Ng = d->n_geno_samples;
  const int K = st->K = d->n_covariates, P = st->P = d->n_phenotypes, C = st->C = d->n_contigs;
  st->n_chunks = ceil_div(Ng, kChunk);
  CUL(cudaDeviceGetAttribute(&st->sm_count, cudaDevAttrMultiProcessorCount, d->device));
  CUL(cudaStreamCreateWithFlags(&st->stream, cudaStreamNonBlocking));

  // sample map (np.delete at log_reg.py:303-304 as a gather)
  int64_t* d_g2r = nullptr;
  std::vector<int64_t> g2r;
  if (d->keep_idx) {
    g2r.assign(Ng, -1);
    for (int64_t i = 0; i < N; ++i) {
      const int64_t g = d->keep_idx[i];
      if (g < 0 || g >= Ng || g2r[g] != -1) return api_fail(GLR_ERR_INVALID, "keep_idx must be unique and in range");
      g2r[g] = i;
    }
    std::vector<int64_t> drop;
    for (int64_t g = 0; g < Ng; ++g)
      if (g2r[g] < 0) drop.push_back(g);
    st->n_drop = (int64_t)drop.size();
    CUL(cudaMalloc(&d_g2r, Ng * sizeof(int64_t)));
    CUL(cudaMemcpy(d_g2r, g2r.data(), Ng * sizeof(int64_t), cudaMemcpyHostToDevice));
    if (st->n_drop) {
      CUL(cudaMalloc(&st->drop_idx, drop.size() * sizeof(int64_t)));
      CUL(cudaMemcpy(st->drop_idx, drop.data(), drop.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    }
  }
  struct Tmp {
    std::vector<void*> v;
    ~Tmp() { for (void* p : v) if (p) cudaFree(p); }
  } tmp;
  tmp.v.push_back(d_g2r);

  st->lin = d->Q ? 1 : 0;
  if (st->lin && K > 64) return api_fail(GLR_ERR_INVALID, "the linear residualisation of approx-firth takes at most 64 covariates");
  st->Mcols = P * (K + 1) + (st->lin ? K : 0);
  const int n_mt = (int)ceil_div(st->Mcols, 8);
  st->MT = std::min(n_mt, kMaxMT);
  st->n_groups = (int)ceil_div(n_mt, kMaxMT);
  if (st->n_groups > 1) st->MT = kMaxMT;
  const int n_mt2 = (int)ceil_div(P, 8);
  st->MT2 = n_mt2 <= kMaxMT ? mt_masked(n_mt2) : kMaxMT;
  st->n_groups2 = n_mt2 <= kMaxMT ? 1 : (int)ceil_div(n_mt2, kMaxMT);
  st->w_contig = (int64_t)st->n_groups * st->n_chunks * kChunkGroups * st->MT * kAtomDoubles;
  st->g_contig = (int64_t)st->n_groups2 * st->n_chunks * kChunkGroups * st->MT2 * kAtomDoubles;
  CUL(cudaMalloc(&st->wpk, (size_t)C * st->w_contig * 8));
  CUL(cudaMemset(st->wpk, 0, (size_t)C * st->w_contig * 8));
  CUL(cudaMalloc(&st->gpk, (size_t)C * st->g_contig * 8));
  CUL(cudaMemset(st->gpk, 0, (size_t)C * st->g_contig * 8));
  const size_t qb_bytes = (size_t)st->n_chunks * kChunkGroups * 1 * kAtomDoubles * 8;
  CUL(cudaMalloc(&st->qb0, qb_bytes));
  CUL(cudaMemset(st->qb0, 0, qb_bytes));

  // per contig: W = [gamma_p o C | Y_res_p]_p (row-major [N x P (K + 1)], built on the host), C^T Gamma C, t = C^T Y_res
  std::vector<double> W((size_t)N * st->Mcols), A((size_t)C * P * K * K, 0.0), T((size_t)C * P * K, 0.0);
  std::vector<double> M1, M2, T2, Rit;
  if (st->lin) {
    M1.assign(A.size(), 0.0);
    M2.assign(A.size(), 0.0);
    T2.assign(T.size(), 0.0);
    // R = Q^T C; R^-T by Gauss-Jordan with partial pivoting on the K x K matrix (host, once)
    std::vector<double> R((size_t)K * K, 0.0), Inv((size_t)K * K, 0.0);
    for (int64_t s = 0; s < N; ++s)
      for (int k = 0; k < K; ++k)
        for (int l = 0; l < K; ++l) R[k * K + l] = fma(d->Q[s * K + k], d->C[s * K + l], R[k * K + l]);
    for (int k = 0; k < K; ++k) Inv[k * K + k] = 1.0;
    for (int col = 0; col < K; ++col) {
      int piv = col;
      for (int r = col + 1; r < K; ++r)
        if (std::fabs(R[r * K + col]) > std::fabs(R[piv * K + col])) piv = r;
      if (R[piv * K + col] == 0.0) return api_fail(GLR_ERR_INVALID, "Q^T C is singular: Q must be the orthonormal basis of C");
      for (int l = 0; l < K; ++l) {
        std::swap(R[col * K + l], R[piv * K + l]);
        std::swap(Inv[col * K + l], Inv[piv * K + l]);
      }
      const double dinv = 1.0 / R[col * K + col];
      for (int l = 0; l < K; ++l) {
        R[col * K + l] *= dinv;
        Inv[col * K + l] *= dinv;
      }
      for (int r = 0; r < K; ++r) {
        if (r == col) continue;
        const double f = R[r * K + col];
        if (f == 0.0) continue;
        for (int l = 0; l < K; ++l) {
          R[r * K + l] -= f * R[col * K + l];
          Inv[r * K + l] -= f * Inv[col * K + l];
        }
      }
    }
    Rit.assign((size_t)K * K, 0.0);
    for (int k = 0; k < K; ++k)
      for (int l = 0; l < K; ++l) Rit[k * K + l] = Inv[l * K + k];  // transpose of R^-1
  }
  // int8 digit images (split8.cu): exact contraction of 2-bit blocks on the tcgen05 tensor cores.  |digit| * 4 * Ng < 2^31
  // bounds the squared-call contraction; at least 4 stages, as in the linear path.
  double* d_colmax = nullptr;
  {
    const char* e = getenv("GLR_LOGREG_SPLIT8");
    if ((!e || atoi(e) != 0) && Ng > 3 * 128 && Ng < (int64_t)4000000) {
      split8_geometry(st->Mcols, 0, 7, &st->s8_ncb, &st->s8_cpb, &st->s8_NC);
      split8_geometry(P, 0, 7, &st->s8g_ncb, &st->s8g_cpb, &st->s8g_NC);
      st->s8_stages = ceil_div(Ng, 128);
      st->s8_contig = (int64_t)st->s8_ncb * st->s8_stages * st->s8_NC * 128;
      st->s8g_contig = (int64_t)st->s8g_ncb * st->s8_stages * st->s8g_NC * 128;
      if ((int64_t)C * (st->s8_contig + st->s8g_contig) <= ((int64_t)16 << 30)) st->s8 = 1;
    }
    if (st->s8) {
      CUL(cudaMalloc(&st->w8, (size_t)C * st->s8_contig));
      CUL(cudaMemset(st->w8, 0, (size_t)C * st->s8_contig));
      CUL(cudaMalloc(&st->g8, (size_t)C * st->s8g_contig));
      CUL(cudaMemset(st->g8, 0, (size_t)C * st->s8g_contig));
      CUL(cudaMalloc(&st->s8_scale, (size_t)C * st->Mcols * 8));
      CUL(cudaMalloc(&st->s8g_scale, (size_t)C * P * 8));
      CUL(cudaMalloc(&d_colmax, (size_t)std::max(st->Mcols, P) * 8));
      tmp.v.push_back(d_colmax);
    }
  }
  double *dW = nullptr, *dG = nullptr;
  CUL(cudaMalloc(&dW, W.size() * 8));
  tmp.v.push_back(dW);
  CUL(cudaMalloc(&dG, (size_t)N * P * 8));
  tmp.v.push_back(dG);
  for (int c = 0; c < C; ++c) {
    const double* gam = d->gamma + (size_t)c * N * P;
    const double* yres = d->Y_res + (size_t)c * N * P;
    double* Ac = A.data() + (size_t)c * P * K * K;
    double* Tc = T.data() + (size_t)c * P * K;
    for (int64_t s = 0; s < N; ++s) {
      const double* cs = d->C + s * K;
      double* ws = W.data() + (size_t)s * st->Mcols;
      for (int p = 0; p < P; ++p) {
        const double gsp = gam[s * P + p], ysp = yres[s * P + p];
        double* wp = ws + (size_t)p * (K + 1);
        for (int k = 0; k < K; ++k) wp[k] = gsp * cs[k];
        wp[K] = ysp;
        double* Ap = Ac + (size_t)p * K * K;
        for (int k = 0; k < K; ++k) {
          const double gk = gsp * cs[k];
          for (int l = 0; l < K; ++l) Ap[k * K + l] = fma(gk, cs[l], Ap[k * K + l]);
          Tc[p * K + k] = fma(cs[k], ysp, Tc[p * K + k]);
        }
        if (st->lin) {
          const double* qs = d->Q + s * K;
          double* M1p = M1.data() + ((size_t)c * P + p) * K * K;
          double* M2p = M2.data() + ((size_t)c * P + p) * K * K;
          double* T2p = T2.data() + ((size_t)c * P + p) * K;
          for (int k = 0; k < K; ++k) {
            const double gc = gsp * cs[k], gq = gsp * qs[k];
            for (int l = 0; l < K; ++l) {
              M1p[k * K + l] = fma(gc, qs[l], M1p[k * K + l]);
              M2p[k * K + l] = fma(gq, qs[l], M2p[k * K + l]);
            }
            T2p[k] = fma(qs[k], ysp, T2p[k]);
          }
        }
      }
      if (st->lin) {
        const double* qs = d->Q + s * K;
        double* wq = ws + (size_t)P * (K + 1);
        for (int k = 0; k < K; ++k) wq[k] = qs[k];
      }
    }
    CUL(cudaMemcpy(dW, W.data(), W.size() * 8, cudaMemcpyHostToDevice));
    CUL(cudaMemcpy(dG, gam, (size_t)N * P * 8, cudaMemcpyHostToDevice));
    launch_pack_a(dW, N, st->Mcols, 0, st->Mcols, d_g2r, Ng, st->wpk + (int64_t)c * st->w_contig, 0, st->MT, 0, st->n_chunks, nullptr);
    launch_pack_a(dG, N, P, 0, P, d_g2r, Ng, st->gpk + (int64_t)c * st->g_contig, 0, st->MT2, 0, st->n_chunks, nullptr);
    if (st->s8) {  // column scale = max|w_c| * 2^-54
      launch_pack_split8(dW, N, st->Mcols, 0, st->Mcols, d_colmax, d_g2r, Ng, st->w8 + (int64_t)c * st->s8_contig, 0, st->s8_cpb,
                         st->s8_NC, st->s8_stages, 0, 7, nullptr);
      CUL(cudaMemcpyAsync(st->s8_scale + (int64_t)c * st->Mcols, d_colmax, (size_t)st->Mcols * 8, cudaMemcpyDeviceToDevice, nullptr));
      launch_pack_split8(dG, N, P, 0, P, d_colmax, d_g2r, Ng, st->g8 + (int64_t)c * st->s8g_contig, 0, st->s8g_cpb, st->s8g_NC,
                         st->s8_stages, 0, 7, nullptr);
      CUL(cudaMemcpyAsync(st->s8g_scale + (int64_t)c * P, d_colmax, (size_t)P * 8, cudaMemcpyDeviceToDevice, nullptr));
    }
    CUL(cudaDeviceSynchronize());
  }
  if (st->s8) {
    scale_exp_kernel<<<(unsigned)ceil_div((int64_t)C * st->Mcols, 256), 256>>>(st->s8_scale, C * st->Mcols, -54);
    scale_exp_kernel<<<(unsigned)ceil_div((int64_t)C * P, 256), 256>>>(st->s8g_scale, C * P, -54);
    CUL(cudaDeviceSynchronize());
  }
  CUL(cudaMalloc(&st->inv, A.size() * 8));
  CUL(cudaMemcpy(st->inv, d->inv_CtGammaC, A.size() * 8, cudaMemcpyHostToDevice));
  CUL(cudaMalloc(&st->ctgc, A.size() * 8));
  CUL(cudaMemcpy(st->ctgc, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
  CUL(cudaMalloc(&st->tvec, T.size() * 8));
  CUL(cudaMemcpy(st->tvec, T.data(), T.size() * 8, cudaMemcpyHostToDevice));
  if (st->lin) {
    CUL(cudaMalloc(&st->m1, M1.size() * 8));
    CUL(cudaMemcpy(st->m1, M1.data(), M1.size() * 8, cudaMemcpyHostToDevice));
    CUL(cudaMalloc(&st->m2, M2.size() * 8));
    CUL(cudaMemcpy(st->m2, M2.data(), M2.size() * 8, cudaMemcpyHostToDevice));
    CUL(cudaMalloc(&st->t2, T2.size() * 8));
    CUL(cudaMemcpy(st->t2, T2.data(), T2.size() * 8, cudaMemcpyHostToDevice));
    CUL(cudaMalloc(&st->rinvt, Rit.size() * 8));
    CUL(cudaMemcpy(st->rinvt, Rit.data(), Rit.size() * 8, cudaMemcpyHostToDevice));
  }
  CUL(cudaGetLastError());
  guard.s = nullptr;
  *out = st;
  return 0;
}

}  // extern "C"

// allow_s8 = false: second run of an int8 block whose values turned out not to be genotype states (see below)
static int logreg_run_block(glr_logreg_state* st, const glr_block_desc* b, double* chisq, double* pvalue, bool allow_s8) {
  if (!st || !b) return api_fail(GLR_ERR_INVALID, "null argument");
  if (b->format < GLR_FMT_F64 || b->format > GLR_FMT_PLINK2) return api_fail(GLR_ERR_INVALID, "unknown genotype format");
  if (b->contig < 0 || b->contig >= st->C) return api_fail(GLR_ERR_INVALID, "contig index out of range");
  if (b->n_variants < 0) return api_fail(GLR_ERR_INVALID, "negative n_variants");
  const int G = b->n_variants;
  if (G == 0) return 0;
  if (!chisq || !pvalue || !b->genotypes) return api_fail(GLR_ERR_INVALID, "missing buffer");
  const size_t esz = b->format == GLR_FMT_F64 ? 8 : (b->format == GLR_FMT_F32 ? 4 : 1);
  const size_t rb = b->format == GLR_FMT_PLINK2 ? (size_t)((st->Ng + 3) / 4) : (size_t)st->Ng * esz;
  if ((size_t)b->row_stride_bytes < rb) return api_fail(GLR_ERR_INVALID, "row stride smaller than a row");
  if ((b->format == GLR_FMT_F64 || b->format == GLR_FMT_F32) && (b->row_stride_bytes % esz || (uintptr_t)b->genotypes % esz))
    return api_fail(GLR_ERR_INVALID, "genotype rows must be aligned to their element size");
  CUL(cudaSetDevice(st->device));
  cudaStream_t s = st->stream;
  const int P = st->P, K = st->K;
  const bool float_fmt = b->format == GLR_FMT_F64 || b->format == GLR_FMT_F32;
  const bool scrub = float_fmt && st->n_drop > 0;

  // (the 2-bit kernels are instantiated as 16 consumer warps up to 4 m-tiles, 8 beyond: gram_p2.cu)
  const int nt = 2, warps = (b->format == GLR_FMT_PLINK2 && st->MT <= 4) ? 16 : 8;
  const int tile = gram_tile_variants(nt, warps);
  const int64_t ldS = std::max(round_up(round_up(G, tile), 8), round_up(G, 256));
  auto split_for = [&](int64_t ctas) {
    int n = 1;
    if (ctas < st->sm_count) {
      n = (int)std::min<int64_t>(ceil_div(2 * st->sm_count, ctas), std::max<int64_t>(1, st->n_chunks / 8));
      n = std::max(1, std::min(n, 64));
    }
    return n;
  };
  const int n_split = split_for(ceil_div(G, tile) * st->n_groups);
  const int n_split2 = split_for(ceil_div(G, gram_tile_variants(2)) * st->n_groups2);
  const int64_t Mpad = (int64_t)st->n_groups * st->MT * 8, Upad = (int64_t)st->n_groups2 * st->MT2 * 8;

  const uint8_t* d_geno = reinterpret_cast<const uint8_t*>(b->genotypes);
  if (b->memspace == GLR_MEM_HOST || scrub) {
    const size_t bytes = (size_t)(G - 1) * b->row_stride_bytes + rb;
    if (int rc = st->geno.ensure(bytes + 64)) return rc;
    CUL(cudaMemcpyAsync(st->geno.p, b->genotypes, bytes, b->memspace == GLR_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s));
    d_geno = st->geno.as<uint8_t>();
  }
  if (int rc = st->s_red.ensure((size_t)Mpad * ldS * 8)) return rc;
  if (n_split > 1)
    if (int rc = st->s_part.ensure((size_t)n_split * Mpad * ldS * 8)) return rc;
  if (int rc = st->mom.ensure((size_t)2 * ldS * 8)) return rc;
  if (int rc = st->mom_part.ensure((size_t)std::max(n_split, 1) * 2 * ldS * 8)) return rc;
  if (int rc = st->v1.ensure((size_t)ldS * 8)) return rc;
  if (int rc = st->d_red.ensure((size_t)Upad * ldS * 8)) return rc;
  if (n_split2 > 1)
    if (int rc = st->d_part.ensure((size_t)n_split2 * Upad * ldS * 8)) return rc;
  if (int rc = st->out.ensure((size_t)2 * P * G * 8)) return rc;

  BlockView x;
  x.base = d_geno;
  x.stride = b->row_stride_bytes;
  x.n_variants = G;
  x.format = b->format;
  x.n_geno = st->Ng;
  x.v1 = st->v1.as<double>();
  if (scrub) launch_zero_dropped(d_geno, b->row_stride_bytes, G, b->format, st->drop_idx, st->n_drop, s);
  // 2-bit blocks: both contractions on the int8 tensor cores.  Pass 1 is the linear path's kernel on W (counts and mean
  // substitute fused); pass 2 is the same kernel with the squared-call table on the gamma columns and mu^2 as the
  // substitute: sum_s gamma_s (xi_s^2 + mu^2 e_s) = x^T Gamma x.
  // (dropped samples carry zero digits, and the mean substitute is taken over all genotype samples -- upstream of the drop
  // in the reference: mean_substitute runs on the whole array -- so the fused counts are the right ones here)
  // int8 genotype states take the same kernels when the rows are 4-byte aligned; the kernel checks every byte and a block
  // with a value outside {-1, 0, 1, 2} is run again on the FP64 path (any int8 value is legal there)
  const bool i8_ok = b->format == GLR_FMT_I8 && ((reinterpret_cast<uintptr_t>(d_geno) | (uintptr_t)b->row_stride_bytes) & 3) == 0;
  const bool use_s8 = st->s8 && allow_s8 && (b->format == GLR_FMT_PLINK2 || i8_ok);
  if (use_s8) {
    if (int rc = st->v1sq.ensure((size_t)ldS * 8)) return rc;
    if (int rc = st->s8_flags.ensure(2 * sizeof(int32_t))) return rc;
    CUL(cudaMemsetAsync(st->s8_flags.p, 0, 2 * sizeof(int32_t), s));
    for (int pass = 0; pass < 2; ++pass) {
      const int ncb = pass ? st->s8g_ncb : st->s8_ncb, NC = pass ? st->s8g_NC : st->s8_NC;
      const Split8Plan plan = split8_plan(G, ncb, st->s8_stages, st->sm_count, GLR_AUTO, 0);
      if (plan.n_split > 1) {
        if (int rc = st->s8_ws.ensure(split8_ws_bytes(ncb, NC, ldS))) return rc;
        if (int rc = st->s8_n00.ensure((size_t)ldS * sizeof(int32_t))) return rc;
      }
      Split8Params sp{};
      sp.x = x;
      sp.w8 = pass ? st->g8 + (int64_t)b->contig * st->s8g_contig : st->w8 + (int64_t)b->contig * st->s8_contig;
      sp.scale = pass ? st->s8g_scale + (int64_t)b->contig * P : st->s8_scale + (int64_t)b->contig * st->Mcols;
      sp.n_cols = pass ? P : st->Mcols;
      sp.n_cb = ncb;
      sp.cpb = pass ? st->s8g_cpb : st->s8_cpb;
      sp.NC = NC;
      sp.lo_digits = 7;
      sp.fused_counts = pass ? 0 : 1;
      sp.missing_policy = b->missing_policy;
      sp.n_stages = st->s8_stages;
      sp.S = pass ? st->d_red.as<double>() : st->s_red.as<double>();
      sp.ldS = ldS;
      sp.v1_out = st->v1.as<double>();
      sp.mom_out = st->mom.as<double>();
      sp.pair = plan.pair;
      sp.n_split = plan.n_split;
      sp.split_stages = plan.split_stages;
      sp.z_ws = st->s8_ws.as<int32_t>();
      sp.n00_ws = st->s8_n00.as<int32_t>();
      sp.flags = st->s8_flags.as<int32_t>();
      if (pass) {
        square_kernel<<<(unsigned)ceil_div(G, 256), 256, 0, s>>>(st->v1.as<double>(), st->v1sq.as<double>(), G);
        sp.x.v1 = st->v1sq.as<double>();
        sp.x_lut = 0x00010004u;  // 00 -> 4, 10 -> 1
      }
      launch_split8(sp, st->sm_count, s);
    }
  }
  if (!float_fmt && !use_s8) launch_prepass(x, b->missing_policy, st->v1.as<double>(), st->mom.as<double>(), ldS, s);

  GramParams gp;
  gp.x = x;
  gp.wpk = st->wpk + (int64_t)b->contig * st->w_contig;
  gp.MT = st->MT;
  gp.tail = 0;
  gp.n_groups = st->n_groups;
  gp.n_chunks = st->n_chunks;
  gp.n_split = n_split;
  gp.nt = nt;
  gp.warps = warps;
  gp.ldS = ldS;
  gp.S = n_split > 1 ? st->s_part.as<double>() : st->s_red.as<double>();
  gp.s_slab = Mpad * ldS;
  gp.mom_part = st->mom_part.as<double>();
  if (!use_s8) {
  launch_gram(gp, s);
  if (n_split > 1) launch_reduce_splits(st->s_part.as<double>(), st->s_red.as<double>(), Mpad, ldS, n_split, s);

  MaskedParams mp;
  mp.x = x;
  mp.qb = st->qb0;
  mp.mpk = st->gpk + (int64_t)b->contig * st->g_contig;
  mp.A = st->s_red.as<double>();  // no coefficients are read: K = 0
  mp.K = 0;
  mp.KT2 = 1;
  mp.MT = st->MT2;
  mp.n_groups = st->n_groups2;
  mp.n_chunks = st->n_chunks;
  mp.n_split = n_split2;
  mp.ldS = ldS;
  mp.D = n_split2 > 1 ? st->d_part.as<double>() : st->d_red.as<double>();
  launch_masked(mp, s);
  if (n_split2 > 1) launch_reduce_splits(st->d_part.as<double>(), st->d_red.as<double>(), Upad, ldS, n_split2, s);
  }

  double* d_chi = st->out.as<double>();
  double* d_p = d_chi + (size_t)P * G;
  const int64_t total = (int64_t)P * G;
  LogregEpi ep;
  ep.S = st->s_red.as<double>();
  ep.D = st->d_red.as<double>();
  ep.ldS = ldS;
  ep.K = K;
  ep.P = P;
  ep.G = G;
  ep.lin = st->lin;
  const int64_t coff = (int64_t)b->contig * P * K * K, toff = (int64_t)b->contig * P * K;
  ep.inv = st->inv + coff;
  ep.ctgc = st->ctgc + coff;
  ep.tvec = st->tvec + toff;
  ep.m1 = st->lin ? st->m1 + coff : nullptr;
  ep.m2 = st->lin ? st->m2 + coff : nullptr;
  ep.t2 = st->lin ? st->t2 + toff : nullptr;
  ep.rinvt = st->rinvt;
  ep.chisq = d_chi;
  ep.pvalue = d_p;
  logreg_epilogue_kernel<<<(unsigned)ceil_div(total, 128), 128, 0, s>>>(ep);
  CUL(cudaGetLastError());
  CUL(cudaMemcpyAsync(chisq, d_chi, (size_t)total * 8, cudaMemcpyDeviceToHost, s));
  CUL(cudaMemcpyAsync(pvalue, d_p, (size_t)total * 8, cudaMemcpyDeviceToHost, s));
  int32_t h_flags[2] = {0, 0};
  if (use_s8 && b->format == GLR_FMT_I8)
    CUL(cudaMemcpyAsync(h_flags, st->s8_flags.p, sizeof(h_flags), cudaMemcpyDeviceToHost, s));
  CUL(cudaStreamSynchronize(s));
  if (h_flags[0] & 2) return logreg_run_block(st, b, chisq, pvalue, false);
  return 0;
}

extern "C" {
int glr_logreg_run_block(glr_logreg_state* st, const glr_block_desc* b, double* chisq, double* pvalue) {
  return logreg_run_block(st, b, chisq, pvalue, true);
}
}  // extern "C"
