// C ABI (include/glow_b200.h): state packing, lanes (streams + staging), block execution.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/glow_b200.h"
#include "api_internal.h"
#include "common.cuh"
#include "kernels.h"

using namespace glr;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CU(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return fail(GLR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));               \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      p = nullptr;
      return fail(GLR_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    }
    cap = want;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

struct Lane {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[6] = {};
  DevBuf geno, out, s_part, s_red, mom_part, mom, v1, d_part, d_red, xx, m8_rmax, m8_z8, m8_Z, s8_ws, s8_n00, flags;
  int32_t* h_flags = nullptr;  // pinned mirror of `flags` ([0] optimistic form met a missing call, [1] missing calls of the block)
  bool busy = false;
  bool has_masked = false, has_prepass = false;
  bool optimistic = false, lo_digits = false;  // the block in flight runs the form without the missing-call operand
  bool counted = false;     // flags[1] of the block in flight is meaningful
  bool checked = false;     // flags[0] must be read back at wait time
  glr_block_desc blk{};     // the block in flight (for the re-run of an optimistic block that had missing calls)
  glr_block_out blk_out{};
  int launches = 0;
  int path = 0;
  glr_timing timing{};
};

// knob = desc field (glr_tristate), overridden by the environment variable when set (tests): env 0 = never, 1 = auto, 2 = always
int tri_knob(int32_t field, const char* env) {
  if (const char* e = getenv(env)) {
    const int v = atoi(e);
    return v <= 0 ? GLR_OFF : (v == 1 ? GLR_AUTO : GLR_FORCE);
  }
  return (field == GLR_OFF || field == GLR_FORCE) ? field : GLR_AUTO;
}
bool flag_knob(int32_t field, const char* env) { return field != 0 || getenv(env) != nullptr; }
int knob_int(int32_t field, const char* env) {
  if (const char* e = getenv(env)) return atoi(e);
  return field;
}

int pick_mt(int n_mtiles) {
  return n_mtiles < 1 ? 1 : (n_mtiles > 8 ? 8 : n_mtiles);
}
int pick_mt_masked(int n_mtiles) {
  const int opts[] = {1, 2, 4, 8};
  for (int o : opts)
    if (n_mtiles <= o) return o;
  return 8;
}

}  // namespace

struct glr_state {
  int device = 0;
  int64_t N = 0, Ng = 0;
  int K = 0, P = 0, C = 1, verbose = 0;
  int64_t dof = 0;
  int64_t n_chunks = 0;
  // pass-1 operand
  int Mcols = 0, MT = 1, n_groups = 1, tail = 0;
  int64_t s_rows = 0;  // rows of the S workspace
  // intercept column of Q handled outside the GEMM: Q^T x for a constant column q0 is q0 * sum(x)
  int has_icpt = 0;
  double icpt_q0 = 0.0;
  int64_t w_contig_doubles = 0;
  double* wpk = nullptr;  // [C][n_groups][n_chunks][8][MT][64]
  // integer tensor-core variant of pass 1 (PLINK2 blocks): digits of W, see split8.cu
  int split8 = 0;  // 0 = off, 1 = built
  int s8_cols = 0, s8_ncb = 0, s8_cpb = 0, s8_NC = 0;  // column blocks, columns per block, digit rows per block
  int64_t s8_stages = 0, s8_contig_bytes = 0;
  int8_t* w8 = nullptr;       // [C][column block][stage][NC][128]
  double* s8_scale = nullptr; // [C][s8_cols]
  // reduced-digit image: the Q columns with lo_digits digits (results certified per test, a block that fails is re-run
  // on the full image); built when one column block holds all columns
  int s8_n_lo = 0, s8_lo_digits = 7, s8_NC_lo = 0;
  int64_t s8_contig_bytes_lo = 0;
  int8_t* w8_lo = nullptr;
  double* s8_scale_lo = nullptr;
  int exact_hold = 0;  // blocks for which the full image is used after a failed certificate
  // masked pass
  int U = 0, MT2 = 1, n_groups2 = 0, KT2 = 0;
  double* qb = nullptr;
  double* mpk = nullptr;
  // masked pass, integer tensor-core variant (mask8.cu): many distinct masks
  int mask8 = 0, m8_nmb = 0, m8_NB = 0;
  int64_t m8_stages = 0;
  double* qg = nullptr;  // [Ng][K] rows of Q in genotype-sample order
  int8_t* m8 = nullptr;  // [mask block][stage][NB][128]
  // masked pass, sparse complement form (masked_sparse.cu): few unobserved phenotype values
  int sparse_masked = 0;
  int64_t sp_total = 0;
  int32_t* sp_idx = nullptr;  // unobserved genotype-sample indices per distinct mask, concatenated
  double* sp_qm = nullptr;    // [sp_total][Kp]
  int64_t* sp_off = nullptr;  // [U + 1]
  // small per-phenotype arrays
  int32_t* mask_id = nullptr;  // [P]
  double* YdotY = nullptr;     // [C][P]
  double* Y_scale = nullptr;   // [P]
  double* n_obs = nullptr;     // [P]
  int64_t* drop_idx = nullptr;
  int64_t n_drop = 0;
  PvalConst pc{};
  std::vector<Lane> lanes;
  int sm_count = 148;
  int force_split = 0;
  int s8_pair_mode = GLR_AUTO;
  int optimistic_mode = GLR_AUTO;
  bool expect_missing = true;  // the last counted PLINK2 block had missing calls (or none has been seen yet)
  int n_lanes_req = 2;
  // persistent device arrays: (offset of the pointer field inside this struct, bytes) -- what glr_comm_broadcast_state
  // replicates to the other devices
  std::vector<std::pair<size_t, size_t>> arrays;
  template <typename T>
  cudaError_t alloc(T** field, size_t bytes) {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(field), bytes ? bytes : 1);
    if (e == cudaSuccess) arrays.emplace_back((size_t)(reinterpret_cast<char*>(field) - reinterpret_cast<char*>(this)), bytes);
    return e;
  }
};

extern "C" {

int glr_abi_version(void) { return GLR_ABI_VERSION; }
const char* glr_last_error(void) { return g_err.c_str(); }

int glr_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int d = 0; d < n; ++d) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
  }
  return ok;
}

static int check_device(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return fail(GLR_ERR_NO_DEVICE, "no CUDA device visible: glow_b200 has no CPU fallback");
  }
  if (device < 0 || device >= n) return fail(GLR_ERR_INVALID, "device ordinal out of range");
  int major = 0;
  CU(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10)
    return fail(GLR_ERR_NO_DEVICE, "device is not compute capability 10.x (kernels are built for sm_100a only)");
  CU(cudaSetDevice(device));
  return 0;
}

int glr_state_destroy(glr_state* st) {
  if (!st) return 0;
  cudaSetDevice(st->device);
  for (auto& l : st->lanes) {
    if (l.stream) cudaStreamSynchronize(l.stream);
    for (auto& e : l.ev)
      if (e) cudaEventDestroy(e);
    for (DevBuf* b : {&l.geno, &l.out, &l.s_part, &l.s_red, &l.mom_part, &l.mom, &l.v1, &l.d_part, &l.d_red, &l.xx, &l.m8_rmax, &l.m8_z8,
                       &l.m8_Z, &l.s8_ws, &l.s8_n00, &l.flags})
      b->release();
    if (l.h_flags) cudaFreeHost(l.h_flags);
    if (l.stream) cudaStreamDestroy(l.stream);
  }
  for (auto& a : st->arrays) {
    void* p = *reinterpret_cast<void**>(reinterpret_cast<char*>(st) + a.first);
    if (p) cudaFree(p);
  }
  delete st;
  return 0;
}

int glr_state_unique_masks(glr_state* st) { return st ? st->U : 0; }

// lanes = independent streams with their own workspaces (current device = st->device)
static int create_lanes(glr_state* st) {
  int n_lanes = st->n_lanes_req > 0 ? st->n_lanes_req : 2;
  if (n_lanes > 8) n_lanes = 8;
  st->lanes.resize(n_lanes);
  for (auto& l : st->lanes) {
    CU(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
    for (auto& e : l.ev) CU(cudaEventCreate(&e));
    if (int rc2 = l.flags.ensure(2 * sizeof(int32_t))) return rc2;
    CU(cudaHostAlloc((void**)&l.h_flags, 2 * sizeof(int32_t), cudaHostAllocDefault));
    l.h_flags[0] = l.h_flags[1] = 0;
  }
  st->expect_missing = st->optimistic_mode != GLR_FORCE;
  return 0;
}

int glr_state_create(const glr_state_desc* d, glr_state** out) {
  if (!d || !out) return fail(GLR_ERR_INVALID, "null argument");
  *out = nullptr;
  if (d->abi_version != GLR_ABI_VERSION) return fail(GLR_ERR_INVALID, "ABI version mismatch");
  if (d->n_samples <= 0 || d->n_covariates < 0 || d->n_phenotypes <= 0 || d->n_contigs <= 0)
    return fail(GLR_ERR_INVALID, "bad dimensions");
  if (d->n_geno_samples < d->n_samples) return fail(GLR_ERR_INVALID, "n_geno_samples < n_samples");
  if (!d->Y || !d->Y_mask || !d->YdotY || !d->Y_scale || (d->n_covariates > 0 && !d->Q))
    return fail(GLR_ERR_INVALID, "missing state array");
  if (d->verbose && !d->Y_raw) return fail(GLR_ERR_INVALID, "verbose output needs Y_raw");
  if (d->n_geno_samples != d->n_samples && !d->keep_idx)
    return fail(GLR_ERR_INVALID, "keep_idx required when genotype rows are dropped");
  int rc = check_device(d->device);
  if (rc) return rc;

  glr_state* st = new glr_state();
  struct Guard {
    glr_state* s;
    ~Guard() {
      if (s) glr_state_destroy(s);
    }
  } guard{st};

  st->device = d->device;
  st->N = d->n_samples;
  st->Ng = d->n_geno_samples;
  st->K = d->n_covariates;
  st->P = d->n_phenotypes;
  st->C = d->n_contigs;
  st->verbose = d->verbose ? 1 : 0;
  st->dof = d->dof;
  st->pc = make_pval_const((double)d->dof);
  st->n_chunks = ceil_div(st->Ng, kChunk);
  CU(cudaDeviceGetAttribute(&st->sm_count, cudaDevAttrMultiProcessorCount, d->device));
  st->force_split = d->opt.sample_split > 0 ? d->opt.sample_split : 0;
  if (const char* e = getenv("GLR_SPLIT")) st->force_split = atoi(e);
  st->s8_pair_mode = tri_knob(d->opt.split8_pair, "GLR_S8_PAIR");
  st->optimistic_mode = tri_knob(d->opt.optimistic_calls, "GLR_OPTIMISTIC");
  {
    const int qd = knob_int(d->opt.q_digits, "GLR_Q_DIGITS");
    if (qd != 0 && (qd < 3 || qd > 7)) return fail(GLR_ERR_INVALID, "q_digits must be 0 (automatic) or 3..7");
  }

  const int64_t N = st->N, Ng = st->Ng;
  const int K = st->K, P = st->P, C = st->C;
  const bool dev_in = d->memspace == GLR_MEM_DEVICE;
  const cudaMemcpyKind in_kind = dev_in ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  cudaStream_t s0 = nullptr;  // default stream for the one-time packing

  // ---- sample map (np.delete at lin_reg.py:228-229 as a gather) ----
  int64_t* d_g2r = nullptr;
  std::vector<int64_t> g2r;
  if (d->keep_idx) {
    g2r.assign(Ng, -1);
    for (int64_t i = 0; i < N; ++i) {
      const int64_t g = d->keep_idx[i];
      if (g < 0 || g >= Ng || g2r[g] != -1) return fail(GLR_ERR_INVALID, "keep_idx must be unique and in range");
      g2r[g] = i;
    }
    std::vector<int64_t> drop;
    for (int64_t g = 0; g < Ng; ++g)
      if (g2r[g] < 0) drop.push_back(g);
    st->n_drop = (int64_t)drop.size();
    CU(cudaMalloc(&d_g2r, Ng * sizeof(int64_t)));
    CU(cudaMemcpy(d_g2r, g2r.data(), Ng * sizeof(int64_t), cudaMemcpyHostToDevice));
    if (st->n_drop) {
      CU(st->alloc(&st->drop_idx, drop.size() * sizeof(int64_t)));
      CU(cudaMemcpy(st->drop_idx, drop.data(), drop.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    }
  }
  struct TmpFree {
    std::vector<void*> v;
    ~TmpFree() {
      for (void* p : v)
        if (p) cudaFree(p);
    }
  } tmp;
  tmp.v.push_back(d_g2r);

  // ---- missingness patterns: phenotypes sharing a mask share one masked sum of squares ----
  std::vector<double> h_mask;
  const double* mask_host = d->Y_mask;
  if (dev_in) {
    h_mask.resize((size_t)N * P);
    CU(cudaMemcpy(h_mask.data(), d->Y_mask, (size_t)N * P * sizeof(double), cudaMemcpyDeviceToHost));
    mask_host = h_mask.data();
  }
  std::vector<uint64_t> h1(P, 1469598103934665603ull), h2(P, 0x9E3779B97F4A7C15ull);
  std::vector<int64_t> nobs(P, 0);
  for (int64_t s = 0; s < N; ++s) {
    const double* row = mask_host + s * P;
    for (int p = 0; p < P; ++p) {
      const uint64_t bit = row[p] != 0.0;
      nobs[p] += (int64_t)bit;
      h1[p] = (h1[p] ^ bit) * 1099511628211ull;
      h2[p] = (h2[p] + bit + 1) * 0xff51afd7ed558ccdull;
      h2[p] ^= h2[p] >> 29;
    }
  }
  std::vector<int32_t> mask_id(P, -1);
  std::vector<int> rep;  // representative phenotype of each unique mask
  {
    struct KeyHash {
      size_t operator()(const std::pair<uint64_t, uint64_t>& k) const { return (size_t)(k.first ^ (k.second * 31)); }
    };
    // hash -> candidate mask ids; a hit is confirmed by comparing the columns, so a collision cannot merge two masks
    std::unordered_map<std::pair<uint64_t, uint64_t>, std::vector<int>, KeyHash> seen;
    auto same_mask = [&](int a, int b) {
      for (int64_t s = 0; s < N; ++s)
        if ((mask_host[s * P + a] != 0.0) != (mask_host[s * P + b] != 0.0)) return false;
      return true;
    };
    for (int p = 0; p < P; ++p) {
      if (nobs[p] == N) continue;
      auto& cands = seen[std::make_pair(h1[p], h2[p])];
      for (int id : cands)
        if (same_mask(p, rep[id])) {
          mask_id[p] = id;
          break;
        }
      if (mask_id[p] < 0) {
        mask_id[p] = (int)rep.size();
        cands.push_back(mask_id[p]);
        rep.push_back(p);
      }
    }
  }
  st->U = (int)rep.size();

  // ---- small arrays ----
  CU(st->alloc(&st->mask_id, P * sizeof(int32_t)));
  CU(cudaMemcpy(st->mask_id, mask_id.data(), P * sizeof(int32_t), cudaMemcpyHostToDevice));
  CU(st->alloc(&st->YdotY, (size_t)C * P * sizeof(double)));
  CU(cudaMemcpy(st->YdotY, d->YdotY, (size_t)C * P * sizeof(double), in_kind));
  CU(st->alloc(&st->Y_scale, P * sizeof(double)));
  CU(cudaMemcpy(st->Y_scale, d->Y_scale, P * sizeof(double), in_kind));
  {
    std::vector<double> nf(P);
    for (int p = 0; p < P; ++p) nf[p] = (double)nobs[p];
    CU(st->alloc(&st->n_obs, P * sizeof(double)));
    CU(cudaMemcpy(st->n_obs, nf.data(), P * sizeof(double), cudaMemcpyHostToDevice));
  }

  // ---- row-major device copies used only for packing ----
  double *dQ = nullptr, *dY = nullptr, *dM = nullptr, *dYraw = nullptr, *dT = nullptr;
  if (K > 0) {
    CU(cudaMalloc(&dQ, (size_t)N * K * sizeof(double)));
    tmp.v.push_back(dQ);
    CU(cudaMemcpy(dQ, d->Q, (size_t)N * K * sizeof(double), in_kind));
    CU(cudaMalloc(&dT, (size_t)K * P * sizeof(double)));
    tmp.v.push_back(dT);
  }
  CU(cudaMalloc(&dY, (size_t)N * P * sizeof(double)));
  tmp.v.push_back(dY);
  if (st->verbose) {
    CU(cudaMalloc(&dM, (size_t)N * P * sizeof(double)));
    tmp.v.push_back(dM);
    CU(cudaMemcpy(dM, d->Y_mask, (size_t)N * P * sizeof(double), in_kind));
    CU(cudaMalloc(&dYraw, (size_t)N * P * sizeof(double)));
    tmp.v.push_back(dYraw);
    CU(cudaMemcpy(dYraw, d->Y_raw, (size_t)N * P * sizeof(double), in_kind));
  }

  // ---- intercept column: np.linalg.qr of [1 | covariates] (lin_reg.py:142-147) gives a constant
  // first column +-1/sqrt(N); its product with x is q0 * sum(x), which the pre-pass already has ----
  if (K > 0 && !flag_knob(d->opt.keep_intercept, "GLR_KEEP_INTERCEPT")) {
    std::vector<double> col0(N);
    CU(cudaMemcpy2D(col0.data(), sizeof(double), dQ, (size_t)K * sizeof(double), sizeof(double), N,
                    cudaMemcpyDeviceToHost));
    long double sum = 0;
    for (int64_t i = 0; i < N; ++i) sum += col0[i];
    const double q0 = (double)(sum / N);
    double dev = 0;
    for (int64_t i = 0; i < N; ++i) dev = std::max(dev, std::fabs(col0[i] - q0));
    if (q0 != 0.0 && dev <= 1e-13 * std::fabs(q0)) {
      st->has_icpt = 1;
      st->icpt_q0 = q0;
    }
  }
  const int Kg = K - st->has_icpt;  // Q columns that go through the GEMM

  // ---- pass-1 operand W_c = [Q | Ytilde_c | mask | Yraw], Ytilde_c = (I - Q Q^T) Y_c ----
  st->Mcols = Kg + P * (st->verbose ? 3 : 1);
  const int n_mtiles = (int)ceil_div(st->Mcols, 8);
  const int full_tiles = st->Mcols / 8, rem_cols = st->Mcols % 8;
  const bool no_tail = flag_knob(d->opt.no_tail, "GLR_NO_TAIL");
  if (!no_tail && full_tiles >= 1 && full_tiles <= 4 && rem_cols >= 1 && rem_cols <= 4) {
    // 8*MT + 1..4 columns: the remainder goes through FP64 FMAs instead of a zero-padded m-tile (an FMA column costs
    // the FP64 datapath what a DMMA column costs, so up to half a tile it is cheaper than padding to 8)
    st->MT = full_tiles;
    st->tail = rem_cols;
    st->n_groups = 1;
  } else if (n_mtiles <= kMaxMT) {
    st->MT = pick_mt(n_mtiles);
    st->n_groups = 1;
  } else {
    st->MT = kMaxMT;
    st->n_groups = (int)ceil_div(n_mtiles, kMaxMT);
  }
  st->s_rows = st->has_icpt + (int64_t)st->n_groups * st->MT * 8 + (st->tail ? 8 : 0);
  st->w_contig_doubles =
      (int64_t)st->n_groups * st->n_chunks * kChunkGroups * (st->MT * kAtomDoubles + st->tail * kGroup);
  const size_t w_bytes = (size_t)C * st->w_contig_doubles * sizeof(double);
  CU(st->alloc(&st->wpk, w_bytes));
  CU(cudaMemsetAsync(st->wpk, 0, w_bytes, s0));
  // integer tensor-core operand (split8.cu): any number of columns (blocks of up to 18)
  double* d_colmax = nullptr;
  const int split8_mode = tri_knob(d->opt.split8, "GLR_SPLIT8");
  if (split8_mode != GLR_OFF && st->Mcols > 0 && Ng < (int64_t)8000000 && Ng > 3 * 128) {  // >= 4 stages: one per issuer warp
    st->s8_cols = st->Mcols;
    split8_geometry(st->s8_cols, 0, 7, &st->s8_ncb, &st->s8_cpb, &st->s8_NC);
    st->s8_stages = ceil_div(Ng, 128);
    st->s8_contig_bytes = (int64_t)st->s8_ncb * st->s8_stages * st->s8_NC * 128;
    // the digit image costs about 7.1 bytes per W entry (the FP64 image costs 8): built while it fits 16 GiB
    if ((int64_t)C * st->s8_contig_bytes <= ((int64_t)16 << 30)) st->split8 = split8_mode == GLR_FORCE ? 2 : 1;
    // reduced digits for the Q columns: 5 when every phenotype is fully observed, 6 when a masked pass follows (its
    // certificate is looser); worth building only when it removes a 16-row step of the MMA N dimension
    int qd = knob_int(d->opt.q_digits, "GLR_Q_DIGITS");
    if (qd == 0) qd = st->U > 0 ? 6 : 5;
    if (st->split8 && st->s8_ncb == 1 && Kg > 0 && qd >= 3 && qd < 7) {
      int ncb, cpb, nc;
      split8_geometry(st->s8_cols, Kg, qd, &ncb, &cpb, &nc);
      if (nc < st->s8_NC && (int64_t)C * (st->s8_contig_bytes + st->s8_stages * nc * 128) <= ((int64_t)16 << 30)) {
        st->s8_n_lo = Kg;
        st->s8_lo_digits = qd;
        st->s8_NC_lo = nc;
        st->s8_contig_bytes_lo = st->s8_stages * nc * 128;
      }
    }
  }
  if (st->split8) {
    CU(st->alloc(&st->w8, (size_t)C * st->s8_contig_bytes));
    CU(cudaMemsetAsync(st->w8, 0, (size_t)C * st->s8_contig_bytes, s0));
    CU(st->alloc(&st->s8_scale, (size_t)C * st->s8_cols * sizeof(double)));
    CU(cudaMalloc(&d_colmax, (size_t)st->s8_cols * sizeof(double)));
    tmp.v.push_back(d_colmax);
    if (st->s8_n_lo) {
      CU(st->alloc(&st->w8_lo, (size_t)C * st->s8_contig_bytes_lo));
      CU(cudaMemsetAsync(st->w8_lo, 0, (size_t)C * st->s8_contig_bytes_lo, s0));
      CU(st->alloc(&st->s8_scale_lo, (size_t)C * st->s8_cols * sizeof(double)));
    }
  }
  for (int c = 0; c < C; ++c) {
    double* dst = st->wpk + (int64_t)c * st->w_contig_doubles;
    CU(cudaMemcpyAsync(dY, d->Y + (int64_t)c * N * P, (size_t)N * P * sizeof(double), in_kind, s0));
    if (K > 0) {
      launch_qty(dQ, dY, N, K, P, dT, s0);
      launch_residualize(dQ, dT, dY, N, K, P, s0);
      if (Kg > 0)
        launch_pack_a(dQ, N, K, st->has_icpt, Kg, d_g2r, Ng, dst, 0, st->MT, st->tail, st->n_chunks, s0);
    }
    launch_pack_a(dY, N, P, 0, P, d_g2r, Ng, dst, Kg, st->MT, st->tail, st->n_chunks, s0);
    if (st->split8) {
      for (int img = 0; img < (st->s8_n_lo ? 2 : 1); ++img) {  // the full image, then the reduced-digit one
        int8_t* d8 = img ? st->w8_lo + (int64_t)c * st->s8_contig_bytes_lo : st->w8 + (int64_t)c * st->s8_contig_bytes;
        const int nc = img ? st->s8_NC_lo : st->s8_NC, n_lo = img ? st->s8_n_lo : 0, lo = st->s8_lo_digits;
        if (Kg > 0)
          launch_pack_split8(dQ, N, K, st->has_icpt, Kg, d_colmax, d_g2r, Ng, d8, 0, st->s8_cpb, nc, st->s8_stages, n_lo, lo, s0);
        launch_pack_split8(dY, N, P, 0, P, d_colmax + Kg, d_g2r, Ng, d8, Kg, st->s8_cpb, nc, st->s8_stages, n_lo, lo, s0);
        if (st->verbose) {  // mask and raw-phenotype columns of the verbose outputs (lin_reg.py:234-237)
          launch_pack_split8(dM, N, P, 0, P, d_colmax + Kg + P, d_g2r, Ng, d8, Kg + P, st->s8_cpb, nc, st->s8_stages, n_lo, lo, s0);
          launch_pack_split8(dYraw, N, P, 0, P, d_colmax + Kg + 2 * P, d_g2r, Ng, d8, Kg + 2 * P, st->s8_cpb, nc,
                             st->s8_stages, n_lo, lo, s0);
        }
      }
      CU(cudaMemcpyAsync(st->s8_scale + (int64_t)c * st->s8_cols, d_colmax, (size_t)st->s8_cols * sizeof(double),
                         cudaMemcpyDeviceToDevice, s0));
    }
    if (st->verbose) {
      launch_pack_a(dM, N, P, 0, P, d_g2r, Ng, dst, Kg + P, st->MT, st->tail, st->n_chunks, s0);
      launch_pack_a(dYraw, N, P, 0, P, d_g2r, Ng, dst, Kg + 2 * P, st->MT, st->tail, st->n_chunks, s0);
    }
  }
  CU(cudaGetLastError());

  // ---- masked-pass operands ----
  if (st->U > 0) {
    const int U = st->U;
    const int KT = (int)ceil_div(std::max(K, 1), 4);
    st->KT2 = (int)ceil_div(KT, 2);
    const bool dmma_ok = 2 * st->KT2 <= 8;  // the FP64 DMMA form keeps K <= 32 coefficients per lane
    const int n_mt2 = (int)ceil_div(U, 8);
    if (n_mt2 <= kMaxMT) {
      st->MT2 = pick_mt_masked(n_mt2);
      st->n_groups2 = 1;
    } else {
      st->MT2 = kMaxMT;
      st->n_groups2 = (int)ceil_div(n_mt2, kMaxMT);
    }
    // unique mask columns, row-major [N x U]
    std::vector<double> mu((size_t)N * U);
    for (int64_t s = 0; s < N; ++s)
      for (int u = 0; u < U; ++u) mu[(size_t)s * U + u] = mask_host[s * P + rep[u]];
    double* dMu = nullptr;
    CU(cudaMalloc(&dMu, mu.size() * sizeof(double)));
    tmp.v.push_back(dMu);
    CU(cudaMemcpy(dMu, mu.data(), mu.size() * sizeof(double), cudaMemcpyHostToDevice));
    // Three forms of the pass; cost per variant in FP64-FMA equivalents (measured rates):
    //   sparse complement (masked_sparse.cu)  T * K            T = unobserved (sample, mask) pairs
    //   FP64 DMMA (masked_kernel)             Ng * (K + U)     K <= 32
    //   int8 tensor cores (mask8.cu)          Ng * (115 + 0.055 U)   K <= 64
    // GLR_MASK_SPARSE / GLR_MASK8: 0 = never, 1 (default) = when cheapest, 2 = always
    // (0 = never, 1 = when cheapest, 2 = always)
    const int tri2mode[3] = {1, 0, 2};
    const int sparse_mode = tri2mode[tri_knob(d->opt.mask_sparse, "GLR_MASK_SPARSE")];
    const int mask8_mode = tri2mode[tri_knob(d->opt.mask_int8, "GLR_MASK8")];
    std::vector<int32_t> sp_idx;
    std::vector<int64_t> sp_rows, sp_off(1, 0);
    for (int u = 0; u < U; ++u) {
      for (int64_t sg = 0; sg < Ng; ++sg) {
        const int64_t row = g2r.empty() ? sg : g2r[sg];
        if (row >= 0 && mask_host[row * P + rep[u]] == 0.0) {
          sp_idx.push_back((int32_t)sg);
          sp_rows.push_back(row);
        }
      }
      sp_off.push_back((int64_t)sp_idx.size());
    }
    const double T = (double)sp_idx.size();
    const double cost_sparse = T * std::max(K, 1);
    const double cost_dmma = dmma_ok ? (double)Ng * (K + U) : 1e300;
    int m8_nmb = 0, m8_NB = 0;
    mask8_geometry(U, &m8_nmb, &m8_NB);
    const bool int8_ok = K <= 64 && Ng < (int64_t)16000000 &&
                         (size_t)m8_nmb * (size_t)ceil_div(Ng, 128) * m8_NB * 128 <= ((size_t)8 << 30);
    const double cost_int8 = int8_ok ? (double)Ng * (115.0 + 0.055 * U) : 1e300;
    // the sparse-complement form has a generic FP64 kernel for any number of covariates: with K > 64 it is the only
    // form, whatever the knobs say (the reference has no limit on K, lin_reg.py:241)
    const bool only_sparse = !dmma_ok && !int8_ok;
    const bool sparse_ok = (sparse_mode > 0 || only_sparse) && Ng < ((int64_t)1 << 31);
    bool use_sparse = sparse_ok && (sparse_mode > 1 || only_sparse || (cost_sparse <= cost_dmma && cost_sparse <= cost_int8));
    bool use_int8 = !use_sparse && mask8_mode > 0 && int8_ok && (mask8_mode > 1 || cost_int8 < cost_dmma);
    if (mask8_mode > 1 && sparse_mode <= 1 && int8_ok && !only_sparse) {  // an explicit request for the int8 form wins over "cheapest"
      use_sparse = false;
      use_int8 = true;
    }
    if (!use_sparse && !use_int8 && !dmma_ok && Ng < ((int64_t)1 << 31)) use_sparse = true;  // the only form left
    if (use_sparse) {
      st->sparse_masked = std::max(sparse_mode, 1);
      st->sp_total = (int64_t)sp_idx.size();
      const int Kp = (K + 1) & ~1;
      int64_t* d_rows = nullptr;
      CU(st->alloc(&st->sp_idx, std::max<size_t>(sp_idx.size(), 1) * sizeof(int32_t)));
      CU(st->alloc(&st->sp_off, sp_off.size() * sizeof(int64_t)));
      CU(st->alloc(&st->sp_qm, std::max<size_t>(sp_idx.size() * Kp, 1) * sizeof(double)));
      CU(cudaMalloc(&d_rows, std::max<size_t>(sp_rows.size(), 1) * sizeof(int64_t)));
      tmp.v.push_back(d_rows);
      CU(cudaMemcpy(st->sp_idx, sp_idx.data(), sp_idx.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
      CU(cudaMemcpy(st->sp_off, sp_off.data(), sp_off.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
      CU(cudaMemcpy(d_rows, sp_rows.data(), sp_rows.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
      launch_gather_rows(dQ, K, Kp, d_rows, st->sp_total, st->sp_qm, s0);
    }
    if (use_int8) {
      mask8_geometry(U, &st->m8_nmb, &st->m8_NB);
      st->m8_stages = ceil_div(Ng, 128);
      const size_t m8_bytes = (size_t)st->m8_nmb * st->m8_stages * st->m8_NB * 128;
      if (m8_bytes <= ((size_t)8 << 30)) {
        st->mask8 = mask8_mode;
        CU(st->alloc(&st->m8, m8_bytes));
        CU(cudaMemsetAsync(st->m8, 0, m8_bytes, s0));
        launch_pack_mask8(dMu, N, U, d_g2r, Ng, st->m8_NB, st->m8_stages, st->m8, s0);
        CU(st->alloc(&st->qg, (size_t)std::max<int64_t>(Ng * K, 1) * sizeof(double)));
        launch_gather_q(dQ, N, K, d_g2r, Ng, st->qg, s0);
      }
    }
    if (!st->mask8 && !st->sparse_masked) {
      if (!dmma_ok)
        return fail(GLR_ERR_INVALID, "masked pass: no applicable form (more than 2^31 genotype samples with more than 32 covariates)");
      const size_t qb_bytes = (size_t)st->n_chunks * kChunkGroups * st->KT2 * kAtomDoubles * sizeof(double);
      CU(st->alloc(&st->qb, qb_bytes));
      CU(cudaMemsetAsync(st->qb, 0, qb_bytes, s0));
      if (K > 0) launch_pack_qb(dQ, N, K, d_g2r, Ng, st->qb, st->KT2, st->n_chunks, s0);
      const size_t m_bytes =
          (size_t)st->n_groups2 * st->n_chunks * kChunkGroups * st->MT2 * kAtomDoubles * sizeof(double);
      CU(st->alloc(&st->mpk, m_bytes));
      CU(cudaMemsetAsync(st->mpk, 0, m_bytes, s0));
      launch_pack_a(dMu, N, U, 0, U, d_g2r, Ng, st->mpk, 0, st->MT2, 0, st->n_chunks, s0);
    }
  }
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  if (st->split8) {  // scale_c = max|w_c| * 2^-54 (2^-(8 d - 2) for a column with d digits)
    std::vector<double> sc((size_t)C * st->s8_cols), lo;
    CU(cudaMemcpy(sc.data(), st->s8_scale, sc.size() * sizeof(double), cudaMemcpyDeviceToHost));
    if (st->s8_n_lo) {
      lo = sc;
      for (size_t i = 0; i < lo.size(); ++i)
        lo[i] = std::ldexp(lo[i], (int)(i % st->s8_cols) < st->s8_n_lo ? -(8 * st->s8_lo_digits - 2) : -54);
      CU(cudaMemcpy(st->s8_scale_lo, lo.data(), lo.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    for (auto& v : sc) v = std::ldexp(v, -54);
    CU(cudaMemcpy(st->s8_scale, sc.data(), sc.size() * sizeof(double), cudaMemcpyHostToDevice));
  }

  st->n_lanes_req = d->n_lanes;
  if (int rc2 = create_lanes(st)) return rc2;
  guard.s = nullptr;
  *out = st;
  return 0;
}

static size_t row_bytes_of(int format, int64_t n) {
  switch (format) {
    case GLR_FMT_F64: return (size_t)n * 8;
    case GLR_FMT_F32: return (size_t)n * 4;
    case GLR_FMT_I8: return (size_t)n;
    default: return (size_t)((n + 3) / 4);
  }
}

// `replay`: re-run of a block whose previous run was void (its genotypes are already on the device), a set of:
//   kReplayFull  the optimistic form met a missing call -> form with the missing-call operand
//   kReplayFp64  the int8 tensor-core form of an int8 block met a value outside {-1, 0, 1, 2} -> FP64 path
//   kReplayExact a test failed the digit certificate of the reduced-digit image -> full 7-digit image
enum { kReplayFull = 1, kReplayFp64 = 2, kReplayExact = 4 };
static int submit_impl(glr_state* st, int lane_id, const glr_block_desc* b, const glr_block_out* o, int replay) {
  if (!st || !b || !o) return fail(GLR_ERR_INVALID, "null argument");
  if (lane_id < 0 || lane_id >= (int)st->lanes.size()) return fail(GLR_ERR_INVALID, "lane out of range");
  Lane& L = st->lanes[lane_id];
  if (L.busy && !replay) return fail(GLR_ERR_STATE, "lane already has a block in flight");
  if (b->format < GLR_FMT_F64 || b->format > GLR_FMT_PLINK2) return fail(GLR_ERR_INVALID, "unknown genotype format");
  if (b->contig < 0 || b->contig >= st->C) return fail(GLR_ERR_INVALID, "contig index out of range");
  if (b->n_variants < 0) return fail(GLR_ERR_INVALID, "negative n_variants");
  const int G = b->n_variants;
  const size_t rb = row_bytes_of(b->format, st->Ng);
  if (G > 0 && (!b->genotypes || (size_t)b->row_stride_bytes < rb))
    return fail(GLR_ERR_INVALID, "genotype buffer missing or row stride smaller than a row");
  if (G > 0 && (!o->effect || !o->stderror || !o->tvalue || !o->pvalue))
    return fail(GLR_ERR_INVALID, "missing output array");
  if (G > 0 && st->verbose && (!o->n || !o->sum_x || !o->y_transpose_x))
    return fail(GLR_ERR_INVALID, "verbose state needs n / sum_x / y_transpose_x outputs");
  if ((b->format == GLR_FMT_F64 && (b->row_stride_bytes % 8 || (uintptr_t)b->genotypes % 8)) ||
      (b->format == GLR_FMT_F32 && (b->row_stride_bytes % 4 || (uintptr_t)b->genotypes % 4)))
    return fail(GLR_ERR_INVALID, "genotype rows must be aligned to their element size");
  CU(cudaSetDevice(st->device));
  L.launches = 0;
  L.has_masked = false;
  L.has_prepass = false;
  L.optimistic = false;
  L.lo_digits = false;
  L.counted = false;
  L.checked = false;
  L.path = replay ? GLR_PATH_REPLAYED : 0;
  if (!replay) {
    L.blk = *b;
    L.blk_out = *o;
  }
  if (G == 0) {
    L.busy = true;
    for (auto& e : L.ev) CU(cudaEventRecord(e, L.stream));
    return 0;
  }
  cudaStream_t s = L.stream;
  const int P = st->P, K = st->K;

  // ---- plan ----
  const int nt = 2;
  const int warps = (b->format == GLR_FMT_PLINK2 && st->MT <= 4) ? 16 : 8;
  const int tile = gram_tile_variants(nt, warps);
  const int64_t ldS = std::max(round_up(round_up(G, tile), 8), round_up(G, 256));
  const int64_t ctas = ceil_div(G, tile) * st->n_groups;
  int n_split = 1;
  if (ctas < st->sm_count) {
    n_split = (int)std::min<int64_t>(ceil_div(2 * st->sm_count, ctas), std::max<int64_t>(1, st->n_chunks / 8));
    n_split = std::max(1, std::min(n_split, 64));
  } else if (ctas < 12 * (int64_t)st->sm_count) {
    // a few waves: the last, partly filled one costs a whole wave.  Splitting the sample range in 2..4 makes the waves
    // shorter (5.3 waves of work run as 6 whole ones, but as 11 halves = 5.5); partial sums are reduced in fixed order
    double best = (double)ceil_div(ctas, st->sm_count);
    for (int ns = 2; ns <= 4; ++ns) {
      if (st->n_chunks / ns < 16) break;
      const double t = (double)ceil_div(ctas * ns, st->sm_count) / ns + 0.03 * ns;  // + reduce pass and shorter pipelines
      if (t < best) {
        best = t;
        n_split = ns;
      }
    }
  }
  if (st->force_split > 0) n_split = (int)std::min<int64_t>(st->force_split, st->n_chunks);
  const int64_t Mpad = st->s_rows;

  // ---- buffers ----
  const uint8_t* d_geno = reinterpret_cast<const uint8_t*>(b->genotypes);
  const bool float_fmt = b->format == GLR_FMT_F64 || b->format == GLR_FMT_F32;
  // float blocks with dropped samples (intersect_samples) are staged so that the dropped entries can be zeroed: the
  // reference deletes those rows before any arithmetic (lin_reg.py:228-229), so a NaN there must not reach the sums
  const bool scrub = float_fmt && st->n_drop > 0;
  if (replay) {
    if (b->memspace == GLR_MEM_HOST) d_geno = L.geno.as<uint8_t>();  // staged by the first run
  } else if (b->memspace == GLR_MEM_HOST || scrub) {
    const size_t bytes = (size_t)(G - 1) * b->row_stride_bytes + rb;
    if (int rc = L.geno.ensure(bytes + 64)) return rc;
    CU(cudaMemcpyAsync(L.geno.p, b->genotypes, bytes,
                       b->memspace == GLR_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s));
    d_geno = L.geno.as<uint8_t>();
  }
  if (int rc = L.s_red.ensure((size_t)Mpad * ldS * 8)) return rc;
  if (n_split > 1)
    if (int rc = L.s_part.ensure((size_t)n_split * Mpad * ldS * 8)) return rc;
  if (int rc = L.mom.ensure((size_t)2 * ldS * 8)) return rc;
  if (int rc = L.v1.ensure((size_t)ldS * 8)) return rc;
  if (int rc = L.xx.ensure((size_t)3 * ldS * 8)) return rc;
  if (float_fmt && n_split > 1)
    if (int rc = L.mom_part.ensure((size_t)n_split * 2 * ldS * 8)) return rc;
  const int n_out = st->verbose ? 7 : 4;
  double* outs[7] = {o->effect, o->stderror, o->tvalue, o->pvalue, o->n, o->sum_x, o->y_transpose_x};
  double* d_outs[7];
  const size_t out_elems = (size_t)P * G;
  if (o->memspace == GLR_MEM_HOST) {
    if (int rc = L.out.ensure(out_elems * 8 * n_out)) return rc;
    for (int i = 0; i < n_out; ++i) d_outs[i] = L.out.as<double>() + i * out_elems;
  } else {
    for (int i = 0; i < n_out; ++i) d_outs[i] = outs[i];
  }

  BlockView x;
  x.base = d_geno;
  x.stride = b->row_stride_bytes;
  x.n_variants = G;
  x.format = b->format;
  x.n_geno = st->Ng;
  x.v1 = L.v1.as<double>();

  // int8 tensor-core first pass when it is the faster one: a work unit (CTA or CTA pair) spends ~540 cycles per 128-sample
  // stage; the plan splits the sample range of small blocks so that every SM has work.  The DMMA kernel does about
  // 30 TFLOP/s on the padded column count plus its pre-pass.
  bool use_split8 = false;
  Split8Plan plan;
  // int8 genotype states take the same kernel (128 bytes per stage instead of 32, values checked on the fly) when the rows
  // are 4-byte aligned
  const bool i8_ok = b->format == GLR_FMT_I8 && !(replay & kReplayFp64) &&
                     ((reinterpret_cast<uintptr_t>(d_geno) | (uintptr_t)b->row_stride_bytes) & 3) == 0;
  if (st->split8 && (b->format == GLR_FMT_PLINK2 || i8_ok)) {
    plan = split8_plan(G, st->s8_ncb, st->s8_stages, st->sm_count, st->s8_pair_mode, st->force_split);
    const double t_s8 = plan.makespan_stages * (i8_ok ? 700.0 : 540.0) / 1.7e9 + (plan.n_split > 1 ? 8e-6 : 0.0);
    const double t_dmma = (double)G * (2.0 * (double)st->Ng * (double)Mpad / 30e12 + (double)st->Ng / 4.0 / 5e12) /
                              std::min(1.0, (double)(ctas * n_split) / st->sm_count) + 10e-6;
    use_split8 = st->split8 >= 2 || t_s8 < t_dmma;
  }
  const bool fused_counts = use_split8 && st->n_drop == 0;
  // optimistic form: no missing-call operand while the stream of blocks has shown no missing call (validated in the
  // kernel; a violation re-runs the block at wait time)
  const bool optimistic = use_split8 && fused_counts && !(replay & (kReplayFull | kReplayFp64)) &&
                          st->optimistic_mode != GLR_OFF && !st->expect_missing;
  // reduced-digit image unless a certificate failed recently
  const bool lo_digits = use_split8 && st->s8_n_lo > 0 && !(replay & kReplayExact) && st->exact_hold == 0;
  if (scrub) {
    launch_zero_dropped(d_geno, b->row_stride_bytes, G, b->format, st->drop_idx, st->n_drop, s);
    ++L.launches;
  }
  if (use_split8 && plan.n_split > 1) {
    if (int rc = L.s8_ws.ensure(split8_ws_bytes(st->s8_ncb, st->s8_NC, ldS))) return rc;
    if (int rc = L.s8_n00.ensure((size_t)ldS * sizeof(int32_t))) return rc;
  }
  CU(cudaMemsetAsync(L.flags.p, 0, 2 * sizeof(int32_t), s));
  CU(cudaEventRecord(L.ev[0], s));
  // ---- pre-pass (A0): missing substitute + sum of squares from integer counts ----
  if (!float_fmt && !fused_counts) {
    launch_prepass(x, b->missing_policy, L.v1.as<double>(), L.mom.as<double>(), ldS, s);
    L.has_prepass = true;
    L.path |= GLR_PATH_PREPASS;
    ++L.launches;
  }
  CU(cudaEventRecord(L.ev[1], s));
  // ---- pass 1 (A1-A3): S = [Q | Ytilde | ...]^T X ----
  GramParams gp;
  gp.x = x;
  gp.wpk = st->wpk + (int64_t)b->contig * st->w_contig_doubles;
  gp.MT = st->MT;
  gp.tail = st->tail;
  gp.n_groups = st->n_groups;
  gp.n_chunks = st->n_chunks;
  gp.n_split = n_split;
  gp.nt = nt;
  gp.warps = warps;
  gp.ldS = ldS;
  // GEMM rows start below the intercept row (row 0 of S) when the intercept is handled outside
  gp.S = (n_split > 1 ? L.s_part.as<double>() : L.s_red.as<double>()) + (int64_t)st->has_icpt * ldS;
  gp.s_slab = Mpad * ldS;
  gp.mom_part = (float_fmt && n_split > 1) ? L.mom_part.as<double>() : L.mom.as<double>();
  if (use_split8) {
    Split8Params sp{};
    sp.x = x;
    sp.w8 = lo_digits ? st->w8_lo + (int64_t)b->contig * st->s8_contig_bytes_lo
                      : st->w8 + (int64_t)b->contig * st->s8_contig_bytes;
    sp.scale = (lo_digits ? st->s8_scale_lo : st->s8_scale) + (int64_t)b->contig * st->s8_cols;
    sp.n_cols = st->s8_cols;
    sp.n_cb = st->s8_ncb;
    sp.cpb = st->s8_cpb;
    sp.NC = lo_digits ? st->s8_NC_lo : st->s8_NC;
    sp.n_lo = lo_digits ? st->s8_n_lo : 0;
    sp.lo_digits = st->s8_lo_digits;
    sp.fused_counts = fused_counts ? 1 : 0;
    sp.missing_policy = b->missing_policy;
    sp.n_stages = st->s8_stages;
    sp.S = L.s_red.as<double>() + (int64_t)st->has_icpt * ldS;
    sp.ldS = ldS;
    sp.v1_out = L.v1.as<double>();
    sp.mom_out = L.mom.as<double>();
    sp.pair = plan.pair;
    sp.no_e = optimistic ? 1 : 0;
    sp.n_split = plan.n_split;
    sp.split_stages = plan.split_stages;
    sp.z_ws = L.s8_ws.as<int32_t>();
    sp.n00_ws = L.s8_n00.as<int32_t>();
    sp.flags = L.flags.as<int32_t>();
    launch_split8(sp, st->sm_count, s);
    L.optimistic = optimistic;
    L.counted = fused_counts;
    L.lo_digits = lo_digits;
    L.checked = fused_counts || b->format == GLR_FMT_I8 || lo_digits;
    L.path |= GLR_PATH_SPLIT8 | (plan.pair ? GLR_PATH_SPLIT8_PAIR : 0) | (plan.n_split > 1 ? GLR_PATH_SPLIT8_NSPLIT : 0) |
              (optimistic ? GLR_PATH_OPTIMISTIC : 0) | (lo_digits ? GLR_PATH_Q_DIGITS_LO : 0);
    if (plan.n_split > 1) L.launches += 1;
  } else {
    launch_gram(gp, s);
    L.path |= GLR_PATH_DMMA;
  }
  ++L.launches;
  if (n_split > 1 && !use_split8) {
    launch_reduce_splits(L.s_part.as<double>(), L.s_red.as<double>(), Mpad, ldS, n_split, s);
    ++L.launches;
    if (float_fmt) {
      launch_reduce_splits(L.mom_part.as<double>(), L.mom.as<double>(), 2, ldS, n_split, s);
      ++L.launches;
    }
  }
  if (st->n_drop > 0 && !scrub) {  // (scrubbed float blocks hold zeros at the dropped samples: nothing to subtract)
    launch_dropped_moments(x, st->drop_idx, st->n_drop, L.mom.as<double>(), ldS, s);
    ++L.launches;
  }
  if (st->has_icpt) {
    launch_intercept_row(L.mom.as<double>() + ldS, st->icpt_q0, L.s_red.as<double>(), G, s);
    ++L.launches;
  }
  CU(cudaEventRecord(L.ev[2], s));
  // ---- pass 2 (A4 for phenotypes with missing samples): D = Mu^T (X - Q A)^2 ----
  const double* d_D = nullptr;
  bool d_is_complement = false;
  if (st->U > 0 && st->sparse_masked) {
    // sparse complement form: sum of r^2 over the unobserved samples of every distinct mask
    const int U = st->U;
    const int64_t Upad = (int64_t)st->n_groups2 * st->MT2 * 8;
    if (int rc = L.d_red.ensure((size_t)Upad * ldS * 8)) return rc;
    const int64_t vblocks = ceil_div(G, 256);
    const int64_t longest = std::max<int64_t>(1, ceil_div(st->sp_total, (int64_t)U));
    int n_parts = (int)std::min<int64_t>(std::max<int64_t>(1, ceil_div(4 * st->sm_count, vblocks * U)), ceil_div(longest, 128));
    n_parts = std::max(1, std::min(n_parts, 64));
    if (n_parts > 1)
      if (int rc = L.d_part.ensure((size_t)n_parts * U * ldS * 8)) return rc;
    SparseMaskedParams sp{};
    sp.x = x;
    sp.A = L.s_red.as<double>();
    sp.ldS = ldS;
    sp.K = K;
    sp.idx = st->sp_idx;
    sp.qm = st->sp_qm;
    sp.off = st->sp_off;
    sp.U = U;
    sp.n_parts = n_parts;
    sp.comp = n_parts > 1 ? L.d_part.as<double>() : L.d_red.as<double>();
    launch_sparse_masked(sp, s);
    ++L.launches;
    if (n_parts > 1) {
      launch_reduce_splits(L.d_part.as<double>(), L.d_red.as<double>(), U, ldS, n_parts, s);
      ++L.launches;
    }
    d_D = L.d_red.as<double>();
    d_is_complement = true;
    L.has_masked = true;
    L.path |= K > 64 ? GLR_PATH_MASK_GENERIC : GLR_PATH_MASK_SPARSE;
  } else if (st->U > 0 && st->mask8) {
    // integer tensor-core form (mask8.cu), in sub-ranges of variants that bound the digit image to 3 GiB
    const int64_t Upad = (int64_t)st->n_groups2 * st->MT2 * 8;
    if (int rc = L.d_red.ensure((size_t)Upad * ldS * 8)) return rc;
    if (int rc = L.m8_rmax.ensure((size_t)ldS * 8)) return rc;
    const int tv = mask8_tile_variants();
    const size_t tile_bytes = mask8_tile_bytes(st->m8_stages);
    int64_t tiles_sub = std::max<int64_t>(2, (int64_t)(((size_t)3 << 30) / tile_bytes) / 2 * 2);
    tiles_sub = std::min<int64_t>(tiles_sub, round_up(ceil_div(G, tv), 2));
    if (int rc = L.m8_z8.ensure((size_t)tiles_sub * tile_bytes)) return rc;
    if (int rc = L.m8_Z.ensure(mask8_z_bytes((int)tiles_sub, st->m8_nmb, st->m8_NB))) return rc;
    for (int64_t v0 = 0; v0 < G; v0 += tiles_sub * tv) {
      Mask8Params mp;
      mp.x = x;
      mp.qg = st->qg;
      mp.A = L.s_red.as<double>();
      mp.ldS = ldS;
      mp.K = K;
      mp.v0 = (int)v0;
      mp.nv = (int)std::min<int64_t>(tiles_sub * tv, G - v0);
      mp.n_stages = st->m8_stages;
      mp.rmax_bits = L.m8_rmax.as<unsigned long long>();
      mp.z8 = L.m8_z8.as<int8_t>();
      mp.m8 = st->m8;
      mp.n_mb = st->m8_nmb;
      mp.NB = st->m8_NB;
      mp.U = st->U;
      mp.Z = L.m8_Z.as<int32_t>();
      mp.D = L.d_red.as<double>();
      launch_mask8(mp, st->sm_count, s);
      L.launches += 4;
    }
    d_D = L.d_red.as<double>();
    L.has_masked = true;
    L.path |= GLR_PATH_MASK_INT8;
  } else if (st->U > 0) {
    const int64_t Upad = (int64_t)st->n_groups2 * st->MT2 * 8;
    const int tile2 = gram_tile_variants(2);
    const int64_t ctas2 = ceil_div(G, tile2) * st->n_groups2;
    int n_split2 = 1;
    if (ctas2 < st->sm_count) {
      n_split2 = (int)std::min<int64_t>(ceil_div(2 * st->sm_count, ctas2), std::max<int64_t>(1, st->n_chunks / 8));
      n_split2 = std::max(1, std::min(n_split2, 64));
    }
    if (st->force_split > 0) n_split2 = (int)std::min<int64_t>(st->force_split, st->n_chunks);
    if (int rc = L.d_red.ensure((size_t)Upad * ldS * 8)) return rc;
    if (n_split2 > 1)
      if (int rc = L.d_part.ensure((size_t)n_split2 * Upad * ldS * 8)) return rc;
    MaskedParams mp;
    mp.x = x;
    mp.qb = st->qb;
    mp.mpk = st->mpk;
    mp.A = L.s_red.as<double>();
    mp.K = K;
    mp.KT2 = st->KT2;
    mp.MT = st->MT2;
    mp.n_groups = st->n_groups2;
    mp.n_chunks = st->n_chunks;
    mp.n_split = n_split2;
    mp.ldS = ldS;
    mp.D = n_split2 > 1 ? L.d_part.as<double>() : L.d_red.as<double>();
    launch_masked(mp, s);
    ++L.launches;
    if (n_split2 > 1) {
      launch_reduce_splits(L.d_part.as<double>(), L.d_red.as<double>(), Upad, ldS, n_split2, s);
      ++L.launches;
    }
    d_D = L.d_red.as<double>();
    L.has_masked = true;
    L.path |= GLR_PATH_MASK_DMMA;
  }
  CU(cudaEventRecord(L.ev[3], s));
  // ---- epilogue (A5-A7) ----
  EpilogueParams ep;
  ep.S = L.s_red.as<double>();
  ep.sxx = L.mom.as<double>();
  ep.D = d_D;
  ep.d_is_complement = d_is_complement ? 1 : 0;
  ep.mask_id = st->mask_id;
  ep.YdotY = st->YdotY + (int64_t)b->contig * P;
  ep.Y_scale = st->Y_scale;
  ep.n_obs = st->n_obs;
  ep.ldS = ldS;
  ep.K = K;
  ep.P = P;
  ep.G = G;
  ep.verbose = st->verbose;
  ep.dof = (double)st->dof;
  ep.pc = st->pc;
  ep.effect = d_outs[0];
  ep.stderror = d_outs[1];
  ep.tvalue = d_outs[2];
  ep.pvalue = d_outs[3];
  ep.n = st->verbose ? d_outs[4] : nullptr;
  ep.sum_x = st->verbose ? d_outs[5] : nullptr;
  ep.ytx = st->verbose ? d_outs[6] : nullptr;
  if (L.lo_digits) {
    ep.lo_scale = st->s8_scale_lo + (int64_t)b->contig * st->s8_cols;
    ep.sum_x_row = L.mom.as<double>() + ldS;
    ep.n_lo = st->s8_n_lo;
    ep.lo_row0 = st->has_icpt;
    ep.flags = L.flags.as<int32_t>();
  }
  launch_epilogue(ep, L.xx.as<double>(), s);
  L.launches += 2;
  CU(cudaEventRecord(L.ev[4], s));
  CU(cudaGetLastError());
  if (o->memspace == GLR_MEM_HOST) {
    for (int i = 0; i < n_out; ++i)
      CU(cudaMemcpyAsync(outs[i], d_outs[i], out_elems * 8, cudaMemcpyDeviceToHost, s));
  }
  if (L.checked) CU(cudaMemcpyAsync(L.h_flags, L.flags.p, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  CU(cudaEventRecord(L.ev[5], s));
  L.busy = true;
  return 0;
}

int glr_block_submit(glr_state* st, int lane_id, const glr_block_desc* b, const glr_block_out* o) {
  return submit_impl(st, lane_id, b, o, 0);
}

int glr_block_wait(glr_state* st, int lane_id) {
  if (!st) return fail(GLR_ERR_INVALID, "null state");
  if (lane_id < 0 || lane_id >= (int)st->lanes.size()) return fail(GLR_ERR_INVALID, "lane out of range");
  Lane& L = st->lanes[lane_id];
  if (!L.busy) return fail(GLR_ERR_STATE, "no block in flight on this lane");
  CU(cudaSetDevice(st->device));
  CU(cudaStreamSynchronize(L.stream));
  if (L.checked) {
    // void results: an int8 block with values outside {-1, 0, 1, 2} goes to the FP64 path (any int8 value is legal there);
    // an optimistic block that had missing calls after all is run again in the full form; a block with a test whose digit
    // certificate failed is run again on the 7-digit image.  Each re-run can only add conditions: at most 3 rounds.
    int mode = 0;
    for (int round = 0; round < 3 && L.checked; ++round) {
      int need = mode;
      if (L.h_flags[0] & 2) need |= kReplayFp64;
      if (L.optimistic && (L.h_flags[0] & 1)) need |= kReplayFull;
      if (L.lo_digits && (L.h_flags[0] & 4)) need |= kReplayExact;
      if (need == mode) break;
      mode = need;
      if (mode & kReplayFull) st->expect_missing = true;
      if (mode & kReplayExact) st->exact_hold = 32;
      if (int rc = submit_impl(st, lane_id, &L.blk, &L.blk_out, mode)) {
        L.busy = false;
        return rc;
      }
      CU(cudaStreamSynchronize(L.stream));
    }
    if (L.counted && !L.optimistic) st->expect_missing = L.h_flags[1] != 0;
  }
  if (!(L.path & GLR_PATH_REPLAYED) && st->exact_hold > 0) --st->exact_hold;
  L.busy = false;
  glr_timing t{};
  cudaEventElapsedTime(&t.total_ms, L.ev[0], L.ev[4]);
  cudaEventElapsedTime(&t.prepass_ms, L.ev[0], L.ev[1]);
  cudaEventElapsedTime(&t.gram_ms, L.ev[1], L.ev[2]);
  cudaEventElapsedTime(&t.masked_ms, L.ev[2], L.ev[3]);
  cudaEventElapsedTime(&t.epilogue_ms, L.ev[3], L.ev[4]);
  t.kernel_launches = L.launches;
  t.path = L.path;
  L.timing = t;
  CU(cudaGetLastError());
  return 0;
}

int glr_run_block(glr_state* st, const glr_block_desc* blk, const glr_block_out* out) {
  int rc = glr_block_submit(st, 0, blk, out);
  if (rc) return rc;
  return glr_block_wait(st, 0);
}

int glr_block_timing(glr_state* st, int lane_id, glr_timing* t) {
  if (!st || !t) return fail(GLR_ERR_INVALID, "null argument");
  if (lane_id < 0 || lane_id >= (int)st->lanes.size()) return fail(GLR_ERR_INVALID, "lane out of range");
  *t = st->lanes[lane_id].timing;
  return 0;
}

int glr_lane_stream(glr_state* st, int lane_id, void** stream) {
  if (!st || !stream) return fail(GLR_ERR_INVALID, "null argument");
  if (lane_id < 0 || lane_id >= (int)st->lanes.size()) return fail(GLR_ERR_INVALID, "lane out of range");
  *stream = (void*)st->lanes[lane_id].stream;
  return 0;
}

int glr_host_alloc(size_t bytes, void** ptr) {
  if (!ptr) return fail(GLR_ERR_INVALID, "null argument");
  CU(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return 0;
}
int glr_host_free(void* ptr) {
  if (ptr) CU(cudaFreeHost(ptr));
  return 0;
}

}  // extern "C"
namespace glr {
int api_fail(int code, const std::string& msg) { return fail(code, msg); }
int api_check_device(int device) { return check_device(device); }
}  // namespace glr
extern "C" {

int glr_device_alloc(int device, size_t bytes, void** ptr) {
  if (!ptr) return fail(GLR_ERR_INVALID, "null argument");
  if (int rc = check_device(device)) return rc;
  CU(cudaMalloc(ptr, bytes ? bytes : 1));
  return 0;
}
int glr_device_free(int device, void* ptr) {
  if (!ptr) return 0;
  if (int rc = check_device(device)) return rc;
  CU(cudaFree(ptr));
  return 0;
}
int glr_device_upload(int device, void* dst, const void* src_host, size_t bytes) {
  if (bytes && (!dst || !src_host)) return fail(GLR_ERR_INVALID, "null argument");
  if (int rc = check_device(device)) return rc;
  CU(cudaMemcpy(dst, src_host, bytes, cudaMemcpyHostToDevice));
  return 0;
}

int glr_bgen_to_plink2(int device, const void* probs_host, const uint8_t* ploidy_host, int32_t n_variants, int64_t n_samples,
                       int32_t bits, double hard_call_threshold, void* out_device, int64_t out_stride_bytes) {
  if (n_variants < 0 || n_samples <= 0 || (bits != 8 && bits != 16 && bits != 32)) return fail(GLR_ERR_INVALID, "bad arguments");
  if (n_variants == 0) return 0;
  if (!probs_host || !ploidy_host || !out_device || out_stride_bytes < (n_samples + 3) / 4)
    return fail(GLR_ERR_INVALID, "missing buffer or output stride smaller than a row");
  if (int rc = check_device(device)) return rc;
  const size_t pbytes = (size_t)n_variants * 2 * n_samples * (bits / 8), lbytes = (size_t)n_variants * n_samples;
  void *dp = nullptr, *dl = nullptr;
  CU(cudaMalloc(&dp, pbytes));
  cudaError_t e = cudaMalloc(&dl, lbytes);
  if (e == cudaSuccess) e = cudaMemcpy(dp, probs_host, pbytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dl, ploidy_host, lbytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    launch_bgen_to_plink2(dp, (const uint8_t*)dl, n_variants, n_samples, bits, hard_call_threshold, (uint8_t*)out_device,
                          out_stride_bytes, nullptr);
    e = cudaDeviceSynchronize();
  }
  cudaFree(dp);
  if (dl) cudaFree(dl);
  if (e != cudaSuccess) return fail(GLR_ERR_CUDA, std::string("glr_bgen_to_plink2: ") + cudaGetErrorString(e));
  return 0;
}

int glr_loco_predict(int device, int64_t n_samples, int32_t n_headers, int32_t n_sample_blocks, int32_t n_chromosomes,
                     const double* X_raw, const double* mu, const double* sig, const double* B, const int32_t* header_chromosome,
                     const int32_t* sample_block, const double* cov, int32_t n_cov, const double* B_cov, double* out) {
  if (n_samples <= 0 || n_headers < 0 || n_sample_blocks <= 0 || n_chromosomes <= 0 || n_cov < 0)
    return fail(GLR_ERR_INVALID, "bad dimensions");
  if ((n_headers > 0 && (!X_raw || !mu || !sig || !B || !header_chromosome)) || !sample_block || !out || (n_cov > 0 && (!cov || !B_cov)))
    return fail(GLR_ERR_INVALID, "missing array");
  const int64_t N = n_samples;
  const int H = n_headers, C = n_chromosomes;
  // headers grouped by chromosome (stable: the order inside a chromosome is the caller's)
  std::vector<int32_t> perm, seg(C + 1, 0);
  for (int h = 0; h < H; ++h) {
    if (header_chromosome[h] < 0 || header_chromosome[h] >= C) return fail(GLR_ERR_INVALID, "header chromosome index out of range");
    if (sig[h] == 0.0) return fail(GLR_ERR_INVALID, "Standard deviation cannot be 0.");  // model_functions.py:140-141
    ++seg[header_chromosome[h] + 1];
  }
  for (int c = 0; c < C; ++c) seg[c + 1] += seg[c];
  perm.resize(std::max(H, 1));
  {
    std::vector<int32_t> fill(seg.begin(), seg.end() - 1);
    for (int h = 0; h < H; ++h) perm[fill[header_chromosome[h]]++] = h;
  }
  for (int64_t s = 0; s < N; ++s)
    if (sample_block[s] < 0 || sample_block[s] >= n_sample_blocks) return fail(GLR_ERR_INVALID, "sample block index out of range");
  if (int rc = check_device(device)) return rc;
  struct Tmp {
    std::vector<void*> v;
    ~Tmp() { for (void* p : v) if (p) cudaFree(p); }
    cudaError_t up(const void* src, size_t bytes, void** dst) {
      cudaError_t e = cudaMalloc(dst, bytes ? bytes : 1);
      if (e != cudaSuccess) return e;
      v.push_back(*dst);
      return bytes ? cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) : cudaSuccess;
    }
  } tmp;
  void *dX, *dmu, *dsig, *dB, *dperm, *dseg, *dsb, *dcov = nullptr, *dBc = nullptr, *dout;
  CU(tmp.up(X_raw, (size_t)N * H * 8, &dX));
  CU(tmp.up(mu, (size_t)H * 8, &dmu));
  CU(tmp.up(sig, (size_t)H * 8, &dsig));
  CU(tmp.up(B, (size_t)n_sample_blocks * H * 8, &dB));
  CU(tmp.up(perm.data(), (size_t)H * 4, &dperm));
  CU(tmp.up(seg.data(), (size_t)(C + 1) * 4, &dseg));
  CU(tmp.up(sample_block, (size_t)N * 4, &dsb));
  if (n_cov > 0) {
    CU(tmp.up(cov, (size_t)N * n_cov * 8, &dcov));
    CU(tmp.up(B_cov, (size_t)n_sample_blocks * n_cov * 8, &dBc));
  }
  CU(cudaMalloc(&dout, (size_t)C * N * 8));
  tmp.v.push_back(dout);
  launch_loco_predict((const double*)dX, N, H, (const double*)dmu, (const double*)dsig, (const double*)dB, (const int32_t*)dperm,
                      (const int32_t*)dseg, C, (const int32_t*)dsb, (const double*)dcov, n_cov, (const double*)dBc, (double*)dout, nullptr);
  CU(cudaGetLastError());
  CU(cudaMemcpy(out, dout, (size_t)C * N * 8, cudaMemcpyDeviceToHost));
  return 0;
}

int glr_tensor_i8_issue_rate(int device, double milliseconds, double* tops) {
  if (!tops || !(milliseconds > 0.0)) return fail(GLR_ERR_INVALID, "bad arguments");
  int rc = check_device(device);
  if (rc) return rc;
  CU(cudaSetDevice(device));
  int sm_count = 0;
  CU(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
  const double r = measure_s8_issue_rate(sm_count, milliseconds, nullptr);
  if (r < 0.0) return fail(GLR_ERR_CUDA, "tensor-core issue-rate probe failed");
  *tops = r;
  return 0;
}

int glr_t_two_sided_pvalues(int device, const double* tvalues, int64_t n, double dof, double* pvalues) {
  if (n < 0 || (n > 0 && (!tvalues || !pvalues))) return fail(GLR_ERR_INVALID, "bad arguments");
  int rc = check_device(device);
  if (rc) return rc;
  if (n == 0) return 0;
  double *dt = nullptr, *dp = nullptr;
  CU(cudaMalloc(&dt, n * sizeof(double)));
  CU(cudaMalloc(&dp, n * sizeof(double)));
  CU(cudaMemcpy(dt, tvalues, n * sizeof(double), cudaMemcpyHostToDevice));
  launch_pvalues(dt, n, make_pval_const(dof), dp, nullptr);
  cudaError_t e = cudaMemcpy(pvalues, dp, n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(dt);
  cudaFree(dp);
  if (e != cudaSuccess) return fail(GLR_ERR_CUDA, cudaGetErrorString(e));
  return 0;
}

int glr_decode_block(int device, const glr_block_desc* b, int64_t n_geno, double* out_values) {
  if (!b || !out_values || n_geno <= 0 || b->n_variants < 0) return fail(GLR_ERR_INVALID, "bad arguments");
  if (b->memspace != GLR_MEM_HOST) return fail(GLR_ERR_INVALID, "glr_decode_block takes host buffers");
  int rc = check_device(device);
  if (rc) return rc;
  const int G = b->n_variants;
  if (G == 0) return 0;
  const size_t rb = row_bytes_of(b->format, n_geno);
  if ((size_t)b->row_stride_bytes < rb) return fail(GLR_ERR_INVALID, "row stride smaller than a row");
  const size_t bytes = (size_t)(G - 1) * b->row_stride_bytes + rb;
  uint8_t* dg = nullptr;
  double *dv1 = nullptr, *dsxx = nullptr, *dout = nullptr;
  CU(cudaMalloc(&dg, bytes + 64));
  CU(cudaMalloc(&dv1, G * sizeof(double)));
  CU(cudaMalloc(&dsxx, 2 * G * sizeof(double)));
  CU(cudaMalloc(&dout, (size_t)G * n_geno * sizeof(double)));
  CU(cudaMemcpy(dg, b->genotypes, bytes, cudaMemcpyHostToDevice));
  BlockView x;
  x.base = dg;
  x.stride = b->row_stride_bytes;
  x.n_variants = G;
  x.format = b->format;
  x.n_geno = n_geno;
  x.v1 = dv1;
  if (b->format == GLR_FMT_I8 || b->format == GLR_FMT_PLINK2)
    launch_prepass(x, b->missing_policy, dv1, dsxx, G, nullptr);
  launch_decode(x, dout, nullptr);
  cudaError_t e = cudaMemcpy(out_values, dout, (size_t)G * n_geno * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(dg);
  cudaFree(dv1);
  cudaFree(dsxx);
  cudaFree(dout);
  if (e != cudaSuccess) return fail(GLR_ERR_CUDA, cudaGetErrorString(e));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Multi-GPU inside one process (SURVEY.md section 8e / 8b export group 5).  Variants shard across devices, the state
// is replicated ONCE: it is packed on the first device and every packed array is broadcast to the others in one grouped
// NCCL call (ncclBroadcast over NVLink / NVSwitch).  NCCL is loaded at run time (dlopen of libnccl.so.2; GLR_NCCL_LIB
// overrides the name) so that the library has no link-time dependency on it; without NCCL the same copies go through
// cudaMemcpyPeerAsync.  There is no collective in the block loop.
// ------------------------------------------------------------------------------------------------
namespace {
struct NcclApi {
  void* lib = nullptr;
  int (*CommInitAll)(void** comms, int ndev, const int* devlist) = nullptr;
  int (*CommDestroy)(void* comm) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Broadcast)(const void* send, void* recv, size_t count, int dtype, int root, void* comm, cudaStream_t stream) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok() const { return CommInitAll && CommDestroy && GroupStart && GroupEnd && Broadcast; }
};
NcclApi load_nccl() {
  NcclApi a;
  const char* names[] = {getenv("GLR_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n || !*n) continue;
    a.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (a.lib) break;
  }
  if (!a.lib) return a;
  a.CommInitAll = reinterpret_cast<decltype(a.CommInitAll)>(dlsym(a.lib, "ncclCommInitAll"));
  a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(a.lib, "ncclCommDestroy"));
  a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(dlsym(a.lib, "ncclGroupStart"));
  a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(dlsym(a.lib, "ncclGroupEnd"));
  a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(dlsym(a.lib, "ncclBroadcast"));
  a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(a.lib, "ncclGetErrorString"));
  return a;
}
}  // namespace

struct glr_comm {
  std::vector<int> devices;
  std::vector<cudaStream_t> streams;
  std::vector<void*> nccl;  // ncclComm_t per device, empty when the peer-copy transport is used
  NcclApi api;
  std::string transport;
};

int glr_comm_destroy(glr_comm* c) {
  if (!c) return 0;
  for (size_t i = 0; i < c->devices.size(); ++i) {
    cudaSetDevice(c->devices[i]);
    if (i < c->nccl.size() && c->nccl[i]) c->api.CommDestroy(c->nccl[i]);
    if (i < c->streams.size() && c->streams[i]) cudaStreamDestroy(c->streams[i]);
  }
  delete c;
  return 0;
}

int glr_comm_init(const int32_t* devices, int32_t n_devices, glr_comm** out) {
  if (!devices || n_devices <= 0 || !out) return fail(GLR_ERR_INVALID, "bad arguments");
  *out = nullptr;
  for (int i = 0; i < n_devices; ++i) {
    if (int rc = check_device(devices[i])) return rc;
    for (int j = 0; j < i; ++j)
      if (devices[j] == devices[i]) return fail(GLR_ERR_INVALID, "duplicate device in glr_comm_init");
  }
  glr_comm* c = new glr_comm();
  c->devices.assign(devices, devices + n_devices);
  c->streams.resize(n_devices, nullptr);
  for (int i = 0; i < n_devices; ++i) {
    if (cudaSetDevice(devices[i]) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking) != cudaSuccess) {
      glr_comm_destroy(c);
      return fail(GLR_ERR_CUDA, "cannot create a stream on a device of the communicator");
    }
  }
  c->transport = "p2p";
  const char* no_nccl = getenv("GLR_COMM_NO_NCCL");
  if (n_devices > 1 && !(no_nccl && atoi(no_nccl))) {
    c->api = load_nccl();
    if (c->api.ok()) {
      c->nccl.assign(n_devices, nullptr);
      if (c->api.CommInitAll(c->nccl.data(), n_devices, c->devices.data()) == 0) c->transport = "nccl";
      else c->nccl.clear();
    }
  }
  if (c->transport == "p2p") {
    for (int i = 1; i < n_devices; ++i) {  // direct NVLink copies from the first device
      int can = 0;
      cudaDeviceCanAccessPeer(&can, devices[i], devices[0]);
      if (can) {
        cudaSetDevice(devices[i]);
        if (cudaDeviceEnablePeerAccess(devices[0], 0) != cudaSuccess) cudaGetLastError();
      }
    }
  }
  *out = c;
  return 0;
}

int glr_comm_size(glr_comm* c) { return c ? (int)c->devices.size() : 0; }
const char* glr_comm_transport(glr_comm* c) { return c ? c->transport.c_str() : ""; }

int glr_comm_broadcast_state(glr_comm* c, const glr_state_desc* desc, glr_state** out_states) {
  if (!c || !desc || !out_states) return fail(GLR_ERR_INVALID, "null argument");
  const int n = (int)c->devices.size();
  for (int i = 0; i < n; ++i) out_states[i] = nullptr;
  glr_state_desc d0 = *desc;
  d0.device = c->devices[0];
  if (int rc = glr_state_create(&d0, &out_states[0])) return rc;
  glr_state* src = out_states[0];
  auto fail_all = [&](int code, const std::string& msg) {
    for (int i = 0; i < n; ++i) {
      if (out_states[i]) glr_state_destroy(out_states[i]);
      out_states[i] = nullptr;
    }
    return fail(code, msg);
  };
  // replicas: same metadata, own device arrays (same sizes), own lanes
  for (int i = 1; i < n; ++i) {
    glr_state* r = new glr_state(*src);
    r->lanes.clear();
    r->device = c->devices[i];
    for (auto& a : r->arrays) *reinterpret_cast<void**>(reinterpret_cast<char*>(r) + a.first) = nullptr;
    out_states[i] = r;
    if (cudaSetDevice(r->device) != cudaSuccess) return fail_all(GLR_ERR_CUDA, "cudaSetDevice failed");
    cudaDeviceGetAttribute(&r->sm_count, cudaDevAttrMultiProcessorCount, r->device);
    for (auto& a : r->arrays) {
      void* p = nullptr;
      if (cudaMalloc(&p, a.second ? a.second : 1) != cudaSuccess) return fail_all(GLR_ERR_CUDA, "cudaMalloc of a state replica failed");
      *reinterpret_cast<void**>(reinterpret_cast<char*>(r) + a.first) = p;
    }
    if (create_lanes(r)) return fail_all(GLR_ERR_CUDA, g_err);
  }
  if (n == 1) return 0;
  auto field = [](glr_state* s, size_t off) { return *reinterpret_cast<void**>(reinterpret_cast<char*>(s) + off); };
  if (c->transport == "nccl") {
    // ONE group: every array, every device (root = the first device, in place there)
    int e = c->api.GroupStart();
    for (auto& a : src->arrays) {
      if (!a.second) continue;
      for (int i = 0; i < n && e == 0; ++i) {
        cudaSetDevice(c->devices[i]);
        e = c->api.Broadcast(field(src, a.first), field(out_states[i], a.first), a.second, /*ncclUint8*/ 1, 0, c->nccl[i], c->streams[i]);
      }
    }
    const int e2 = c->api.GroupEnd();
    if (e || e2)
      return fail_all(GLR_ERR_CUDA, std::string("ncclBroadcast: ") + (c->api.GetErrorString ? c->api.GetErrorString(e ? e : e2) : "error"));
  } else {
    for (int i = 1; i < n; ++i) {
      cudaSetDevice(c->devices[i]);
      for (auto& a : src->arrays)
        if (a.second &&
            cudaMemcpyPeerAsync(field(out_states[i], a.first), c->devices[i], field(src, a.first), c->devices[0], a.second, c->streams[i]) !=
                cudaSuccess)
          return fail_all(GLR_ERR_CUDA, "cudaMemcpyPeerAsync of a state array failed");
    }
  }
  for (int i = 0; i < n; ++i) {
    cudaSetDevice(c->devices[i]);
    if (cudaStreamSynchronize(c->streams[i]) != cudaSuccess) return fail_all(GLR_ERR_CUDA, "state broadcast failed");
  }
  return 0;
}

}  // extern "C"
