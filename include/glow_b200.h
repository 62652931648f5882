/*
 * glow_b200.h -- C ABI of the B200-native linear-regression GWAS block kernel.
 *
 * The reference (projectglow/glow) has no FFI for this path: the seam is the Python call
 *   _linear_regression_inner(genotype_pdf, Y_state, Y_mask, Y_scale, Q, dof, phenotype_names,
 *                            Y_for_verbose_output, verbose_output, gt_indices_to_drop)
 * at python/glow/gwas/lin_reg.py:203-254, reached from map_func (lin_reg.py:277-282) through
 * _loco_dispatch (python/glow/gwas/functions.py:92-101).  Each entry point below cites the
 * reference lines it replaces.  All functions are extern "C", take plain pointers and sizes, never
 * throw, and return 0 on success or a negative glr_status; glr_last_error() gives the message of
 * the last failure on the calling thread.
 *
 * Ownership: the caller owns every buffer it passes; the library owns all device memory behind a
 * handle.  Calls on one handle must be serialised by the caller; distinct handles (one per GPU)
 * may be driven from distinct host threads.
 */
#ifndef GLOW_B200_H
#define GLOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLR_ABI_VERSION 3

typedef enum glr_status {
  GLR_OK = 0,
  GLR_ERR_INVALID = -1,  /* bad argument (maps to Python ValueError) */
  GLR_ERR_CUDA = -2,     /* CUDA runtime failure */
  GLR_ERR_NO_DEVICE = -3,/* no usable sm_100 device: the product path has no CPU fallback */
  GLR_ERR_STATE = -4     /* call sequence error (e.g. wait on an idle lane) */
} glr_status;

/* Genotype block encodings.  One block = G_b variants, variant-major: variant g starts at
 * genotypes + g * row_stride_bytes. */
typedef enum glr_format {
  /* N doubles per variant: the `_glow_regression_values` array<double> column
   * (functions.py:56-68), i.e. the rows np.column_stack gathers at lin_reg.py:226-227. */
  GLR_FMT_F64 = 0,
  /* N floats per variant (dt=np.float32 input arrays, functions.py:47-53); widened to f64. */
  GLR_FMT_F32 = 1,
  /* N int8 per variant: genotype_states output, -1 = missing
   * (core/.../sql/expressions/VariantUtilExprs.scala:37-59). */
  GLR_FMT_I8 = 2,
  /* PLINK .bed variant-major block, ceil(N/4) bytes per variant, sample s in byte s/4 bits
   * 2(s%4): code 00 -> 2, 01 -> missing, 10 -> 1, 11 -> 0
   * (core/.../plink/PlinkRowToInternalRowConverter.scala:40-47,56-66; PlinkFileFormat.scala:282-284). */
  GLR_FMT_PLINK2 = 3
} glr_format;

typedef enum glr_memspace { GLR_MEM_HOST = 0, GLR_MEM_DEVICE = 1 } glr_memspace;

/* How missing calls (-1 in I8, code 01 in PLINK2) become numbers. */
typedef enum glr_missing_policy {
  GLR_MISSING_AS_MINUS_ONE = 0, /* genotype_states semantics: the value -1 enters the regression */
  GLR_MISSING_MEAN = 1          /* mean_substitute(genotype_states(..)):
                                   core/.../sql/expressions/MeanSubstitute.scala:39-108; all-missing -> -1 */
} glr_missing_policy;

/* Kernel-selection knobs.  All zero = the defaults (every choice made by the library's cost models). */
typedef enum glr_tristate { GLR_AUTO = 0, GLR_OFF = 1, GLR_FORCE = 2 } glr_tristate;

typedef struct glr_options {
  int32_t split8;          /* glr_tristate: exact int8 tensor-core contraction (tcgen05) for PLINK2 blocks: when a time
                              model says it beats the FP64 DMMA kernel | never | always */
  int32_t split8_pair;     /* glr_tristate: its CTA-pair form (cta_group::2): when the pairs fill the GPU | never | always */
  int32_t mask_sparse;     /* glr_tristate: sparse-complement form of the masked second pass: when cheapest | never | always */
  int32_t mask_int8;       /* glr_tristate: int8 tensor-core form of the masked second pass: when cheapest | never | always */
  int32_t keep_intercept;  /* 1: keep a constant first column of Q inside the GEMM */
  int32_t no_tail;         /* 1: pad the column count to a multiple of 8 instead of FP64-FMA tail columns */
  int32_t sample_split;    /* > 0: force this split of the sample range across CTAs (small blocks); 0 = automatic */
  int32_t optimistic_calls;/* glr_tristate: run PLINK2 blocks without the missing-call operand while the previous block of
                              the state had no missing call, re-running a block that turns out to have one (results are
                              identical either way): automatic | never | (FORCE = start optimistic at the first block) */
  int32_t q_digits;        /* radix-256 digits of the covariate columns in the int8 contraction: 0 = automatic (5, or 6
                              when a phenotype has missing samples), 7 = always the full precision.  With fewer than 7
                              digits every test carries a certified bound of the error of XdotX; a block with a test
                              whose bound exceeds 1e-11 XdotX is re-run with 7 digits, so results stay within 1e-11 of
                              the 7-digit ones (which are exact to FP64 rounding) */
} glr_options;

typedef struct glr_state glr_state; /* opaque */
typedef struct glr_comm glr_comm;   /* opaque: a set of devices of one process */

/* Driver-side state: what map_func closes over at lin_reg.py:277-282.
 * All matrices are row-major float64 in PHENOTYPE sample order. */
typedef struct glr_state_desc {
  int32_t abi_version;     /* GLR_ABI_VERSION */
  int32_t device;          /* CUDA device ordinal */
  int64_t n_samples;       /* N  = rows of Q / Y (after intersect_samples) */
  int64_t n_geno_samples;  /* length of each genotype array (>= N); == N unless rows are dropped */
  int32_t n_covariates;    /* K  = columns of Q (covariates + intercept) */
  int32_t n_phenotypes;    /* P */
  int32_t n_contigs;       /* C  = 1 unless LOCO offsets (one YState per contig, lin_reg.py:166-199) */
  int32_t verbose;         /* also produce n / sum_x / y_transpose_x (lin_reg.py:234-237) */
  const double* Q;         /* [N x K]   lin_reg.py:147 */
  const double* Y;         /* [C][N x P] YState.Y per contig, lin_reg.py:194-199 */
  const double* Y_mask;    /* [N x P]   0/1, lin_reg.py:149 */
  const double* YdotY;     /* [C][P]    YState.YdotY */
  const double* Y_scale;   /* [P]       lin_reg.py:154 */
  const double* Y_raw;     /* [N x P]   Y_for_verbose_output (zero-filled raw phenotypes) or NULL */
  const int64_t* keep_idx; /* [N] genotype-array index of each phenotype row (np.delete at
                              lin_reg.py:228-229 expressed as a gather), or NULL for identity */
  int64_t dof;             /* N - K - 1, lin_reg.py:159 */
  int32_t max_block_variants; /* hint for workspace sizing (grown on demand) */
  int32_t n_lanes;         /* independent in-flight blocks (streams); 0 -> default 2 */
  int32_t memspace;        /* glr_memspace of Q, Y, Y_mask, YdotY, Y_scale, Y_raw (keep_idx is always
                              host): GLR_MEM_DEVICE lets ranks != 0 build their replica straight
                              from the buffer an NCCL broadcast filled, without a host round trip */
  int32_t reserved;
  glr_options opt;         /* kernel-selection knobs; all zero = defaults.  For tests every knob can also be overridden
                              from the environment at state creation: GLR_SPLIT8, GLR_S8_PAIR, GLR_MASK_SPARSE, GLR_MASK8,
                              GLR_OPTIMISTIC = 0 (never) | 1 (automatic) | 2 (always); GLR_KEEP_INTERCEPT, GLR_NO_TAIL,
                              GLR_SPLIT = n */
} glr_state_desc;

/* One genotype block in, = the `genotype_pdf` argument of _linear_regression_inner. */
typedef struct glr_block_desc {
  int32_t format;           /* glr_format */
  int32_t memspace;         /* glr_memspace of `genotypes` */
  const void* genotypes;
  int64_t row_stride_bytes; /* distance between consecutive variants */
  int32_t n_variants;       /* G_b */
  int32_t contig;           /* index into the C YStates (0 when C == 1): _loco_dispatch, functions.py:92-101 */
  int32_t missing_policy;   /* glr_missing_policy (I8 / PLINK2 only) */
  int32_t reserved;
} glr_block_desc;

/* Results out: four (seven if verbose) [P x G_b] row-major arrays = the P x G_b matrices the
 * reference ravels into its output columns (lin_reg.py:247-252): phenotype-major, variant-minor. */
typedef struct glr_block_out {
  int32_t memspace; /* glr_memspace of all output pointers */
  int32_t reserved;
  double* effect;   /* betas * Y_scale */
  double* stderror; /* standard_error * Y_scale */
  double* tvalue;
  double* pvalue;
  double* n;             /* verbose only, may be NULL */
  double* sum_x;         /* verbose only */
  double* y_transpose_x; /* verbose only */
} glr_block_out;

typedef struct glr_timing {
  float total_ms;    /* first kernel start -> last kernel end of the block (CUDA events) */
  float prepass_ms;  /* count / impute pre-pass (0 when the int8 tensor-core kernel counts on the fly) */
  float gram_ms;     /* contraction kernel(s): FP64 DMMA, or the exact int8 digit-split kernel (2-bit blocks) */
  float masked_ms;   /* masked sum-of-squares kernel (0 when every phenotype is fully observed) */
  float epilogue_ms; /* statistics + p-values */
  int32_t kernel_launches;
  int32_t path;      /* glr_path_flags of the block: which kernels ran (tests assert the path they mean to cover) */
} glr_timing;

typedef enum glr_path_flags {
  GLR_PATH_SPLIT8 = 1,        /* first pass on the int8 tensor cores (split8_kernel) */
  GLR_PATH_SPLIT8_PAIR = 2,   /*   ... in its CTA-pair form */
  GLR_PATH_SPLIT8_NSPLIT = 4, /*   ... with the sample range split across CTAs (small blocks) */
  GLR_PATH_PREPASS = 8,       /* separate count / impute pre-pass */
  GLR_PATH_DMMA = 16,         /* first pass on the FP64 tensor cores (gram_kernel) */
  GLR_PATH_MASK_SPARSE = 32,  /* masked pass: sparse complement */
  GLR_PATH_MASK_INT8 = 64,    /* masked pass: int8 tensor cores */
  GLR_PATH_MASK_DMMA = 128,   /* masked pass: FP64 DMMA */
  GLR_PATH_OPTIMISTIC = 256,  /* block ran without the missing-call operand */
  GLR_PATH_REPLAYED = 512,    /* ... and was re-run because it had missing calls */
  GLR_PATH_MASK_GENERIC = 1024, /* masked pass: generic FP64 form (more than 64 covariates) */
  GLR_PATH_Q_DIGITS_LO = 2048   /* first pass with reduced-digit covariate columns (results certified) */
} glr_path_flags;

int glr_abi_version(void);
const char* glr_last_error(void);
/* Number of visible CUDA devices of compute capability 10.x; <= 0 means the library cannot run. */
int glr_device_count(void);

/* Replaces the closure capture + broadcast of (Y_state, Y_mask, Y_scale, Q, dof) at
 * lin_reg.py:157-163,277-284.  Copies the arrays to `device` and packs them for the kernels. */
int glr_state_create(const glr_state_desc* desc, glr_state** out);
int glr_state_destroy(glr_state* st);
/* Number of distinct non-trivial missingness patterns among the phenotypes (0 = every phenotype
 * fully observed, the masked pass is skipped). */
int glr_state_unique_masks(glr_state* st);

/* ---- logistic-regression score test --------------------------------------------------
 * The block kernel _logistic_regression_inner of python/glow/gwas/log_reg.py:278-350 with correction = 'none':
 * genotypes residualised against the null-model fit (log_reg.py:263-275: X - C (C^T Gamma C)^-1 C^T Gamma X), score
 * statistic chisq = (sum X_res Y_res)^2 / (sum X_res^2 gamma) (log_reg.py:316-321), pvalue = chi2.sf(chisq, 1).  The
 * null-model fit (log_reg.py:160-200) stays on the host like the QR of the linear path; its results are the state. */
typedef struct glr_logreg_state glr_logreg_state; /* opaque */
typedef struct glr_logreg_desc {
  int32_t abi_version;       /* GLR_ABI_VERSION */
  int32_t device;
  int64_t n_samples;         /* N */
  int64_t n_geno_samples;    /* length of each genotype array (>= N) */
  int32_t n_covariates;      /* K = columns of C (covariates + intercept), log_reg.py:126-128 */
  int32_t n_phenotypes;      /* P */
  int32_t n_contigs;         /* LogRegState per contig with LOCO offsets, else 1 */
  int32_t reserved;
  const double* C;           /* [N x K] covariate matrix (NOT orthogonalised) */
  const double* gamma;       /* [contigs][N x P]  LogRegState.gamma = yhat (1 - yhat), 0 where the phenotype is missing */
  const double* Y_res;       /* [contigs][N x P]  LogRegState.Y_res = y - yhat, 0 where missing */
  const double* inv_CtGammaC;/* [contigs][P][K x K] LogRegState.inv_CtGammaC */
  const int64_t* keep_idx;   /* [N] genotype-array index of each phenotype row (np.delete, log_reg.py:303-304) or NULL */
  const double* Q;           /* [N x K] orthonormal basis of C (np.linalg.qr(C)[0], log_reg.py:134) or NULL.  Non-NULL =
                                correction 'approx-firth': the genotypes are first residualised linearly, X - Q Q^T X
                                (log_reg.py:312-313), and the score test runs on that (K <= 64) */
} glr_logreg_desc;
int glr_logreg_state_create(const glr_logreg_desc* desc, glr_logreg_state** out);
int glr_logreg_state_destroy(glr_logreg_state* st);
/* One block (same glr_block_desc as the linear path; synchronous): chisq and pvalue are [P x G_b] host arrays,
 * phenotype-major like the reference's np.ravel (log_reg.py:322-328). */
int glr_logreg_run_block(glr_logreg_state* st, const glr_block_desc* blk, double* chisq, double* pvalue);

/* Multi-GPU inside one process.  The reference parallelises over Spark partitions of variant rows
 * and ships the state to every worker by closure capture (lin_reg.py:277-284); here variants shard across the devices
 * of the communicator and the state is replicated ONCE: glr_comm_broadcast_state builds it on devices[0] and sends every
 * packed array to the other devices in one grouped ncclBroadcast (NCCL loaded at run time; cudaMemcpyPeerAsync when it
 * is absent -- glr_comm_transport says which).  out_states[i] lives on devices[i]; each is driven and destroyed like a
 * state from glr_state_create.  There is no collective in the block loop. */
int glr_comm_init(const int32_t* devices, int32_t n_devices, glr_comm** out);
int glr_comm_destroy(glr_comm* comm);
int glr_comm_size(glr_comm* comm);
const char* glr_comm_transport(glr_comm* comm); /* "nccl" | "p2p" */
int glr_comm_broadcast_state(glr_comm* comm, const glr_state_desc* desc, glr_state** out_states /* [n_devices] */);

/* Replaces one call of _linear_regression_inner (lin_reg.py:203-254).  Synchronous:
 * submit on lane 0 + wait. */
int glr_run_block(glr_state* st, const glr_block_desc* blk, const glr_block_out* out);
/* Asynchronous form: lanes are independent streams with their own staging buffers so the H2D
 * copy of block i+1 overlaps the kernels of block i.  Host buffers must stay valid (and should
 * be pinned, see glr_host_alloc) until glr_block_wait returns. */
int glr_block_submit(glr_state* st, int lane, const glr_block_desc* blk, const glr_block_out* out);
int glr_block_wait(glr_state* st, int lane);
/* CUDA-event timings of the last completed block on `lane`. */
int glr_block_timing(glr_state* st, int lane, glr_timing* t);
/* The CUDA stream (cudaStream_t) a lane launches on, for callers that time with their own events. */
int glr_lane_stream(glr_state* st, int lane, void** stream);

/* Pinned host memory for block staging. */
int glr_host_alloc(size_t bytes, void** ptr);
int glr_host_free(void* ptr);

/* Device memory for genotype blocks that stay RESIDENT in HBM across calls (one shard of the genotype matrix is loaded
 * once -- 15.6 GB per GPU for the UK-Biobank shape on 8 GPUs -- and many phenotype sets are then regressed against it
 * with glr_block_desc.memspace = GLR_MEM_DEVICE, so that the steady state is not bound by the PCIe copy of 125 kB per
 * variant).  glr_device_upload is synchronous. */
int glr_device_alloc(int device, size_t bytes, void** ptr);
int glr_device_free(int device, void* ptr);
int glr_device_upload(int device, void* dst, const void* src_host, size_t bytes);

/* BGEN v1.2 layout-2 probability blocks -> Glow hard calls -> PLINK 2-bit rows in device memory (SURVEY.md section 8f
 * row N2; reference: core/.../bgen/BgenFileIterator.scala, hardCallThreshold BgenSchemaInferrer.scala:42, genotype_states
 * VariantUtilExprs.scala:37-59).  probs_host: [n_variants][n_samples][2] stored probabilities P(hom ref), P(het) of
 * `bits` bits each (unphased, diploid, biallelic, as inflated from the file); ploidy_host: [n_variants][n_samples] ploidy
 * bytes (top bit = missing).  The rows land at out_device + g * out_stride_bytes and are then run like any
 * GLR_FMT_PLINK2 block with memspace = GLR_MEM_DEVICE: hard calls are exact small integers, so BGEN input takes the
 * int8 tensor-core path.  Synchronous. */
int glr_bgen_to_plink2(int device, const void* probs_host, const uint8_t* ploidy_host, int32_t n_variants, int64_t n_samples,
                       int32_t bits, double hard_call_threshold, void* out_device, int64_t out_stride_bytes);

/* GloWGR LOCO apply step for ONE label: the offsets RidgeRegression.transform_loco
 * (python/glow/wgr/ridge_regression.py:199-234) produces by applying the fitted model with the headers of one
 * chromosome removed, once per chromosome.  Per (sample block, label) the reference's apply_model
 * (python/glow/wgr/ridge_udfs.py:259-347) computes X @ B with X = [covariates | (X_raw - mu) / sig]
 * (assemble_block, model_functions.py:118-158).  Here
 *   out[c][s] = cov[s] . B_cov[block(s)] + sum over headers h with header_chromosome[h] != c of
 *               (X_raw[s][h] - mu[h]) / sig[h] * B[block(s)][h]
 * X_raw [n_samples x n_headers] row-major (the reduced block matrix of this label, samples in label order),
 * B [n_sample_blocks x n_headers] (coefficients of the chosen alpha for the model that excludes each sample block),
 * sample_block [n_samples], cov [n_samples x n_cov] / B_cov [n_sample_blocks x n_cov] (optional), out
 * [n_chromosomes x n_samples].  Host in / host out, synchronous.  sig == 0 is an error as in the reference. */
int glr_loco_predict(int device, int64_t n_samples, int32_t n_headers, int32_t n_sample_blocks, int32_t n_chromosomes,
                     const double* X_raw, const double* mu, const double* sig, const double* B, const int32_t* header_chromosome,
                     const int32_t* sample_block, const double* cov, int32_t n_cov, const double* B_cov, double* out);

/* Standalone A6: p = 2 * t.sf(|t|, dof) (lin_reg.py:245) for n device-computed values; host in/out.
 * Exposed so the special-function kernel can be tested against scipy/mpmath directly. */
int glr_t_two_sided_pvalues(int device, const double* tvalues, int64_t n, double dof, double* pvalues);
/* Standalone A0: decode a packed block to float64 [G_b x N] on the device (host in/out);
 * tests the fused decode against PlinkReaderSuite / MeanSubstitute semantics. */
int glr_decode_block(int device, const glr_block_desc* blk, int64_t n_geno_samples, double* out_values);

/* Measurement aid (no reference counterpart): the int8 tensor-core rate of this device right now, in TOP/s
 * (2 operations per multiply-accumulate), from a run of about `milliseconds` of back-to-back
 * tcgen05.mma kind::i8 with resident operands on every SM -- the upper bound of the int8 first pass.
 * bench.py reports it next to the roofline of that kernel. */
int glr_tensor_i8_issue_rate(int device, double milliseconds, double* tops);

#ifdef __cplusplus
}
#endif
#endif /* GLOW_B200_H */
