/* floppity_b200.h -- C ABI of the B200-native neural-spline-flow hot path.
 *
 * This is the drop-in boundary.  The reference (FlopPITy) reaches the path through
 * sbi/nflows Python objects:
 *     modules.py:63      posterior_nn(model='nsf', ...)                  (build)
 *     modules.py:64      SNPE_C(prior, density_estimator, device)        (trainer)
 *     floppity.py:216    inference.append_simulations(theta, x, proposal)
 *     floppity.py:219    inference.train(...)                            (fwd+bwd+Adam loop)
 *     floppity.py:225    inference.build_posterior(net).set_default_x(x)
 *     floppity.py:252    proposal.sample((n,))                           (inverse flow)
 *     floppityFUN.py:35  posterior.log_prob(theta, x=default_x)          (forward flow)
 * The host side of this repo (floppity_deprecated_b200/*.py) keeps those Python names and
 * calls the functions below through ctypes.  Conventions:
 *   - plain pointers + sizes, no torch types; every pointer is DEVICE memory owned by the
 *     caller (row-major, contiguous, fp32 unless stated), 16-byte aligned;
 *   - nothing is allocated or freed here; kernels are enqueued on `stream` (a cudaStream_t
 *     passed as void*) and the call returns without synchronising;
 *   - return value: 0 = ok, >0 = cudaError_t of a failed launch, <0 = FP_E_* below;
 *   - `status` words are device int32[4]: {n_nonfinite, n_negative_discriminant, n_fp16_range (operand pieces of the
 *     fused conditioner that reached the edge of fp16's range in 3xFP16 mode; re-run with FLOPPITY_FUSED_FP16=0), 0},
 *     accumulated with atomics so the Python wrapper can raise like upstream's
 *     assert_all_finite / `assert (discriminant >= 0).all()`.
 */
#ifndef FLOPPITY_B200_H
#define FLOPPITY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FP_E_BADARG   (-1)
#define FP_E_TOOLARGE (-2)
#define FP_E_NOTREADY (-3)

/* Hyper-parameters of build_nsf (sbi.neural_nets.flow.build_nsf as called through
 * posterior_nn at modules.py:63; flags floppity.py:34,35,48,50,57). */
typedef struct fp_nsf_config {
  int32_t D;              /* dim(theta)                                   */
  int32_t C;              /* dim(context) = spectral bins after embedding */
  int32_t T;              /* num_transforms (coupling + LU pairs)         */
  int32_t K;              /* num_bins                                     */
  int32_t H;              /* hidden_features                              */
  int32_t Bk;             /* num_blocks                                   */
  float tail_bound;       /* 3.0 (sbi default; floppity's -tail_bound is NOT forwarded) */
  float min_bin_width;    /* 1e-3 */
  float min_bin_height;   /* 1e-3 */
  float min_derivative;   /* 1e-3 */
  float lu_eps;           /* 1e-3 */
  float dropout_p;        /* floppity.py:57 */
} fp_nsf_config;

/* ---- parameter layout (one flat fp32 buffer; grads / Adam moments share it) ---------- */
enum fp_param_kind {
  FP_P_CTX_W = 0,   /* [H, C]   slot `block`: 0 = context columns of initial_layer, 1+b = blocks.b.context_layer */
  FP_P_CTX_B = 1,   /* [H]      slot `block`: 0 = initial_layer.bias,               1+b = blocks.b.context_layer.bias */
  FP_P_W0X   = 2,   /* [H, D]   theta columns of initial_layer, expanded to all D (transform columns are structural zeros) */
  FP_P_W1    = 3,   /* [H, H]   blocks.b.linear_layers.0.weight */
  FP_P_B1    = 4,   /* [H] */
  FP_P_W2    = 5,   /* [H, H]   blocks.b.linear_layers.1.weight */
  FP_P_B2    = 6,   /* [H] */
  FP_P_WF    = 7,   /* [d_tr*(3K-1), H] final_layer.weight */
  FP_P_BF    = 8,   /* [d_tr*(3K-1)] */
  FP_P_LU_LOWER = 9,  /* [D(D-1)/2] row-major strict lower */
  FP_P_LU_UPPER = 10, /* [D(D-1)/2] row-major strict upper */
  FP_P_LU_DIAG  = 11, /* [D] unconstrained_upper_diag */
  FP_P_LU_BIAS  = 12  /* [D] */
};
int64_t fp_nsf_param_count(const fp_nsf_config* cfg);
/* gradient buffer = parameter layout + a tail holding the LU product-matrix gradients and sum(d_logp) */
int64_t fp_nsf_grad_count(const fp_nsf_config* cfg);
int64_t fp_nsf_param_offset(const fp_nsf_config* cfg, int32_t layer, int32_t kind, int32_t block);
int64_t fp_nsf_param_size(const fp_nsf_config* cfg, int32_t layer, int32_t kind);
/* number of transformed features of coupling layer `layer` (mask alternates, even layer -> features 0,2,4..) */
int32_t fp_nsf_dtr(const fp_nsf_config* cfg, int32_t layer);
/* floats of the per-step derived buffer (W=L*U, L, U, logdet per layer) filled by fp_nsf_prepare */
int64_t fp_nsf_prepared_count(const fp_nsf_config* cfg);
/* floats of workspace for R flow rows sharing Rc context rows; training=1 keeps activations for backward */
int64_t fp_nsf_workspace_count(const fp_nsf_config* cfg, int64_t R, int64_t Rc, int32_t training);

/* ---- spline kernels: nflows rational_quadratic_spline behind
 *      PiecewiseRationalQuadraticCouplingTransform (reached from modules.py:63) ---------- */
/* theta_in/out: [R, D]; params: [R, d_tr, 3K-1] raw conditioner output (width/height logits are
 * divided by sqrt(H) inside); logdet[R] is ACCUMULATED (+=); bin_idx (may be NULL): int32 [R, d_tr],
 * -1 for inputs in the linear tails. Features f with f%2 == parity are transformed, the others copied. */
int fp_rqs_forward(const fp_nsf_config* cfg, int32_t parity, const float* theta_in, const float* params,
                   float* theta_out, float* logdet, int32_t* bin_idx, int64_t R, int32_t* status, void* stream);
int fp_rqs_inverse(const fp_nsf_config* cfg, int32_t parity, const float* theta_in, const float* params,
                   float* theta_out, float* logdet, int32_t* bin_idx, int64_t R, int32_t* status, void* stream);
/* d_theta_out [R, D], d_logdet [R] -> d_params [R, d_tr, 3K-1] (overwritten), d_theta_in [R, D] (overwritten:
 * transformed features get the spline input-gradient, identity features a copy of d_theta_out). */
int fp_rqs_backward(const fp_nsf_config* cfg, int32_t parity, const float* theta_in, const float* params,
                    const float* d_theta_out, const float* d_logdet, float* d_params, float* d_theta_in,
                    int64_t R, void* stream);

/* ---- fp32 SIMT GEMM with fused conditioner epilogues (nflows ResidualNet, reached from modules.py:63) ----
 * C[M,N] = op(A)[M,K] * op(B)[N,K]^T ; a_kmajor: A(m,k) at A[m*lda+k] else A[k*lda+m]; same for B.
 * flags: see FP_G_* ; aux pointers per epilogue. */
enum fp_gemm_flags {
  FP_G_RELU_A      = 1,   /* A := max(A,0) on load */
  FP_G_RELU_B      = 2,   /* B := max(B,0) on load */
  FP_G_BIAS        = 4,   /* + bias[n] */
  FP_G_ADD_CTX     = 8,   /* + G[m/ctx_div (0 if ctx_div==0), n]  (ldg) */
  FP_G_GLU_RES     = 16,  /* U[m,n]=v (if U); C = Rsd[m,n] + v*sigmoid(G[..]) */
  FP_G_RELUMASK    = 32,  /* C = v * (Msk[m,n] > 0) */
  FP_G_ACCUM       = 64,  /* C += (after the above) */
  FP_G_ATOMIC      = 128, /* split-K: atomicAdd into C */
  FP_G_TF32X3      = 256, /* use the tcgen05 3xTF32 path when shapes allow (fp32-accurate) */
  FP_G_COLSUM_A    = 4096, /* weight-gradient form only: U[m] += sum_k A[k,m] (the bias gradient rides along) */
  FP_G_DROP_OUT    = 8192  /* inverted dropout on the RESULT (after bias): C = keep(m,n) ? v/(1-drop_p) : 0 with the
                              counter-based mask of fp_dropout_mask(drop_seed, width N).  nflows ResidualBlock applies
                              dropout to relu(t); storing t' = dropout(t) instead gives relu(t') = dropout(relu(t)) to the
                              next GEMM's relu prologue and (t' > 0) = relu mask AND keep mask to the backward pass, so no
                              kernel ever has to re-draw the mask (floppity.py:57 -dropout, modules.py:63). */
};
typedef struct fp_gemm_args {
  const float* A; int64_t lda; int32_t a_kmajor;
  const float* B; int64_t ldb; int32_t b_kmajor;
  float* C; int64_t ldc;
  int64_t M; int32_t N; int64_t K;
  int32_t flags;
  const float* bias;
  const float* G; int64_t ldg; int32_t ctx_div;
  const float* Rsd; int64_t ldr;
  float* U; int64_t ldu;
  const float* Msk; int64_t ldm;
  int32_t split_k;
  float mask_scale;      /* FP_G_RELUMASK: C = Msk > 0 ? v * mask_scale : 0 ; 0 is read as 1 */
  float drop_p;          /* FP_G_DROP_OUT */
  uint64_t drop_seed;    /* FP_G_DROP_OUT */
  const float* a_scale;  /* fp_gemm_tc16 / fp_gemm_tc16_tn: optional DEVICE scalar s (a power of two): the A operand is
                          * converted as s * A and the result scaled back by 1/s -- brings gradient operands (~1/batch)
                          * into fp16's range; NULL = 1 */
  int32_t* status;       /* fp_gemm_tc16*: optional status words; [2] counts converting threads that met |s * A| > 6e4 */
  const void* pre_hi16;  /* fp_gemm_tc16: optional fp16 pieces (fp_split_f16) of A [M, K]; fp_gemm_tc16_tn: of B [K, N] -- an */
  const void* pre_lo16;  /* operand many tiles share is converted once by the caller instead of once per tile; leading      */
  int64_t pre_ld;        /* dimension pre_ld in halves (a multiple of 8); no relu prologue / operand scale on that operand  */
} fp_gemm_args;
int fp_gemm(const fp_gemm_args* args, void* stream);
/* Same contract on the tcgen05 tensor cores with the fp32-accurate 3xTF32 split (A split in-kernel, B_hi/B_lo
 * pre-split with fp_split_tf32).  Supports K-major A and B with 16-byte-aligned rows; returns FP_E_NOTREADY for
 * anything else (the caller then uses fp_gemm).  Set FLOPPITY_TC=0 to disable the tensor-core path. */
int fp_gemm_tc(const fp_gemm_args* args, const float* B_hi, const float* B_lo, void* stream);
int fp_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream);
/* The same GEMM with 16-bit operand pieces ("3xFP16": x = hi + lo * 2^-11, hi = fp16(x), lo = fp16((x - hi) * 2^11); the
 * same 22 mantissa bits as 3xTF32 at half the tensor-core time and half the shared-memory bytes per k-block, so twice the
 * k-blocks are in flight).  B_hi16 / B_lo16 [N, K] fp16 with leading dimension args->ldb (HALVES, a multiple of 8) from
 * fp_split_f16; A is converted in the kernel.  fp16's range is guarded: |a_scale * A| > 6e4 is counted in status[2]
 * (values clamp).  Returns FP_E_NOTREADY when the shape is not supported. */
int fp_gemm_tc16(const fp_gemm_args* args, const void* B_hi16, const void* B_lo16, void* stream);
int fp_split_f16(const float* x, void* hi16, void* lo16, int64_t n, void* stream);
/* fp_gemm_tc_tn with 3xFP16 operand pieces (both operands converted in the kernel; a_scale applies to A = the gradient) */
int fp_gemm_tc16_tn(const fp_gemm_args* args, void* stream);
/* weight-gradient form: C[M,N] += A[K,M]^T * B[K,N] (a_kmajor = b_kmajor = 0, flags = FP_G_ATOMIC [| FP_G_RELU_B]
 * [| FP_G_COLSUM_A with the bias-gradient vector [M] in `U`]); contraction split across CTAs, partial tiles added
 * with vector atomics.  FP_E_NOTREADY -> use fp_gemm (which ignores FP_G_COLSUM_A: the caller sums columns itself). */
int fp_gemm_tc_tn(const fp_gemm_args* args, void* stream);
/* keep[r, c] (uint8, row-major [R, width]) = 1 where the dropout mask keeps element (r, c): the same counter-based draw
 * FP_G_DROP_OUT and the fused conditioner use (one 32-bit hash of (seed, element pair) gives two 16-bit uniforms; an
 * element is dropped when its uniform is below round(p * 65536)).  fp_nsf_logprob_forward(training=1, dropout_seed=s)
 * draws block `block` of coupling layer `layer` with seed fp_dropout_layer_seed(s, layer, block). */
int fp_dropout_mask(uint8_t* keep, int64_t R, int32_t width, float p, uint64_t seed, void* stream);
uint64_t fp_dropout_layer_seed(uint64_t seed, int32_t layer, int32_t block);
/* Run-time switches of the conditioner paths (the FLOPPITY_TC / FLOPPITY_FUSED / FLOPPITY_FUSED_FP16 / FLOPPITY_SMALLH
 * environment variables give the initial values): tc = tcgen05 GEMMs (else exact fp32 SIMT), fused = one persistent
 * kernel per coupling layer for hidden 128, fused_fp16 = 3xFP16 operands in it (else 3xTF32), smallh = shared-memory-
 * resident fused MLP for hidden <= 64.  A negative argument leaves that switch as it is.  Returns the switches as bits
 * (1 tc, 2 fused, 4 fused_fp16, 8 smallh). */
int fp_set_mode(int32_t tc, int32_t fused, int32_t fused_fp16, int32_t smallh);
/* REDUCED-PRECISION variant of the stand-alone tensor-core GEMMs (the wide-conditioner path, c3; FLOPPITY_SINGLE_PASS=1):
 * one TF32 product per GEMM instead of the fp32-accurate three (operands rounded to 10 mantissa bits, fp32 accumulation).
 * NOT within the 1e-4 parity bar: log_prob agrees with the fp32 path to ~1e-3 relative (tests state the tolerance) and is
 * reported apart in bench.py.  Negative argument = query.  Returns the current setting. */
int fp_set_precision(int32_t single_pass);
/* Operand format of the stand-alone tensor-core GEMMs (the wide-conditioner path and every unfused backward GEMM;
 * FLOPPITY_GEMM_FP16): 1 = 3xFP16 pieces (default; fp32-accurate like 3xTF32 -- 11 + 11 mantissa bits -- with fp16's range
 * guarded by status[2]), 0 = 3xTF32.  Negative argument = query.  Returns the current setting. */
int fp_set_gemm_format(int32_t fp16);
/* debug: SM-clock stamps of CTA 0 of the next fp_gemm_tc launches into a device int64[512] buffer (NULL = off) */
int fp_tc_set_trace(void* buf);
/* debug: ablation switches for timing experiments on the tensor-core GEMM kernels (1 skip the operand conversion, 2 skip the
 * epilogue's stores / reductions, 4 two of three products, 8 (NT) do not load B_lo: results are wrong while any of these
 * is set), 32 = run long-contraction NT GEMMs as CTA pairs (cta_group::2; correct, measured slower, off by default) and
 * 128 = the tile order that ignores the L2 (the round-1 order; results stay correct) */
int fp_tc_set_debug(int32_t bits);
/* debug: SM-clock stamps of CTA 0 of the fused conditioner kernel (cond_fused.cu) into a device int64[512] (NULL = off) */
int fp_cf_set_trace(void* buf);

/* ---- whole-flow entry points --------------------------------------------------------------- */
/* Derive per-step constants from the parameters (W = L*U, L, U, sum log diag U). */
int fp_nsf_prepare(const fp_nsf_config* cfg, const float* params, float* prepared, void* stream);

/* Flow.log_prob (nflows Flow._log_prob; floppityFUN.py:35, and the training loss floppity.py:219).
 * theta [R, D] raw; ctx [Rc, C] ALREADY standardised; row r uses context row r / (R/Rc).
 * zshift/zscale [D] = the theta z-score affine; zlogdet = sum log|zscale|.
 * logp [R] out.  workspace from fp_nsf_workspace_count(cfg, R, Rc, training). */
int fp_nsf_logprob_forward(const fp_nsf_config* cfg, const float* params, const float* prepared,
                           const float* theta, const float* ctx, int64_t R, int64_t Rc,
                           const float* zshift, const float* zscale, float zlogdet,
                           float* logp, float* workspace, int32_t training, uint64_t dropout_seed,
                           int32_t* status, void* stream);
/* Backward of the above given d_logp [R]; ACCUMULATES into grads (same layout as params);
 * d_theta (may be NULL) [R, D] receives dL/dtheta (raw theta). Must follow a training=1 forward
 * with the same workspace.  status (may be NULL): status[2] counts gradient operands beyond fp16's range in the 3xFP16
 * GEMMs (they are scaled by one power of two derived from max|d_logp|, so this needs gradients 6e4 times the upstream one). */
int fp_nsf_logprob_backward(const fp_nsf_config* cfg, const float* params, const float* prepared,
                            const float* ctx, int64_t R, int64_t Rc, const float* zscale,
                            const float* d_logp, float* grads, float* d_theta, float* workspace,
                            uint64_t dropout_seed, int32_t* status, void* stream);
/* d_ctx [Rc, C] = gradient wrt the context rows (for an embedding net trained jointly in front of the flow: sbi
 * FCEmbedding built at /root/reference/modules.py:21-45).  Call after fp_nsf_logprob_backward with the same workspace. */
int fp_nsf_ctx_backward(const fp_nsf_config* cfg, const float* params, int64_t R, int64_t Rc, const float* workspace,
                        float* d_ctx, void* stream);
/* Fold d(W=L*U) etc. accumulated by the backward into the LU parameter gradients and zero the
 * structural zeros of W0X. Call once after all backward passes of a step. */
int fp_nsf_finish_grads(const fp_nsf_config* cfg, const float* params, const float* prepared,
                        float* grads, void* stream);

/* Flow.sample inverse pass (nflows Flow._sample; floppity.py:252): z [R, D] base noise -> theta [R, D].
 * ctx [Rc, C] standardised; Rc==1 is the hoisted single-observation case. */
int fp_nsf_sample(const fp_nsf_config* cfg, const float* params, const float* prepared,
                  const float* z, const float* ctx, int64_t R, int64_t Rc,
                  const float* zshift, const float* zscale, float* theta_out, float* workspace,
                  int32_t* status, void* stream);

/* ---- SNPE-C atomic loss (sbi SNPE_C._log_prob_proposal_posterior_atomic; floppity.py:219-221) ---- */
/* choices int32 [B, A-1]: A-1 distinct indices != b, uniform (Floyd sampling, counter RNG). */
int fp_atom_choices(int32_t* choices, int64_t B, int32_t A, uint64_t seed, void* stream);
/* same, with the draw position taken from a device counter (seed + step_counter[0]): the launch can sit in a CUDA
 * graph and still draw fresh atoms on every replay */
int fp_atom_choices_step(int32_t* choices, int64_t B, int32_t A, uint64_t seed, const int32_t* step_counter, void* stream);
/* rows [B*A, D] = cat(theta[b], theta[choices[b,:]]). */
int fp_atom_gather(const float* theta, const int32_t* choices, float* rows, int64_t B, int32_t A, int32_t D, void* stream);
/* logp [B, A]; atom_theta [B*A, D]; low/high [D]; masks [B] (may be NULL => 0);
 * out_lp [B] = u0 - logsumexp(u) (+ masks*logp[:,0] if combined); d_logp [B, A] = d(-mean out_lp)/d logp (may be NULL). */
int fp_atomic_loss(const float* logp, const float* atom_theta, const float* low, const float* high,
                   const float* masks, int32_t combined, float* out_lp, float* d_logp,
                   int64_t B, int32_t A, int32_t D, int32_t* status, void* stream);

/* ---- optimiser: clip_grad_norm_(5.0) + Adam (sbi PosteriorEstimator.train; floppity.py:219) ------ */
/* sumsq[0] must be zeroed by the caller before fp_grad_sumsq; fp_adam_step reads it (after an
 * optional all-reduce) and applies clip coefficient min(1, max_norm/(sqrt(sumsq)+1e-6)) * grad_scale. */
int fp_grad_sumsq(const float* grads, int64_t n, float* sumsq, void* stream);
int fp_adam_step(float* params, const float* grads, float* m, float* v, int64_t n, float lr, float beta1,
                 float beta2, float eps, const int32_t* step_dev, float max_norm, float grad_scale,
                 const float* sumsq, void* stream);

/* ---- small utilities on the path -------------------------------------------------------------- */
int fp_gather_rows(const float* src, const int64_t* idx, float* dst, int64_t n, int32_t width, void* stream);
/* out = (x - mean)/std  (sbi Standardize in front of the flow) */
int fp_standardize(const float* x, const float* mean, const float* std, float* out, int64_t n, int32_t width, void* stream);
/* out[i] = scale*sum(x) accumulated into out[0] */
int fp_sum(const float* x, int64_t n, float scale, float* out, void* stream);
/* BoxUniform support test: inside[r] = all(low <= theta[r] <= high) (uint8) */
/* ---- data-side neighbours of the flow ---------------------------------------------
 * Noise augmentation + scaling of simulated spectra into fp32 context rows: /root/reference/floppityFUN.py:383-428
 * (`preprocess`: np.repeat(spec, naug) + noise*randn, then rm_mean / sigma_res_scale, then the fitted StandardScaler
 * or the PCA centring as (v - shift) / scale).  spec [n_spec, C] fp64, z [n_spec*naug, C] fp32 standard-normal draw,
 * noise / obs [C] fp64, shift / scale [C (+1 when mode 1)] fp64 or NULL, out [n_spec*naug, ldo] fp32.
 * mode 0 plain, 1 remove the row mean and append it as column C (floppityFUN.py:229-244), 2 residuals
 * (v - obs) / noise (floppityFUN.py:196-204). */
int fp_preprocess_rows(const double* spec, const float* z, const double* noise, const double* obs, const double* shift,
                       const double* scale, float* out, int64_t n_spec, int32_t C, int32_t naug, int32_t mode,
                       int64_t ldo, void* stream);
/* Row-wise Gaussian log-likelihood of model spectra x [rows, C] against the observation (floppityFUN.py:24-28, the
 * inner loop of IS / evidence, floppityFUN.py:30-91): out[i] = sum_j -log(sqrt(2 pi) err_j) - (obs_j - x_ij)^2 / (2 err_j^2) */
int fp_gauss_loglik(const double* obs, const double* err, const double* x, double* out, int64_t rows, int32_t C, void* stream);

/* out = max(x, 0) ; out = d * (z > 0): the output relu of the embedding net (its inner relus are GEMM prologues / masks) */
int fp_relu(const float* x, float* out, int64_t n, void* stream);
int fp_relu_bwd(const float* d, const float* z, float* out, int64_t n, void* stream);

/* ---- convolutional embedding net in front of the flow: sbi CNNEmbedding's cnn_subnet, built by
 * createEmbedding at /root/reference/modules.py:26-33 for -embedding_type CNN (the default type, floppity.py:46).
 * One layer = Conv1d(stride 1, zero padding `pad`) -> ReLU -> MaxPool1d(pool):
 *   x [B, Cin, L], w [Cout, Cin, k], b [Cout] -> z [B, Cout, Lc] pre-activation (Lc = L + 2 pad - k + 1, kept for the
 *   backward), y [B, Cout, Lc / pool].
 * Backward: dy [B, Cout, Lc / pool] -> dz [B, Cout, Lc] (scratch, overwritten), dx [B, Cin, L] (overwritten; NULL for the
 * first layer), dw / db ACCUMULATED.  Limits: Cout*Cin*k <= 4096, Cin*k <= 64 (FP_E_TOOLARGE beyond). */
int fp_conv1d_relu_pool_fwd(const float* x, const float* w, const float* b, float* z, float* y, int64_t B, int32_t Cin, int32_t L,
                            int32_t Cout, int32_t k, int32_t pad, int32_t pool, void* stream);
int fp_conv1d_relu_pool_bwd(const float* x, const float* w, const float* z, const float* dy, float* dz, float* dx, float* dw,
                            float* db, int64_t B, int32_t Cin, int32_t L, int32_t Cout, int32_t k, int32_t pad, int32_t pool,
                            void* stream);

int fp_within_box(const float* theta, const float* low, const float* high, uint8_t* inside, int64_t R, int32_t D, void* stream);

/* ---- launch accounting (bench.py: gpu_launches and the live per-kernel roofline numbers) -----------------------
 * Kernels are tagged individually (fp_prof_ntags() tags, fp_prof_tag_name(i) = the __global__ function's name):
 * k_gemm, k_gemm_tc_nt, k_gemm_tc_tn, k_cond_fwd, k_rqs_apply_fwd, k_rqs_apply_inv, k_rqs_backward, k_cond_small_fwd,
 * k_cond_small_bwd, k_embed_conv, k_small_wgrad, other.  fp_prof_collect synchronises the device and fills, per tag, the elapsed
 * milliseconds (CUDA events recorded around every launch on the launching stream), the launch count and the
 * ALGORITHMIC flops and HBM bytes those launches were asked to do, then clears the records. */
/* Rejection step of accept_reject_sample: append the rows of theta [R, D] that lie inside [low, high] to out (row
 * capacity `cap`) behind the device counter *counter (rows accepted so far, also counts rows beyond cap). */
int fp_compact_inside(const float* theta, const float* low, const float* high, float* out, uint64_t* counter, int64_t R,
                      int32_t D, int64_t cap, void* stream);
int64_t fp_launch_count(void);
int fp_prof_enable(int32_t on);
int32_t fp_prof_ntags(void);
const char* fp_prof_tag_name(int32_t tag);
int fp_prof_collect(double* ms, int64_t* launches, double* flops, double* bytes);

const char* fp_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FLOPPITY_B200_H */
