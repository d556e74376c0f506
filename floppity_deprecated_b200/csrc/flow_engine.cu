// flow_engine.cu -- whole-flow entry points: Flow.log_prob forward/backward and Flow.sample.
//
// Stands for nflows Flow._log_prob / Flow._sample over CompositeTransform[ affine z-score,
// T x (PiecewiseRationalQuadraticCouplingTransform(ResidualNet), LULinear) ] + StandardNormal, as built by
// sbi build_nsf (called from /root/reference/modules.py:63) and driven by the training loop at
// floppity.py:219, posterior.log_prob at floppityFUN.py:35 and posterior.sample at floppity.py:252.
//
// Upstream executes ~100 eager kernels per coupling layer.  Here a layer is
// 2*Bk+2 GEMM launches with fused epilogues + ONE spline/LU kernel, the context projection of ALL
// layers and blocks is ONE GEMM, there is no host synchronisation anywhere, and the backward pass is
// written out by hand (no autograd graph), so the whole training step is a fixed launch sequence that
// can be captured in a CUDA graph.
#include "common.cuh"
#include "layout.h"
#include "prof.h"
#include <math.h>
#include <stdlib.h>

namespace fp {

// from the other translation units
int launch_gemm(const fp_gemm_args* a, cudaStream_t stream);
int launch_gemm_tc(const fp_gemm_args* a, const float* B_hi, const float* B_lo, cudaStream_t stream);   // gemm_tc.cu
int launch_gemm_tc_tn(const fp_gemm_args* a, cudaStream_t stream);
int launch_gemm_tc16(const fp_gemm_args* a, const void* B_hi16, const void* B_lo16, cudaStream_t stream);
int launch_gemm_tc16_tn(const fp_gemm_args* a, cudaStream_t stream);
int launch_prepare_gemm16(const fp_nsf_config* cfg, const float* params, float* prepared, cudaStream_t stream);
int launch_grad_scale(const float* d_logp, int64_t n, float* out, cudaStream_t s);
int launch_split_f16(const float* x, void* hi16, void* lo16, int64_t n, cudaStream_t stream, int32_t* status);
int launch_pad_rows(const float* in, int D, float* out, int Dp, int64_t R, cudaStream_t s);
int launch_split_tf32(const float* x, float* hi, float* lo, int64_t n, cudaStream_t stream);
int launch_transpose_split_weights(const fp_nsf_config* cfg, const float* params, float* t_hi, float* t_lo, cudaStream_t stream);
bool tc_enabled();
int launch_cond_fwd(const fp_nsf_config* cfg, const Layout& L, const float* params, const float* prepared, int i,
                    const float* theta_i, const float* ctx_all, int64_t R, int ctx_div, float* const* h, float* const* t,
                    float* const* u, float* prm, bool training, float drop_p, uint64_t drop_seed, int32_t* status,
                    cudaStream_t stream);       // cond_fused.cu
int launch_rqs_apply(bool inverse, const fp_nsf_config* cfg, int parity, const float* theta_in, const float* params,
                     float* theta_out, float* theta_sp, float* logdet, int32_t* bin_idx, int64_t R, int32_t* status,
                     const float* luM, const float* luPre, float pre_sign, const float* luPost,
                     const float* lu_logdet, int ldp, cudaStream_t stream);
int launch_rqs_backward(const fp_nsf_config* cfg, int parity, const float* theta_in, const float* params,
                        const float* d_theta_out, const float* d_logdet, float* d_params, float* d_theta_in,
                        int64_t R, int ldp, cudaStream_t stream, const float* luM, const float* theta_sp, float* dM,
                        float* dbias);
int launch_affine_fwd(const float*, const float*, const float*, float, float*, float*, int64_t, int, cudaStream_t);
int launch_affine_inv(const float*, const float*, const float*, float*, int64_t, int, cudaStream_t);
int launch_affine_bwd(const float*, const float*, float*, int64_t, int, cudaStream_t);
int launch_base_logprob(const float*, const float*, float*, int64_t, int, int32_t*, cudaStream_t);
int launch_base_bwd(const float*, const float*, float*, int64_t, int, cudaStream_t);
int launch_rows_affine(const float*, const float*, const float*, float, const float*, float*, int64_t, int, cudaStream_t);
int launch_glu_bwd(const float*, const float*, const float*, int64_t, float*, float*, int64_t, int, int, int, int, cudaStream_t);
int launch_gate_sigmoid(float*, int64_t, int64_t, int, int, cudaStream_t);
bool cond_fwd_supported(const Layout& L, const float* prepared);      // cond_fused.cu
bool cond_small_supported(const Layout& L, const float* prepared);    // cond_small.cu
int launch_small_prepare(const fp_nsf_config* cfg, const float* params, float* prepared, cudaStream_t stream);
int launch_cond_small_fwd(const fp_nsf_config* cfg, const Layout& L, const float* params, const float* prepared, int i,
                          const float* theta_i, const float* ctx_all, int64_t R, int ctx_div, float* const* h, float* const* t,
                          float* const* u, float* prm, bool training, float drop_p, uint64_t drop_seed, cudaStream_t stream);
int launch_cond_small_bwd(const fp_nsf_config* cfg, const Layout& L, const float* prepared, int i, const float* dprm,
                          const float* const* h, const float* const* t, const float* const* u, const float* ctx_all,
                          int ctx_div, int64_t R, float* const* du, float* const* dt, float* dh0, float* dctx_all, float* dth,
                          float mask_scale, cudaStream_t stream);
int launch_cond_prepare(const fp_nsf_config* cfg, const float* params, float* prepared, cudaStream_t stream);
int launch_ctx_reduce(const float*, float*, int64_t, int64_t, int, int, int, cudaStream_t);
int launch_colsum(const float*, int64_t, int64_t, int, float*, cudaStream_t);
int launch_sum(const float*, int64_t, float, float*, cudaStream_t);
__global__ void k_lu_prepare(const float*, float*, fp_nsf_config);
__global__ void k_lu_finish(const float*, const float*, float*, fp_nsf_config);

// ---- run-time switches of the conditioner paths (fp_set_mode; the environment gives the initial values) ----
static int g_mode[6] = {-1, -1, -1, -1, -1, -1};     // tc, fused, fused_fp16, smallh, single_pass, gemm_fp16
static const char* const g_mode_env[6] = {"FLOPPITY_TC", "FLOPPITY_FUSED", "FLOPPITY_FUSED_FP16", "FLOPPITY_SMALLH",
                                          "FLOPPITY_SINGLE_PASS", "FLOPPITY_GEMM_FP16"};
static bool mode_on(int i) {
  if (g_mode[i] < 0) {
    const char* e = getenv(g_mode_env[i]);
    g_mode[i] = (i == 4) ? ((e && e[0] == '1') ? 1 : 0) : ((e && e[0] == '0') ? 0 : 1);      // single_pass defaults to OFF
  }
  return g_mode[i] == 1;
}
bool tc_single_pass() { return mode_on(4); }
bool tc_enabled() { return mode_on(0); }
bool fused_enabled() { return mode_on(1); }
bool cond_fused_f16() { return mode_on(2); }
bool smallh_enabled() { return mode_on(3); }
bool gemm_fp16() { return mode_on(5); }
static bool pad_theta() {      // FLOPPITY_PAD_THETA=0: first-layer GEMMs of D % 8 != 0 flows stay on the SIMT kernel (A/B timing)
  static int v = -1;
  if (v < 0) { const char* e = getenv("FLOPPITY_PAD_THETA"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}     // 3xFP16 operand pieces in the stand-alone tensor-core GEMMs (else 3xTF32)

// per-call context of the 3xFP16 GEMMs: the device scalar that scales gradient operands into fp16's range (backward pass
// only) and the status words that count operands beyond it
struct GemmCtx { const float* gscale; int32_t* status; };
static thread_local GemmCtx g_gctx = {nullptr, nullptr};

#define FP_TRY(expr)            \
  do {                          \
    int _r = (expr);            \
    if (_r != 0) return _r;     \
  } while (0)

static int check_cfg(const fp_nsf_config* c) {
  if (!c) return FP_E_BADARG;
  if (c->D < 2 || c->D > 64 || c->C < 1 || c->T < 1 || c->T > 64 || c->K < 2 || c->K > 64 || c->H < 1 || c->Bk < 0 ||
      c->Bk > 16)
    return FP_E_BADARG;
  if (c->dropout_p < 0.f || c->dropout_p >= 1.f) return FP_E_BADARG;
  return 0;
}

static fp_gemm_args gemm_nt(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t M,
                            int N, int64_t K, int flags) {
  fp_gemm_args g{};
  g.A = A; g.lda = lda; g.a_kmajor = 1; g.B = B; g.ldb = ldb; g.b_kmajor = 1;
  g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K; g.flags = flags; g.split_k = 1;
  return g;
}

static int pick_split(int64_t M, int N, int64_t K) {
  // wgrad: output is tiny (M x N), contraction K is the batch; spread over ~2 waves of CTAs
  int64_t tiles = ((M + 127) / 128) * ((N + 63) / 64);
  int64_t want = (2 * (int64_t)num_sms() + tiles - 1) / tiles;
  int64_t maxs = (K + 255) / 256;  // at least 256 contraction rows per split
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  if (want > 4096) want = 4096;
  return (int)want;
}

// dW[N_out, K_in] += dOut[R, N_out]^T @ Act[R, K_in]   (both operands row-major over the batch)
// db[N_out] += colsum(dOut) rides along when `db` is given (fused into the tensor-core kernel, else its own launch)
static int wgrad(const float* dOut, int64_t ldo, int n_out, const float* act, int64_t lda, int k_in, int64_t R,
                 float* dW, int64_t ldw, int flags_extra, cudaStream_t s, float* db = nullptr, const void* act_hi16 = nullptr,
                 const void* act_lo16 = nullptr) {
  fp_gemm_args g{};
  g.pre_hi16 = act_hi16; g.pre_lo16 = act_lo16; g.pre_ld = lda;      // fp16 pieces of `act`, when the caller has them
  g.A = dOut; g.lda = ldo; g.a_kmajor = 0;   // logical A'[n_out, R], element (n, r) at dOut[r*ldo + n]
  g.B = act; g.ldb = lda; g.b_kmajor = 0;    // logical B'[k_in, R], element (k, r) at act[r*lda + k]
  g.C = dW; g.ldc = ldw; g.M = n_out; g.N = k_in; g.K = R;
  g.flags = FP_G_ATOMIC | flags_extra | (db ? FP_G_COLSUM_A : 0);
  g.U = db;
  if (gemm_fp16()) {
    g.a_scale = g_gctx.gscale; g.status = g_gctx.status;
    int r = launch_gemm_tc16_tn(&g, s);
    if (r != FP_E_NOTREADY) return r;
  }
  {
    int r = launch_gemm_tc_tn(&g, s);
    if (r != FP_E_NOTREADY) return r;
  }
  if (db) FP_TRY(launch_colsum(dOut, ldo, R, n_out, db, s));
  g.flags &= ~FP_G_COLSUM_A; g.U = nullptr;
  g.split_k = pick_split(n_out, k_in, R);
  return launch_gemm(&g, s);
}

// dIn[R, K_in] (=|+=) dOut[R, N_out] @ W[N_out, K_in]
static fp_gemm_args dgrad_args(const float* dOut, int64_t ldo, int n_out, const float* W, int64_t ldw, int k_in,
                               int64_t R, float* dIn, int64_t ldi, int flags) {
  fp_gemm_args g{};
  g.A = dOut; g.lda = ldo; g.a_kmajor = 1;
  g.B = W; g.ldb = ldw; g.b_kmajor = 0;      // logical B[k_in, n_out], element (k, n) at W[n*ldw + k]
  g.C = dIn; g.ldc = ldi; g.M = R; g.N = k_in; g.K = n_out; g.flags = flags; g.split_k = 1;
  return g;
}

// forward GEMM against a weight matrix living at `woff` in the parameter buffer: tcgen05 3xTF32 when the shape
// allows (pre-split weights in `prepared`), else the fp32 SIMT kernel
static int gemm_weight(const Layout& L, const float* prepared, int64_t woff, fp_gemm_args* g, cudaStream_t s) {
  if (prepared && gemm_fp16() && !L.smallH()) {      // fp16 copy of the same [N, K] matrix, same leading dimension (layout.h prepG16Hi)
    fp_gemm_args g16 = *g;
    g16.status = g_gctx.status;
    int r = launch_gemm_tc16(&g16, reinterpret_cast<const uint16_t*>(prepared + L.prepG16Hi()) + 2 * woff,
                             reinterpret_cast<const uint16_t*>(prepared + L.prepG16Lo()) + 2 * woff, s);
    if (r != FP_E_NOTREADY) return r;
  }
  if (prepared) {
    int r = launch_gemm_tc(g, prepared + L.prepHi() + woff, prepared + L.prepLo() + woff, s);
    if (r != FP_E_NOTREADY) return r;
  }
  return launch_gemm(g, s);
}

// dIn[R, k_in] (=|+=) dOut[R, n_out] @ W[n_out, k_in] with the epilogue flags in `flags`:
// tcgen05 against the pre-transposed + split copy W^T [k_in, ld_t] at `toff`, else SIMT against W itself
static int dgrad_weight(const Layout& L, const float* prepared, int64_t toff, int64_t ld_t, const float* dOut, int64_t ldo,
                        int n_out, const float* W, int64_t ldw, int k_in, int64_t R, float* dIn, int64_t ldi, int flags,
                        const float* Msk, int64_t ldm, float mask_scale, cudaStream_t s) {
  if (prepared) {
    fp_gemm_args g{};
    g.A = dOut; g.lda = ldo; g.a_kmajor = 1;
    g.B = prepared + L.prepTHi() + toff; g.ldb = ld_t; g.b_kmajor = 1;
    g.C = dIn; g.ldc = ldi; g.M = R; g.N = k_in; g.K = n_out; g.flags = flags; g.split_k = 1;
    g.Msk = Msk; g.ldm = ldm; g.mask_scale = mask_scale;
    if (gemm_fp16() && !L.smallH()) {
      fp_gemm_args g16 = g;
      g16.ldb = Layout::ld16((int)ld_t);
      g16.a_scale = g_gctx.gscale; g16.status = g_gctx.status;
      int r = launch_gemm_tc16(&g16, reinterpret_cast<const uint16_t*>(prepared + L.prepGT16Hi()) + 2 * toff,
                               reinterpret_cast<const uint16_t*>(prepared + L.prepGT16Lo()) + 2 * toff, s);
      if (r != FP_E_NOTREADY) return r;
    }
    int r = launch_gemm_tc(&g, prepared + L.prepTHi() + toff, prepared + L.prepTLo() + toff, s);
    if (r != FP_E_NOTREADY) return r;
  }
  fp_gemm_args g = dgrad_args(dOut, ldo, n_out, W, ldw, k_in, R, dIn, ldi, flags);
  g.Msk = Msk; g.ldm = ldm; g.mask_scale = mask_scale;
  return launch_gemm(&g, s);
}

uint64_t layer_seed(uint64_t seed, int layer, int block) {
  return seed * 0x9E3779B97F4A7C15ull + (uint64_t)(layer * 64 + block + 1) * 0xD6E8FEB86659FD93ull;
}

// all 2 Bk + 2 weight (and bias) gradients of layer i's hidden <= 64 conditioner in one launch (cond_small.cu)
int launch_small_wgrad_layer(const Layout& L, int i, const float* dprm, const float* const* hs, const float* const* ts,
                             float* const* dus, float* const* dts, const float* dh0, const float* theta_i, int64_t R,
                             float* grads, cudaStream_t s);
static int small_wgrads(const Layout& L, int i, const float* dprm, const float* const* hs, const float* const* ts,
                        float* const* dus, float* const* dts, const float* dh0, const float* theta_i, int64_t R, float* grads,
                        cudaStream_t s) {
  int r = launch_small_wgrad_layer(L, i, dprm, hs, ts, dus, dts, dh0, theta_i, R, grads, s);
  if (r != FP_E_NOTREADY) return r;
  const int H = L.H, lH = L.ldH(), D = L.D;      // shapes the batched kernel does not take: one GEMM per matrix
  FP_TRY(wgrad(dprm, L.ldP(i), L.dP(i), hs[L.Bk], lH, H, R, grads + L.wf(i), H, 0, s, grads + L.bf(i)));
  for (int k = L.Bk - 1; k >= 0; --k) {
    FP_TRY(wgrad(dus[k], lH, H, ts[k], lH, H, R, grads + L.w2(i, k), H, FP_G_RELU_B, s, grads + L.b2(i, k)));
    FP_TRY(wgrad(dts[k], lH, H, hs[k], lH, H, R, grads + L.w1(i, k), H, FP_G_RELU_B, s, grads + L.b1(i, k)));
  }
  return wgrad(dh0, lH, H, theta_i, D, D, R, grads + L.w0x(i), D, 0, s);
}

// One decision per flow pass: does every layer's conditioner run as ONE kernel -- the tcgen05 kernel for hidden 128
// (cond_fused.cu, returns 1) or the shared-memory-resident kernel for hidden <= 64 (cond_small.cu, returns 2)?  Both read
// sigmoid(gate) from the gate slots of the context projection (k_gate_sigmoid).
static int use_fused(const fp_nsf_config* cfg, const Layout& L, const float* prepared, bool training) {
  (void)cfg; (void)training;
  if (cond_fwd_supported(L, prepared)) return 1;
  if (cond_small_supported(L, prepared)) return 2;
  return 0;
}

// conditioner forward of layer i: theta_i [R, D] -> spline params [R, dP_i]
// h buffers: h(b) for b = 0..Bk (may alias when not training), t(b), u(b) (NULL when not training)
struct CondBufs {
  float* h[17];
  float* t[16];
  float* u[16];
  float* prm;
};

static int conditioner_forward(const fp_nsf_config* cfg, const Layout& L, const float* params, const float* prepared, int i,
                               const float* theta_i,
                               const float* ctx_all, int64_t R, int ctx_div, const CondBufs& b, bool training,
                               uint64_t seed, int fused, int32_t* status, cudaStream_t s, float* th_pad = nullptr) {
  const int H = L.H, D = L.D, lH = L.ldH();      // lH: row stride of the hidden-activation buffers
  const float drop = training ? cfg->dropout_p : 0.f;
  if (fused == 2)   // hidden <= 64: the whole chain in one persistent kernel with the layer's weights resident in shared memory
    return launch_cond_small_fwd(cfg, L, params, prepared, i, theta_i, ctx_all, R, ctx_div, b.h, b.t, b.u, b.prm, training, drop,
                                 seed, s);
  if (fused) {   // hidden 128: the whole chain in one persistent tcgen05 kernel (activations stay in TMEM); the gate
                 // slots of ctx_all already hold sigmoid(gate), so there is no falling back to the GEMM chain from here
    return launch_cond_fwd(cfg, L, params, prepared, i, theta_i, ctx_all, R, ctx_div, b.h, b.t, b.u, b.prm, training, drop, seed,
                           status, s);
  }
  {  // initial layer: h0 = theta_i @ W0x^T + (ctx @ W0ctx^T + b0)   [the bracket is a column block of ctx_all]
    fp_gemm_args g = gemm_nt(theta_i, D, params + L.w0x(i), D, b.h[0], lH, R, H, D, FP_G_ADD_CTX);
    g.G = ctx_all + L.ctxCol(i, 0); g.ldg = L.Nctx; g.ctx_div = ctx_div;
    int r0 = FP_E_NOTREADY;
    if (th_pad && prepared && tc_enabled() && gemm_fp16() && pad_theta()) {
      // D = 30: 120-byte theta rows are not TMA-addressable -- a copy padded to 32 floats (and the padded fp16 W0x) puts this
      // GEMM on the tensor cores; the copy stays in the workspace for the weight gradient
      const int Dp = Layout::ld16(D);
      FP_TRY(launch_pad_rows(theta_i, D, th_pad, Dp, R, s));
      fp_gemm_args g16 = g;
      g16.A = th_pad; g16.lda = Dp; g16.ldb = Dp; g16.status = g_gctx.status;
      r0 = launch_gemm_tc16(&g16, reinterpret_cast<const uint16_t*>(prepared + L.prepG16Hi()) + 2 * L.w0x(i),
                            reinterpret_cast<const uint16_t*>(prepared + L.prepG16Lo()) + 2 * L.w0x(i), s);
      if (r0 != 0 && r0 != FP_E_NOTREADY) return r0;
    }
    if (r0 != 0) FP_TRY(gemm_weight(L, prepared, L.w0x(i), &g, s));
  }
  for (int k = 0; k < L.Bk; ++k) {
    {  // t' = dropout(relu(h) @ W1^T + b1): relu(t') = dropout(relu(t)), (t' > 0) = relu mask AND keep mask
      fp_gemm_args g = gemm_nt(b.h[k], lH, params + L.w1(i, k), H, b.t[k], lH, R, H, H,
                               FP_G_RELU_A | FP_G_BIAS | (drop > 0.f ? FP_G_DROP_OUT : 0));
      g.bias = params + L.b1(i, k);
      g.drop_p = drop; g.drop_seed = layer_seed(seed, i, k);
      FP_TRY(gemm_weight(L, prepared, L.w1(i, k), &g, s));
    }
    {  // u = relu(t') @ W2^T + b2 ; h_next = h + u * sigmoid(ctx gate)
      int fl = FP_G_RELU_A | FP_G_BIAS | FP_G_GLU_RES;
      fp_gemm_args g = gemm_nt(b.t[k], lH, params + L.w2(i, k), H, b.h[k + 1], lH, R, H, H, fl);
      g.bias = params + L.b2(i, k);
      g.G = ctx_all + L.ctxCol(i, 1 + k); g.ldg = L.Nctx; g.ctx_div = ctx_div;
      g.Rsd = b.h[k]; g.ldr = lH;
      g.U = b.u[k]; g.ldu = lH;
      FP_TRY(gemm_weight(L, prepared, L.w2(i, k), &g, s));
    }
  }
  {  // params = h @ Wf^T + bf
    fp_gemm_args g = gemm_nt(b.h[L.Bk], lH, params + L.wf(i), H, b.prm, L.ldP(i), R, L.dP(i), H, FP_G_BIAS);
    g.bias = params + L.bf(i);
    FP_TRY(gemm_weight(L, prepared, L.wf(i), &g, s));
  }
  return 0;
}

// The context rows are the A operand of every column tile of the context projection (300 tiles at c3) and the B operand of
// every row tile of its weight gradient: with 3xFP16 GEMMs they are converted to fp16 pieces ONCE per pass (workspace ctx16:
// hi [Rc, C] then lo [Rc, C], halves) instead of once per tile.
static bool use_ctx16(const Layout& L, const float* prepared, int64_t Rc) {
  return prepared && tc_enabled() && gemm_fp16() && !L.smallH() && (L.C & 7) == 0 && Rc >= 256 && L.Nctx >= 512;
}
static int context_projection(const Layout& L, const float* params, const float* prepared, const float* ctx, int64_t Rc,
                              float* ctx_all, float* ctx16, cudaStream_t s) {
  fp_gemm_args g = gemm_nt(ctx, L.C, params + L.ctx_w, L.C, ctx_all, L.Nctx, Rc, (int)L.Nctx, L.C, FP_G_BIAS);
  g.bias = params + L.ctx_b;
  if (ctx16 && use_ctx16(L, prepared, Rc)) {
    uint16_t* hi = reinterpret_cast<uint16_t*>(ctx16);
    uint16_t* lo = hi + Rc * L.C;
    FP_TRY(launch_split_f16(ctx, hi, lo, Rc * L.C, s, g_gctx.status));
    g.pre_hi16 = hi; g.pre_lo16 = lo; g.pre_ld = L.C;
  }
  return gemm_weight(L, prepared, L.ctx_w, &g, s);
}

}  // namespace fp

using namespace fp;

extern "C" {

int fp_set_mode(int32_t tc, int32_t fused, int32_t fused_fp16, int32_t smallh) {
  const int32_t v[4] = {tc, fused, fused_fp16, smallh};
  int bits = 0;
  if (mode_on(4)) bits |= 16;
  for (int i = 0; i < 4; ++i) {
    if (v[i] >= 0) g_mode[i] = v[i] ? 1 : 0;
    if (mode_on(i)) bits |= 1 << i;
  }
  return bits;
}
int fp_set_precision(int32_t single_pass) {
  if (single_pass >= 0) g_mode[4] = single_pass ? 1 : 0;
  return mode_on(4) ? 1 : 0;
}
int fp_set_gemm_format(int32_t fp16) {
  if (fp16 >= 0) g_mode[5] = fp16 ? 1 : 0;
  return mode_on(5) ? 1 : 0;
}
uint64_t fp_dropout_layer_seed(uint64_t seed, int32_t layer, int32_t block) { return layer_seed(seed, layer, block); }

int64_t fp_nsf_param_count(const fp_nsf_config* cfg) {
  if (check_cfg(cfg)) return FP_E_BADARG;
  return Layout(cfg).total();
}
int64_t fp_nsf_grad_count(const fp_nsf_config* cfg) {
  if (check_cfg(cfg)) return FP_E_BADARG;
  return Layout(cfg).gradTotal();
}
int32_t fp_nsf_dtr(const fp_nsf_config* cfg, int32_t layer) {
  if (check_cfg(cfg)) return FP_E_BADARG;
  return Layout(cfg).dtr(layer);
}
int64_t fp_nsf_param_offset(const fp_nsf_config* cfg, int32_t layer, int32_t kind, int32_t block) {
  if (check_cfg(cfg) || layer < 0 || layer >= cfg->T) return FP_E_BADARG;
  Layout L(cfg);
  switch (kind) {
    case FP_P_CTX_W: return (block < 0 || block > L.Bk) ? FP_E_BADARG : L.ctxW(layer, block);
    case FP_P_CTX_B: return (block < 0 || block > L.Bk) ? FP_E_BADARG : L.ctxB(layer, block);
    case FP_P_W0X: return L.w0x(layer);
    case FP_P_W1: return (block < 0 || block >= L.Bk) ? FP_E_BADARG : L.w1(layer, block);
    case FP_P_B1: return (block < 0 || block >= L.Bk) ? FP_E_BADARG : L.b1(layer, block);
    case FP_P_W2: return (block < 0 || block >= L.Bk) ? FP_E_BADARG : L.w2(layer, block);
    case FP_P_B2: return (block < 0 || block >= L.Bk) ? FP_E_BADARG : L.b2(layer, block);
    case FP_P_WF: return L.wf(layer);
    case FP_P_BF: return L.bf(layer);
    case FP_P_LU_LOWER: return L.luLower(layer);
    case FP_P_LU_UPPER: return L.luUpper(layer);
    case FP_P_LU_DIAG: return L.luDiag(layer);
    case FP_P_LU_BIAS: return L.luBias(layer);
  }
  return FP_E_BADARG;
}
int64_t fp_nsf_param_size(const fp_nsf_config* cfg, int32_t layer, int32_t kind) {
  if (check_cfg(cfg) || layer < 0 || layer >= cfg->T) return FP_E_BADARG;
  Layout L(cfg);
  switch (kind) {
    case FP_P_CTX_W: return (int64_t)L.H * L.C;
    case FP_P_CTX_B: return L.H;
    case FP_P_W0X: return (int64_t)L.H * L.D;
    case FP_P_W1: case FP_P_W2: return (int64_t)L.H * L.H;
    case FP_P_B1: case FP_P_B2: return L.H;
    case FP_P_WF: return (int64_t)L.dP(layer) * L.H;
    case FP_P_BF: return L.dP(layer);
    case FP_P_LU_LOWER: case FP_P_LU_UPPER: return L.ntri;
    case FP_P_LU_DIAG: case FP_P_LU_BIAS: return L.D;
  }
  return FP_E_BADARG;
}
int64_t fp_nsf_prepared_count(const fp_nsf_config* cfg) {
  if (check_cfg(cfg)) return FP_E_BADARG;
  return Layout(cfg).prepAll();
}
int64_t fp_nsf_workspace_count(const fp_nsf_config* cfg, int64_t R, int64_t Rc, int32_t training) {
  if (check_cfg(cfg) || R < 0 || Rc < 0) return FP_E_BADARG;
  Layout L(cfg);
  return Workspace(L, R, Rc, training != 0).total;
}

int fp_nsf_prepare(const fp_nsf_config* cfg, const float* params, float* prepared, void* stream) {
  FP_TRY(check_cfg(cfg));
  size_t sm = (size_t)3 * cfg->D * cfg->D * 4;
  ProfScope ps(TAG_OTHER, 0.0, (cudaStream_t)stream);
  k_lu_prepare<<<cfg->T, 128, sm, (cudaStream_t)stream>>>(params, prepared, *cfg);
  FP_LAUNCH_CHECK();
  if (tc_enabled()) {   // operand copies of every weight, once per optimiser step
    Layout L(cfg);
    if (gemm_fp16() && !L.smallH()) {
      // 3xFP16 GEMMs (launch_prepare_gemm16 below): the TF32 copies are only the fall-back of matrices whose contraction length
      // is not a multiple of 8 halves (c2's 500 spectral bins), so only those are made; the transposed fp16 copies are padded
      // to 8 and never fall back
      if (L.C & 7) FP_TRY(launch_split_tf32(params, prepared + L.prepHi(), prepared + L.prepLo(), L.layers, (cudaStream_t)stream));
      if ((L.H & 7) || L.D < 4 || ((L.D & 7) && !pad_theta()) || (fused_enabled() && !cond_fused_f16()))      // (the fused kernel's 3xTF32 mode reads them too)
        FP_TRY(launch_split_tf32(params + L.layers, prepared + L.prepHi() + L.layers, prepared + L.prepLo() + L.layers,
                                 L.total() - L.layers, (cudaStream_t)stream));
    } else {
      FP_TRY(launch_split_tf32(params, prepared + L.prepHi(), prepared + L.prepLo(), L.total(), (cudaStream_t)stream));
      FP_TRY(launch_transpose_split_weights(cfg, params, prepared + L.prepTHi(), prepared + L.prepTLo(), (cudaStream_t)stream));
    }
    FP_TRY(launch_cond_prepare(cfg, params, prepared, (cudaStream_t)stream));     // fp16 operand copies for the fused conditioner
    if (!L.smallH()) FP_TRY(launch_prepare_gemm16(cfg, params, prepared, (cudaStream_t)stream));   // and for the 3xFP16 GEMMs
  }
  FP_TRY(launch_small_prepare(cfg, params, prepared, (cudaStream_t)stream));      // shared-memory images for hidden <= 64
  return 0;
}

int fp_nsf_finish_grads(const fp_nsf_config* cfg, const float* params, const float* prepared, float* grads, void* stream) {
  FP_TRY(check_cfg(cfg));
  size_t sm = (size_t)3 * cfg->D * cfg->D * 4;
  ProfScope ps(TAG_OTHER, 0.0, (cudaStream_t)stream);
  k_lu_finish<<<cfg->T, 128, sm, (cudaStream_t)stream>>>(params, prepared, grads, *cfg);
  FP_LAUNCH_CHECK();
  return 0;
}

int fp_nsf_logprob_forward(const fp_nsf_config* cfg, const float* params, const float* prepared, const float* theta,
                           const float* ctx, int64_t R, int64_t Rc, const float* zshift, const float* zscale,
                           float zlogdet, float* logp, float* workspace, int32_t training, uint64_t dropout_seed,
                           int32_t* status, void* stream) {
  FP_TRY(check_cfg(cfg));
  if (R <= 0) return 0;
  if (Rc <= 0 || R % Rc != 0) return FP_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  Layout L(cfg);
  Workspace W(L, R, Rc, training != 0);
  float* ws = workspace;
  const int ctx_div = (int)(R / Rc);
  const int D = L.D;

  float* ctx_all = ws + W.ctx_all;
  float* logdet = ws + W.logdet;
  g_gctx = GemmCtx{nullptr, status};
  FP_TRY(context_projection(L, params, prepared, ctx, Rc, ctx_all, ws + W.ctx16, s));
  const int fused = use_fused(cfg, L, prepared, training != 0);
  if (fused) FP_TRY(launch_gate_sigmoid(ctx_all, Rc, L.Nctx, L.H, L.S, s));

  float* th_cur = training ? ws + W.thst(0) : ws + W.th_a;
  FP_TRY(launch_affine_fwd(theta, zshift, zscale, zlogdet, th_cur, logdet, R, D, s));

  for (int i = 0; i < L.T; ++i) {
    CondBufs b{};
    float* th_next; float* th_sp = nullptr;
    if (training) {
      for (int k = 0; k <= L.Bk; ++k) b.h[k] = ws + W.hbuf(i, k);
      for (int k = 0; k < L.Bk; ++k) { b.t[k] = ws + W.tbuf(i, k); b.u[k] = ws + W.ubuf(i, k); }
      b.prm = ws + W.prmbuf(i);
      th_next = ws + W.thst(i + 1);
      th_sp = ws + W.thsp(i);
    } else {
      for (int k = 0; k <= L.Bk; ++k) b.h[k] = ws + W.h;     // in place: each element is read then written by one thread
      for (int k = 0; k < L.Bk; ++k) { b.t[k] = ws + W.t; b.u[k] = nullptr; }
      b.prm = ws + W.prm;
      th_next = (th_cur == ws + W.th_a) ? ws + W.th_b : ws + W.th_a;
    }
    FP_TRY(conditioner_forward(cfg, L, params, prepared, i, th_cur, ctx_all, R, ctx_div, b, training != 0, dropout_seed, fused, status, s,
                               (W.thp_sz > 0 && !fused) ? ws + W.thpad(training ? i : 0) : nullptr));
    FP_TRY(launch_rqs_apply(false, cfg, i & 1, th_cur, b.prm, th_next, th_sp, logdet, nullptr, R, status,
                            prepared + L.prepM(i), nullptr, 1.f, params + L.luBias(i), prepared + L.prepLogdet(i), L.ldP(i), s));
    th_cur = th_next;
  }
  FP_TRY(launch_base_logprob(th_cur, logdet, logp, R, D, status, s));
  return 0;
}

int fp_nsf_logprob_backward(const fp_nsf_config* cfg, const float* params, const float* prepared, const float* ctx,
                            int64_t R, int64_t Rc, const float* zscale, const float* d_logp, float* grads,
                            float* d_theta, float* workspace, uint64_t dropout_seed, int32_t* status, void* stream) {
  FP_TRY(check_cfg(cfg));
  if (R <= 0) return 0;
  if (Rc <= 0 || R % Rc != 0) return FP_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  Layout L(cfg);
  Workspace W(L, R, Rc, true);
  float* ws = workspace;
  const int A = (int)(R / Rc);
  const int D = L.D, H = L.H, lH = L.ldH();
  (void)dropout_seed;      // the mask is folded into the saved activations by the forward pass; nothing is re-drawn here
  const float drop_scale = cfg->dropout_p > 0.f ? 1.f / (1.f - cfg->dropout_p) : 1.f;
  const int fused = use_fused(cfg, L, prepared, true);      // same decision as the forward pass that filled ws
  const int sig_given = fused ? 1 : 0;

  float* dth_a = ws + W.dth_a; float* dth_b = ws + W.dth_b;
  float* dh = ws + W.dh; float* dt = ws + W.dt; float* du = ws + W.du; float* dprm = ws + W.dprm;
  float* dctx_all = ws + W.dctx_all;
  const float* ctx_all = ws + W.ctx_all;

  // 3xFP16 GEMMs: every gradient tensor of this pass is linear in d_logp (~ 1 / batch): one power-of-two scale brings them
  // all into fp16's range (undone in the GEMM epilogues); operands beyond the range are counted in status[2]
  g_gctx = GemmCtx{nullptr, status};
  if (gemm_fp16() && tc_enabled() && fused != 2) {
    FP_TRY(launch_grad_scale(d_logp, R, ws + W.gscale, s));
    g_gctx.gscale = ws + W.gscale;
  }
  // the shared-memory-resident backward adds its context gradients with reductions: start from zero
  if (fused == 2) FP_CUDA_OK(cudaMemsetAsync(dctx_all, 0, sizeof(float) * (size_t)Rc * (size_t)L.Nctx, s));

  // d logp / d z_T  and  sum(d_logp) for the batch-constant LU log-dets
  FP_TRY(launch_base_bwd(ws + W.thst(L.T), d_logp, dth_a, R, D, s));
  FP_TRY(launch_sum(d_logp, R, 1.f, grads + L.gradTailSum(), s));

  for (int i = L.T - 1; i >= 0; --i) {
    // ---- LULinear backward (y = M x + b, x = theta_sp[i]: dx = M^T dy, dM += dy^T x, db += colsum dy) fused into the
    //      spline backward: d params, d theta_in (identity features pass through)
    const int dP = L.dP(i), ldP = L.ldP(i);
    FP_TRY(launch_rqs_backward(cfg, i & 1, ws + W.thst(i), ws + W.prmbuf(i), dth_a, d_logp, dprm, dth_b, R, ldP, s,
                               prepared + L.prepM(i), ws + W.thsp(i), grads + L.gradTailDM(i), grads + L.luBias(i)));
    { float* t = dth_a; dth_a = dth_b; dth_b = t; }
    if (fused == 2) {
      // ---- hidden <= 64: the dgrad chain of the whole conditioner in one kernel (dh in registers), then the weight
      //      gradients from the operands it leaves (du_k, dt_k, dh0)
      const float* hs[17]; const float* ts[16]; const float* us[16]; float* dus[16]; float* dts[16];
      for (int k = 0; k <= L.Bk; ++k) hs[k] = ws + W.hbuf(i, k);
      for (int k = 0; k < L.Bk; ++k) { ts[k] = ws + W.tbuf(i, k); us[k] = ws + W.ubuf(i, k); dus[k] = ws + W.dubuf(k); dts[k] = ws + W.dtbuf(k); }
      FP_TRY(launch_cond_small_bwd(cfg, L, prepared, i, dprm, hs, ts, us, ctx_all, A, R, dus, dts, dh, dctx_all, dth_a, drop_scale, s));
      FP_TRY(small_wgrads(L, i, dprm, hs, ts, dus, dts, dh, ws + W.thst(i), R, grads, s));
      continue;
    }
    // ---- final layer
    FP_TRY(wgrad(dprm, ldP, dP, ws + W.hbuf(i, L.Bk), lH, H, R, grads + L.wf(i), H, 0, s, grads + L.bf(i)));
    FP_TRY(dgrad_weight(L, prepared, L.tWf(i), ldP, dprm, ldP, dP, params + L.wf(i), H, H, R, dh, lH, 0, nullptr, 0, 1.f, s));
    // ---- residual blocks, last to first
    for (int k = L.Bk - 1; k >= 0; --k) {
      FP_TRY(launch_glu_bwd(dh, ws + W.ubuf(i, k), ctx_all + L.ctxCol(i, 1 + k), L.Nctx, du,
                            dctx_all + L.ctxCol(i, 1 + k), Rc, A, H, lH, sig_given, s));
      // the saved t is t' = dropout(t) (forward): relu(t') is the W2 operand, (t' > 0) * 1/(1-p) the mask of its gradient
      FP_TRY(wgrad(du, lH, H, ws + W.tbuf(i, k), lH, H, R, grads + L.w2(i, k), H, FP_G_RELU_B, s, grads + L.b2(i, k)));
      FP_TRY(dgrad_weight(L, prepared, L.tW2(i, k), H, du, lH, H, params + L.w2(i, k), H, H, R, dt, lH,
                          FP_G_RELUMASK, ws + W.tbuf(i, k), lH, drop_scale, s));
      FP_TRY(wgrad(dt, lH, H, ws + W.hbuf(i, k), lH, H, R, grads + L.w1(i, k), H, FP_G_RELU_B, s, grads + L.b1(i, k)));
      FP_TRY(dgrad_weight(L, prepared, L.tW1(i, k), H, dt, lH, H, params + L.w1(i, k), H, H, R, dh, lH,
                          FP_G_RELUMASK | FP_G_ACCUM, ws + W.hbuf(i, k), lH, 1.f, s));
    }
    // ---- initial layer: h0 = theta_i @ W0x^T + ctx slot 0
    FP_TRY(launch_ctx_reduce(dh, dctx_all + L.ctxCol(i, 0), L.Nctx, Rc, A, H, lH, s));
    if (W.thp_sz > 0 && !fused && prepared && tc_enabled() && gemm_fp16() && pad_theta())      // the padded theta copy the forward pass left (see there)
      FP_TRY(wgrad(dh, lH, H, ws + W.thpad(i), Layout::ld16(D), D, R, grads + L.w0x(i), D, 0, s));
    else
      FP_TRY(wgrad(dh, lH, H, ws + W.thst(i), D, D, R, grads + L.w0x(i), D, 0, s));
    FP_TRY(dgrad_weight(L, prepared, L.tW0(i), H, dh, lH, H, params + L.w0x(i), D, D, R, dth_a, D, FP_G_ACCUM, nullptr, 0, 1.f, s));
  }
  if (d_theta) FP_TRY(launch_affine_bwd(dth_a, zscale, d_theta, R, D, s));

  // ---- context projection weights: dWctx += dctx_all^T @ ctx ; db += colsum(dctx_all)
  if (use_ctx16(L, prepared, Rc)) {      // the forward pass left the fp16 pieces of the context rows in the workspace
    const uint16_t* hi = reinterpret_cast<const uint16_t*>(ws + W.ctx16);
    FP_TRY(wgrad(dctx_all, L.Nctx, (int)L.Nctx, ctx, L.C, L.C, Rc, grads + L.ctx_w, L.C, 0, s, grads + L.ctx_b, hi, hi + Rc * L.C));
  } else {
    FP_TRY(wgrad(dctx_all, L.Nctx, (int)L.Nctx, ctx, L.C, L.C, Rc, grads + L.ctx_w, L.C, 0, s, grads + L.ctx_b));
  }
  return 0;
}

// d_ctx[Rc, C] = dctx_all[Rc, Nctx] @ Wctx[Nctx, C]: the gradient wrt the (embedded) context rows, for an embedding
// net trained jointly in front of the flow (sbi FCEmbedding, /root/reference/modules.py:21-45).  Reads the context-
// projection gradient the backward pass left in the workspace.
int fp_nsf_ctx_backward(const fp_nsf_config* cfg, const float* params, int64_t R, int64_t Rc, const float* workspace,
                        float* d_ctx, void* stream) {
  FP_TRY(check_cfg(cfg));
  if (Rc <= 0) return 0;
  if (R % Rc != 0 || !d_ctx) return FP_E_BADARG;
  Layout L(cfg);
  Workspace W(L, R, Rc, true);
  fp_gemm_args g = dgrad_args(workspace + W.dctx_all, L.Nctx, (int)L.Nctx, params + L.ctx_w, L.C, L.C, Rc, d_ctx, L.C, 0);
  return launch_gemm(&g, (cudaStream_t)stream);
}

int fp_nsf_sample(const fp_nsf_config* cfg, const float* params, const float* prepared, const float* z, const float* ctx,
                  int64_t R, int64_t Rc, const float* zshift, const float* zscale, float* theta_out, float* workspace,
                  int32_t* status, void* stream) {
  FP_TRY(check_cfg(cfg));
  if (R <= 0) return 0;
  if (Rc <= 0 || R % Rc != 0) return FP_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  Layout L(cfg);
  Workspace W(L, R, Rc, false);
  float* ws = workspace;
  const int ctx_div = (Rc == 1) ? 0 : (int)(R / Rc);
  const int D = L.D;
  float* ctx_all = ws + W.ctx_all;
  g_gctx = GemmCtx{nullptr, status};
  FP_TRY(context_projection(L, params, prepared, ctx, Rc, ctx_all, ws + W.ctx16, s));
  const int fused = use_fused(cfg, L, prepared, false);
  if (fused) FP_TRY(launch_gate_sigmoid(ctx_all, Rc, L.Nctx, L.H, L.S, s));

  // inverse of the last LULinear on the base noise
  float* th_cur = ws + W.th_a; float* th_next = ws + W.th_b;
  FP_TRY(launch_rows_affine(z, prepared + L.prepMinv(L.T - 1), params + L.luBias(L.T - 1), -1.f, nullptr, th_cur, R, D, s));
  for (int i = L.T - 1; i >= 0; --i) {
    CondBufs b{};
    for (int k = 0; k <= L.Bk; ++k) b.h[k] = ws + W.h;
    for (int k = 0; k < L.Bk; ++k) { b.t[k] = ws + W.t; b.u[k] = nullptr; }
    b.prm = ws + W.prm;
    FP_TRY(conditioner_forward(cfg, L, params, prepared, i, th_cur, ctx_all, R, ctx_div, b, false, 0, fused, status, s,
                               (W.thp_sz > 0 && !fused) ? ws + W.thpad(0) : nullptr));
    // spline inverse of layer i fused with the inverse LULinear of layer i-1
    const float* M = (i > 0) ? prepared + L.prepMinv(i - 1) : nullptr;
    const float* pre = (i > 0) ? params + L.luBias(i - 1) : nullptr;
    FP_TRY(launch_rqs_apply(true, cfg, i & 1, th_cur, b.prm, th_next, nullptr, nullptr, nullptr, R, status, M, pre, -1.f,
                            nullptr, nullptr, L.ldP(i), s));
    float* tmp = th_cur; th_cur = th_next; th_next = tmp;
  }
  FP_TRY(launch_affine_inv(th_cur, zshift, zscale, theta_out, R, D, s));
  return 0;
}

}  // extern "C"
