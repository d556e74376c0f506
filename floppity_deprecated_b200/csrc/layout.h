// layout.h -- flat parameter / prepared / workspace layouts of one NSF (host-side arithmetic only).
//
// Mirrors what sbi build_nsf (called at /root/reference/modules.py:63) instantiates: T x [coupling
// ResidualNet + LULinear].  All context-facing weights of all layers are stored contiguously
// ([T*(1+Bk)*H, C]) so the context projection of the whole flow is ONE GEMM.
#pragma once
#include <stdint.h>
#include "../../include/floppity_b200.h"

#if defined(__CUDACC__)
#define FP_LHD __host__ __device__ inline
#else
#define FP_LHD inline
#endif

namespace fp {

FP_LHD int64_t al4_(int64_t x) { return (x + 3) & ~int64_t(3); }
inline int64_t al4(int64_t x) { return (x + 3) & ~int64_t(3); }
inline int64_t al32(int64_t x) { return (x + 31) & ~int64_t(31); }

struct Layout {
  int D, C, T, K, H, Bk, P, S;  // S = 1 + Bk context slots per layer
  int64_t Nctx;                 // T * S * H
  int64_t ctx_w, ctx_b, layers; // offsets of the regions
  int64_t ntri;

  FP_LHD explicit Layout(const fp_nsf_config* c)
      : D(c->D), C(c->C), T(c->T), K(c->K), H(c->H), Bk(c->Bk), P(3 * c->K - 1), S(1 + c->Bk) {
    Nctx = (int64_t)T * S * H;
    ctx_w = 0;
    ctx_b = al4_(Nctx * C);
    layers = ctx_b + al4_(Nctx);
    ntri = (int64_t)D * (D - 1) / 2;
  }
  FP_LHD int dtr(int i) const { return (i % 2 == 0) ? (D + 1) / 2 : D / 2; }
  FP_LHD int dP(int i) const { return dtr(i) * P; }
  FP_LHD int dPmax() const { return dtr(0) * P; }
  FP_LHD int ldP(int i) const { return (dP(i) + 3) & ~3; }      // padded row stride of the spline-parameter buffers
  FP_LHD int ldPmax() const { return (dPmax() + 3) & ~3; }
  FP_LHD int ldH() const { return (H + 3) & ~3; }               // padded row stride of the hidden-activation buffers
  FP_LHD int64_t block_sz() const { return al4_((int64_t)H * H) * 2 + al4_(H) * 2; }
  FP_LHD int64_t layer_sz(int i) const {
    return al4_((int64_t)H * D) + Bk * block_sz() + al4_((int64_t)dP(i) * H) + al4_(dP(i)) + 2 * al4_(ntri) + 2 * al4_(D);
  }
  FP_LHD int64_t layer_off(int i) const {
    int64_t o = layers;
    for (int l = 0; l < i; ++l) o += layer_sz(l);
    return o;
  }
  FP_LHD int64_t total() const { return layer_off(T); }
  // ---- per-tensor offsets
  FP_LHD int64_t ctxW(int i, int slot) const { return ctx_w + ((int64_t)(i * S + slot) * H) * C; }
  FP_LHD int64_t ctxB(int i, int slot) const { return ctx_b + (int64_t)(i * S + slot) * H; }
  FP_LHD int64_t ctxCol(int i, int slot) const { return (int64_t)(i * S + slot) * H; }
  FP_LHD int64_t w0x(int i) const { return layer_off(i); }
  FP_LHD int64_t w1(int i, int b) const { return layer_off(i) + al4_((int64_t)H * D) + b * block_sz(); }
  FP_LHD int64_t b1(int i, int b) const { return w1(i, b) + al4_((int64_t)H * H); }
  FP_LHD int64_t w2(int i, int b) const { return b1(i, b) + al4_(H); }
  FP_LHD int64_t b2(int i, int b) const { return w2(i, b) + al4_((int64_t)H * H); }
  FP_LHD int64_t wf(int i) const { return layer_off(i) + al4_((int64_t)H * D) + Bk * block_sz(); }
  FP_LHD int64_t bf(int i) const { return wf(i) + al4_((int64_t)dP(i) * H); }
  FP_LHD int64_t luLower(int i) const { return bf(i) + al4_(dP(i)); }
  FP_LHD int64_t luUpper(int i) const { return luLower(i) + al4_(ntri); }
  FP_LHD int64_t luDiag(int i) const { return luUpper(i) + al4_(ntri); }
  FP_LHD int64_t luBias(int i) const { return luDiag(i) + al4_(D); }
  // ---- gradient buffer = params layout + tail: dM accumulators [T][D*D] + sum(d_logp)
  FP_LHD int64_t gradTailDM(int i) const { return total() + (int64_t)i * al4_((int64_t)D * D); }
  FP_LHD int64_t gradTailSum() const { return total() + (int64_t)T * al4_((int64_t)D * D); }
  FP_LHD int64_t gradTotal() const { return gradTailSum() + 4; }
  // ---- prepared buffer: per layer M=L*U, Minv, L, U, logdet
  FP_LHD int64_t prepSz() const { return 4 * al4_((int64_t)D * D) + 4; }
  FP_LHD int64_t prepM(int i) const { return (int64_t)i * prepSz(); }
  FP_LHD int64_t prepMinv(int i) const { return prepM(i) + al4_((int64_t)D * D); }
  FP_LHD int64_t prepL(int i) const { return prepM(i) + 2 * al4_((int64_t)D * D); }
  FP_LHD int64_t prepU(int i) const { return prepM(i) + 3 * al4_((int64_t)D * D); }
  FP_LHD int64_t prepLogdet(int i) const { return prepM(i) + 4 * al4_((int64_t)D * D); }
  FP_LHD int64_t prepTotal() const { return (int64_t)T * prepSz(); }
  // tensor-core path: TF32 hi / lo copies of the whole parameter buffer (same offsets as the parameters)
  FP_LHD int64_t prepHi() const { return (prepTotal() + 31) & ~int64_t(31); }
  FP_LHD int64_t prepLo() const { return prepHi() + ((total() + 31) & ~int64_t(31)); }
  // transposed (+ split) weight copies for the dgrad GEMMs: per layer W0x^T [D,H], per block W1^T, W2^T [H,H],
  // Wf^T [H, ldP]; region T_hi then T_lo
  FP_LHD int64_t tLayerSz(int i) const { return al4_((int64_t)D * H) + Bk * 2 * al4_((int64_t)H * H) + al4_((int64_t)H * ldP(i)); }
  FP_LHD int64_t tLayerOff(int i) const { int64_t o = 0; for (int l = 0; l < i; ++l) o += tLayerSz(l); return o; }
  FP_LHD int64_t tW0(int i) const { return tLayerOff(i); }
  FP_LHD int64_t tW1(int i, int b) const { return tLayerOff(i) + al4_((int64_t)D * H) + (int64_t)b * 2 * al4_((int64_t)H * H); }
  FP_LHD int64_t tW2(int i, int b) const { return tW1(i, b) + al4_((int64_t)H * H); }
  FP_LHD int64_t tWf(int i) const { return tLayerOff(i) + al4_((int64_t)D * H) + Bk * 2 * al4_((int64_t)H * H); }
  FP_LHD int64_t tTotal() const { return (tLayerOff(T) + 31) & ~int64_t(31); }
  FP_LHD int64_t prepTHi() const { return prepLo() + ((total() + 31) & ~int64_t(31)); }
  FP_LHD int64_t prepTLo() const { return prepTHi() + tTotal(); }
  // fused conditioner in 3xFP16 mode (cond_fused.cu): per layer, fp16 hi and scaled-residual lo copies of
  //   [ W1_0 | W2_0 | W1_1 | ... (2*Bk matrices of H x H) | Wf (dP rows padded to a multiple of 8, H columns) | W0x (H x 64) ]
  // (units below: halves; the regions are addressed as float offsets into `prepared`, two halves per float)
  FP_LHD int64_t f16WfRows(int i) const { return (dP(i) + 7) & ~7; }
  FP_LHD int64_t f16LayerHalves(int i) const { return (int64_t)2 * Bk * H * H + f16WfRows(i) * H + (int64_t)H * 64; }
  FP_LHD int64_t f16LayerOff(int i) const { int64_t o = 0; for (int l = 0; l < i; ++l) o += f16LayerHalves(l); return o; }   // halves
  FP_LHD int64_t f16W(int i, int k, int g) const { return f16LayerOff(i) + (int64_t)(2 * k + g) * H * H; }
  FP_LHD int64_t f16Wf(int i) const { return f16LayerOff(i) + (int64_t)2 * Bk * H * H; }
  FP_LHD int64_t f16W0(int i) const { return f16Wf(i) + f16WfRows(i) * H; }
  FP_LHD int64_t f16Floats() const { return ((f16LayerOff(T) / 2) + 63) & ~int64_t(63); }
  FP_LHD int64_t prepF16Hi() const { return ((prepTLo() + tTotal()) + 63) & ~int64_t(63); }
  FP_LHD int64_t prepF16Lo() const { return prepF16Hi() + f16Floats(); }
  // shared-memory-resident conditioner for small hidden widths (cond_small.cu, H <= 64): per layer the exact byte image
  // the kernels copy into shared memory, fp16 hi and scaled-residual lo of
  //   forward : W0x [64][K0+8] | 2*Bk x (W1_k, W2_k) [64][72] | Wf [NF][72]                 (rows = output unit, k contiguous)
  //   backward: Wf^T [64][KF+8] | 2*Bk x (W1_k^T, W2_k^T) [64][72] | W0x^T [D8][72]
  // row strides are (contraction + 8) halves = an odd multiple of 16 bytes: ldmatrix reads are bank-conflict-free
  FP_LHD bool smallH() const { return H <= 64; }
  FP_LHD int shK0() const { return (D + 15) & ~15; }
  FP_LHD int shNF(int i) const { return (dP(i) + 7) & ~7; }
  FP_LHD int shKF(int i) const { return (dP(i) + 15) & ~15; }
  FP_LHD int shD8() const { return (D + 7) & ~7; }
  FP_LHD int64_t shFwdHalves(int i) const { return (int64_t)64 * (shK0() + 8) + (int64_t)2 * Bk * 64 * 72 + (int64_t)shNF(i) * 72; }
  FP_LHD int64_t shBwdHalves(int i) const { return (int64_t)64 * (shKF(i) + 8) + (int64_t)2 * Bk * 64 * 72 + (int64_t)shD8() * 72; }
  FP_LHD int64_t shLayerHalves(int i) const { return 2 * (shFwdHalves(i) + shBwdHalves(i)); }      // [F hi | F lo | B hi | B lo]
  FP_LHD int64_t shLayerOff(int i) const { int64_t o = 0; for (int l = 0; l < i; ++l) o += shLayerHalves(l); return o; }   // halves
  FP_LHD int64_t prepSmall() const { return (prepF16Lo() + f16Floats() + 63) & ~int64_t(63); }
  FP_LHD int64_t prepSmallEnd() const { return prepSmall() + (smallH() ? ((shLayerOff(T) / 2 + 63) & ~int64_t(63)) : 0); }
  // 3xFP16 stand-alone GEMMs (gemm_tc.cu, fp_gemm_tc16): fp16 hi / scaled-residual lo copies of the weights.  The copy of the
  // tensor at float offset `woff` starts at BYTE offset 4 * woff of its region (half index 2 * woff: 16-byte aligned like
  // the tensor itself) with the tensor's own leading dimension; the transposed copies likewise at 4 * toff with the
  // leading dimension rounded up to 8 halves (ld16).
  FP_LHD int64_t prepG16Hi() const { return (prepSmallEnd() + 63) & ~int64_t(63); }
  FP_LHD int64_t prepG16Lo() const { return prepG16Hi() + ((total() + 63) & ~int64_t(63)); }
  FP_LHD int64_t prepGT16Hi() const { return prepG16Lo() + ((total() + 63) & ~int64_t(63)); }
  FP_LHD int64_t prepGT16Lo() const { return prepGT16Hi() + tTotal(); }
  FP_LHD static int ld16(int ld) { return (ld + 7) & ~7; }
  FP_LHD int64_t prepAll() const { return prepGT16Lo() + tTotal(); }
};

// workspace carve; identical arithmetic in forward and backward so saved activations are found again
struct Workspace {
  int64_t ctx_all, logdet, ctx16, thp, thp_sz;
  // inference
  int64_t th_a, th_b, h, t, prm;
  // training
  int64_t th_st, th_sp, act;  // act: per-layer region
  int64_t dth_a, dth_b, dh, dt, du, dprm, dctx_all, dub, gscale;
  int64_t total;
  int64_t R, Rc;
  int64_t act_layer_sz[64];
  int64_t act_layer_off[64];
  const Layout& L;

  Workspace(const Layout& l, int64_t R_, int64_t Rc_, bool training) : R(R_), Rc(Rc_), L(l) {
    int64_t o = 0;
    auto take = [&](int64_t n) { int64_t r = o; o += al32(n); return r; };
    ctx_all = take(Rc * L.Nctx);
    logdet = take(R);
    // theta rows padded to a multiple of 8 floats (D = 30 -> 32: 120-byte rows are not TMA-addressable): the A operand of the
    // first conditioner GEMM and the B operand of its weight gradient; one buffer per layer when training
    thp_sz = ((L.D & 7) != 0 && L.D >= 4 && !L.smallH()) ? al32(R * Layout::ld16(L.D)) : 0;
    thp = take((training ? L.T : 1) * thp_sz);
    ctx16 = take(Rc * L.C);      // fp16 hi | lo pieces of the context rows (3xFP16 context projection and its weight gradient)
    th_a = th_b = h = t = prm = th_st = th_sp = act = dth_a = dth_b = dh = dt = du = dprm = dctx_all = dub = gscale = 0;
    if (!training) {
      th_a = take(R * L.D); th_b = take(R * L.D);
      h = take(R * L.ldH()); t = take(R * L.ldH()); prm = take(R * L.ldPmax());
    } else {
      th_st = take((int64_t)(L.T + 1) * al32(R * L.D));
      th_sp = take((int64_t)L.T * al32(R * L.D));
      act = o;
      for (int i = 0; i < L.T && i < 64; ++i) {
        act_layer_off[i] = o;
        take((int64_t)(L.Bk + 1) * al32(R * L.ldH()));  // h_0..h_Bk
        take((int64_t)L.Bk * al32(R * L.ldH()));        // t_b (pre-relu output of linear 0, after dropout)
        take((int64_t)L.Bk * al32(R * L.ldH()));        // u_b (pre-gate output of linear 1)
        take(R * L.ldP(i));                   // spline params (padded rows)
        act_layer_sz[i] = o - act_layer_off[i];
      }
      dth_a = take(R * L.D); dth_b = take(R * L.D);
      dh = take(R * L.ldH()); dt = take(R * L.ldH()); du = take(R * L.ldH());
      dprm = take(R * L.ldPmax());
      dctx_all = take(Rc * L.Nctx);
      // the fused small-hidden backward keeps every block's du / dt for the weight gradients that follow it
      dub = L.smallH() ? take((int64_t)2 * L.Bk * al32(R * L.ldH())) : 0;
      gscale = take(4);      // power-of-two scale of the gradient operands of the 3xFP16 GEMMs (k_grad_scale)
    }
    total = o;
  }
  int64_t thpad(int i) const { return thp + (int64_t)i * thp_sz; }      // valid when thp_sz > 0
  int64_t thst(int i) const { return th_st + (int64_t)i * al32(R * L.D); }
  int64_t thsp(int i) const { return th_sp + (int64_t)i * al32(R * L.D); }
  int64_t hsz() const { return al32(R * L.ldH()); }
  int64_t hbuf(int i, int b) const { return act_layer_off[i] + (int64_t)b * hsz(); }
  int64_t tbuf(int i, int b) const { return act_layer_off[i] + (int64_t)(L.Bk + 1) * hsz() + (int64_t)b * hsz(); }
  int64_t ubuf(int i, int b) const { return act_layer_off[i] + (int64_t)(2 * L.Bk + 1) * hsz() + (int64_t)b * hsz(); }
  int64_t prmbuf(int i) const { return act_layer_off[i] + (int64_t)(3 * L.Bk + 1) * hsz(); }
  int64_t dubuf(int b) const { return dub + (int64_t)(2 * b) * hsz(); }
  int64_t dtbuf(int b) const { return dub + (int64_t)(2 * b + 1) * hsz(); }
};

}  // namespace fp
