// rqs_math.cuh -- per-element rational-quadratic spline arithmetic (forward, inverse, backward).
//
// Follows nflows 0.14 `rational_quadratic_spline` / `unconstrained_rational_quadratic_spline`
// (third-party, reached from /root/reference/modules.py:63 through sbi build_nsf; spec in
// SURVEY.md Appendix A.3-A.5).  Written as __host__ __device__ so the same arithmetic can be
// compiled with g++ for a CPU-side arithmetic check in the test-suite -- the product only ever runs
// the nvcc build.
//
// Op order on the knot path (softmax -> min-width mix -> sequential cumsum -> affine -> pinned
// ends) is kept identical to upstream, with contraction into FMA disabled there (fp_mul/fp_add),
// because the bin index is decided by those values.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define FP_HD __host__ __device__ __forceinline__
#else
#define FP_HD inline
#endif

namespace fp {

#if defined(__CUDA_ARCH__)
FP_HD float fp_mul(float a, float b) { return __fmul_rn(a, b); }
FP_HD float fp_add(float a, float b) { return __fadd_rn(a, b); }
#else
FP_HD float fp_mul(float a, float b) { volatile float r = a * b; return r; }
FP_HD float fp_add(float a, float b) { volatile float r = a + b; return r; }
#endif

struct RqsConsts {
  int K;          // bins
  int P;          // 3K-1 params per element
  float tb;       // tail bound
  float span;     // (float)(2*tb)
  float mbw, mbh, md;
  float mixw, mixh;  // (float)(1 - min_bin_width*K), (float)(1 - min_bin_height*K), evaluated in double like upstream
  float sqrtH;    // (float)sqrt(hidden_features)
  float inv_sqrtH; // 1/sqrtH (fp32 reciprocal)
  float dconst;   // (float)log(exp(1-md)-1): unnormalised boundary derivative
};

#if defined(__CUDA_ARCH__) && !defined(FP_RQS_PRECISE_EXP)
// Device fast paths (each keeps ~1e-6 RELATIVE accuracy where the value feeds a product, ~2e-7 ABSOLUTE where it
// feeds a sum): ex2.approx / lg2.approx / rcp.approx instead of the ~8-30 instruction libdevice sequences.
FP_HD float fp_ex2(float t) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t)); return r; }
FP_HD float fp_div(float a, float b) { return __fdividef(a, b); }
FP_HD float fp_log_abs(float a) { return __logf(a); }     // absolute error ~2e-7: only used where the result is summed
// softplus(x) = max(x,0) + log1p(exp(-|x|)), log1p(t) = 2 atanh(t/(2+t)) as an odd series in s = t/(2+t) <= 1/3:
// uniformly RELATIVE accurate (~2e-7), which log(1+t) with an approximate log is not for small t.
FP_HD float softplus_f(float x) {
  if (x > 20.f) return x;
  float t = fp_ex2(-fabsf(x) * 1.4426950408889634f);
  float s = __fdividef(t, 2.f + t);
  float s2 = s * s;
  float p = fmaf(s2, 1.f / 11.f, 1.f / 9.f);
  p = fmaf(s2, p, 1.f / 7.f);
  p = fmaf(s2, p, 1.f / 5.f);
  p = fmaf(s2, p, 1.f / 3.f);
  p = fmaf(s2, p, 1.f);
  return fmaxf(x, 0.f) + 2.f * s * p;
}
FP_HD float sigmoid_f(float x) { return __fdividef(1.f, 1.f + fp_ex2(-x * 1.4426950408889634f)); }
#else
FP_HD float fp_div(float a, float b) { return a / b; }
FP_HD float fp_log_abs(float a) { return logf(a); }
FP_HD float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
FP_HD float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }
#endif

// ---- in-place parameterisation --------------------------------------------------------------------
// `u` points at K logits in (shared) memory.  They are replaced by the UNNORMALISED softmax numerators
// exp(u_k/sqrtH - max) -- computed exactly once per logit -- and 1/sum is returned.  (u * (1/sqrtH) and
// e * (1/sum) stand for upstream's two divisions; they differ from a true division by at most 1 ulp, which can
// move a bin edge by ~1e-7: the parity tests enumerate such near-knot elements with the fp64 oracle.)
FP_HD float softmax_numerators_inplace(float* u, int K, float inv_sqrtH) {
  float mx = -INFINITY;
  for (int k = 0; k < K; ++k) mx = fmaxf(mx, fp_mul(u[k], inv_sqrtH));
  float sum = 0.f;
  for (int k = 0; k < K; ++k) {
    float e = expf(fp_add(fp_mul(u[k], inv_sqrtH), -mx));
    u[k] = e;
    sum += e;
  }
  return 1.f / sum;
}

// Scan the K+1 knots built from numerators e (see above).  want_bin < 0: search the bin of `x` (count of knots
// <= x, minus one; the last knot is nudged by 1e-6 upstream, which can never matter for x <= tb).  Else: pick
// bin want_bin.  lo = knot[bin], hi = knot[bin+1].
FP_HD int knot_scan(const float* e, int K, float rsum, float minb, float mix, float tb, float span,
                    float x, int want_bin, float& lo, float& hi) {
  float c = 0.f, prev = -tb;
  int bin = 0;
  lo = -tb; hi = tb;
  for (int k = 0; k < K; ++k) {
    float wk = fp_add(minb, fp_mul(mix, fp_mul(e[k], rsum)));
    c = fp_add(c, wk);
    float knot = (k == K - 1) ? tb : fp_add(fp_mul(span, c), -tb);
    bool take = (want_bin < 0) ? (x >= prev) : (k == want_bin);
    if (take) { bin = k; lo = prev; hi = knot; }
    prev = knot;
  }
  return bin;
}

struct RqsElem {   // everything the evaluation needs about one element's bin
  int bin;         // -1: outside [-tb, tb] (linear tail)
  float cw, w, ch, h, d0, d1;
  float rsw, rsh;  // 1/sum of the width / height numerators (left in memory by rqs_locate)
};

FP_HD float boundary_or_raw(const float* ud, int j, int K, const RqsConsts& c) {
  return (j == 0 || j == K) ? c.dconst : ud[j - 1];
}

// p points at the 3K-1 raw conditioner outputs of this element: [uw(K) | uh(K) | ud(K-1)]; uw and uh are
// overwritten by their softmax numerators.
FP_HD void rqs_locate(float* p, const RqsConsts& c, float x, bool inverse, RqsElem& e) {
  float* uw = p; float* uh = p + c.K; const float* ud = p + 2 * c.K;
  e.rsw = softmax_numerators_inplace(uw, c.K, c.inv_sqrtH);
  e.rsh = softmax_numerators_inplace(uh, c.K, c.inv_sqrtH);
  float hiw, hih;
  if (!inverse) {
    e.bin = knot_scan(uw, c.K, e.rsw, c.mbw, c.mixw, c.tb, c.span, x, -1, e.cw, hiw);
    knot_scan(uh, c.K, e.rsh, c.mbh, c.mixh, c.tb, c.span, 0.f, e.bin, e.ch, hih);
  } else {
    e.bin = knot_scan(uh, c.K, e.rsh, c.mbh, c.mixh, c.tb, c.span, x, -1, e.ch, hih);
    knot_scan(uw, c.K, e.rsw, c.mbw, c.mixw, c.tb, c.span, 0.f, e.bin, e.cw, hiw);
  }
  e.w = fp_add(hiw, -e.cw);
  e.h = fp_add(hih, -e.ch);
  e.d0 = c.md + softplus_f(boundary_or_raw(ud, e.bin, c.K, c));
  e.d1 = c.md + softplus_f(boundary_or_raw(ud, e.bin + 1, c.K, c));
}

// forward: y, logabsdet, bin
FP_HD void rqs_forward_elem(float* p, const RqsConsts& c, float x, float& y, float& lad, int& bin) {
  if (!(x >= -c.tb && x <= c.tb)) { y = x; lad = 0.f; bin = -1; return; }
  RqsElem e; rqs_locate(p, c, x, false, e);
  bin = e.bin;
  float delta = e.h / e.w;
  float th = (x - e.cw) / e.w;
  float m = th * (1.f - th);
  float s = e.d0 + e.d1 - 2.f * delta;
  float num = e.h * (delta * th * th + e.d0 * m);
  float den = delta + s * m;
  y = e.ch + num / den;
  float omt = 1.f - th;
  float dnum = delta * delta * (e.d1 * th * th + 2.f * delta * m + e.d0 * omt * omt);
  lad = logf(dnum) - 2.f * logf(den);
}

// inverse: x from y; returns the (already negated) logabsdet of the inverse map; neg_disc set if b^2-4ac < 0
FP_HD void rqs_inverse_elem(float* p, const RqsConsts& c, float yin, float& x, float& lad, int& bin,
                            bool& neg_disc) {
  neg_disc = false;
  if (!(yin >= -c.tb && yin <= c.tb)) { x = yin; lad = 0.f; bin = -1; return; }
  RqsElem e; rqs_locate(p, c, yin, true, e);
  bin = e.bin;
  float delta = e.h / e.w;
  float dy = yin - e.ch;
  float s = e.d0 + e.d1 - 2.f * delta;
  float a = dy * s + e.h * (delta - e.d0);
  float b = e.h * e.d0 - dy * s;
  float cc = -delta * dy;
  float disc = b * b - 4.f * a * cc;
  if (disc < 0.f) neg_disc = true;
  float root = (2.f * cc) / (-b - sqrtf(disc));
  x = root * e.w + e.cw;
  float m = root * (1.f - root);
  float den = delta + s * m;
  float omr = 1.f - root;
  float dnum = delta * delta * (e.d1 * root * root + 2.f * delta * m + e.d0 * omr * omr);
  lad = -(logf(dnum) - 2.f * logf(den));
}

// backward of the forward map, IN PLACE: p holds the 3K-1 raw params on entry and dL/dp on exit.
// gy = dL/dy, gl = dL/dlogabsdet.  Returns dL/dx.
FP_HD float rqs_backward_elem(float* p, const RqsConsts& c, float x, float gy, float gl) {
  const int K = c.K;
  if (!(x >= -c.tb && x <= c.tb)) {
    for (int k = 0; k < c.P; ++k) p[k] = 0.f;
    return gy;
  }
  float* uw = p; float* uh = p + K; float* ud = p + 2 * K;
  RqsElem e; rqs_locate(p, c, x, false, e);
  const int bin = e.bin;
  const float cw = e.cw, w = e.w, h = e.h, d0 = e.d0, d1 = e.d1;
  float raw0 = boundary_or_raw(ud, bin, K, c);
  float raw1 = boundary_or_raw(ud, bin + 1, K, c);

  float delta = h / w;
  float th = (x - cw) / w;
  float omt = 1.f - th;
  float m = th * omt;
  float s = d0 + d1 - 2.f * delta;
  float pp = delta * th * th + d0 * m;
  float num = h * pp;
  float den = delta + s * m;
  float q = d1 * th * th + 2.f * delta * m + d0 * omt * omt;
  float dn = delta * delta * q;

  float g_num = gy / den;
  float g_den = -gy * num / (den * den) - 2.f * gl / den;
  float g_dn = gl / dn;
  float g_delta = g_dn * (2.f * delta * q + delta * delta * 2.f * m);
  float g_q = g_dn * delta * delta;
  float g_d1 = g_q * th * th;
  float g_d0 = g_q * omt * omt;
  float g_m = g_q * 2.f * delta;
  float g_th = g_q * (2.f * d1 * th - 2.f * d0 * omt);
  g_delta += g_den;
  float g_s = g_den * m;
  g_m += g_den * s;
  g_d0 += g_s; g_d1 += g_s; g_delta -= 2.f * g_s;
  float g_h = g_num * pp;
  float g_pp = g_num * h;
  g_delta += g_pp * th * th;
  g_th += g_pp * 2.f * delta * th;
  g_d0 += g_pp * m;
  g_m += g_pp * d0;
  g_th += g_m * (1.f - 2.f * th);
  g_h += g_delta / w;
  float g_w = -g_delta * delta / w;
  float g_x = g_th / w;
  float g_cw = -g_th / w;
  g_w -= g_th * th / w;
  float g_ch = gy;

  // knot gradients: value knots bin and bin+1; knots 0 and K are pinned constants
  float GWlo = (bin >= 1) ? (g_cw - g_w) : 0.f;
  float GWhi = (bin + 1 <= K - 1) ? g_w : 0.f;
  float GHlo = (bin >= 1) ? (g_ch - g_h) : 0.f;
  float GHhi = (bin + 1 <= K - 1) ? g_h : 0.f;
  // d/d softmax-prob_k: (1 - minb K) * span * (sum of knot grads for knots j > k)
  float aw = c.mixw * c.span * (GWlo + GWhi);   // k < bin
  float bw = c.mixw * c.span * GWhi;            // k == bin
  float ah = c.mixh * c.span * (GHlo + GHhi);
  float bh = c.mixh * c.span * GHhi;
  // softmax backward needs sum_j prob_j * g_j = a * (sum of probs below the bin) + b * prob_bin
  float prefw = 0.f, prefh = 0.f;
  for (int k = 0; k < K; ++k) {
    if (k < bin) { prefw += uw[k]; prefh += uh[k]; }
  }
  float dotw = (aw * prefw + bw * uw[bin]) * e.rsw;
  float doth = (ah * prefh + bh * uh[bin]) * e.rsh;
  for (int k = 0; k < K; ++k) {
    float gsw = (k < bin) ? aw : ((k == bin) ? bw : 0.f);
    float gsh = (k < bin) ? ah : ((k == bin) ? bh : 0.f);
    uw[k] = uw[k] * e.rsw * (gsw - dotw) * c.inv_sqrtH;
    uh[k] = uh[k] * e.rsh * (gsh - doth) * c.inv_sqrtH;
  }
  float gd0 = g_d0 * (raw0 > 20.f ? 1.f : sigmoid_f(raw0));
  float gd1 = g_d1 * (raw1 > 20.f ? 1.f : sigmoid_f(raw1));
  for (int k = 0; k < K - 1; ++k) {
    float g = 0.f;
    if (k == bin - 1) g += gd0;
    if (k == bin) g += gd1;
    ud[k] = g;
  }
  return g_x;
}


// =====================================================================================================
// Compile-time-K variants: identical arithmetic and op order on the knot path, but fully unrolled with the
// softmax numerators held in registers (no shared-memory read-modify-write, no loop overhead).  The kernels
// dispatch to these for the bin counts FlopPITy uses (floppity.py:48 default 8; BASELINE configs 8/10/16).
// FP_RQS_FAST_EXP (default on for the device build) evaluates the 2K softmax exponentials as
// ex2.approx((u/sqrtH - max) * log2 e): 2 instructions instead of ~8, error <= ~3 ulp instead of <= 2 ulp.
// =====================================================================================================
#if defined(__CUDA_ARCH__) && !defined(FP_RQS_PRECISE_EXP)
#define FP_RQS_FAST_EXP 1
#endif

template <int K>
FP_HD void numerators_t(const float* u, const RqsConsts& c, float (&e)[K], float& rsum) {
  float v[K];
  float mx = -INFINITY;
#if defined(FP_RQS_FAST_EXP)
  const float sl = c.inv_sqrtH * 1.4426950408889634f;
#pragma unroll
  for (int k = 0; k < K; ++k) { v[k] = u[k] * sl; mx = fmaxf(mx, v[k]); }
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) { e[k] = fp_ex2(v[k] - mx); sum += e[k]; }
#else
#pragma unroll
  for (int k = 0; k < K; ++k) { v[k] = fp_mul(u[k], c.inv_sqrtH); mx = fmaxf(mx, v[k]); }
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) { e[k] = expf(fp_add(v[k], -mx)); sum += e[k]; }
#endif
  rsum = fp_div(1.f, sum);
}

// search: bin = (number of knots <= x) - 1 over the sequentially accumulated knots; lo/hi = its edges
template <int K>
FP_HD int knot_search_t(const float (&e)[K], float rsum, float minb, float mix, float tb, float span, float x,
                        float& lo, float& hi) {
  float c = 0.f, prev = -tb;
  int bin = 0;
  lo = -tb; hi = tb;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float wk = fp_add(minb, fp_mul(mix, fp_mul(e[k], rsum)));
    c = fp_add(c, wk);
    float knot = (k == K - 1) ? tb : fp_add(fp_mul(span, c), -tb);
    if (x >= prev) { bin = k; lo = prev; hi = knot; }
    prev = knot;
  }
  return bin;
}
// pick: edges of a given bin (same sequential cumsum)
template <int K>
FP_HD void knot_pick_t(const float (&e)[K], float rsum, float minb, float mix, float tb, float span, int bin,
                       float& lo, float& hi) {
  float c = 0.f, wsel = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float wk = fp_add(minb, fp_mul(mix, fp_mul(e[k], rsum)));
    if (k < bin) c = fp_add(c, wk);
    if (k == bin) wsel = wk;
  }
  lo = (bin == 0) ? -tb : fp_add(fp_mul(span, c), -tb);
  hi = (bin == K - 1) ? tb : fp_add(fp_mul(span, fp_add(c, wsel)), -tb);
}

template <int K>
FP_HD void rqs_locate_t(const float* p, const RqsConsts& c, float x, bool inverse, RqsElem& e, float (&ew)[K],
                        float (&eh)[K]) {
  numerators_t<K>(p, c, ew, e.rsw);
  numerators_t<K>(p + K, c, eh, e.rsh);
  float hiw, hih;
  if (!inverse) {
    e.bin = knot_search_t<K>(ew, e.rsw, c.mbw, c.mixw, c.tb, c.span, x, e.cw, hiw);
    knot_pick_t<K>(eh, e.rsh, c.mbh, c.mixh, c.tb, c.span, e.bin, e.ch, hih);
  } else {
    e.bin = knot_search_t<K>(eh, e.rsh, c.mbh, c.mixh, c.tb, c.span, x, e.ch, hih);
    knot_pick_t<K>(ew, e.rsw, c.mbw, c.mixw, c.tb, c.span, e.bin, e.cw, hiw);
  }
  e.w = fp_add(hiw, -e.cw);
  e.h = fp_add(hih, -e.ch);
  const float* ud = p + 2 * K;
  e.d0 = c.md + softplus_f(boundary_or_raw(ud, e.bin, K, c));
  e.d1 = c.md + softplus_f(boundary_or_raw(ud, e.bin + 1, K, c));
}

FP_HD void rqs_eval_forward(const RqsElem& e, float x, float& y, float& lad) {
  float rw = fp_div(1.f, e.w);
  float delta = e.h * rw;
  float th = (x - e.cw) * rw;
  float m = th * (1.f - th);
  float s = e.d0 + e.d1 - 2.f * delta;
  float num = e.h * (delta * th * th + e.d0 * m);
  float den = delta + s * m;
  float rden = fp_div(1.f, den);
  y = e.ch + num * rden;
  float omt = 1.f - th;
  float dnum = delta * delta * (e.d1 * th * th + 2.f * delta * m + e.d0 * omt * omt);
  lad = fp_log_abs(dnum * rden * rden);
}
FP_HD void rqs_eval_inverse(const RqsElem& e, float yin, float& x, float& lad, bool& neg_disc) {
  float delta = fp_div(e.h, e.w);
  float dy = yin - e.ch;
  float s = e.d0 + e.d1 - 2.f * delta;
  float a = dy * s + e.h * (delta - e.d0);
  float b = e.h * e.d0 - dy * s;
  float cc = -delta * dy;
  float disc = b * b - 4.f * a * cc;
  neg_disc = disc < 0.f;
  float root = fp_div(2.f * cc, -b - sqrtf(disc));
  x = root * e.w + e.cw;
  float m = root * (1.f - root);
  float den = delta + s * m;
  float rden = fp_div(1.f, den);
  float omr = 1.f - root;
  float dnum = delta * delta * (e.d1 * root * root + 2.f * delta * m + e.d0 * omr * omr);
  lad = -fp_log_abs(dnum * rden * rden);
}

template <int K>
FP_HD void rqs_forward_elem_t(const float* p, const RqsConsts& c, float x, float& y, float& lad, int& bin) {
  if (!(x >= -c.tb && x <= c.tb)) { y = x; lad = 0.f; bin = -1; return; }
  RqsElem e; float ew[K], eh[K];
  rqs_locate_t<K>(p, c, x, false, e, ew, eh);
  bin = e.bin;
  rqs_eval_forward(e, x, y, lad);
}
template <int K>
FP_HD void rqs_inverse_elem_t(const float* p, const RqsConsts& c, float yin, float& x, float& lad, int& bin,
                              bool& neg_disc) {
  neg_disc = false;
  if (!(yin >= -c.tb && yin <= c.tb)) { x = yin; lad = 0.f; bin = -1; return; }
  RqsElem e; float ew[K], eh[K];
  rqs_locate_t<K>(p, c, yin, true, e, ew, eh);
  bin = e.bin;
  rqs_eval_inverse(e, yin, x, lad, neg_disc);
}

// shared tail of the backward: from the located bin to (g_x, knot-edge gradients, derivative gradients)
struct RqsBwdOut { float g_x, aw, bw, ah, bh, gd0, gd1; };
FP_HD RqsBwdOut rqs_backward_core(const RqsElem& e, const RqsConsts& c, int K, float x, float gy, float gl,
                                  float raw0, float raw1) {
  const int bin = e.bin;
  const float cw = e.cw, w = e.w, h = e.h, d0 = e.d0, d1 = e.d1;
  const float rw = fp_div(1.f, w);
  float delta = h * rw;
  float th = (x - cw) * rw;
  float omt = 1.f - th;
  float m = th * omt;
  float s = d0 + d1 - 2.f * delta;
  float pp = delta * th * th + d0 * m;
  float num = h * pp;
  float den = delta + s * m;
  float q = d1 * th * th + 2.f * delta * m + d0 * omt * omt;
  float dn = delta * delta * q;
  const float rden = fp_div(1.f, den);
  float g_num = gy * rden;
  float g_den = (-gy * num * rden - 2.f * gl) * rden;
  float g_dn = fp_div(gl, dn);
  float g_delta = g_dn * (2.f * delta * q + delta * delta * 2.f * m);
  float g_q = g_dn * delta * delta;
  float g_d1 = g_q * th * th;
  float g_d0 = g_q * omt * omt;
  float g_m = g_q * 2.f * delta;
  float g_th = g_q * (2.f * d1 * th - 2.f * d0 * omt);
  g_delta += g_den;
  float g_s = g_den * m;
  g_m += g_den * s;
  g_d0 += g_s; g_d1 += g_s; g_delta -= 2.f * g_s;
  float g_h = g_num * pp;
  float g_pp = g_num * h;
  g_delta += g_pp * th * th;
  g_th += g_pp * 2.f * delta * th;
  g_d0 += g_pp * m;
  g_m += g_pp * d0;
  g_th += g_m * (1.f - 2.f * th);
  g_h += g_delta * rw;
  float g_w = -g_delta * delta * rw;
  RqsBwdOut o;
  o.g_x = g_th * rw;
  float g_cw = -g_th * rw;
  g_w -= g_th * th * rw;
  float g_ch = gy;
  float GWlo = (bin >= 1) ? (g_cw - g_w) : 0.f;
  float GWhi = (bin + 1 <= K - 1) ? g_w : 0.f;
  float GHlo = (bin >= 1) ? (g_ch - g_h) : 0.f;
  float GHhi = (bin + 1 <= K - 1) ? g_h : 0.f;
  o.aw = c.mixw * c.span * (GWlo + GWhi);
  o.bw = c.mixw * c.span * GWhi;
  o.ah = c.mixh * c.span * (GHlo + GHhi);
  o.bh = c.mixh * c.span * GHhi;
  o.gd0 = g_d0 * (raw0 > 20.f ? 1.f : sigmoid_f(raw0));
  o.gd1 = g_d1 * (raw1 > 20.f ? 1.f : sigmoid_f(raw1));
  return o;
}

template <int K>
FP_HD float rqs_backward_elem_t(float* p, const RqsConsts& c, float x, float gy, float gl) {
  constexpr int P = 3 * K - 1;
  if (!(x >= -c.tb && x <= c.tb)) {
#pragma unroll
    for (int k = 0; k < P; ++k) p[k] = 0.f;
    return gy;
  }
  RqsElem e; float ew[K], eh[K];
  rqs_locate_t<K>(p, c, x, false, e, ew, eh);
  const int bin = e.bin;
  float* ud = p + 2 * K;
  float raw0 = boundary_or_raw(ud, bin, K, c);
  float raw1 = boundary_or_raw(ud, bin + 1, K, c);
  RqsBwdOut o = rqs_backward_core(e, c, K, x, gy, gl, raw0, raw1);
  float prefw = 0.f, prefh = 0.f, selw = 0.f, selh = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    if (k < bin) { prefw += ew[k]; prefh += eh[k]; }
    if (k == bin) { selw = ew[k]; selh = eh[k]; }
  }
  float dotw = (o.aw * prefw + o.bw * selw) * e.rsw;
  float doth = (o.ah * prefh + o.bh * selh) * e.rsh;
  const float sw = e.rsw * c.inv_sqrtH, sh = e.rsh * c.inv_sqrtH;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float gsw = (k < bin) ? o.aw : ((k == bin) ? o.bw : 0.f);
    float gsh = (k < bin) ? o.ah : ((k == bin) ? o.bh : 0.f);
    p[k] = ew[k] * sw * (gsw - dotw);
    p[K + k] = eh[k] * sh * (gsh - doth);
  }
#pragma unroll
  for (int k = 0; k < K - 1; ++k) {
    float g = 0.f;
    if (k == bin - 1) g += o.gd0;
    if (k == bin) g += o.gd1;
    ud[k] = g;
  }
  return o.g_x;
}

}  // namespace fp
