// rqs_math.cuh -- per-element rational-quadratic spline arithmetic (forward, inverse, backward).
//
// Follows nflows 0.14 `rational_quadratic_spline` / `unconstrained_rational_quadratic_spline`
// (third-party, reached from /root/reference/modules.py:63 through sbi build_nsf; spec in
// SURVEY.md Appendix A.3-A.5).  Written as __host__ __device__ so the same arithmetic can be
// compiled with g++ for the CPU-side math test (tests/hostmath) -- the product only ever runs
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
  float dconst;   // (float)log(exp(1-md)-1): unnormalised boundary derivative
};

FP_HD float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
FP_HD float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// max and sum-of-exp of u[k]/sqrtH  (torch softmax: exp(x - max) / sum)
FP_HD void softmax_stats(const float* u, int K, float sqrtH, float& mx, float& sum) {
  mx = -INFINITY;
  for (int k = 0; k < K; ++k) mx = fmaxf(mx, u[k] / sqrtH);
  sum = 0.f;
  for (int k = 0; k < K; ++k) sum += expf(u[k] / sqrtH - mx);
}

// Scan the K+1 knots built from logits u.  If want_bin < 0: search the bin of `x`
// (count of knots <= x, minus one, last knot nudged by 1e-6 as upstream).  Else: select bin want_bin.
// Returns bin; lo = knot[bin], width = knot[bin+1]-knot[bin], soft = softmax prob of the bin,
// prefix = sum of softmax probs of bins < bin (used by the backward pass).
FP_HD int knot_scan(const float* u, int K, float sqrtH, float mx, float sum, float minb, float mix, float tb,
                    float span, float x, int want_bin, float& lo, float& width, float& soft, float& prefix) {
  float c = 0.f, prev = -tb, pacc = 0.f;
  int bin = 0;
  lo = -tb; width = 0.f; soft = 0.f; prefix = 0.f;
  for (int k = 0; k < K; ++k) {
    float e = expf(u[k] / sqrtH - mx) / sum;
    float wk = fp_add(minb, fp_mul(mix, e));
    c = fp_add(c, wk);
    float knot = (k == K - 1) ? tb : fp_add(fp_mul(span, c), -tb);
    bool take = (want_bin < 0) ? (x >= prev) : (k == want_bin);
    if (take) {
      bin = k; lo = prev; width = fp_add(knot, -prev); soft = e; prefix = pacc;
    }
    pacc += e;
    prev = knot;
  }
  return bin;
}

struct RqsElem {   // everything the evaluation needs about one element's bin
  int bin;         // -1: outside [-tb, tb] (linear tail)
  float cw, w, ch, h, d0, d1;
};

FP_HD float boundary_or_softplus(const float* ud, int j, int K, const RqsConsts& c) {
  // derivative at knot j (0..K); knots 0 and K carry the constant that makes it exactly ~1
  float raw = (j == 0 || j == K) ? c.dconst : ud[j - 1];
  return c.md + softplus_f(raw);
}

// p points at the 3K-1 raw conditioner outputs of this element: [uw(K) | uh(K) | ud(K-1)]
FP_HD void rqs_locate(const float* p, const RqsConsts& c, float x, bool inverse, RqsElem& e) {
  const float* uw = p; const float* uh = p + c.K; const float* ud = p + 2 * c.K;
  float mw, sw, mh, sh, soft, prefix;
  softmax_stats(uw, c.K, c.sqrtH, mw, sw);
  softmax_stats(uh, c.K, c.sqrtH, mh, sh);
  if (!inverse) {
    e.bin = knot_scan(uw, c.K, c.sqrtH, mw, sw, c.mbw, c.mixw, c.tb, c.span, x, -1, e.cw, e.w, soft, prefix);
    knot_scan(uh, c.K, c.sqrtH, mh, sh, c.mbh, c.mixh, c.tb, c.span, 0.f, e.bin, e.ch, e.h, soft, prefix);
  } else {
    e.bin = knot_scan(uh, c.K, c.sqrtH, mh, sh, c.mbh, c.mixh, c.tb, c.span, x, -1, e.ch, e.h, soft, prefix);
    knot_scan(uw, c.K, c.sqrtH, mw, sw, c.mbw, c.mixw, c.tb, c.span, 0.f, e.bin, e.cw, e.w, soft, prefix);
  }
  e.d0 = boundary_or_softplus(ud, e.bin, c.K, c);
  e.d1 = boundary_or_softplus(ud, e.bin + 1, c.K, c);
}

// forward: y, logabsdet, bin
FP_HD void rqs_forward_elem(const float* p, const RqsConsts& c, float x, float& y, float& lad, int& bin) {
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
FP_HD void rqs_inverse_elem(const float* p, const RqsConsts& c, float yin, float& x, float& lad, int& bin,
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

// backward of the forward map.  gy = dL/dy, gl = dL/dlogabsdet.  Writes dL/dp for all 3K-1 params into
// gp (stride gstride between consecutive params) and returns dL/dx.
FP_HD float rqs_backward_elem(const float* p, const RqsConsts& c, float x, float gy, float gl,
                              float* gp, int gstride) {
  const int K = c.K;
  if (!(x >= -c.tb && x <= c.tb)) {
    for (int k = 0; k < c.P; ++k) gp[k * gstride] = 0.f;
    return gy;
  }
  const float* uw = p; const float* uh = p + K; const float* ud = p + 2 * K;
  float mw, sw, mh, sh;
  softmax_stats(uw, K, c.sqrtH, mw, sw);
  softmax_stats(uh, K, c.sqrtH, mh, sh);
  float cw, w, ch, h, softw, prefw, softh, prefh;
  int bin = knot_scan(uw, K, c.sqrtH, mw, sw, c.mbw, c.mixw, c.tb, c.span, x, -1, cw, w, softw, prefw);
  knot_scan(uh, K, c.sqrtH, mh, sh, c.mbh, c.mixh, c.tb, c.span, 0.f, bin, ch, h, softh, prefh);
  float raw0 = (bin == 0) ? c.dconst : ud[bin - 1];
  float raw1 = (bin == K - 1) ? c.dconst : ud[bin];
  float d0 = c.md + softplus_f(raw0), d1 = c.md + softplus_f(raw1);

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
  float dotw = aw * prefw + bw * softw;
  float doth = ah * prefh + bh * softh;
  for (int k = 0; k < K; ++k) {
    float ew = expf(uw[k] / c.sqrtH - mw) / sw;
    float eh = expf(uh[k] / c.sqrtH - mh) / sh;
    float gsw = (k < bin) ? aw : ((k == bin) ? bw : 0.f);
    float gsh = (k < bin) ? ah : ((k == bin) ? bh : 0.f);
    gp[k * gstride] = ew * (gsw - dotw) / c.sqrtH;
    gp[(K + k) * gstride] = eh * (gsh - doth) / c.sqrtH;
  }
  for (int k = 0; k < K - 1; ++k) {
    float g = 0.f;
    if (k == bin - 1) g += g_d0 * (raw0 > 20.f ? 1.f : sigmoid_f(raw0));
    if (k == bin) g += g_d1 * (raw1 > 20.f ? 1.f : sigmoid_f(raw1));
    gp[(2 * K + k) * gstride] = g;
  }
  return g_x;
}

}  // namespace fp
