// TEST INFRASTRUCTURE: compiles the product's __host__ __device__ spline arithmetic
// (floppity_deprecated_b200/csrc/rqs_math.cuh) with g++ so the per-element math can be checked
// against the oracle on machines without a GPU.  Never loaded by the package.
#include "../../floppity_deprecated_b200/csrc/rqs_math.cuh"
#include <cmath>
#include <vector>
#include <algorithm>
using namespace fp;

static RqsConsts make_consts(int K, int H, float tb) {
  RqsConsts c;
  c.K = K; c.P = 3 * K - 1; c.tb = tb; c.span = (float)(2.0 * (double)tb);
  c.mbw = 1e-3f; c.mbh = 1e-3f; c.md = 1e-3f;
  c.mixw = (float)(1.0 - 1e-3 * K); c.mixh = (float)(1.0 - 1e-3 * K);
  c.sqrtH = (float)std::sqrt((double)H);
  c.inv_sqrtH = 1.0f / c.sqrtH;
  c.dconst = (float)std::log(std::exp(1.0 - 1e-3) - 1.0);
  return c;
}

extern "C" {
// x[n], params[n, 3K-1] -> y[n], lad[n], bin[n]
void hm_rqs_forward(const float* x, const float* params, int n, int K, int H, float tb, float* y, float* lad, int* bin) {
  RqsConsts c = make_consts(K, H, tb);
  std::vector<float> tmp(c.P);
  for (int i = 0; i < n; ++i) {
    std::copy(params + (long)i * c.P, params + (long)(i + 1) * c.P, tmp.begin());   // the kernels work in place on a smem copy
    rqs_forward_elem(tmp.data(), c, x[i], y[i], lad[i], bin[i]);
  }
}
void hm_rqs_inverse(const float* x, const float* params, int n, int K, int H, float tb, float* y, float* lad, int* bin) {
  RqsConsts c = make_consts(K, H, tb);
  bool nd;
  std::vector<float> tmp(c.P);
  for (int i = 0; i < n; ++i) {
    std::copy(params + (long)i * c.P, params + (long)(i + 1) * c.P, tmp.begin());
    rqs_inverse_elem(tmp.data(), c, x[i], y[i], lad[i], bin[i], nd);
  }
}
void hm_rqs_backward(const float* x, const float* params, const float* gy, const float* gl, int n, int K, int H,
                     float tb, float* gparams, float* gx) {
  RqsConsts c = make_consts(K, H, tb);
  for (int i = 0; i < n; ++i) {
    float* g = gparams + (long)i * c.P;
    std::copy(params + (long)i * c.P, params + (long)(i + 1) * c.P, g);
    gx[i] = rqs_backward_elem(g, c, x[i], gy[i], gl[i]);
  }
}
}
