// retrieval_kernels.cu -- the data-side neighbours of the flow:
//   * fp_preprocess_rows : noise augmentation + mean removal / residual scaling + per-bin affine scaling of simulated
//                          spectra, written as the fp32 context rows the flow trains on
//                          (/root/reference/floppityFUN.py:383-428 `preprocess`, :196-204 `sigma_res_scale`,
//                           :229-244 `rm_mean`; sklearn StandardScaler / PCA-centering as the affine step);
//   * fp_gauss_loglik    : row-wise Gaussian log-likelihood of model spectra against the observation
//                          (/root/reference/floppityFUN.py:24-28 `likelihood`, the inner loop of IS / evidence,
//                           floppityFUN.py:30-91).
// Both are pure streaming passes (HBM-bound): one CTA per output row walks the row with coalesced accesses, fp64
// arithmetic as in the reference (numpy), fp32 only at the final store of the context rows.
#include "common.cuh"
#include "prof.h"
#include <algorithm>

namespace fp {

__device__ __forceinline__ double block_sum(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();                 // red may still be read from a previous call
  if (l == 0) red[w] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

// mode 0: v = spec + noise z ; mode 1: v - rowmean(v), extra last column = rowmean ; mode 2: (v - obs) / noise
// then (v - shift) / scale when given (over the C (+1) output columns).
__global__ void __launch_bounds__(128)
k_preprocess_rows(const double* __restrict__ spec, const float* __restrict__ z, const double* __restrict__ noise,
                  const double* __restrict__ obs, const double* __restrict__ shift, const double* __restrict__ scale,
                  float* __restrict__ out, int64_t rows, int C, int naug, int mode, int64_t ldo) {
  __shared__ double red[4];
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const double* sp = spec + (r / naug) * (int64_t)C;
    const float* zr = z + r * (int64_t)C;
    float* o = out + r * ldo;
    double xbar = 0.0;
    if (mode == 1) {
      double s = 0.0;
      for (int j = threadIdx.x; j < C; j += blockDim.x) s += sp[j] + noise[j] * (double)zr[j];
      xbar = block_sum(s, red) / (double)C;
    }
    for (int j = threadIdx.x; j < C; j += blockDim.x) {
      double v = sp[j] + noise[j] * (double)zr[j];
      if (mode == 1) v -= xbar;
      else if (mode == 2) v = (v - obs[j]) / noise[j];
      if (shift) v -= shift[j];
      if (scale) v /= scale[j];
      o[j] = (float)v;
    }
    if (mode == 1 && threadIdx.x == 0) {
      double v = xbar;
      if (shift) v -= shift[C];
      if (scale) v /= scale[C];
      o[C] = (float)v;
    }
  }
}

// out[i] = sum_j [ -log(sqrt(2 pi) err_j) - (obs_j - x_ij)^2 / (2 err_j^2) ]
__global__ void __launch_bounds__(128)
k_gauss_loglik(const double* __restrict__ obs, const double* __restrict__ err, const double* __restrict__ x,
               double* __restrict__ out, int64_t rows, int C) {
  __shared__ double red[4];
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const double* xr = x + r * (int64_t)C;
    double s = 0.0;
    for (int j = threadIdx.x; j < C; j += blockDim.x) {
      const double e = err[j], d = obs[j] - xr[j];
      s += -log(2.5066282746310002 * e) - d * d / (2.0 * e * e);
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[r] = s;
  }
}

// relu forward / backward for the embedding net's output layer (every other relu of the net is a GEMM prologue / mask)
__global__ void k_relu(const float* __restrict__ x, float* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = fmaxf(x[i], 0.f);
}
__global__ void k_relu_bwd(const float* __restrict__ d, const float* __restrict__ z, float* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = z[i] > 0.f ? d[i] : 0.f;
}

}  // namespace fp

using namespace fp;

extern "C" {

int fp_preprocess_rows(const double* spec, const float* z, const double* noise, const double* obs, const double* shift,
                       const double* scale, float* out, int64_t n_spec, int32_t C, int32_t naug, int32_t mode,
                       int64_t ldo, void* stream) {
  if (n_spec < 0 || C < 1 || naug < 1 || mode < 0 || mode > 2 || ldo < C + (mode == 1 ? 1 : 0)) return FP_E_BADARG;
  if (mode == 2 && !obs) return FP_E_BADARG;
  const int64_t rows = n_spec * naug;
  if (rows == 0) return 0;
  ProfScope ps(TAG_OTHER, 0.0, (cudaStream_t)stream);
  const unsigned grid = (unsigned)std::min<int64_t>(rows, 16 * (int64_t)num_sms());
  k_preprocess_rows<<<grid, 128, 0, (cudaStream_t)stream>>>(spec, z, noise, obs, shift, scale, out, rows, C, naug, mode, ldo);
  FP_LAUNCH_CHECK();
  return 0;
}

int fp_gauss_loglik(const double* obs, const double* err, const double* x, double* out, int64_t rows, int32_t C, void* stream) {
  if (rows < 0 || C < 1) return FP_E_BADARG;
  if (rows == 0) return 0;
  ProfScope ps(TAG_OTHER, 0.0, (cudaStream_t)stream);
  const unsigned grid = (unsigned)std::min<int64_t>(rows, 16 * (int64_t)num_sms());
  k_gauss_loglik<<<grid, 128, 0, (cudaStream_t)stream>>>(obs, err, x, out, rows, C);
  FP_LAUNCH_CHECK();
  return 0;
}

int fp_relu(const float* x, float* out, int64_t n, void* stream) {
  if (n <= 0) return 0;
  ProfScope ps(TAG_OTHER, 0.0, (cudaStream_t)stream);
  k_relu<<<(unsigned)std::min<int64_t>((n + 255) / 256, 8 * (int64_t)num_sms()), 256, 0, (cudaStream_t)stream>>>(x, out, n);
  FP_LAUNCH_CHECK();
  return 0;
}
int fp_relu_bwd(const float* d, const float* z, float* out, int64_t n, void* stream) {
  if (n <= 0) return 0;
  ProfScope ps(TAG_OTHER, 0.0, (cudaStream_t)stream);
  k_relu_bwd<<<(unsigned)std::min<int64_t>((n + 255) / 256, 8 * (int64_t)num_sms()), 256, 0, (cudaStream_t)stream>>>(d, z, out, n);
  FP_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
