// embed_conv.cu -- the convolutional feature extractor of sbi CNNEmbedding (1-D), forward and backward.
//
// FlopPITy builds it with `-embedding -embedding_type CNN` (the default type, /root/reference/floppity.py:46-47):
// createEmbedding (/root/reference/modules.py:26-33) -> sbi.neural_nets.embedding_nets.CNNEmbedding(input_shape=(nwvl,),
// out_channels_per_layer, num_conv_layers, kernel_size, ...), whose cnn_subnet is num_conv_layers x
// [Conv1d(stride 1, padding 1) -> ReLU -> MaxPool1d(pool_kernel_size)] in front of an FCEmbedding (embedding.py runs that
// part on the GEMM kernels).  Upstream is three cuDNN / ATen launches per layer and an autograd graph; here a layer is
// ONE forward kernel (conv + bias + relu + pool, the pre-activation kept for the backward) and three backward kernels
// (pool/relu routing, input gradient, weight + bias gradient with block reductions).  Channel counts are tiny (4..24),
// so these are bandwidth-bound element-wise kernels, not GEMMs: weights sit in shared memory, rows are read coalesced.
#include "common.cuh"
#include "prof.h"
#include <algorithm>

namespace fp {

constexpr int CONV_MAX_W = 4096;      // Cout * Cin * k floats of weights in shared memory
constexpr int CONV_MAX_CK = 64;       // Cin * k accumulators per thread in the weight-gradient kernel

// y[b, co, lp] = max_{q < pool} relu(z[b, co, lp*pool + q]),  z[b, co, l] = bias[co] + sum_{ci, j} w[co, ci, j] x[b, ci, l + j - pad]
__global__ void k_conv_fwd(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                           float* __restrict__ z, float* __restrict__ y, int64_t B, int Cin, int L, int Cout, int k, int pad,
                           int pool, int Lc, int Lp) {
  extern __shared__ float sw[];
  for (int i = threadIdx.x; i < Cout * Cin * k; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int64_t n = B * Cout * (int64_t)Lp;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int lp = (int)(e % Lp);
    const int co = (int)((e / Lp) % Cout);
    const int64_t b = e / ((int64_t)Lp * Cout);
    const float* xb = x + b * Cin * (int64_t)L;
    const float* wc = sw + co * Cin * k;
    float best = 0.f;                                   // relu floor
    for (int q = 0; q < pool; ++q) {
      const int l = lp * pool + q;
      float acc = bias[co];
      for (int ci = 0; ci < Cin; ++ci) {
        const float* xr = xb + ci * (int64_t)L;
        for (int j = 0; j < k; ++j) {
          const int m = l + j - pad;
          if (m >= 0 && m < L) acc = fmaf(wc[ci * k + j], __ldg(xr + m), acc);
        }
      }
      z[(b * Cout + co) * (int64_t)Lc + l] = acc;
      best = fmaxf(best, acc);
    }
    y[e] = best;
    if (lp == Lp - 1) {                                 // conv positions the pooling drops: still part of z (zeros in dz)
      for (int l = Lp * pool; l < Lc; ++l) z[(b * Cout + co) * (int64_t)Lc + l] = 0.f;
    }
  }
}

// dz[b, co, l] = dy[b, co, l / pool] where l is the (first) arg-max of its pool window and z > 0, else 0
__global__ void k_conv_dz(const float* __restrict__ z, const float* __restrict__ dy, float* __restrict__ dz, int64_t B, int Cout,
                          int pool, int Lc, int Lp) {
  const int64_t n = B * Cout * (int64_t)Lp;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int lp = (int)(e % Lp);
    const int64_t row = e / Lp;
    const float* zr = z + row * Lc + (int64_t)lp * pool;
    float* dr = dz + row * Lc + (int64_t)lp * pool;
    int arg = 0;
    float best = fmaxf(zr[0], 0.f);
    for (int q = 1; q < pool; ++q) {
      const float v = fmaxf(zr[q], 0.f);
      if (v > best) { best = v; arg = q; }
    }
    const float g = dy[e];
    for (int q = 0; q < pool; ++q) dr[q] = (q == arg && zr[q] > 0.f) ? g : 0.f;
    if (lp == Lp - 1)
      for (int l = Lp * pool; l < Lc; ++l) dz[row * Lc + l] = 0.f;
  }
}

// dx[b, ci, m] = sum_{co, j} dz[b, co, m + pad - j] w[co, ci, j]
__global__ void k_conv_dx(const float* __restrict__ dz, const float* __restrict__ w, float* __restrict__ dx, int64_t B, int Cin,
                          int L, int Cout, int k, int pad, int Lc) {
  extern __shared__ float sw[];
  for (int i = threadIdx.x; i < Cout * Cin * k; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int64_t n = B * Cin * (int64_t)L;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(e % L);
    const int ci = (int)((e / L) % Cin);
    const int64_t b = e / ((int64_t)L * Cin);
    float acc = 0.f;
    for (int co = 0; co < Cout; ++co) {
      const float* dr = dz + (b * Cout + co) * (int64_t)Lc;
      const float* wc = sw + (co * Cin + ci) * k;
      for (int j = 0; j < k; ++j) {
        const int l = m + pad - j;
        if (l >= 0 && l < Lc) acc = fmaf(__ldg(dr + l), wc[j], acc);
      }
    }
    dx[e] = acc;
  }
}

// dw[co, ci, j] += sum_{b, l} dz[b, co, l] x[b, ci, l + j - pad] ; db[co] += sum dz[b, co, l]       (blockIdx.y = co)
__global__ void __launch_bounds__(256) k_conv_dw(const float* __restrict__ x, const float* __restrict__ dz, float* __restrict__ dw,
                                                 float* __restrict__ db, int64_t B, int Cin, int L, int Cout, int k, int pad, int Lc) {
  const int co = blockIdx.y;
  const int ck = Cin * k;
  float acc[CONV_MAX_CK];
#pragma unroll
  for (int i = 0; i < CONV_MAX_CK; ++i) acc[i] = 0.f;
  float accb = 0.f;
  const int64_t n = B * (int64_t)Lc;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int l = (int)(e % Lc);
    const int64_t b = e / Lc;
    const float d = dz[(b * Cout + co) * (int64_t)Lc + l];
    if (d == 0.f) continue;                              // most positions: routed away by the pooling or killed by the relu
    accb += d;
    const float* xb = x + b * Cin * (int64_t)L;
#pragma unroll
    for (int i = 0; i < CONV_MAX_CK; ++i) {
      if (i < ck) {
        const int ci = i / k, j = i - ci * k;
        const int m = l + j - pad;
        if (m >= 0 && m < L) acc[i] = fmaf(d, __ldg(xb + ci * (int64_t)L + m), acc[i]);
      }
    }
  }
  __shared__ float red[8][CONV_MAX_CK + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < CONV_MAX_CK; ++i) {
    if (i < ck) {
      float v = acc[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][i] = v;
    }
  }
  {
    float v = accb;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][CONV_MAX_CK] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= CONV_MAX_CK; i += blockDim.x) {
    if (i < ck || i == CONV_MAX_CK) {
      float v = 0.f;
      for (int wv = 0; wv < 8; ++wv) v += red[wv][i];
      if (v != 0.f) atomicAdd(i == CONV_MAX_CK ? db + co : dw + (int64_t)co * ck + i, v);
    }
  }
}

static int conv_grid(int64_t n) { return (int)std::min<int64_t>((n + 255) / 256, 16 * (int64_t)num_sms()); }

}  // namespace fp

using namespace fp;

extern "C" {

int fp_conv1d_relu_pool_fwd(const float* x, const float* w, const float* b, float* z, float* y, int64_t B, int32_t Cin, int32_t L,
                            int32_t Cout, int32_t k, int32_t pad, int32_t pool, void* stream) {
  if (B <= 0) return 0;
  const int Lc = L + 2 * pad - k + 1;
  if (Cin < 1 || Cout < 1 || k < 1 || pad < 0 || pool < 1 || Lc < pool) return FP_E_BADARG;
  if (Cout * Cin * k > CONV_MAX_W) return FP_E_TOOLARGE;
  const int Lp = Lc / pool;
  ProfScope ps(TAG_EMBED, 2.0 * (double)B * Cout * (double)Lp * pool * Cin * k, (cudaStream_t)stream,
               4.0 * (double)B * ((double)Cin * L + (double)Cout * (Lc + Lp)));
  k_conv_fwd<<<conv_grid(B * Cout * (int64_t)Lp), 256, sizeof(float) * Cout * Cin * k, (cudaStream_t)stream>>>(
      x, w, b, z, y, B, Cin, L, Cout, k, pad, pool, Lc, Lp);
  FP_LAUNCH_CHECK();
  return 0;
}

int fp_conv1d_relu_pool_bwd(const float* x, const float* w, const float* z, const float* dy, float* dz, float* dx, float* dw,
                            float* db, int64_t B, int32_t Cin, int32_t L, int32_t Cout, int32_t k, int32_t pad, int32_t pool,
                            void* stream) {
  if (B <= 0) return 0;
  const int Lc = L + 2 * pad - k + 1;
  if (Cin < 1 || Cout < 1 || k < 1 || pad < 0 || pool < 1 || Lc < pool) return FP_E_BADARG;
  if (Cout * Cin * k > CONV_MAX_W || Cin * k > CONV_MAX_CK) return FP_E_TOOLARGE;
  const int Lp = Lc / pool;
  cudaStream_t s = (cudaStream_t)stream;
  {
    ProfScope ps(TAG_EMBED, 0.0, s, 4.0 * (double)B * Cout * (2.0 * Lc + Lp));
    k_conv_dz<<<conv_grid(B * Cout * (int64_t)Lp), 256, 0, s>>>(z, dy, dz, B, Cout, pool, Lc, Lp);
    FP_LAUNCH_CHECK();
  }
  if (dx) {
    ProfScope ps(TAG_EMBED, 2.0 * (double)B * Cin * (double)L * Cout * k, s, 4.0 * (double)B * ((double)Cin * L + (double)Cout * Lc));
    k_conv_dx<<<conv_grid(B * Cin * (int64_t)L), 256, sizeof(float) * Cout * Cin * k, s>>>(dz, w, dx, B, Cin, L, Cout, k, pad, Lc);
    FP_LAUNCH_CHECK();
  }
  {
    ProfScope ps(TAG_EMBED, 2.0 * (double)B * Cout * (double)Lc * Cin * k, s, 4.0 * (double)B * ((double)Cin * L + (double)Cout * Lc));
    dim3 grid((unsigned)std::min<int64_t>((B * (int64_t)Lc + 255) / 256, 4 * (int64_t)num_sms()), (unsigned)Cout);
    k_conv_dw<<<grid, 256, 0, s>>>(x, dz, dw, db, B, Cin, L, Cout, k, pad, Lc);
    FP_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
