// misc_kernels.cu -- the small fused kernels around the GEMMs and the spline:
// z-score affine, base-density log-prob, GLU backward, LU-linear prepare / finish, column sums,
// SNPE-C atomic loss, clip+Adam, row gathers.  Each cites the upstream op it stands for
// (third-party sbi 0.22 / nflows 0.14 reached from /root/reference/modules.py:63-64,
// floppity.py:216-225,252; spec SURVEY.md Appendix A).
#include "common.cuh"
#include "layout.h"
#include "prof.h"
#include "gemm_epilogue.cuh"
#include <math.h>
#include <algorithm>

namespace fp {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float softplus_d(float x) { return x > 20.f ? x : log1pf(expf(x)); }

// ---- nflows PointwiseAffineTransform forward (sbi standardizing_transform) + logdet init ----------
__global__ void k_affine_fwd(const float* __restrict__ theta, const float* __restrict__ shift,
                             const float* __restrict__ scale, float zlogdet, float* __restrict__ out,
                             float* __restrict__ logdet, int64_t R, int D) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < R * D) {
    int f = (int)(i % D);
    out[i] = __fadd_rn(__fmul_rn(theta[i], scale[f]), shift[f]);
  }
  if (i < R) logdet[i] = zlogdet;
}
// inverse: (y - shift) / scale
__global__ void k_affine_inv(const float* __restrict__ y, const float* __restrict__ shift,
                             const float* __restrict__ scale, float* __restrict__ out, int64_t R, int D) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < R * D) {
    int f = (int)(i % D);
    out[i] = (y[i] - shift[f]) / scale[f];
  }
}
// d_theta_raw = d_theta0 * scale
__global__ void k_affine_bwd(const float* __restrict__ d0, const float* __restrict__ scale, float* __restrict__ out,
                             int64_t R, int D) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < R * D) out[i] = d0[i] * scale[(int)(i % D)];
}

// ---- nflows StandardNormal._log_prob + Flow._log_prob: logp = -0.5*|z|^2 - D/2 log(2pi) + logdet ----
__global__ void k_base_logprob(const float* __restrict__ z, const float* __restrict__ logdet, float* __restrict__ logp,
                               int64_t R, int D, float log_z, int32_t* status) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float s = 0.f;
  for (int f = 0; f < D; ++f) { float v = z[r * D + f]; s += v * v; }
  float lp = (-0.5f * s - log_z) + logdet[r];
  logp[r] = lp;
  if (status && !isfinite(lp)) atomicAdd(&status[0], 1);
}
// d z = -z * d_logp
__global__ void k_base_bwd(const float* __restrict__ z, const float* __restrict__ d_logp, float* __restrict__ dz,
                           int64_t R, int D) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < R * D) dz[i] = -z[i] * d_logp[i / D];
}

// ---- y = M (x + sgn*pre) + post for rows (LULinear forward/inverse with the product matrix) ---------
__global__ void k_rows_affine(const float* __restrict__ x, const float* __restrict__ M, const float* __restrict__ pre,
                              float sgn, const float* __restrict__ post, float* __restrict__ y, int64_t R, int D) {
  extern __shared__ float sm[];
  float* sM = sm; float* sPre = sm + D * D;
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) sM[i] = M[i];
  for (int i = threadIdx.x; i < D; i += blockDim.x) sPre[i] = pre ? sgn * pre[i] : 0.f;
  __syncthreads();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * D) return;
  int64_t r = i / D; int f = (int)(i - r * D);
  const float* xr = x + r * D;
  float acc = 0.f;
  for (int k = 0; k < D; ++k) acc = fmaf(sM[f * D + k], xr[k] + sPre[k], acc);
  y[i] = acc + (post ? post[f] : 0.f);
}

// ---- GLU backward (nflows ResidualBlock: temps = glu(cat(temps, context_layer(ctx)))) -----------------
// h_next = h + u*sigmoid(G).  dh is the gradient wrt h_next (and, unchanged, wrt h via the residual).
// du = dh*sig ; dG[cr] = sum over the `A` rows sharing context row cr of dh*u*sig*(1-sig).
// sig_given: G already holds sigmoid(gate) (the fused conditioner path applies k_gate_sigmoid to the context projection).
__global__ void k_glu_bwd(const float* __restrict__ dh, const float* __restrict__ u, const float* __restrict__ G,
                          int64_t ldg, float* __restrict__ du, float* __restrict__ dG, int64_t Rc, int A, int H, int ldh,
                          int sig_given) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Rc * H) return;
  int64_t cr = i / H; int j = (int)(i - cr * H);
  float g = G[cr * ldg + j];
  float sg = sig_given ? g : 1.f / (1.f + expf(-g));
  float acc = 0.f;
  for (int a = 0; a < A; ++a) {
    int64_t r = cr * A + a;
    float d = dh[r * ldh + j];
    du[r * ldh + j] = d * sg;
    acc = fmaf(d, u[r * ldh + j], acc);
  }
  dG[cr * ldg + j] = acc * sg * (1.f - sg);
}
// float4 version (H % 4 == 0, ldh % 4 == 0, ldg % 4 == 0, 16-byte aligned bases): four columns per thread
__global__ void k_glu_bwd4(const float* __restrict__ dh, const float* __restrict__ u, const float* __restrict__ G,
                           int64_t ldg, float* __restrict__ du, float* __restrict__ dG, int64_t Rc, int A, int H, int ldh,
                           int sig_given) {
  const int H4 = H >> 2;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Rc * H4) return;
  const int64_t cr = i / H4; const int j = (int)(i - cr * H4) * 4;
  const float4 g = *reinterpret_cast<const float4*>(G + cr * ldg + j);
  float4 sg = g;
  if (!sig_given) {
    sg.x = 1.f / (1.f + expf(-g.x)); sg.y = 1.f / (1.f + expf(-g.y)); sg.z = 1.f / (1.f + expf(-g.z)); sg.w = 1.f / (1.f + expf(-g.w));
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int a = 0; a < A; ++a) {
    const int64_t r = cr * A + a;
    const float4 d = *reinterpret_cast<const float4*>(dh + r * ldh + j);
    const float4 uu = *reinterpret_cast<const float4*>(u + r * ldh + j);
    *reinterpret_cast<float4*>(du + r * ldh + j) = make_float4(d.x * sg.x, d.y * sg.y, d.z * sg.z, d.w * sg.w);
    acc.x = fmaf(d.x, uu.x, acc.x); acc.y = fmaf(d.y, uu.y, acc.y); acc.z = fmaf(d.z, uu.z, acc.z); acc.w = fmaf(d.w, uu.w, acc.w);
  }
  *reinterpret_cast<float4*>(dG + cr * ldg + j) = make_float4(acc.x * sg.x * (1.f - sg.x), acc.y * sg.y * (1.f - sg.y),
                                                              acc.z * sg.z * (1.f - sg.z), acc.w * sg.w * (1.f - sg.w));
}
// in place: gate slots (every slot but the first of each layer) of the context projection [Rc, T*S*H] -> sigmoid
__global__ void k_gate_sigmoid(float* __restrict__ ctx_all, int64_t n4, int H, int S) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    // rows are a whole number of slots, so the slot index of flat element e is simply (e / H) % S
    const int64_t e = i * 4;
    if (((e / H) % S) == 0) continue;
    float4 v = reinterpret_cast<float4*>(ctx_all)[i];
    v.x = 1.f / (1.f + expf(-v.x)); v.y = 1.f / (1.f + expf(-v.y));
    v.z = 1.f / (1.f + expf(-v.z)); v.w = 1.f / (1.f + expf(-v.w));
    reinterpret_cast<float4*>(ctx_all)[i] = v;
  }
}
__global__ void k_gate_sigmoid1(float* __restrict__ ctx_all, int64_t n, int H, int S) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    if (((e / H) % S) != 0) ctx_all[e] = 1.f / (1.f + expf(-ctx_all[e]));
}
// first-layer context gradient: dG[cr, j] = sum_a dh[cr*A+a, j]
__global__ void k_ctx_reduce(const float* __restrict__ dh, float* __restrict__ dG, int64_t ldg, int64_t Rc, int A, int H, int ldh) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Rc * H) return;
  int64_t cr = i / H; int j = (int)(i - cr * H);
  float acc = 0.f;
  for (int a = 0; a < A; ++a) acc += dh[(cr * A + a) * ldh + j];
  dG[cr * ldg + j] = acc;
}

// ---- column sums (bias gradients): out[n] += sum_m X[m, n] ----------------------------------------------
// Block = 32 column lanes x 8 row lanes; a block covers 32*VEC columns x rows_per_block rows.  Each thread keeps
// eight independent loads in flight (the old one-thread-per-column loop was a serial chain of dependent loads),
// row lanes are combined through shared memory, one atomic per column per block.
template <int VEC>
__global__ void __launch_bounds__(256) k_colsum(const float* __restrict__ X, int64_t ldx, int64_t M, int N,
                                                float* __restrict__ out, int rows_per_block) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = (blockIdx.x * 32 + tx) * VEC;
  const int64_t m0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t m1 = min(M, m0 + rows_per_block);
  float acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
  if (n < N) {
#pragma unroll 8
    for (int64_t m = m0 + ty; m < m1; m += 8) {
      if (VEC == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(X + m * ldx + n));
        acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
      } else {
        acc[0] += __ldg(X + m * ldx + n);
      }
    }
  }
  __shared__ float red[8][32 * VEC + 1];
#pragma unroll
  for (int j = 0; j < VEC; ++j) red[ty][tx * VEC + j] = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * VEC; c += 256) {
    const int nn = blockIdx.x * 32 * VEC + c;
    if (nn < N) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) t += red[r][c];
      atomicAdd(&out[nn], t);
    }
  }
}

// ---- sum reduction: out[0] += scale * sum(x) ----------------------------------------------------------------
__global__ void k_sum(const float* __restrict__ x, int64_t n, float scale, float* __restrict__ out) {
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc += x[i];
  acc = warp_sum(acc);
  __shared__ float ws[32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) ws[w] = acc;
  __syncthreads();
  if (w == 0) {
    acc = (lane < (blockDim.x + 31) / 32) ? ws[lane] : 0.f;
    acc = warp_sum(acc);
    if (lane == 0) atomicAdd(out, acc * scale);
  }
}
__global__ void k_sumsq(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = x[i]; acc = fmaf(v, v, acc);
  }
  acc = warp_sum(acc);
  __shared__ float ws[32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) ws[w] = acc;
  __syncthreads();
  if (w == 0) {
    acc = (lane < (blockDim.x + 31) / 32) ? ws[lane] : 0.f;
    acc = warp_sum(acc);
    if (lane == 0) atomicAdd(out, acc);
  }
}

// ---- LULinear: build L, U, M = L*U, Minv, logdet from the packed entries (nflows LULinear._create_lower_upper)
// one block per layer
__global__ void k_lu_prepare(const float* __restrict__ params, float* __restrict__ prepared, fp_nsf_config cfg) {
  extern __shared__ float sm[];
  const int D = cfg.D; const float eps = cfg.lu_eps;
  float* L = sm; float* U = sm + D * D; float* Y = sm + 2 * D * D;  // Y: scratch for the inverse
  const int layer = blockIdx.x;
  const Layout lay(&cfg);
  const float* lower = params + lay.luLower(layer);
  const float* upper = params + lay.luUpper(layer);
  const float* udiag = params + lay.luDiag(layer);
  float* out = prepared + lay.prepM(layer);
  const int DD = (D * D + 3) & ~3;
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
    int r = i / D, c = i - r * D;
    float l = 0.f, u = 0.f;
    if (r > c) l = lower[r * (r - 1) / 2 + c];
    else if (r == c) { l = 1.f; u = softplus_d(udiag[r]) + eps; }
    else {
      // row-major strict upper: row r holds D-1-r entries, preceded by sum_{q<r} (D-1-q)
      int base = r * (D - 1) - r * (r - 1) / 2;
      u = upper[base + (c - r - 1)];
    }
    L[i] = l; U[i] = u;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
    int r = i / D, c = i - r * D;
    float acc = 0.f;
    for (int k = 0; k < D; ++k) acc = fmaf(L[r * D + k], U[k * D + c], acc);
    out[i] = acc;                // M
    out[2 * DD + i] = L[i];
    out[3 * DD + i] = U[i];
  }
  // Minv column j: solve L y = e_j, U x = y
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    float* y = Y + j * D;
    for (int r = 0; r < D; ++r) {
      float acc = (r == j) ? 1.f : 0.f;
      for (int k = 0; k < r; ++k) acc -= L[r * D + k] * y[k];
      y[r] = acc;
    }
    for (int r = D - 1; r >= 0; --r) {
      float acc = y[r];
      for (int k = r + 1; k < D; ++k) acc -= U[r * D + k] * y[k];
      y[r] = acc / U[r * D + r];
    }
    for (int r = 0; r < D; ++r) out[DD + r * D + j] = y[r];
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int r = 0; r < D; ++r) s += logf(U[r * D + r]);
    out[4 * DD] = s;
  }
}

// ---- fold dM (accumulated by the LU wgrad GEMM) into d lower / d upper / d diag; zero structural zeros of W0X
// dL = strict_lower(dM U^T), dU = upper(L^T dM); d udiag = dU_jj*sigmoid(udiag) + sum(d_logp)/U_jj*sigmoid(udiag)
__global__ void k_lu_finish(const float* __restrict__ params, const float* __restrict__ prepared, float* __restrict__ grads,
                            fp_nsf_config cfg) {
  extern __shared__ float sm[];
  const int D = cfg.D, H = cfg.H;
  float* L = sm; float* U = sm + D * D; float* dM = sm + 2 * D * D;
  const int layer = blockIdx.x;
  const Layout lay(&cfg);
  const int DD = (D * D + 3) & ~3;
  const float* prep = prepared + lay.prepM(layer);
  const float* gdm = grads + lay.gradTailDM(layer);
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) { L[i] = prep[2 * DD + i]; U[i] = prep[3 * DD + i]; dM[i] = gdm[i]; }
  __syncthreads();
  float* glower = grads + lay.luLower(layer);
  float* gupper = grads + lay.luUpper(layer);
  float* gdiag = grads + lay.luDiag(layer);
  const float* udiag = params + lay.luDiag(layer);
  const float sum_dlogp = grads[lay.gradTailSum()];
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
    int r = i / D, c = i - r * D;
    if (r > c) {
      float acc = 0.f;  // (dM U^T)[r,c] = sum_k dM[r,k] U[c,k]
      for (int k = 0; k < D; ++k) acc = fmaf(dM[r * D + k], U[c * D + k], acc);
      glower[r * (r - 1) / 2 + c] += acc;
    } else {
      float acc = 0.f;  // (L^T dM)[r,c] = sum_k L[k,r] dM[k,c]
      for (int k = 0; k < D; ++k) acc = fmaf(L[k * D + r], dM[k * D + c], acc);
      if (r == c) {
        float x = udiag[r];
        float sg = x > 20.f ? 1.f : 1.f / (1.f + expf(-x));
        gdiag[r] += (acc + sum_dlogp / U[r * D + r]) * sg;
      } else {
        int base = r * (D - 1) - r * (r - 1) / 2;
        gupper[base + (c - r - 1)] += acc;
      }
    }
  }
  // structural zeros of the expanded first-layer weight: column f is transformed in this layer iff f%2 == layer%2
  float* gw0 = grads + lay.w0x(layer);
  for (int i = threadIdx.x; i < H * D; i += blockDim.x) {
    int f = i % D;
    if ((f & 1) == (layer & 1)) gw0[i] = 0.f;
  }
}

// ---- SNPE-C atomic loss ------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix(uint64_t& s) {
  s += 0x9E3779B97F4A7C15ull;
  uint64_t z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// Floyd's algorithm: A-1 distinct values from {0..B-2}, then skip self.
__global__ void k_atom_choices(int32_t* __restrict__ choices, int64_t B, int A, uint64_t seed, const int32_t* __restrict__ step) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (step) seed += (uint64_t)(uint32_t)step[0] * 0x9E3779B97F4A7C15ull;   // device-side stream position (CUDA-graph replay)
  uint64_t s = seed * 0xD1342543DE82EF95ull + (uint64_t)b * 0x2545F4914F6CDD1Dull + 1;
  int n = A - 1;
  int64_t Nn = B - 1;
  int32_t* out = choices + b * n;
  int cnt = 0;
  for (int64_t j = Nn - n; j < Nn; ++j) {
    uint64_t r = splitmix(s);
    int64_t t = (int64_t)(r % (uint64_t)(j + 1));
    bool found = false;
    for (int q = 0; q < cnt; ++q) found |= (out[q] == (int32_t)t);
    out[cnt++] = found ? (int32_t)j : (int32_t)t;
  }
  for (int q = 0; q < n; ++q) { int32_t v = out[q]; out[q] = v + (v >= b ? 1 : 0); }
}
__global__ void k_atom_gather(const float* __restrict__ theta, const int32_t* __restrict__ choices, float* __restrict__ rows,
                              int64_t B, int A, int D) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * A * D) return;
  int f = (int)(i % D);
  int64_t ra = i / D;
  int64_t b = ra / A; int a = (int)(ra - b * A);
  int64_t src = (a == 0) ? b : (int64_t)choices[b * (A - 1) + (a - 1)];
  rows[i] = theta[src * D + f];
}
// one thread per batch row
__global__ void k_atomic_loss(const float* __restrict__ logp, const float* __restrict__ atom_theta,
                              const float* __restrict__ low, const float* __restrict__ high,
                              const float* __restrict__ masks, int combined, float* __restrict__ out_lp,
                              float* __restrict__ d_logp, int64_t B, int A, int D, int32_t* status) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float lp_box = 0.f;
  for (int f = 0; f < D; ++f) lp_box -= logf(high[f] - low[f]);
  float mx = -INFINITY;
  for (int a = 0; a < A; ++a) {
    const float* th = atom_theta + (b * A + a) * D;
    bool inside = true;
    for (int f = 0; f < D; ++f) inside &= (th[f] >= low[f]) && (th[f] < high[f]);
    float prior = inside ? lp_box : -INFINITY;
    float lp = logp[b * A + a];
    if (status && (!isfinite(lp) || !inside)) atomicAdd(&status[0], 1);
    float u = lp - prior;
    mx = fmaxf(mx, u);
  }
  float se = 0.f;
  for (int a = 0; a < A; ++a) {
    // prior is the same constant for every atom inside the box
    float u = logp[b * A + a] - lp_box;
    se += expf(u - mx);
  }
  float lse = mx + logf(se);
  float u0 = logp[b * A] - lp_box;
  float mk = masks ? masks[b] : 0.f;
  float out = u0 - lse;
  if (combined) out = mk * logp[b * A] + out;
  out_lp[b] = out;
  if (d_logp) {
    float invB = 1.f / (float)B;
    for (int a = 0; a < A; ++a) {
      float u = logp[b * A + a] - lp_box;
      float p = expf(u - lse);
      float g = ((a == 0) ? 1.f : 0.f) - p;
      if (a == 0 && combined) g += mk;
      d_logp[b * A + a] = -g * invB;  // d(-mean out)/d logp
    }
  }
}

// ---- clip_grad_norm_ + Adam (torch.optim.Adam defaults; sbi train loop floppity.py:219) --------------------
__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                       int64_t n, float lr, float b1, float b2, float eps, const int32_t* __restrict__ step_dev,
                       float max_norm, float grad_scale, const float* __restrict__ sumsq) {
  float clip = 1.f;
  if (max_norm > 0.f && sumsq) {
    float total = sqrtf(sumsq[0]) * fabsf(grad_scale);
    float c = max_norm / (total + 1e-6f);
    clip = c < 1.f ? c : 1.f;
  }
  const float gs = clip * grad_scale;
  const int step = step_dev ? *step_dev : 1;
  const float bc1 = 1.f - powf(b1, (float)step);
  const float bc2 = 1.f - powf(b2, (float)step);
  const float step_size = lr / bc1;
  const float bc2s = sqrtf(bc2);
  auto upd = [&](float& pi, float gr, float& mi, float& vi) {
    const float gi = gr * gs;
    mi = mi + (1.f - b1) * (gi - mi);                      // torch: exp_avg.lerp_(grad, 1-beta1)
    vi = b2 * vi + (1.f - b2) * gi * gi;
    const float denom = sqrtf(vi) / bc2s + eps;
    pi -= step_size * (mi / denom);
  };
  // four 16-byte streams in, three out per thread and iteration (the buffers come from the caching allocator: 16-byte aligned)
  const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0;
  const int64_t n4 = vec ? n >> 2 : 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 P = reinterpret_cast<float4*>(p)[i], M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
    const float4 G = reinterpret_cast<const float4*>(g)[i];
    upd(P.x, G.x, M.x, V.x); upd(P.y, G.y, M.y, V.y); upd(P.z, G.z, M.z, V.z); upd(P.w, G.w, M.w, V.w);
    reinterpret_cast<float4*>(m)[i] = M; reinterpret_cast<float4*>(v)[i] = V; reinterpret_cast<float4*>(p)[i] = P;
  }
  for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float pi = p[i], mi = m[i], vi = v[i];
    upd(pi, g[i], mi, vi);
    m[i] = mi; v[i] = vi; p[i] = pi;
  }
}

__global__ void k_gather_rows(const float* __restrict__ src, const int64_t* __restrict__ idx, float* __restrict__ dst,
                              int64_t n, int width) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * width) return;
  int64_t r = i / width; int c = (int)(i - r * width);
  dst[i] = src[idx[r] * width + c];
}
__global__ void k_standardize(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ sd,
                              float* __restrict__ out, int64_t n, int width) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * width) return;
  int c = (int)(i % width);
  out[i] = (x[i] - mean[c]) / sd[c];
}
__global__ void k_within_box(const float* __restrict__ theta, const float* __restrict__ low, const float* __restrict__ high,
                             uint8_t* __restrict__ inside, int64_t R, int D) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  bool ok = true;
  for (int f = 0; f < D; ++f) { float v = theta[r * D + f]; ok &= (v >= low[f]) && (v <= high[f]); }
  inside[r] = ok ? 1 : 0;
}

// Rejection step of accept_reject_sample (sbi.samplers.rejection, reached from floppity.py:252): rows inside the prior
// box are appended to `out` behind a device counter -- one warp-aggregated atomic per 32 candidates -- until `cap`
// rows are stored.  The order of accepted rows is not the candidate order (they are i.i.d. draws).
__global__ void k_compact_inside(const float* __restrict__ theta, const float* __restrict__ low, const float* __restrict__ high,
                                 float* __restrict__ out, unsigned long long* __restrict__ counter, int64_t R, int D,
                                 int64_t cap) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = r < R;
  if (ok) for (int f = 0; f < D; ++f) { const float v = theta[r * D + f]; ok &= (v >= low[f]) && (v <= high[f]); }
  const unsigned m = __ballot_sync(0xffffffffu, ok);
  if (m == 0u) return;
  const int lane = threadIdx.x & 31;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(counter, (unsigned long long)__popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (!ok) return;
  const int64_t pos = (int64_t)base + __popc(m & ((1u << lane) - 1u));
  if (pos >= cap) return;
  for (int f = 0; f < D; ++f) out[pos * D + f] = theta[r * D + f];
}

// ---- launch wrappers used by the engine ----------------------------------------------------------------------
static inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

int launch_affine_fwd(const float* theta, const float* shift, const float* scale, float zlogdet, float* out, float* logdet,
                      int64_t R, int D, cudaStream_t s) {
  int64_t n = R * D > R ? R * D : R;
  ProfScope ps(TAG_OTHER, 0.0, s);
  k_affine_fwd<<<nblk(n, 256), 256, 0, s>>>(theta, shift, scale, zlogdet, out, logdet, R, D);
  FP_LAUNCH_CHECK(); return 0;
}
int launch_affine_inv(const float* y, const float* shift, const float* scale, float* out, int64_t R, int D, cudaStream_t s) {
  ProfScope ps(TAG_OTHER, 0.0, s);
  k_affine_inv<<<nblk(R * D, 256), 256, 0, s>>>(y, shift, scale, out, R, D);
  FP_LAUNCH_CHECK(); return 0;
}
int launch_affine_bwd(const float* d0, const float* scale, float* out, int64_t R, int D, cudaStream_t s) {
  ProfScope ps(TAG_OTHER, 0.0, s);
  k_affine_bwd<<<nblk(R * D, 256), 256, 0, s>>>(d0, scale, out, R, D);
  FP_LAUNCH_CHECK(); return 0;
}
int launch_base_logprob(const float* z, const float* logdet, float* logp, int64_t R, int D, int32_t* status, cudaStream_t s) {
  float log_z = (float)(0.5 * D * log(2.0 * M_PI));
  ProfScope ps(TAG_OTHER, 0.0, s);
  k_base_logprob<<<nblk(R, 256), 256, 0, s>>>(z, logdet, logp, R, D, log_z, status);
  FP_LAUNCH_CHECK(); return 0;
}
int launch_base_bwd(const float* z, const float* d_logp, float* dz, int64_t R, int D, cudaStream_t s) {
  ProfScope ps(TAG_OTHER, 0.0, s);
  k_base_bwd<<<nblk(R * D, 256), 256, 0, s>>>(z, d_logp, dz, R, D);
  FP_LAUNCH_CHECK(); return 0;
}
int launch_rows_affine(const float* x, const float* M, const float* pre, float sgn, const float* post, float* y, int64_t R,
                       int D, cudaStream_t s) {
  size_t sm = (size_t)(D * D + D) * 4;
  ProfScope ps(TAG_OTHER, 0.0, s);
  k_rows_affine<<<nblk(R * D, 256), 256, sm, s>>>(x, M, pre, sgn, post, y, R, D);
  FP_LAUNCH_CHECK(); return 0;
}
int launch_glu_bwd(const float* dh, const float* u, const float* G, int64_t ldg, float* du, float* dG, int64_t Rc, int A,
                   int H, int ldh, int sig_given, cudaStream_t s) {
  ProfScope ps(TAG_OTHER, 0.0, s);
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if ((H & 3) == 0 && (ldh & 3) == 0 && (ldg & 3) == 0 && al16(dh) && al16(u) && al16(G) && al16(du) && al16(dG)) {
    k_glu_bwd4<<<nblk(Rc * (H / 4), 256), 256, 0, s>>>(dh, u, G, ldg, du, dG, Rc, A, H, ldh, sig_given);
    FP_LAUNCH_CHECK(); return 0;
  }
  k_glu_bwd<<<nblk(Rc * H, 256), 256, 0, s>>>(dh, u, G, ldg, du, dG, Rc, A, H, ldh, sig_given);
  FP_LAUNCH_CHECK(); return 0;
}
// float4 path: H % 4 == 0 and a 16-byte aligned buffer; any other hidden width goes element by element
int launch_gate_sigmoid(float* ctx_all, int64_t Rc, int64_t Nctx, int H, int S, cudaStream_t s) {
  if ((H & 3) != 0 || (reinterpret_cast<uintptr_t>(ctx_all) & 15) != 0) {
    const int64_t n = Rc * Nctx;
    if (n <= 0) return 0;
    ProfScope ps(TAG_OTHER, 0.0, s);
    k_gate_sigmoid1<<<(unsigned)std::min<int64_t>((n + 255) / 256, 16 * (int64_t)num_sms()), 256, 0, s>>>(ctx_all, n, H, S);
    FP_LAUNCH_CHECK(); return 0;
  }
  const int64_t n4 = Rc * Nctx / 4;
  if (n4 <= 0) return 0;
  ProfScope ps(TAG_OTHER, 0.0, s);
  k_gate_sigmoid<<<(unsigned)std::min<int64_t>((n4 + 255) / 256, 16 * (int64_t)num_sms()), 256, 0, s>>>(ctx_all, n4, H, S);
  FP_LAUNCH_CHECK(); return 0;
}
// four columns per thread (16-byte accesses) when the rows allow it
__global__ void k_ctx_reduce4(const float* __restrict__ dh, float* __restrict__ dG, int64_t ldg, int64_t Rc, int A, int H4, int ldh) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < Rc * H4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t cr = i / H4; const int j = (int)(i - cr * H4) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int a = 0; a < A; ++a) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(dh + (cr * A + a) * ldh + j));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(dG + cr * ldg + j) = acc;
  }
}
int launch_ctx_reduce(const float* dh, float* dG, int64_t ldg, int64_t Rc, int A, int H, int ldh, cudaStream_t s) {
  ProfScope ps(TAG_OTHER, 0.0, s);
  if ((H & 3) == 0 && (ldh & 3) == 0 && (ldg & 3) == 0 && ((reinterpret_cast<uintptr_t>(dh) | reinterpret_cast<uintptr_t>(dG)) & 15) == 0) {
    k_ctx_reduce4<<<(int)std::min<int64_t>((Rc * (H / 4) + 255) / 256, 16 * num_sms()), 256, 0, s>>>(dh, dG, ldg, Rc, A, H / 4, ldh);
    FP_LAUNCH_CHECK(); return 0;
  }
  k_ctx_reduce<<<nblk(Rc * H, 256), 256, 0, s>>>(dh, dG, ldg, Rc, A, H, ldh);
  FP_LAUNCH_CHECK(); return 0;
}
int launch_colsum(const float* X, int64_t ldx, int64_t M, int N, float* out, cudaStream_t s) {
  if (M <= 0 || N <= 0) return 0;
  // float4 columns when every row segment is 16-byte aligned and N is a multiple of 4
  const bool vec = ((reinterpret_cast<uintptr_t>(X) & 15) == 0) && ((ldx & 3) == 0) && ((N & 3) == 0);
  const int cols = vec ? 128 : 32;
  const int bx = (N + cols - 1) / cols;
  // enough blocks for ~4 per SM, at least 64 rows each
  int rpb = 64;
  while ((M + rpb - 1) / rpb * bx > 8 * (int64_t)num_sms() && rpb < (1 << 20)) rpb *= 2;
  int64_t by = (M + rpb - 1) / rpb;
  while (by > 65535) { rpb *= 2; by = (M + rpb - 1) / rpb; }
  dim3 grid(bx, (unsigned)by);
  ProfScope ps(TAG_OTHER, 0.0, s);
  if (vec) k_colsum<4><<<grid, 256, 0, s>>>(X, ldx, M, N, out, rpb);
  else k_colsum<1><<<grid, 256, 0, s>>>(X, ldx, M, N, out, rpb);
  FP_LAUNCH_CHECK(); return 0;
}
int launch_sum(const float* x, int64_t n, float scale, float* out, cudaStream_t s) {
  int blocks = (int)std::min<int64_t>((n + 1023) / 1024, 4 * num_sms());
  if (blocks < 1) blocks = 1;
  ProfScope ps(TAG_OTHER, 0.0, s);
  k_sum<<<blocks, 256, 0, s>>>(x, n, scale, out);
  FP_LAUNCH_CHECK(); return 0;
}

// out[r, 0..Dp) = in[r, 0..D) followed by zeros (rows padded to a TMA-addressable stride)
__global__ void k_pad_rows(const float* __restrict__ in, int D, float* __restrict__ out, int Dp, int64_t R) {
  const int64_t n = R * Dp;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / Dp; const int c = (int)(e - r * Dp);
    out[e] = c < D ? in[r * D + c] : 0.f;
  }
}
int launch_pad_rows(const float* in, int D, float* out, int Dp, int64_t R, cudaStream_t s) {
  if (R <= 0) return 0;
  ProfScope ps(TAG_OTHER, 0.0, s);
  k_pad_rows<<<(int)std::min<int64_t>((R * Dp + 255) / 256, 8 * num_sms()), 256, 0, s>>>(in, D, out, Dp, R);
  FP_LAUNCH_CHECK(); return 0;
}

// Power-of-two scale for the gradient operands of the 3xFP16 GEMMs: s = 2^-e with 2^(e-1) <= max|d_logp| < 2^e, so the
// upstream gradient enters the backward pass with its largest element in [0.5, 1) (every later gradient tensor is linear
// in it); s = 1 for an all-zero or non-finite upstream gradient.  One block.
__global__ void k_grad_scale(const float* __restrict__ d_logp, int64_t n, float* __restrict__ out) {
  __shared__ float red[32];
  float m = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(d_logp[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
    float sc = 1.f;
    if (m > 0.f && m < 3.0e38f) {
      int e;
      frexpf(m, &e);                 // m = f * 2^e, f in [0.5, 1)
      e = max(-100, min(100, e));
      sc = ldexpf(1.f, -e);
    }
    out[0] = sc;
  }
}
int launch_grad_scale(const float* d_logp, int64_t n, float* out, cudaStream_t s) {
  ProfScope ps(TAG_OTHER, 0.0, s);
  k_grad_scale<<<1, 1024, 0, s>>>(d_logp, n, out);
  FP_LAUNCH_CHECK(); return 0;
}

// keep flags of the dropout mask, for the parity tests (the oracle is handed the very mask the kernels drew)
__global__ void k_dropout_mask(uint8_t* __restrict__ keep, int64_t R, int width, DropKey dk) {
  const int64_t n = R * (int64_t)width;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    keep[e] = drop_keep(dk, e / width, (int)(e % width), width) ? 1 : 0;
}

}  // namespace fp

using namespace fp;

extern "C" {

int fp_dropout_mask(uint8_t* keep, int64_t R, int32_t width, float p, uint64_t seed, void* stream) {
  if (R <= 0) return 0;
  if (width < 1 || p < 0.f || p >= 1.f || !keep) return FP_E_BADARG;
  ProfScope ps(TAG_OTHER, 0.0, (cudaStream_t)stream);
  k_dropout_mask<<<(int)std::min<int64_t>((R * width + 255) / 256, 4 * num_sms()), 256, 0, (cudaStream_t)stream>>>(
      keep, R, width, make_drop_key(seed, p));
  FP_LAUNCH_CHECK(); return 0;
}

int fp_atom_choices(int32_t* choices, int64_t B, int32_t A, uint64_t seed, void* stream) {
  if (A < 2 || A > B || A > 4096) return FP_E_BADARG;
  ProfScope ps(TAG_OTHER, 0.0, (cudaStream_t)stream);
  k_atom_choices<<<nblk(B, 128), 128, 0, (cudaStream_t)stream>>>(choices, B, A, seed, nullptr);
  FP_LAUNCH_CHECK(); return 0;
}
int fp_atom_choices_step(int32_t* choices, int64_t B, int32_t A, uint64_t seed, const int32_t* step_counter, void* stream) {
  if (A < 2 || A > B || A > 4096 || !step_counter) return FP_E_BADARG;
  ProfScope ps(TAG_OTHER, 0.0, (cudaStream_t)stream);
  k_atom_choices<<<nblk(B, 128), 128, 0, (cudaStream_t)stream>>>(choices, B, A, seed, step_counter);
  FP_LAUNCH_CHECK(); return 0;
}
int fp_atom_gather(const float* theta, const int32_t* choices, float* rows, int64_t B, int32_t A, int32_t D, void* stream) {
  ProfScope ps(TAG_OTHER, 0.0, (cudaStream_t)stream);
  k_atom_gather<<<nblk(B * A * D, 256), 256, 0, (cudaStream_t)stream>>>(theta, choices, rows, B, A, D);
  FP_LAUNCH_CHECK(); return 0;
}
int fp_atomic_loss(const float* logp, const float* atom_theta, const float* low, const float* high, const float* masks,
                   int32_t combined, float* out_lp, float* d_logp, int64_t B, int32_t A, int32_t D, int32_t* status,
                   void* stream) {
  ProfScope ps(TAG_OTHER, 0.0, (cudaStream_t)stream);
  k_atomic_loss<<<nblk(B, 128), 128, 0, (cudaStream_t)stream>>>(logp, atom_theta, low, high, masks, combined, out_lp,
                                                                d_logp, B, A, D, status);
  FP_LAUNCH_CHECK(); return 0;
}
int fp_grad_sumsq(const float* grads, int64_t n, float* sumsq, void* stream) {
  int blocks = (int)std::min<int64_t>((n + 1023) / 1024, 4 * num_sms());
  if (blocks < 1) blocks = 1;
  ProfScope ps(TAG_OTHER, 0.0, (cudaStream_t)stream);
  k_sumsq<<<blocks, 256, 0, (cudaStream_t)stream>>>(grads, n, sumsq);
  FP_LAUNCH_CHECK(); return 0;
}
int fp_adam_step(float* params, const float* grads, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                 float eps, const int32_t* step_dev, float max_norm, float grad_scale, const float* sumsq, void* stream) {
  int blocks = (int)std::min<int64_t>((n / 4 + 255) / 256 + 1, 8 * num_sms());
  if (blocks < 1) blocks = 1;
  ProfScope ps(TAG_OTHER, 0.0, (cudaStream_t)stream);
  k_adam<<<blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, m, v, n, lr, beta1, beta2, eps, step_dev, max_norm,
                                                    grad_scale, sumsq);
  FP_LAUNCH_CHECK(); return 0;
}
int fp_gather_rows(const float* src, const int64_t* idx, float* dst, int64_t n, int32_t width, void* stream) {
  if (n <= 0) return 0;
  ProfScope ps(TAG_OTHER, 0.0, (cudaStream_t)stream);
  k_gather_rows<<<nblk(n * width, 256), 256, 0, (cudaStream_t)stream>>>(src, idx, dst, n, width);
  FP_LAUNCH_CHECK(); return 0;
}
int fp_standardize(const float* x, const float* mean, const float* std, float* out, int64_t n, int32_t width, void* stream) {
  if (n <= 0) return 0;
  ProfScope ps(TAG_OTHER, 0.0, (cudaStream_t)stream);
  k_standardize<<<nblk(n * width, 256), 256, 0, (cudaStream_t)stream>>>(x, mean, std, out, n, width);
  FP_LAUNCH_CHECK(); return 0;
}
int fp_sum(const float* x, int64_t n, float scale, float* out, void* stream) {
  return launch_sum(x, n, scale, out, (cudaStream_t)stream);
}
int fp_within_box(const float* theta, const float* low, const float* high, uint8_t* inside, int64_t R, int32_t D, void* stream) {
  if (R <= 0) return 0;
  ProfScope ps(TAG_OTHER, 0.0, (cudaStream_t)stream);
  k_within_box<<<nblk(R, 256), 256, 0, (cudaStream_t)stream>>>(theta, low, high, inside, R, D);
  FP_LAUNCH_CHECK(); return 0;
}
int fp_compact_inside(const float* theta, const float* low, const float* high, float* out, uint64_t* counter, int64_t R,
                      int32_t D, int64_t cap, void* stream) {
  if (R <= 0) return 0;
  if (D < 1 || cap < 0 || !counter) return FP_E_BADARG;
  ProfScope ps(TAG_OTHER, 0.0, (cudaStream_t)stream);
  k_compact_inside<<<nblk(R, 256), 256, 0, (cudaStream_t)stream>>>(theta, low, high, out,
                                                                   reinterpret_cast<unsigned long long*>(counter), R, D, cap);
  FP_LAUNCH_CHECK(); return 0;
}
const char* fp_version(void) { return "floppity_b200 0.1 (sm_100a)"; }

}  // extern "C"
