// gemm_simt.cu -- fp32 SIMT GEMM with the ResidualNet epilogues fused in.
//
// The conditioner of every coupling layer (nflows ResidualNet/ResidualBlock, built by sbi build_nsf,
// reached from /root/reference/modules.py:63) is a chain of small dense layers.  Upstream runs each as
// addmm + separate relu / dropout / cat / glu / add kernels; here each dense layer is ONE launch whose
// prologue applies relu to the activation operand and whose epilogue applies bias (+ dropout on the result), the
// context term, the GLU gate + residual, or the relu mask of the backward pass.  The same kernel serves
// forward (NT), dgrad (NN) and wgrad (TN, split-K with atomics) through the two "k-major" switches.
// This is the exact-fp32 path; the tcgen05 path (gemm_tc.cu) takes over the large context projection.
#include "common.cuh"
#include "prof.h"
#include "gemm_epilogue.cuh"
#include <algorithm>

namespace fp {

constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4;
constexpr int GEMM_THREADS = (BM / TM) * (BN / TN);  // 256
constexpr int APAD = 4, BPAD = 4;

template <bool A_KMAJ, bool B_KMAJ>
__global__ void __launch_bounds__(GEMM_THREADS) k_gemm(fp_gemm_args g, int64_t k_chunk) {
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN + BPAD];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN);  // 0..15
  const int ty = tid / (BN / TN);  // 0..15
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * k_chunk;
  const int64_t kend = min(g.K, kbeg + k_chunk);
  if (kbeg >= kend) return;
  const int nk = (int)((kend - kbeg + BK - 1) / BK);

  const bool relu_a = g.flags & FP_G_RELU_A, relu_b = g.flags & FP_G_RELU_B;
  const bool a_vec = ((g.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0);
  const bool b_vec = ((g.ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0);

  float ra[2][4];
  float rb[4];

  auto load_a = [&](int kt) {
    const int64_t k0 = kbeg + (int64_t)kt * BK;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int idx = tid + i * GEMM_THREADS;
      if (A_KMAJ) {
        int row = idx >> 2, kq = (idx & 3) * 4;
        int64_t m = m0 + row, k = k0 + kq;
        const float* src = g.A + m * g.lda + k;
        if (m < g.M && k + 3 < kend && a_vec && ((k & 3) == 0)) {
          float4 v = *reinterpret_cast<const float4*>(src);
          ra[i][0] = v.x; ra[i][1] = v.y; ra[i][2] = v.z; ra[i][3] = v.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) ra[i][j] = (m < g.M && k + j < kend) ? src[j] : 0.f;
        }
      } else {
        int krow = idx >> 5, mq = (idx & 31) * 4;
        int64_t k = k0 + krow, m = m0 + mq;
        const float* src = g.A + k * g.lda + m;
        if (k < kend && m + 3 < g.M && a_vec && ((m & 3) == 0)) {
          float4 v = *reinterpret_cast<const float4*>(src);
          ra[i][0] = v.x; ra[i][1] = v.y; ra[i][2] = v.z; ra[i][3] = v.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) ra[i][j] = (k < kend && m + j < g.M) ? src[j] : 0.f;
        }
      }
      if (relu_a) {
#pragma unroll
        for (int j = 0; j < 4; ++j) ra[i][j] = fmaxf(ra[i][j], 0.f);
      }
    }
  };
  auto load_b = [&](int kt) {
    const int64_t k0 = kbeg + (int64_t)kt * BK;
    if (B_KMAJ) {
      int row = tid >> 2, kq = (tid & 3) * 4;
      int n = n0 + row;
      int64_t k = k0 + kq;
      const float* src = g.B + (int64_t)n * g.ldb + k;
      if (n < g.N && k + 3 < kend && b_vec && ((k & 3) == 0)) {
        float4 v = *reinterpret_cast<const float4*>(src);
        rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) rb[j] = (n < g.N && k + j < kend) ? src[j] : 0.f;
      }
    } else {
      int krow = tid >> 4, nq = (tid & 15) * 4;
      int64_t k = k0 + krow;
      int n = n0 + nq;
      const float* src = g.B + k * g.ldb + n;
      if (k < kend && n + 3 < g.N && b_vec && ((n & 3) == 0)) {
        float4 v = *reinterpret_cast<const float4*>(src);
        rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) rb[j] = (k < kend && n + j < g.N) ? src[j] : 0.f;
      }
    }
    if (relu_b) {
#pragma unroll
      for (int j = 0; j < 4; ++j) rb[j] = fmaxf(rb[j], 0.f);
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int idx = tid + i * GEMM_THREADS;
      if (A_KMAJ) {
        int row = idx >> 2, kq = (idx & 3) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) As[buf][kq + j][row] = ra[i][j];
      } else {
        int krow = idx >> 5, mq = (idx & 31) * 4;
        *reinterpret_cast<float4*>(&As[buf][krow][mq]) = make_float4(ra[i][0], ra[i][1], ra[i][2], ra[i][3]);
      }
    }
    if (B_KMAJ) {
      int row = tid >> 2, kq = (tid & 3) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) Bs[buf][kq + j][row] = rb[j];
    } else {
      int krow = tid >> 4, nq = (tid & 15) * 4;
      *reinterpret_cast<float4*>(&Bs[buf][krow][nq]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  load_a(0); load_b(0);
  store_smem(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) { load_a(kt + 1); load_b(kt + 1); }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM + 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * TN]);
      float av[TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[TN] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_smem(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int64_t m = m0 + ty * TM + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n >= g.N) continue;
      gemm_epilogue_store(g, g.flags, m, n, acc[i][j]);
    }
  }
}

int launch_gemm(const fp_gemm_args* a, cudaStream_t stream) {
  if (a->M <= 0 || a->N <= 0) return 0;
  if (a->K <= 0) return FP_E_BADARG;
  if (a->M > 2147483647LL && (a->flags & (FP_G_ADD_CTX | FP_G_GLU_RES))) return FP_E_TOOLARGE;
  int split = a->split_k > 0 ? a->split_k : 1;
  if (split > 1 && !(a->flags & FP_G_ATOMIC)) return FP_E_BADARG;
  int64_t kc = (a->K + split - 1) / split;
  kc = (kc + BK - 1) / BK * BK;
  split = (int)((a->K + kc - 1) / kc);
  int64_t mt = (a->M + BM - 1) / BM;
  int nt = (a->N + BN - 1) / BN;
  if (nt > 65535 || split > 65535 || mt > 2147483647LL) return FP_E_TOOLARGE;
  dim3 grid((unsigned)mt, (unsigned)nt, (unsigned)split);
  ProfScope ps(TAG_GEMM_SIMT, 2.0 * (double)a->M * (double)a->N * (double)a->K, stream, gemm_algo_bytes(a));
  if (a->a_kmajor && a->b_kmajor) k_gemm<true, true><<<grid, GEMM_THREADS, 0, stream>>>(*a, kc);
  else if (a->a_kmajor && !a->b_kmajor) k_gemm<true, false><<<grid, GEMM_THREADS, 0, stream>>>(*a, kc);
  else if (!a->a_kmajor && a->b_kmajor) k_gemm<false, true><<<grid, GEMM_THREADS, 0, stream>>>(*a, kc);
  else k_gemm<false, false><<<grid, GEMM_THREADS, 0, stream>>>(*a, kc);
  FP_LAUNCH_CHECK();
  return 0;
}

}  // namespace fp

extern "C" int fp_gemm(const fp_gemm_args* args, void* stream) {
  if (!args) return FP_E_BADARG;
  return fp::launch_gemm(args, (cudaStream_t)stream);
}
