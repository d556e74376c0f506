// cond_small.cu -- the ResidualNet conditioner of one coupling layer as a shared-memory-resident fused MLP, for the
// small hidden widths the reference defaults to (-hidden 50, /root/reference/floppity.py:34; BASELINE.json configs[0]).
//
// Stands for nflows ResidualNet.forward / its autograd (initial layer, Bk x ResidualBlock with context GLU, final
// layer), reached from /root/reference/modules.py:63 via sbi build_nsf.  At hidden <= 64 a layer's weights are a few
// tens of KB: they are not a "real dense contraction" worth tcgen05 tiles fed by TMA (200-byte rows are not even
// TMA-addressable), and as separate GEMM launches the chain was 86 fp32 SIMT launches per c1 step at 10 % of HBM
// (profiles/r1j_bench_c1.json).  Here:
//   * a persistent CTA copies the layer's weights ONCE into shared memory (fp16 hi + 2^11-scaled fp16 residual images
//     prepared by fp_nsf_prepare, rows padded to an odd multiple of 16 bytes so ldmatrix is bank-conflict-free);
//   * every warp takes 16 flow rows through the WHOLE chain with the activations in registers: the accumulator
//     fragment of one warp-level mma.sync.m16n8k16 is, element for element, the A fragment of the next layer
//     (two adjacent 8-column accumulator blocks = one 16-deep k block), so nothing is staged between layers;
//   * products are fp32-accurate: a*b ~= a_hi*b_hi + (a_lo*b_hi + a_hi*b_lo) with the cross terms in their own
//     accumulator (3xFP16, the same split the tcgen05 conditioner uses), and every activation row is scaled by a
//     power of two so its largest element sits in [1, 2) -- fp16's range can then neither overflow nor flush the
//     significant elements, for activations and gradients alike;
//   * HBM sees theta in and spline parameters out (+ the saved activations when training); the backward kernel carries
//     dh in registers through GLU-backward -> W2^T -> relu/dropout mask -> W1^T and leaves du / dt / dh0 for the
//     weight-gradient GEMMs, context gradients go out as reductions over the atoms sharing a context row.
#include "common.cuh"
#include "prof.h"
#include "layout.h"
#include "tc_common.cuh"
#include "gemm_epilogue.cuh"
#include <algorithm>

namespace fp {

bool smallh_enabled();                                        // flow_engine.cu (fp_set_mode)
uint64_t layer_seed(uint64_t seed, int layer, int block);

constexpr int CS_WARPS = 8;
constexpr int CS_THREADS = CS_WARPS * 32;
constexpr int CS_MAXBK = 4;
constexpr int CS_HP = 64;                // padded hidden width (8 accumulator blocks of 8 columns)
constexpr int CS_SH = CS_HP + 8;         // row stride (halves) of the [*, 64]-contraction matrices
constexpr int CS_SMEM_LIMIT = 220 * 1024;

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}

// power-of-two scale that brings a row maximum into [1, 2), and its inverse (exact; 0 -> result forced to 0)
__device__ __forceinline__ void row_scale(float amax, float& s, float& un) {
  uint32_t e = (__float_as_uint(amax) >> 23) & 0xffu;
  e = e > 254u ? 254u : e;
  s = __uint_as_float((254u - e) << 23);
  un = __uint_as_float(e << 23);
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}

// acc (+)= A[16 x 16*kbn] * W^T for n blocks nb0 .. nb0+nbn-1 of the matrix whose hi image starts at shared address `w`
// (rows = output unit, `stride` bytes apart, contraction contiguous; the lo image sits `lo_off` bytes further).  One
// ldmatrix.x4 fetches {hi, lo} x {k 0..7, k 8..15} of one 8-row block: three MMAs per fetch.
template <int KB, int NB>
__device__ __forceinline__ void tile_gemm(float (&hh)[NB][4], float (&xx)[NB][4], const uint32_t (&ahi)[KB][4],
                                          const uint32_t (&alo)[KB][4], int kbn, int kb0, int nbn, uint32_t w, uint32_t lo_off,
                                          uint32_t stride, int lane) {
  const uint32_t lane_off = (uint32_t)(lane & 7) * stride + ((lane >> 4) ? lo_off : 0u) + (((lane >> 3) & 1) ? 16u : 0u);
#pragma unroll
  for (int kb = 0; kb < KB; ++kb) {
    if (kb < kbn) {
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        if (nb < nbn) {
          uint32_t bh0, bh1, bl0, bl1;
          ldsm4(w + lane_off + (uint32_t)nb * 8u * stride + (uint32_t)(kb0 + kb) * 32u, bh0, bh1, bl0, bl1);
          mma16816(hh[nb], ahi[kb], bh0, bh1);
          mma16816(xx[nb], alo[kb], bh0, bh1);
          mma16816(xx[nb], ahi[kb], bl0, bl1);
        }
      }
    }
  }
}

template <int NB>
__device__ __forceinline__ void zero_acc(float (&hh)[NB][4], float (&xx)[NB][4]) {
#pragma unroll
  for (int nb = 0; nb < NB; ++nb)
#pragma unroll
    for (int j = 0; j < 4; ++j) { hh[nb][j] = 0.f; xx[nb][j] = 0.f; }
}

// accumulator-layout values (thread holds rows g, g+8 x columns 8 nb + 2t, +1) -> the A fragments of the next GEMM,
// row-scaled; un0 / un1 undo the scaling of rows g / g+8 on that GEMM's result
template <int NB>
__device__ __forceinline__ void make_afrags(const float (&x)[NB][4], uint32_t (&ahi)[NB / 2][4], uint32_t (&alo)[NB / 2][4],
                                            float& un0, float& un1) {
  float m0 = 0.f, m1 = 0.f;
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    m0 = fmaxf(m0, fmaxf(fabsf(x[nb][0]), fabsf(x[nb][1])));
    m1 = fmaxf(m1, fmaxf(fabsf(x[nb][2]), fabsf(x[nb][3])));
  }
  float s0, s1;
  row_scale(quad_max(m0), s0, un0);
  row_scale(quad_max(m1), s1, un1);
#pragma unroll
  for (int kb = 0; kb < NB / 2; ++kb) {
    split_f16x2(x[2 * kb][0] * s0, x[2 * kb][1] * s0, ahi[kb][0], alo[kb][0]);
    split_f16x2(x[2 * kb][2] * s1, x[2 * kb][3] * s1, ahi[kb][1], alo[kb][1]);
    split_f16x2(x[2 * kb + 1][0] * s0, x[2 * kb + 1][1] * s0, ahi[kb][2], alo[kb][2]);
    split_f16x2(x[2 * kb + 1][2] * s1, x[2 * kb + 1][3] * s1, ahi[kb][3], alo[kb][3]);
  }
}

// one 16-deep k block of A fragments straight from a row-major global matrix (rows r0 / r1 of this thread, columns
// c0 + {2t, 2t+1, 2t+8, 2t+9}), zero beyond `ncols`; the values are multiplied by the row scales before the split
__device__ __forceinline__ void load_a8(const float* __restrict__ p0, const float* __restrict__ p1, int c, int ncols, float (&v)[8]) {
  v[0] = c < ncols ? __ldg(p0 + c) : 0.f;         v[1] = c + 1 < ncols ? __ldg(p0 + c + 1) : 0.f;
  v[2] = c < ncols ? __ldg(p1 + c) : 0.f;         v[3] = c + 1 < ncols ? __ldg(p1 + c + 1) : 0.f;
  v[4] = c + 8 < ncols ? __ldg(p0 + c + 8) : 0.f; v[5] = c + 9 < ncols ? __ldg(p0 + c + 9) : 0.f;
  v[6] = c + 8 < ncols ? __ldg(p1 + c + 8) : 0.f; v[7] = c + 9 < ncols ? __ldg(p1 + c + 9) : 0.f;
}
__device__ __forceinline__ void split_a8(const float (&v)[8], float s0, float s1, uint32_t (&ahi)[4], uint32_t (&alo)[4]) {
  split_f16x2(v[0] * s0, v[1] * s0, ahi[0], alo[0]);
  split_f16x2(v[2] * s1, v[3] * s1, ahi[1], alo[1]);
  split_f16x2(v[4] * s0, v[5] * s0, ahi[2], alo[2]);
  split_f16x2(v[6] * s1, v[7] * s1, ahi[3], alo[3]);
}

// two consecutive columns of one row: 8-byte access when the pair is whole and aligned (even column, even stride)
__device__ __forceinline__ void st_pair(float* row, int col, int ncols, float a, float b) {
  if (col + 1 < ncols) *reinterpret_cast<float2*>(row + col) = make_float2(a, b);
  else if (col < ncols) row[col] = a;
}
__device__ __forceinline__ float2 ld_pair(const float* row, int col, int ncols) {
  if (col + 1 < ncols) return *reinterpret_cast<const float2*>(row + col);
  return make_float2(col < ncols ? row[col] : 0.f, 0.f);
}
__device__ __forceinline__ float2 ldg_pair_unaligned(const float* row, int col, int ncols) {
  return make_float2(col < ncols ? __ldg(row + col) : 0.f, col + 1 < ncols ? __ldg(row + col + 1) : 0.f);
}

// The thread's 32 values (rows q0 / q1 x columns 8 nb + 2t, +1) of a row-major operand, REQUESTED EARLY: the kernels were
// waiting on global loads for ~60 % of their stall samples (ncu source page: the first consumer of every operand loaded
// at its point of use), so each phase's operand is fetched into this buffer before the GEMM that precedes its use
__device__ __forceinline__ void preload_rows(float (&b)[8][4], const float* __restrict__ r0, const float* __restrict__ r1,
                                             int ncols, int t) {
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) {
    const int col = nb * 8 + 2 * t;
    const float2 a = ldg_pair_unaligned(r0, col, ncols), c = ldg_pair_unaligned(r1, col, ncols);
    b[nb][0] = a.x; b[nb][1] = a.y; b[nb][2] = c.x; b[nb][3] = c.y;
  }
}

struct CsFwdParams {
  const uint4* img; int32_t img_vec;        // forward image (hi | lo), 16-byte vectors
  int32_t fwd_halves;                       // halves of ONE image (lo image offset)
  const float* theta; int32_t D, K0;
  const float* ctx; int64_t ldg; int32_t ctx_div;     // ctx_all + first column of this layer's slot 0; slot k at + H k
  const float* b1[CS_MAXBK]; const float* b2[CS_MAXBK]; const float* bf;
  float* h[CS_MAXBK + 1]; float* t[CS_MAXBK]; float* u[CS_MAXBK]; int32_t ldh;   // saved activations (training)
  float* prm; int32_t ldp, dP, NF;
  int64_t R; int32_t H, Bk, training;
  DropKey dk[CS_MAXBK];
};

template <bool DROP>
__global__ void __launch_bounds__(CS_THREADS, 1) k_cond_small_fwd(const CsFwdParams p) {
  extern __shared__ __align__(16) uint8_t cs_smem[];
  const uint32_t sm = smem_u32(cs_smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int H = p.H, Bk = p.Bk;
  // ---- the layer's weights and biases into shared memory, once per CTA
  {
    uint4* dst = reinterpret_cast<uint4*>(cs_smem);
    for (int i = threadIdx.x; i < p.img_vec; i += CS_THREADS) dst[i] = __ldg(p.img + i);
    float* sb = reinterpret_cast<float*>(cs_smem + (size_t)p.img_vec * 16);
    for (int i = threadIdx.x; i < Bk * 2 * CS_HP; i += CS_THREADS) {
      const int k = i / (2 * CS_HP), j = i % (2 * CS_HP), c = j % CS_HP;
      sb[i] = c < H ? (j < CS_HP ? p.b1[k][c] : p.b2[k][c]) : 0.f;
    }
    for (int i = threadIdx.x; i < p.NF; i += CS_THREADS) sb[Bk * 2 * CS_HP + i] = i < p.dP ? p.bf[i] : 0.f;
  }
  __syncthreads();
  const float* sb = reinterpret_cast<const float*>(cs_smem + (size_t)p.img_vec * 16);
  const uint32_t lo_off = (uint32_t)p.fwd_halves * 2u;
  const uint32_t s0b = (uint32_t)(p.K0 + 8) * 2u, shb = (uint32_t)CS_SH * 2u;
  const uint32_t w0_a = sm, wblk_a = sm + (uint32_t)CS_HP * s0b, wf_a = wblk_a + (uint32_t)(2 * Bk) * CS_HP * shb;
  const int kb0n = p.K0 >> 4;
  const int nfb = p.NF >> 3;
  const uint32_t cdiv = p.ctx_div <= 0 ? 0xFFFFFFFFu : (uint32_t)p.ctx_div;
  const bool training = p.training != 0;
  const int64_t n_tiles = (p.R + 15) >> 4;

  for (int64_t tile = (int64_t)blockIdx.x * CS_WARPS + warp; tile < n_tiles; tile += (int64_t)gridDim.x * CS_WARPS) {
    const int64_t r0 = tile * 16 + g, r1 = r0 + 8;
    const bool v0 = r0 < p.R, v1 = r1 < p.R;
    const int64_t q0 = v0 ? r0 : p.R - 1, q1 = v1 ? r1 : p.R - 1;
    const float* c0 = p.ctx + (int64_t)((uint32_t)q0 / cdiv) * p.ldg;
    const float* c1 = p.ctx + (int64_t)((uint32_t)q1 / cdiv) * p.ldg;
    float hh[8][4], xx[8][4], hres[8][4];
    uint32_t ahi[4][4], alo[4][4];
    float un0, un1;

    {  // ---- initial layer: h0 = theta W0x^T + ctx slot 0
      const float* t0 = p.theta + q0 * p.D;
      const float* t1 = p.theta + q1 * p.D;
      float tv[4][8];
      float m0 = 0.f, m1 = 0.f;
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        if (kb < kb0n) {
          load_a8(t0, t1, kb * 16 + 2 * t, p.D, tv[kb]);
          m0 = fmaxf(m0, fmaxf(fmaxf(fabsf(tv[kb][0]), fabsf(tv[kb][1])), fmaxf(fabsf(tv[kb][4]), fabsf(tv[kb][5]))));
          m1 = fmaxf(m1, fmaxf(fmaxf(fabsf(tv[kb][2]), fabsf(tv[kb][3])), fmaxf(fabsf(tv[kb][6]), fabsf(tv[kb][7]))));
        }
      }
      float s0, s1;
      row_scale(quad_max(m0), s0, un0);
      row_scale(quad_max(m1), s1, un1);
#pragma unroll
      for (int kb = 0; kb < 4; ++kb)
        if (kb < kb0n) split_a8(tv[kb], s0, s1, ahi[kb], alo[kb]);
      float cb[8][4];
      preload_rows(cb, c0, c1, H, t);                  // context slot 0, requested before the GEMM
      zero_acc(hh, xx);
      tile_gemm<4, 8>(hh, xx, ahi, alo, kb0n, 0, 8, w0_a, lo_off, s0b, lane);
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const int col = nb * 8 + 2 * t;
        hres[nb][0] = un0 * (hh[nb][0] + FP16_LO_UNSCALE * xx[nb][0]) + cb[nb][0];
        hres[nb][1] = un0 * (hh[nb][1] + FP16_LO_UNSCALE * xx[nb][1]) + cb[nb][1];
        hres[nb][2] = un1 * (hh[nb][2] + FP16_LO_UNSCALE * xx[nb][2]) + cb[nb][2];
        hres[nb][3] = un1 * (hh[nb][3] + FP16_LO_UNSCALE * xx[nb][3]) + cb[nb][3];
        if (training) {
          if (v0) st_pair(p.h[0] + r0 * p.ldh, col, H, hres[nb][0], hres[nb][1]);
          if (v1) st_pair(p.h[0] + r1 * p.ldh, col, H, hres[nb][2], hres[nb][3]);
        }
      }
    }

    for (int k = 0; k < Bk; ++k) {
      float x[8][4];
      // ---- t' = dropout(relu(h) W1^T + b1)
#pragma unroll
      for (int nb = 0; nb < 8; ++nb)
#pragma unroll
        for (int j = 0; j < 4; ++j) x[nb][j] = fmaxf(hres[nb][j], 0.f);
      make_afrags<8>(x, ahi, alo, un0, un1);
      zero_acc(hh, xx);
      tile_gemm<4, 8>(hh, xx, ahi, alo, 4, 0, 8, wblk_a + (uint32_t)(2 * k) * CS_HP * shb, lo_off, shb, lane);
      const float* bb1 = sb + k * 2 * CS_HP;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const int col = nb * 8 + 2 * t;
        float v[4];
        v[0] = un0 * (hh[nb][0] + FP16_LO_UNSCALE * xx[nb][0]) + bb1[col];
        v[1] = un0 * (hh[nb][1] + FP16_LO_UNSCALE * xx[nb][1]) + bb1[col + 1];
        v[2] = un1 * (hh[nb][2] + FP16_LO_UNSCALE * xx[nb][2]) + bb1[col];
        v[3] = un1 * (hh[nb][3] + FP16_LO_UNSCALE * xx[nb][3]) + bb1[col + 1];
        if (DROP) {
          const float ds = p.dk[k].scale;
          v[0] = (col < H && drop_keep(p.dk[k], r0, col, H)) ? v[0] * ds : 0.f;
          v[1] = (col + 1 < H && drop_keep(p.dk[k], r0, col + 1, H)) ? v[1] * ds : 0.f;
          v[2] = (col < H && drop_keep(p.dk[k], r1, col, H)) ? v[2] * ds : 0.f;
          v[3] = (col + 1 < H && drop_keep(p.dk[k], r1, col + 1, H)) ? v[3] * ds : 0.f;
        }
        if (training) {
          if (v0) st_pair(p.t[k] + r0 * p.ldh, col, H, v[0], v[1]);
          if (v1) st_pair(p.t[k] + r1 * p.ldh, col, H, v[2], v[3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) x[nb][j] = fmaxf(v[j], 0.f);
      }
      // ---- u = relu(t') W2^T + b2 ; h += u * sigmoid(gate)   (slot 1+k of ctx_all holds sigmoid(gate): k_gate_sigmoid)
      make_afrags<8>(x, ahi, alo, un0, un1);
      float gb[8][4];
      preload_rows(gb, c0 + (int64_t)H * (1 + k), c1 + (int64_t)H * (1 + k), H, t);      // sigmoid(gate), before the GEMM
      zero_acc(hh, xx);
      tile_gemm<4, 8>(hh, xx, ahi, alo, 4, 0, 8, wblk_a + (uint32_t)(2 * k + 1) * CS_HP * shb, lo_off, shb, lane);
      const float* bb2 = bb1 + CS_HP;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const int col = nb * 8 + 2 * t;
        const float2 a = make_float2(gb[nb][0], gb[nb][1]), b = make_float2(gb[nb][2], gb[nb][3]);
        float v[4];
        v[0] = un0 * (hh[nb][0] + FP16_LO_UNSCALE * xx[nb][0]) + bb2[col];
        v[1] = un0 * (hh[nb][1] + FP16_LO_UNSCALE * xx[nb][1]) + bb2[col + 1];
        v[2] = un1 * (hh[nb][2] + FP16_LO_UNSCALE * xx[nb][2]) + bb2[col];
        v[3] = un1 * (hh[nb][3] + FP16_LO_UNSCALE * xx[nb][3]) + bb2[col + 1];
        hres[nb][0] = fmaf(v[0], a.x, hres[nb][0]); hres[nb][1] = fmaf(v[1], a.y, hres[nb][1]);
        hres[nb][2] = fmaf(v[2], b.x, hres[nb][2]); hres[nb][3] = fmaf(v[3], b.y, hres[nb][3]);
        if (training) {
          if (v0) { st_pair(p.u[k] + r0 * p.ldh, col, H, v[0], v[1]); st_pair(p.h[k + 1] + r0 * p.ldh, col, H, hres[nb][0], hres[nb][1]); }
          if (v1) { st_pair(p.u[k] + r1 * p.ldh, col, H, v[2], v[3]); st_pair(p.h[k + 1] + r1 * p.ldh, col, H, hres[nb][2], hres[nb][3]); }
        }
      }
    }

    // ---- final layer: params = h Wf^T + bf, 64 output columns at a time
    make_afrags<8>(hres, ahi, alo, un0, un1);
    const float* sbf = sb + Bk * 2 * CS_HP;
    for (int c = 0; c * 8 < nfb; ++c) {
      zero_acc(hh, xx);
      const int nbn = min(8, nfb - c * 8);
      tile_gemm<4, 8>(hh, xx, ahi, alo, 4, 0, nbn, wf_a + (uint32_t)(c * 64) * shb, lo_off, shb, lane);
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        if (nb < nbn) {
          const int col = c * 64 + nb * 8 + 2 * t;
          const float ba = sbf[col], bbv = sbf[col + 1];
          if (v0) st_pair(p.prm + r0 * p.ldp, col, p.dP, un0 * (hh[nb][0] + FP16_LO_UNSCALE * xx[nb][0]) + ba,
                          un0 * (hh[nb][1] + FP16_LO_UNSCALE * xx[nb][1]) + bbv);
          if (v1) st_pair(p.prm + r1 * p.ldp, col, p.dP, un1 * (hh[nb][2] + FP16_LO_UNSCALE * xx[nb][2]) + ba,
                          un1 * (hh[nb][3] + FP16_LO_UNSCALE * xx[nb][3]) + bbv);
        }
      }
    }
  }
}

struct CsBwdParams {
  const uint4* img; int32_t img_vec;        // backward image (hi | lo)
  int32_t bwd_halves;
  const float* dprm; int32_t ldp, dP, KF;
  const float* h[CS_MAXBK + 1]; const float* t[CS_MAXBK]; const float* u[CS_MAXBK]; int32_t ldh;
  const float* ctx; int64_t ldg; int32_t ctx_div;     // sigmoid(gate) of block k at ctx + H (1+k)
  float* du[CS_MAXBK]; float* dt[CS_MAXBK]; float* dh0;    // gradient operands of the weight-gradient GEMMs (stride ldh)
  float* dctx;                                        // dctx_all + first column of this layer's slot 0 (stride ldg), += by reductions
  float* dth; int32_t D, D8;                          // d theta_i [R, D] +=
  float mask_scale;
  int64_t R; int32_t H, Bk;
};

__global__ void __launch_bounds__(CS_THREADS, 1) k_cond_small_bwd(const CsBwdParams p) {
  extern __shared__ __align__(16) uint8_t cs_smem[];
  const uint32_t sm = smem_u32(cs_smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int H = p.H, Bk = p.Bk;
  {
    uint4* dst = reinterpret_cast<uint4*>(cs_smem);
    for (int i = threadIdx.x; i < p.img_vec; i += CS_THREADS) dst[i] = __ldg(p.img + i);
  }
  __syncthreads();
  const uint32_t lo_off = (uint32_t)p.bwd_halves * 2u;
  const uint32_t sfb = (uint32_t)(p.KF + 8) * 2u, shb = (uint32_t)CS_SH * 2u;
  const uint32_t wf_a = sm, wblk_a = sm + (uint32_t)CS_HP * sfb, w0_a = wblk_a + (uint32_t)(2 * Bk) * CS_HP * shb;
  const int kbfn = p.KF >> 4;
  const uint32_t cdiv = p.ctx_div <= 0 ? 0xFFFFFFFFu : (uint32_t)p.ctx_div;
  const int64_t n_tiles = (p.R + 15) >> 4;

  for (int64_t tile = (int64_t)blockIdx.x * CS_WARPS + warp; tile < n_tiles; tile += (int64_t)gridDim.x * CS_WARPS) {
    const int64_t r0 = tile * 16 + g, r1 = r0 + 8;
    const bool v0 = r0 < p.R, v1 = r1 < p.R;
    const int64_t q0 = v0 ? r0 : p.R - 1, q1 = v1 ? r1 : p.R - 1;
    const int64_t cr0 = (int64_t)((uint32_t)q0 / cdiv), cr1 = (int64_t)((uint32_t)q1 / cdiv);
    const float* c0 = p.ctx + cr0 * p.ldg;
    const float* c1 = p.ctx + cr1 * p.ldg;
    float* dc0 = p.dctx + cr0 * p.ldg;
    float* dc1 = p.dctx + cr1 * p.ldg;
    float hh[8][4], xx[8][4], dh[8][4];
    float un0, un1;
    float tb[8][4], gb[8][4];            // the next phase's saved activation / gate, requested one GEMM ahead
    preload_rows(tb, p.u[Bk - 1] + q0 * p.ldh, p.u[Bk - 1] + q1 * p.ldh, H, t);
    preload_rows(gb, c0 + (int64_t)H * Bk, c1 + (int64_t)H * Bk, H, t);

    {  // ---- final layer: dh = dprm Wf  (contraction over the d_tr (3K-1) spline parameters, streamed from HBM)
      const float* d0 = p.dprm + q0 * p.ldp;
      const float* d1 = p.dprm + q1 * p.ldp;
      float m0 = 0.f, m1 = 0.f;
      for (int kb = 0; kb < kbfn; ++kb) {
        float v[8];
        load_a8(d0, d1, kb * 16 + 2 * t, p.dP, v);
        m0 = fmaxf(m0, fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[4]), fabsf(v[5]))));
        m1 = fmaxf(m1, fmaxf(fmaxf(fabsf(v[2]), fabsf(v[3])), fmaxf(fabsf(v[6]), fabsf(v[7]))));
      }
      float s0, s1;
      row_scale(quad_max(v0 ? m0 : 0.f), s0, un0);
      row_scale(quad_max(v1 ? m1 : 0.f), s1, un1);
      zero_acc(hh, xx);
      for (int kb = 0; kb < kbfn; ++kb) {
        float v[8];
        uint32_t a1h[1][4], a1l[1][4];
        load_a8(d0, d1, kb * 16 + 2 * t, p.dP, v);       // second pass over the row: L1 / L2 hits
        split_a8(v, s0, s1, a1h[0], a1l[0]);
        tile_gemm<1, 8>(hh, xx, a1h, a1l, 1, kb, 8, wf_a, lo_off, sfb, lane);
      }
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        dh[nb][0] = un0 * (hh[nb][0] + FP16_LO_UNSCALE * xx[nb][0]); dh[nb][1] = un0 * (hh[nb][1] + FP16_LO_UNSCALE * xx[nb][1]);
        dh[nb][2] = un1 * (hh[nb][2] + FP16_LO_UNSCALE * xx[nb][2]); dh[nb][3] = un1 * (hh[nb][3] + FP16_LO_UNSCALE * xx[nb][3]);
      }
    }

    uint32_t ahi[4][4], alo[4][4];
    for (int k = Bk - 1; k >= 0; --k) {
      float x[8][4];
      // ---- GLU backward: du = dh * sig ; d gate += dh * u * sig (1 - sig), reduced over the rows sharing a context row
      float* dg0 = dc0 + (int64_t)H * (1 + k);
      float* dg1 = dc1 + (int64_t)H * (1 + k);
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const int col = nb * 8 + 2 * t;
        const float2 sa = make_float2(gb[nb][0], gb[nb][1]), sb = make_float2(gb[nb][2], gb[nb][3]);
        const float2 ua = make_float2(tb[nb][0], tb[nb][1]), ub = make_float2(tb[nb][2], tb[nb][3]);
        x[nb][0] = dh[nb][0] * sa.x; x[nb][1] = dh[nb][1] * sa.y;
        x[nb][2] = dh[nb][2] * sb.x; x[nb][3] = dh[nb][3] * sb.y;
        if (v0) {
          st_pair(p.du[k] + r0 * p.ldh, col, H, x[nb][0], x[nb][1]);
          if (col < H) red_add_f32(dg0 + col, dh[nb][0] * ua.x * sa.x * (1.f - sa.x));
          if (col + 1 < H) red_add_f32(dg0 + col + 1, dh[nb][1] * ua.y * sa.y * (1.f - sa.y));
        }
        if (v1) {
          st_pair(p.du[k] + r1 * p.ldh, col, H, x[nb][2], x[nb][3]);
          if (col < H) red_add_f32(dg1 + col, dh[nb][2] * ub.x * sb.x * (1.f - sb.x));
          if (col + 1 < H) red_add_f32(dg1 + col + 1, dh[nb][3] * ub.y * sb.y * (1.f - sb.y));
        }
      }
      // ---- dt = (du W2) * (t' > 0) * 1/(1-p)        (t' = dropout(t): relu mask and keep mask in one)
      make_afrags<8>(x, ahi, alo, un0, un1);
      preload_rows(tb, p.t[k] + q0 * p.ldh, p.t[k] + q1 * p.ldh, H, t);
      zero_acc(hh, xx);
      tile_gemm<4, 8>(hh, xx, ahi, alo, 4, 0, 8, wblk_a + (uint32_t)(2 * k + 1) * CS_HP * shb, lo_off, shb, lane);
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const int col = nb * 8 + 2 * t;
        x[nb][0] = tb[nb][0] > 0.f ? un0 * p.mask_scale * (hh[nb][0] + FP16_LO_UNSCALE * xx[nb][0]) : 0.f;
        x[nb][1] = tb[nb][1] > 0.f ? un0 * p.mask_scale * (hh[nb][1] + FP16_LO_UNSCALE * xx[nb][1]) : 0.f;
        x[nb][2] = tb[nb][2] > 0.f ? un1 * p.mask_scale * (hh[nb][2] + FP16_LO_UNSCALE * xx[nb][2]) : 0.f;
        x[nb][3] = tb[nb][3] > 0.f ? un1 * p.mask_scale * (hh[nb][3] + FP16_LO_UNSCALE * xx[nb][3]) : 0.f;
        if (v0) st_pair(p.dt[k] + r0 * p.ldh, col, H, x[nb][0], x[nb][1]);
        if (v1) st_pair(p.dt[k] + r1 * p.ldh, col, H, x[nb][2], x[nb][3]);
      }
      // ---- dh += (dt W1) * (h_k > 0)
      make_afrags<8>(x, ahi, alo, un0, un1);
      preload_rows(tb, p.h[k] + q0 * p.ldh, p.h[k] + q1 * p.ldh, H, t);
      zero_acc(hh, xx);
      tile_gemm<4, 8>(hh, xx, ahi, alo, 4, 0, 8, wblk_a + (uint32_t)(2 * k) * CS_HP * shb, lo_off, shb, lane);
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        if (tb[nb][0] > 0.f) dh[nb][0] += un0 * (hh[nb][0] + FP16_LO_UNSCALE * xx[nb][0]);
        if (tb[nb][1] > 0.f) dh[nb][1] += un0 * (hh[nb][1] + FP16_LO_UNSCALE * xx[nb][1]);
        if (tb[nb][2] > 0.f) dh[nb][2] += un1 * (hh[nb][2] + FP16_LO_UNSCALE * xx[nb][2]);
        if (tb[nb][3] > 0.f) dh[nb][3] += un1 * (hh[nb][3] + FP16_LO_UNSCALE * xx[nb][3]);
      }
      if (k > 0) {       // for the next block's GLU backward
        preload_rows(tb, p.u[k - 1] + q0 * p.ldh, p.u[k - 1] + q1 * p.ldh, H, t);
        preload_rows(gb, c0 + (int64_t)H * k, c1 + (int64_t)H * k, H, t);
      }
    }

    // ---- initial layer: dh0 out (weight gradient operand), d ctx slot 0 += dh0, d theta_i += dh0 W0x
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const int col = nb * 8 + 2 * t;
      if (v0) {
        st_pair(p.dh0 + r0 * p.ldh, col, H, dh[nb][0], dh[nb][1]);
        if (col < H) red_add_f32(dc0 + col, dh[nb][0]);
        if (col + 1 < H) red_add_f32(dc0 + col + 1, dh[nb][1]);
      }
      if (v1) {
        st_pair(p.dh0 + r1 * p.ldh, col, H, dh[nb][2], dh[nb][3]);
        if (col < H) red_add_f32(dc1 + col, dh[nb][2]);
        if (col + 1 < H) red_add_f32(dc1 + col + 1, dh[nb][3]);
      }
    }
    make_afrags<8>(dh, ahi, alo, un0, un1);
    zero_acc(hh, xx);
    const int nbd = p.D8 >> 3;
    tile_gemm<4, 8>(hh, xx, ahi, alo, 4, 0, nbd, w0_a, lo_off, shb, lane);
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      if (nb < nbd) {
        const int col = nb * 8 + 2 * t;
        if (v0) {
          float* d = p.dth + r0 * p.D;
          if (col < p.D) d[col] += un0 * (hh[nb][0] + FP16_LO_UNSCALE * xx[nb][0]);
          if (col + 1 < p.D) d[col + 1] += un0 * (hh[nb][1] + FP16_LO_UNSCALE * xx[nb][1]);
        }
        if (v1) {
          float* d = p.dth + r1 * p.D;
          if (col < p.D) d[col] += un1 * (hh[nb][2] + FP16_LO_UNSCALE * xx[nb][2]);
          if (col + 1 < p.D) d[col + 1] += un1 * (hh[nb][3] + FP16_LO_UNSCALE * xx[nb][3]);
        }
      }
    }
  }
}

// ---- all weight gradients of one coupling layer's conditioner in ONE launch ---------------------------------------
// dW_j[M_j, N_j] += dOut_j[R, M_j]^T act_j[R, N_j] (relu on act_j where the forward applied one), db_j[M_j] += colsum(dOut_j)
// for the 2 Bk + 2 linear layers of a hidden <= 64 ResidualNet.  As separate launches these were 5 tcgen05 calls whose
// 128x128 tiles held a 50x50 result (39 us each, a ring hand-shake every 32 rows) plus an fp32 SIMT call for the 10-column
// initial layer (77 us).  Here a CTA owns one (matrix, 64-row slab of it, range of batch rows): 64-row tiles of both
// operands are converted to fp16 hi + scaled residual with ONE POWER-OF-TWO SCALE PER COLUMN (per output unit and per
// input unit: the scales factor out of the contraction exactly), staged row-major in shared memory and read back
// TRANSPOSED by ldmatrix.trans -- the contraction runs over the batch rows -- into warp-level m16n8k16 MMAs; every warp
// keeps its 16 x 32 slice of the result in fp32 registers across all tiles and adds it to dW once at the end.
constexpr int SW_MAXPAIRS = 12;
struct SwPair {
  const float* A; int32_t lda, M;            // dOut [R, M]
  const float* B; int32_t ldb, N, relu_b;    // act [R, N], N <= 64
  float* dW; int32_t ldw; float* db;         // dW [M, N] +=, db [M] += (may be NULL)
};
struct SwParams {
  SwPair pair[SW_MAXPAIRS];
  int32_t unit_pair[4 * SW_MAXPAIRS], unit_m0[4 * SW_MAXPAIRS];      // blockIdx.y -> (pair, first output row of its 64-row slab)
  int64_t R; int32_t rows_per_cta;
};
constexpr int SW_TS = 72;                    // row stride (halves) of the staged tiles: 144 B, conflict-free for ldmatrix

__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}

constexpr int SW_RAW_FLOATS = 64 * 64;                       // one raw fp32 operand tile
constexpr int SW_SMEM_BYTES = 2 * 2 * SW_RAW_FLOATS * 4 + 4 * 64 * SW_TS * 2;     // 2 buffers x (A, B) raw + 4 fp16 tiles

__global__ void __launch_bounds__(256, 2) k_small_wgrad(const SwParams p) {
  extern __shared__ __align__(16) uint8_t sw_smem[];
  float* raw = reinterpret_cast<float*>(sw_smem);                                         // [buf][A | B][row][64]
  uint16_t* tA = reinterpret_cast<uint16_t*>(sw_smem + 2 * 2 * SW_RAW_FLOATS * 4);        // [hi | lo][row][SW_TS]
  uint16_t* tB = tA + 2 * 64 * SW_TS;
  __shared__ int cmaxA[64], cmaxB[64];
  __shared__ float unA[64], unB[64], scA[64], scB[64], dbs[64];
  const SwPair& pr = p.pair[p.unit_pair[blockIdx.y]];
  const int m0 = p.unit_m0[blockIdx.y];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int mt = warp & 3, nh = warp >> 2;                   // this warp's 16 output rows / 32 output columns
  const int c4 = (tid & 15) * 4, rr = tid >> 4;              // staging: columns c4..c4+3 of rows rr + 16 j
  const int64_t r_begin = (int64_t)blockIdx.x * p.rows_per_cta;
  const int64_t r_end = min(p.R, r_begin + p.rows_per_cta);
  if (r_begin >= r_end) return;
  const float* A = pr.A; const float* B = pr.B;
  const int lda = pr.lda, ldb = pr.ldb, M = pr.M, N = pr.N;
  const bool relu_b = pr.relu_b != 0;
  const bool a_vec = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((m0 & 3) == 0);
  const bool b_vec = ((ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);

  float master[4][4], dbacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int nb = 0; nb < 4; ++nb)
#pragma unroll
    for (int j = 0; j < 4; ++j) master[nb][j] = 0.f;

  // Raw fp32 tiles arrive by cp.async (16-byte chunks, zero-filled beyond the matrix) into a two-deep buffer: the copy of
  // tile i+1 is issued BEFORE tile i is converted and multiplied, so a whole iteration hides its latency (the kernel was
  // waiting on its global loads: ncu long_scoreboard).  Unaligned operands take plain loads into the same buffer.
  auto issue_tile = [&](int64_t row0, int buf) {
    float* ra = raw + buf * 2 * SW_RAW_FLOATS;
    float* rb = ra + SW_RAW_FLOATS;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = rr + 16 * j;
      const int64_t r = row0 + row;
      const bool rv = r < r_end;
      const int na = rv ? max(0, min(4, M - (m0 + c4))) : 0;
      const int nb = rv ? max(0, min(4, N - c4)) : 0;
      if (a_vec) {
        const float* src = na ? A + r * lda + m0 + c4 : A;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(ra + row * 64 + c4)), "l"(src), "r"(na * 4) : "memory");
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) ra[row * 64 + c4 + e] = e < na ? __ldg(A + r * lda + m0 + c4 + e) : 0.f;
      }
      if (b_vec) {
        const float* src = nb ? B + r * ldb + c4 : B;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(rb + row * 64 + c4)), "l"(src), "r"(nb * 4) : "memory");
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) rb[row * 64 + c4 + e] = e < nb ? __ldg(B + r * ldb + c4 + e) : 0.f;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  float va[4][4], vb[4][4];
  issue_tile(r_begin, 0);
  int it = 0;
  for (int64_t row0 = r_begin; row0 < r_end; row0 += 64, ++it) {
    if (tid < 64) { cmaxA[tid] = 0; cmaxB[tid] = 0; }
    const bool more = row0 + 64 < r_end;
    if (more) issue_tile(row0 + 64, (it + 1) & 1);           // lands while this tile is converted and multiplied
    if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                        // tile `it` is in shared memory (everybody's chunks); the previous
                                                            // tile's MMAs are done with the staged fp16 tiles
    {
      const float* ra = raw + (it & 1) * 2 * SW_RAW_FLOATS;
      const float* rb = ra + SW_RAW_FLOATS;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 a4 = *reinterpret_cast<const float4*>(ra + (rr + 16 * j) * 64 + c4);
        float4 b4 = *reinterpret_cast<const float4*>(rb + (rr + 16 * j) * 64 + c4);
        if (relu_b) { b4.x = fmaxf(b4.x, 0.f); b4.y = fmaxf(b4.y, 0.f); b4.z = fmaxf(b4.z, 0.f); b4.w = fmaxf(b4.w, 0.f); }
        va[j][0] = a4.x; va[j][1] = a4.y; va[j][2] = a4.z; va[j][3] = a4.w;
        vb[j][0] = b4.x; vb[j][1] = b4.y; vb[j][2] = b4.z; vb[j][3] = b4.w;
      }
    }
    // ---- per-column maxima of this tile (non-negative floats order like their bit patterns)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float ma = 0.f, mb = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) { ma = fmaxf(ma, fabsf(va[j][e])); mb = fmaxf(mb, fabsf(vb[j][e])); dbacc[e] += va[j][e]; }
      if (ma > 0.f) atomicMax(&cmaxA[c4 + e], __float_as_int(ma));
      if (mb > 0.f) atomicMax(&cmaxB[c4 + e], __float_as_int(mb));
    }
    __syncthreads();
    if (tid < 64) {
      float sc, un;
      row_scale(__int_as_float(cmaxA[tid]), sc, un); scA[tid] = sc; unA[tid] = un;
      row_scale(__int_as_float(cmaxB[tid]), sc, un); scB[tid] = sc; unB[tid] = un;
    }
    __syncthreads();
    // ---- convert and stage: [row][col] halves, hi and scaled residual
    {
      const float4 sa = *reinterpret_cast<const float4*>(&scA[c4]);
      const float4 sb = *reinterpret_cast<const float4*>(&scB[c4]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int row = rr + 16 * j;
        uint32_t h0, l0, h1, l1;
        split_f16x2(va[j][0] * sa.x, va[j][1] * sa.y, h0, l0);
        split_f16x2(va[j][2] * sa.z, va[j][3] * sa.w, h1, l1);
        *reinterpret_cast<uint2*>(&tA[row * SW_TS + c4]) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(&tA[64 * SW_TS + row * SW_TS + c4]) = make_uint2(l0, l1);
        split_f16x2(vb[j][0] * sb.x, vb[j][1] * sb.y, h0, l0);
        split_f16x2(vb[j][2] * sb.z, vb[j][3] * sb.w, h1, l1);
        *reinterpret_cast<uint2*>(&tB[row * SW_TS + c4]) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(&tB[64 * SW_TS + row * SW_TS + c4]) = make_uint2(l0, l1);
      }
    }
    __syncthreads();
    // ---- this warp's 16 x 32 slice: contraction over the tile's 64 batch rows
    float hh[4][4], xx[4][4];
    zero_acc(hh, xx);
    const uint32_t a_hi = smem_u32(tA), a_lo = smem_u32(tA + 64 * SW_TS);
    const uint32_t b_hi = smem_u32(tB), b_lo = smem_u32(tB + 64 * SW_TS);
    // ldmatrix.trans lane addresses: matrices {k 0-7, k 8-15} x {col 0-7, col 8-15}
    const uint32_t a_off = (uint32_t)(((lane & 7) + ((lane >> 4) << 3)) * SW_TS + mt * 16 + (((lane >> 3) & 1) << 3)) * 2u;
    const uint32_t b_off = (uint32_t)(((lane & 7) + (((lane >> 3) & 1) << 3)) * SW_TS + nh * 32 + ((lane >> 4) << 3)) * 2u;
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
      uint32_t ah[4], al[4];
      const uint32_t ko = (uint32_t)(kb * 16 * SW_TS) * 2u;
      // A fragment order: a0 (m 0-7, k 0-7), a1 (m 8-15, k 0-7), a2 (m 0-7, k 8-15), a3 (m 8-15, k 8-15)
      ldsm4t(a_hi + ko + a_off, ah[0], ah[1], ah[2], ah[3]);
      ldsm4t(a_lo + ko + a_off, al[0], al[1], al[2], al[3]);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t bh0, bh1, bh2, bh3, bl0, bl1, bl2, bl3;      // {n-block 2np: k lo, k hi}, {n-block 2np+1: k lo, k hi}
        ldsm4t(b_hi + ko + b_off + (uint32_t)(np * 16) * 2u, bh0, bh1, bh2, bh3);
        ldsm4t(b_lo + ko + b_off + (uint32_t)(np * 16) * 2u, bl0, bl1, bl2, bl3);
        mma16816(hh[2 * np], ah, bh0, bh1); mma16816(xx[2 * np], al, bh0, bh1); mma16816(xx[2 * np], ah, bl0, bl1);
        mma16816(hh[2 * np + 1], ah, bh2, bh3); mma16816(xx[2 * np + 1], al, bh2, bh3); mma16816(xx[2 * np + 1], ah, bl2, bl3);
      }
    }
    const float ua0 = unA[mt * 16 + g], ua1 = unA[mt * 16 + g + 8];
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      const float ub0 = unB[nh * 32 + nb * 8 + 2 * t], ub1 = unB[nh * 32 + nb * 8 + 2 * t + 1];
      master[nb][0] += (hh[nb][0] + FP16_LO_UNSCALE * xx[nb][0]) * ua0 * ub0;
      master[nb][1] += (hh[nb][1] + FP16_LO_UNSCALE * xx[nb][1]) * ua0 * ub1;
      master[nb][2] += (hh[nb][2] + FP16_LO_UNSCALE * xx[nb][2]) * ua1 * ub0;
      master[nb][3] += (hh[nb][3] + FP16_LO_UNSCALE * xx[nb][3]) * ua1 * ub1;
    }
  }
  // ---- add this CTA's partial result
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) {
    const int n = nh * 32 + nb * 8 + 2 * t;
    const int ma = m0 + mt * 16 + g, mb = ma + 8;
    if (ma < M) {
      if (n < N) red_add_f32(pr.dW + (int64_t)ma * pr.ldw + n, master[nb][0]);
      if (n + 1 < N) red_add_f32(pr.dW + (int64_t)ma * pr.ldw + n + 1, master[nb][1]);
    }
    if (mb < M) {
      if (n < N) red_add_f32(pr.dW + (int64_t)mb * pr.ldw + n, master[nb][2]);
      if (n + 1 < N) red_add_f32(pr.dW + (int64_t)mb * pr.ldw + n + 1, master[nb][3]);
    }
  }
  if (pr.db) {
    __syncthreads();
    if (tid < 64) dbs[tid] = 0.f;
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) atomicAdd(&dbs[c4 + e], dbacc[e]);
    __syncthreads();
    if (tid < 64 && m0 + tid < M) red_add_f32(pr.db + m0 + tid, dbs[tid]);
  }
}

// Host side: queue the pairs of a layer, then launch once.
struct SmallWgradBatch {
  SwParams p;
  int npairs = 0, nunits = 0;
  SmallWgradBatch() { p.R = 0; p.rows_per_cta = 0; }
  bool add(const float* dOut, int ldo, int M, const float* act, int lda, int N, bool relu, float* dW, int ldw, float* db) {
    if (npairs >= SW_MAXPAIRS || N > 64 || N < 1 || M < 1) return false;
    const int units = (M + 63) / 64;
    if (nunits + units > 4 * SW_MAXPAIRS) return false;
    SwPair& q = p.pair[npairs];
    q.A = dOut; q.lda = ldo; q.M = M; q.B = act; q.ldb = lda; q.N = N; q.relu_b = relu ? 1 : 0; q.dW = dW; q.ldw = ldw; q.db = db;
    for (int u = 0; u < units; ++u) { p.unit_pair[nunits] = npairs; p.unit_m0[nunits] = u * 64; ++nunits; }
    ++npairs;
    return true;
  }
};

int launch_small_wgrad(SmallWgradBatch& b, int64_t R, cudaStream_t stream) {
  if (b.nunits == 0 || R <= 0) return 0;
  // ~4 CTAs per SM in total, whole 64-row tiles per CTA
  int64_t gx = std::max<int64_t>(1, (4 * (int64_t)num_sms() + b.nunits - 1) / b.nunits);
  int64_t rows = ((R + gx - 1) / gx + 63) / 64 * 64;
  rows = std::max<int64_t>(rows, 256);
  gx = (R + rows - 1) / rows;
  b.p.R = R; b.p.rows_per_cta = (int32_t)rows;
  double flops = 0.0, bytes = 0.0;
  for (int i = 0; i < b.npairs; ++i) {
    flops += 2.0 * (double)R * b.p.pair[i].M * b.p.pair[i].N;
    bytes += 4.0 * (double)R * (b.p.pair[i].M + (double)b.p.pair[i].N * ((b.p.pair[i].M + 63) / 64));
  }
  ProfScope ps(TAG_SMALL_WGRAD, flops, stream, bytes);
  dim3 grid((unsigned)gx, (unsigned)b.nunits);
  static bool attr_set = false;
  if (!attr_set) {
    FP_CUDA_OK(cudaFuncSetAttribute(k_small_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_SMEM_BYTES));
    attr_set = true;
  }
  k_small_wgrad<<<grid, 256, SW_SMEM_BYTES, stream>>>(b.p);
  FP_LAUNCH_CHECK();
  return 0;
}

int launch_small_wgrad_layer(const Layout& L, int i, const float* dprm, const float* const* hs, const float* const* ts,
                             float* const* dus, float* const* dts, const float* dh0, const float* theta_i, int64_t R,
                             float* grads, cudaStream_t s) {
  const int H = L.H, lH = L.ldH(), D = L.D;
  SmallWgradBatch b;
  bool ok = b.add(dprm, L.ldP(i), L.dP(i), hs[L.Bk], lH, H, false, grads + L.wf(i), H, grads + L.bf(i));
  for (int k = L.Bk - 1; k >= 0 && ok; --k) {
    ok = ok && b.add(dus[k], lH, H, ts[k], lH, H, true, grads + L.w2(i, k), H, grads + L.b2(i, k));
    ok = ok && b.add(dts[k], lH, H, hs[k], lH, H, true, grads + L.w1(i, k), H, grads + L.b1(i, k));
  }
  ok = ok && b.add(dh0, lH, H, theta_i, D, D, false, grads + L.w0x(i), D, nullptr);
  if (!ok) return FP_E_NOTREADY;
  return launch_small_wgrad(b, R, s);
}

// ---- fp16 hi / scaled-residual images of one layer's matrices in the kernels' shared-memory layout (layout.h sh*) ----
__global__ void k_prep_small(const float* __restrict__ params, uint16_t* __restrict__ img, fp_nsf_config cfg) {
  const Layout L(&cfg);
  const int i = blockIdx.y;
  const int H = L.H, D = L.D, dP = L.dP(i), Bk = L.Bk;
  const int64_t nF = L.shFwdHalves(i), nB = L.shBwdHalves(i);
  uint16_t* fhi = img + L.shLayerOff(i);
  uint16_t* flo = fhi + nF;
  uint16_t* bhi = flo + nF;
  uint16_t* blo = bhi + nB;
  const int s0 = L.shK0() + 8, sf = L.shKF(i) + 8;
  const int64_t nW0 = (int64_t)CS_HP * s0, nBlk = (int64_t)CS_HP * CS_SH, nWfT = (int64_t)CS_HP * sf;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nF + nB; e += (int64_t)gridDim.x * blockDim.x) {
    float v = 0.f;
    if (e < nF) {
      if (e < nW0) {
        const int n = (int)(e / s0), k = (int)(e % s0);
        if (n < H && k < D) v = params[L.w0x(i) + (int64_t)n * D + k];
      } else if (e < nW0 + 2 * Bk * nBlk) {
        const int64_t r = e - nW0;
        const int m = (int)(r / nBlk), n = (int)((r % nBlk) / CS_SH), k = (int)(r % CS_SH);
        if (n < H && k < H) v = params[((m & 1) ? L.w2(i, m >> 1) : L.w1(i, m >> 1)) + (int64_t)n * H + k];
      } else {
        const int64_t r = e - nW0 - 2 * Bk * nBlk;
        const int n = (int)(r / CS_SH), k = (int)(r % CS_SH);
        if (n < dP && k < H) v = params[L.wf(i) + (int64_t)n * H + k];
      }
    } else {
      const int64_t eb = e - nF;
      if (eb < nWfT) {                 // Wf^T: row = hidden unit, contraction over the spline parameters
        const int n = (int)(eb / sf), k = (int)(eb % sf);
        if (n < H && k < dP) v = params[L.wf(i) + (int64_t)k * H + n];
      } else if (eb < nWfT + 2 * Bk * nBlk) {
        const int64_t r = eb - nWfT;
        const int m = (int)(r / nBlk), n = (int)((r % nBlk) / CS_SH), k = (int)(r % CS_SH);
        if (n < H && k < H) v = params[((m & 1) ? L.w2(i, m >> 1) : L.w1(i, m >> 1)) + (int64_t)k * H + n];
      } else {                         // W0x^T: row = theta column, contraction over the hidden units
        const int64_t r = eb - nWfT - 2 * Bk * nBlk;
        const int n = (int)(r / CS_SH), k = (int)(r % CS_SH);
        if (n < D && k < H) v = params[L.w0x(i) + (int64_t)k * D + n];
      }
    }
    uint32_t h2, l2;
    split_f16x2(v, 0.f, h2, l2);
    if (e < nF) { fhi[e] = (uint16_t)(h2 & 0xffffu); flo[e] = (uint16_t)(l2 & 0xffffu); }
    else { bhi[e - nF] = (uint16_t)(h2 & 0xffffu); blo[e - nF] = (uint16_t)(l2 & 0xffffu); }
  }
}

static int cs_fwd_smem(const Layout& L, int i) { return (int)(L.shFwdHalves(i) * 4 + (L.Bk * 2 * CS_HP + L.shNF(i)) * 4); }
static int cs_bwd_smem(const Layout& L, int i) { return (int)(L.shBwdHalves(i) * 4); }

// Shapes the shared-memory-resident kernels cover (a property of the flow, identical for every pass)
bool cond_small_supported(const Layout& L, const float* prepared) {
  if (!smallh_enabled() || !prepared) return false;
  if (!(L.H >= 2 && L.H <= CS_HP && L.D <= 64 && L.Bk >= 1 && L.Bk <= CS_MAXBK)) return false;
  return cs_fwd_smem(L, 0) <= CS_SMEM_LIMIT && cs_bwd_smem(L, 0) <= CS_SMEM_LIMIT;
}

int launch_small_prepare(const fp_nsf_config* cfg, const float* params, float* prepared, cudaStream_t stream) {
  Layout L(cfg);
  if (!cond_small_supported(L, prepared)) return 0;
  ProfScope ps(TAG_OTHER, 0.0, stream);
  dim3 grid(32, L.T);
  k_prep_small<<<grid, 256, 0, stream>>>(params, reinterpret_cast<uint16_t*>(prepared + L.prepSmall()), *cfg);
  FP_LAUNCH_CHECK();
  return 0;
}

static int cs_grid(int64_t R) {
  const int64_t tiles = (R + 15) / 16;
  return (int)std::min<int64_t>((tiles + CS_WARPS - 1) / CS_WARPS, num_sms());
}

int launch_cond_small_fwd(const fp_nsf_config* cfg, const Layout& L, const float* params, const float* prepared, int i,
                          const float* theta_i, const float* ctx_all, int64_t R, int ctx_div, float* const* h, float* const* t,
                          float* const* u, float* prm, bool training, float drop_p, uint64_t drop_seed, cudaStream_t stream) {
  if (!cond_small_supported(L, prepared)) return FP_E_NOTREADY;
  if (R < 1 || R > 2147483647LL - 256) return FP_E_TOOLARGE;
  CsFwdParams p{};
  const uint16_t* img = reinterpret_cast<const uint16_t*>(prepared + L.prepSmall()) + L.shLayerOff(i);
  p.img = reinterpret_cast<const uint4*>(img);
  p.fwd_halves = (int32_t)L.shFwdHalves(i);
  p.img_vec = (int32_t)(L.shFwdHalves(i) * 4 / 16);
  p.theta = theta_i; p.D = L.D; p.K0 = L.shK0();
  p.ctx = ctx_all + L.ctxCol(i, 0); p.ldg = L.Nctx; p.ctx_div = ctx_div;
  for (int k = 0; k < L.Bk; ++k) { p.b1[k] = params + L.b1(i, k); p.b2[k] = params + L.b2(i, k); }
  p.bf = params + L.bf(i);
  if (training) {
    for (int k = 0; k <= L.Bk; ++k) p.h[k] = h[k];
    for (int k = 0; k < L.Bk; ++k) { p.t[k] = t[k]; p.u[k] = u[k]; }
  }
  p.ldh = L.ldH();
  p.prm = prm; p.ldp = L.ldP(i); p.dP = L.dP(i); p.NF = L.shNF(i);
  p.R = R; p.H = L.H; p.Bk = L.Bk; p.training = training ? 1 : 0;
  const bool drop = training && drop_p > 0.f;
  for (int k = 0; k < L.Bk; ++k) p.dk[k] = make_drop_key(layer_seed(drop_seed, i, k), drop_p);
  const int smem = cs_fwd_smem(L, i);
  static int attr_set = 0;
  if (attr_set < smem) {
    FP_CUDA_OK(cudaFuncSetAttribute(k_cond_small_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM_LIMIT));
    FP_CUDA_OK(cudaFuncSetAttribute(k_cond_small_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM_LIMIT));
    attr_set = CS_SMEM_LIMIT;
  }
  const double flops = 2.0 * (double)R * ((double)L.D * L.H + (double)L.Bk * 2 * L.H * L.H + (double)L.H * L.dP(i));
  const double bytes = 4.0 * (double)R * (L.D + L.dP(i) + (training ? (3.0 * L.Bk + 1) * L.H : 0.0));
  ProfScope ps(TAG_COND_SMALL_FWD, flops, stream, bytes);
  if (drop) k_cond_small_fwd<true><<<cs_grid(R), CS_THREADS, smem, stream>>>(p);
  else k_cond_small_fwd<false><<<cs_grid(R), CS_THREADS, smem, stream>>>(p);
  FP_LAUNCH_CHECK();
  return 0;
}

int launch_cond_small_bwd(const fp_nsf_config* cfg, const Layout& L, const float* prepared, int i, const float* dprm,
                          const float* const* h, const float* const* t, const float* const* u, const float* ctx_all,
                          int ctx_div, int64_t R, float* const* du, float* const* dt, float* dh0, float* dctx_all, float* dth,
                          float mask_scale, cudaStream_t stream) {
  (void)cfg;
  if (!cond_small_supported(L, prepared)) return FP_E_NOTREADY;
  if (R < 1 || R > 2147483647LL - 256) return FP_E_TOOLARGE;
  CsBwdParams p{};
  const uint16_t* img = reinterpret_cast<const uint16_t*>(prepared + L.prepSmall()) + L.shLayerOff(i) + 2 * L.shFwdHalves(i);
  p.img = reinterpret_cast<const uint4*>(img);
  p.bwd_halves = (int32_t)L.shBwdHalves(i);
  p.img_vec = (int32_t)(L.shBwdHalves(i) * 4 / 16);
  p.dprm = dprm; p.ldp = L.ldP(i); p.dP = L.dP(i); p.KF = L.shKF(i);
  for (int k = 0; k <= L.Bk; ++k) p.h[k] = h[k];
  for (int k = 0; k < L.Bk; ++k) { p.t[k] = t[k]; p.u[k] = u[k]; p.du[k] = du[k]; p.dt[k] = dt[k]; }
  p.ldh = L.ldH();
  p.ctx = ctx_all + L.ctxCol(i, 0); p.ldg = L.Nctx; p.ctx_div = ctx_div;
  p.dh0 = dh0; p.dctx = dctx_all + L.ctxCol(i, 0);
  p.dth = dth; p.D = L.D; p.D8 = L.shD8();
  p.mask_scale = mask_scale;
  p.R = R; p.H = L.H; p.Bk = L.Bk;
  const int smem = cs_bwd_smem(L, i);
  static bool attr_set = false;
  if (!attr_set) {
    FP_CUDA_OK(cudaFuncSetAttribute(k_cond_small_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM_LIMIT));
    attr_set = true;
  }
  const double flops = 2.0 * (double)R * ((double)L.D * L.H + (double)L.Bk * 2 * L.H * L.H + (double)L.H * L.dP(i));
  const double bytes = 4.0 * (double)R * (L.D + L.dP(i) + (5.0 * L.Bk + 2) * L.H);
  ProfScope ps(TAG_COND_SMALL_BWD, 2.0 * flops, stream, bytes);
  k_cond_small_bwd<<<cs_grid(R), CS_THREADS, smem, stream>>>(p);
  FP_LAUNCH_CHECK();
  return 0;
}

}  // namespace fp
