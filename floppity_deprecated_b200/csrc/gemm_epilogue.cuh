// gemm_epilogue.cuh -- the fused conditioner epilogue shared by the SIMT and tcgen05 GEMM kernels.
#pragma once
#include "common.cuh"

namespace fp {

__device__ __forceinline__ uint32_t mix32(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (uint32_t)(z >> 32);
}
// keep-probability test shared by forward prologue and backward mask: element (row, col) of the
// dropped activation matrix with `width` columns
__device__ __forceinline__ float drop_scale(uint64_t seed, int64_t row, int col, int width, float p) {
  uint32_t h = mix32(seed ^ (uint64_t)(row * (int64_t)width + col));
  float u = (float)(h >> 8) * (1.0f / 16777216.0f);
  return u < p ? 0.f : 1.f / (1.f - p);
}

struct GemmDrop {
  float p;          // 0 => off
  uint64_t seed;
  int width;        // columns of the dropped activation matrix
};

// v = accumulator of element (m, n); applies the flags of fp_gemm_args and stores.
__device__ __forceinline__ void gemm_epilogue_store(const fp_gemm_args& g, const GemmDrop& drop, int64_t m, int n, float v) {
  const int fl = g.flags;
  const int64_t cr = (g.ctx_div > 0) ? (m / g.ctx_div) : 0;
  if (fl & FP_G_BIAS) v += g.bias[n];
  if (fl & FP_G_ADD_CTX) v += g.G[cr * g.ldg + n];
  if (fl & FP_G_GLU_RES) {
    if (g.U) g.U[m * g.ldu + n] = v;
    float gate = g.G[cr * g.ldg + n];
    v = g.Rsd[m * g.ldr + n] + v * (1.f / (1.f + expf(-gate)));
  }
  if (fl & FP_G_RELUMASK) {
    float mk = g.Msk[m * g.ldm + n];
    v = mk > 0.f ? v : 0.f;
    if (drop.p > 0.f && (fl & FP_G_DROPMASK)) v *= drop_scale(drop.seed, m, n, drop.width, drop.p);
  }
  float* c = g.C + m * g.ldc + n;
  if (fl & FP_G_ATOMIC) atomicAdd(c, v);
  else if (fl & FP_G_ACCUM) *c += v;
  else *c = v;
}


// Four consecutive columns n..n+3 of row m at once.  Caller guarantees n % 4 == 0, n + 3 < N and that every
// pointer used by the flags is 16-byte aligned with a leading dimension that is a multiple of 4 (no dropout).
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ void gemm_epilogue_store4(const fp_gemm_args& g, int64_t m, int n, float4 v) {
  const int fl = g.flags;
  const int64_t cr = (g.ctx_div > 0) ? (m / g.ctx_div) : 0;
  if (fl & FP_G_BIAS) {
    const float4 b = *reinterpret_cast<const float4*>(g.bias + n);
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  }
  if (fl & FP_G_ADD_CTX) {
    const float4 c = *reinterpret_cast<const float4*>(g.G + cr * g.ldg + n);
    v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;
  }
  if (fl & FP_G_GLU_RES) {
    const float4 gt = *reinterpret_cast<const float4*>(g.G + cr * g.ldg + n);
    const float4 rs = *reinterpret_cast<const float4*>(g.Rsd + m * g.ldr + n);
    if (g.U) *reinterpret_cast<float4*>(g.U + m * g.ldu + n) = v;
    v.x = rs.x + v.x * sigmoid_acc(gt.x); v.y = rs.y + v.y * sigmoid_acc(gt.y);
    v.z = rs.z + v.z * sigmoid_acc(gt.z); v.w = rs.w + v.w * sigmoid_acc(gt.w);
  }
  if (fl & FP_G_RELUMASK) {
    const float4 mk = *reinterpret_cast<const float4*>(g.Msk + m * g.ldm + n);
    v.x = mk.x > 0.f ? v.x : 0.f; v.y = mk.y > 0.f ? v.y : 0.f;
    v.z = mk.z > 0.f ? v.z : 0.f; v.w = mk.w > 0.f ? v.w : 0.f;
  }
  float* c = g.C + m * g.ldc + n;
  if (fl & FP_G_ACCUM) {
    const float4 p = *reinterpret_cast<const float4*>(c);
    v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
  }
  *reinterpret_cast<float4*>(c) = v;
}

}  // namespace fp
