// gemm_epilogue.cuh -- the fused conditioner epilogue shared by the SIMT and tcgen05 GEMM kernels.
#pragma once
#include "common.cuh"

namespace fp {

// ---- dropout mask (nflows ResidualBlock.dropout; -dropout at /root/reference/floppity.py:57, forwarded at modules.py:63) ----
// Counter-based: elements are numbered e = row * width + col; one 32-bit hash of (seed, e >> 1) yields two 16-bit
// uniforms, element e takes half (e & 1); it is DROPPED when its uniform is below thr = round(p * 65536).  The same
// function serves the GEMM epilogue (FP_G_DROP_OUT), the fused conditioner and fp_dropout_mask (which exports the mask
// so the parity tests can hand it to the oracle).
__host__ __device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du;
  x ^= x >> 15; x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
struct DropKey {
  uint32_t k0;          // folded seed
  uint32_t thr;         // drop when u16 < thr
  float scale;          // 1 / (1 - p)
};
__host__ __device__ __forceinline__ DropKey make_drop_key(uint64_t seed, float p) {
  DropKey k;
  k.k0 = lowbias32(lowbias32((uint32_t)seed ^ 0x9E3779B9u) + (uint32_t)(seed >> 32) * 0x85EBCA6Bu);
  float t = p * 65536.f + 0.5f;
  k.thr = t <= 0.f ? 0u : (t >= 65535.f ? 65535u : (uint32_t)t);
  k.scale = 1.f / (1.f - p);
  return k;
}
// the two 16-bit uniforms of element pair `word` (= element index >> 1): low half = even element
__device__ __forceinline__ uint32_t drop_word(const DropKey& k, uint64_t word) {
  return lowbias32(((uint32_t)word * 0x9E3779B1u) ^ k.k0 ^ ((uint32_t)(word >> 32) * 0x85EBCA77u));
}
__device__ __forceinline__ bool drop_keep(const DropKey& k, int64_t row, int col, int width) {
  const uint64_t e = (uint64_t)row * (uint64_t)width + (uint64_t)col;
  const uint32_t w = drop_word(k, e >> 1);
  return ((w >> (16u * (uint32_t)(e & 1u))) & 0xffffu) >= k.thr;
}

// explicit global-space reductions (a plain atomicAdd on a kernel-argument pointer compiles to the generic-address
// path with shared-memory CAS loops)
__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add_f32x4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// algorithmic HBM bytes of one GEMM launch: operands once, result once, plus whatever the epilogue streams
inline double gemm_algo_bytes(const fp_gemm_args* a) {
  const double M = (double)a->M, N = (double)a->N, K = (double)a->K;
  double b = M * K + N * K + M * N;
  if (a->flags & FP_G_GLU_RES) b += M * N * (a->U ? 2.0 : 1.0);      // residual in, pre-gate activation out
  if (a->flags & FP_G_RELUMASK) b += M * N;
  if (a->flags & (FP_G_ACCUM | FP_G_ATOMIC)) b += M * N;
  return 4.0 * b;
}


// context row of flow row m (only evaluated when an epilogue needs it; 32-bit division -- M < 2^31 is enforced at launch)
__device__ __forceinline__ int64_t ctx_row(const fp_gemm_args& g, int64_t m) {
  if (g.ctx_div <= 0) return 0;
  if (g.ctx_div == 1) return m;
  return (int64_t)((uint32_t)m / (uint32_t)g.ctx_div);
}

// v = accumulator of element (m, n); applies the flags of fp_gemm_args and stores.
__device__ __forceinline__ void gemm_epilogue_store(const fp_gemm_args& g, int fl, int64_t m, int n, float v) {
  const int64_t cr = (fl & (FP_G_ADD_CTX | FP_G_GLU_RES)) ? ctx_row(g, m) : 0;
  if (fl & FP_G_BIAS) v += g.bias[n];
  if (fl & FP_G_DROP_OUT) {
    const DropKey dk = make_drop_key(g.drop_seed, g.drop_p);
    v = drop_keep(dk, m, n, g.N) ? v * dk.scale : 0.f;
  }
  if (fl & FP_G_ADD_CTX) v += g.G[cr * g.ldg + n];
  if (fl & FP_G_GLU_RES) {
    if (g.U) g.U[m * g.ldu + n] = v;
    float gate = g.G[cr * g.ldg + n];
    v = g.Rsd[m * g.ldr + n] + v * (1.f / (1.f + expf(-gate)));
  }
  if (fl & FP_G_RELUMASK) {
    float mk = g.Msk[m * g.ldm + n];
    v = mk > 0.f ? v * (g.mask_scale != 0.f ? g.mask_scale : 1.f) : 0.f;
  }
  float* c = g.C + m * g.ldc + n;
  if (fl & FP_G_ATOMIC) red_add_f32(c, v);
  else if (fl & FP_G_ACCUM) *c += v;
  else *c = v;
}

// Register snapshot of the epilogue operands (the tensor-core kernels take it once per role: reading the kernel
// argument struct through a reference costs a generic-address load per access).
struct EpiArgs {
  float* C; int64_t ldc; int64_t M; int N;
  const float* bias;
  const float* G; int64_t ldg; int ctx_div;
  const float* Rsd; int64_t ldr;
  float* U; int64_t ldu;
  const float* Msk; int64_t ldm;
  float mask_scale;
  DropKey dk;
};
__device__ __forceinline__ EpiArgs make_epi(const fp_gemm_args& g) {
  EpiArgs e;
  e.mask_scale = g.mask_scale != 0.f ? g.mask_scale : 1.f;
  e.dk = make_drop_key(g.drop_seed, g.drop_p);
  e.C = g.C; e.ldc = g.ldc; e.M = g.M; e.N = g.N; e.bias = g.bias; e.G = g.G; e.ldg = g.ldg; e.ctx_div = g.ctx_div;
  e.Rsd = g.Rsd; e.ldr = g.ldr; e.U = g.U; e.ldu = g.ldu; e.Msk = g.Msk; e.ldm = g.ldm;
  return e;
}
__device__ __forceinline__ int64_t ctx_row(const EpiArgs& g, int64_t m) {
  if (g.ctx_div <= 0) return 0;
  if (g.ctx_div == 1) return m;
  return (int64_t)((uint32_t)m / (uint32_t)g.ctx_div);
}
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

// out-of-line scalar fallback for up to 4 consecutive columns (ragged N, unaligned rows): keeps the hot vector
// path of the tensor-core epilogues small.  Same semantics as gemm_epilogue_store without dropout / atomics.
static __device__ __noinline__ void gemm_epilogue_store_slow(EpiArgs g, int fl, int64_t m, int n, float4 v4, int cnt) {
  const float e[4] = {v4.x, v4.y, v4.z, v4.w};
  const int64_t cr = (fl & (FP_G_ADD_CTX | FP_G_GLU_RES)) ? ctx_row(g, m) : 0;
  for (int j = 0; j < cnt; ++j) {
    const int nn = n + j;
    float v = e[j];
    if (fl & FP_G_BIAS) v += g.bias[nn];
    if (fl & FP_G_DROP_OUT) v = drop_keep(g.dk, m, nn, g.N) ? v * g.dk.scale : 0.f;
    if (fl & FP_G_ADD_CTX) v += g.G[cr * g.ldg + nn];
    if (fl & FP_G_GLU_RES) {
      if (g.U) g.U[m * g.ldu + nn] = v;
      v = g.Rsd[m * g.ldr + nn] + v * sigmoid_acc(g.G[cr * g.ldg + nn]);
    }
    if (fl & FP_G_RELUMASK) v = g.Msk[m * g.ldm + nn] > 0.f ? v * g.mask_scale : 0.f;
    float* c = g.C + m * g.ldc + nn;
    if (fl & FP_G_ACCUM) *c += v; else *c = v;
  }
}

// ---- compile-time-specialised vector epilogue (tensor-core kernels) ---------------------------------------
// The run-time-flag version above costs ~200 issue slots per float4 (every operand's address arithmetic is
// emitted and predicated off); the skinny conditioner GEMMs are epilogue-bound, so the tcgen05 kernels are
// instantiated per flag combination (FL >= 0) and only fall back to run-time flags (FL < 0) for unusual ones.
constexpr int FP_EPI_MASK = FP_G_BIAS | FP_G_ADD_CTX | FP_G_GLU_RES | FP_G_RELUMASK | FP_G_ACCUM | FP_G_DROP_OUT;

__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

// The auxiliary operands of one float4 (context/gate row, residual, relu mask, old C).  They are loaded for a
// whole batch of rows BEFORE any of the batch is stored: a load placed after a store that may alias it cannot be
// hoisted by the compiler, and one exposed L2/HBM round trip per row made the fused epilogues latency-bound.
struct EpiAux { float4 ctx, rsd, msk, old; };

template <int FL>
__device__ __forceinline__ void epi_load4(const EpiArgs& g, int fl_rt, int64_t m, int64_t cr, int n, EpiAux& a) {
  const int fl = (FL >= 0) ? FL : fl_rt;
  if (fl & (FP_G_ADD_CTX | FP_G_GLU_RES)) a.ctx = __ldg(reinterpret_cast<const float4*>(g.G + cr * g.ldg + n));
  if (fl & FP_G_GLU_RES) a.rsd = *reinterpret_cast<const float4*>(g.Rsd + m * g.ldr + n);   // may alias C (in-place eval)
  if (fl & FP_G_RELUMASK) a.msk = __ldg(reinterpret_cast<const float4*>(g.Msk + m * g.ldm + n));
  if (fl & FP_G_ACCUM) a.old = *reinterpret_cast<const float4*>(g.C + m * g.ldc + n);
}

// v = four consecutive accumulator columns n..n+3 of row m; bias4 = bias[n..n+3] (hoisted by the caller)
template <int FL>
__device__ __forceinline__ void epi_store4(const EpiArgs& g, int fl_rt, int64_t m, int n, float4 bias4, const EpiAux& a, float4 v) {
  const int fl = (FL >= 0) ? FL : fl_rt;
  if (fl & FP_G_BIAS) { v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w; }
  if (fl & FP_G_DROP_OUT) {       // n % 4 == 0: elements n..n+3 of row m are two whole element pairs when N is even
    const uint64_t e = (uint64_t)m * (uint64_t)g.N + (uint64_t)n;
    if ((g.N & 1) == 0) {
      const uint32_t w0 = drop_word(g.dk, e >> 1), w1 = drop_word(g.dk, (e >> 1) + 1);
      v.x = (w0 & 0xffffu) >= g.dk.thr ? v.x * g.dk.scale : 0.f; v.y = (w0 >> 16) >= g.dk.thr ? v.y * g.dk.scale : 0.f;
      v.z = (w1 & 0xffffu) >= g.dk.thr ? v.z * g.dk.scale : 0.f; v.w = (w1 >> 16) >= g.dk.thr ? v.w * g.dk.scale : 0.f;
    } else {
      v.x = drop_keep(g.dk, m, n, g.N) ? v.x * g.dk.scale : 0.f; v.y = drop_keep(g.dk, m, n + 1, g.N) ? v.y * g.dk.scale : 0.f;
      v.z = drop_keep(g.dk, m, n + 2, g.N) ? v.z * g.dk.scale : 0.f; v.w = drop_keep(g.dk, m, n + 3, g.N) ? v.w * g.dk.scale : 0.f;
    }
  }
  if (fl & FP_G_ADD_CTX) { v.x += a.ctx.x; v.y += a.ctx.y; v.z += a.ctx.z; v.w += a.ctx.w; }
  if (fl & FP_G_GLU_RES) {
    if (g.U) *reinterpret_cast<float4*>(g.U + m * g.ldu + n) = v;
    v.x = fmaf(v.x, sigmoid_fast(a.ctx.x), a.rsd.x); v.y = fmaf(v.y, sigmoid_fast(a.ctx.y), a.rsd.y);
    v.z = fmaf(v.z, sigmoid_fast(a.ctx.z), a.rsd.z); v.w = fmaf(v.w, sigmoid_fast(a.ctx.w), a.rsd.w);
  }
  if (fl & FP_G_RELUMASK) {
    const float ms = g.mask_scale;
    v.x = a.msk.x > 0.f ? v.x * ms : 0.f; v.y = a.msk.y > 0.f ? v.y * ms : 0.f;
    v.z = a.msk.z > 0.f ? v.z * ms : 0.f; v.w = a.msk.w > 0.f ? v.w * ms : 0.f;
  }
  if (fl & FP_G_ACCUM) { v.x += a.old.x; v.y += a.old.y; v.z += a.old.z; v.w += a.old.w; }
  *reinterpret_cast<float4*>(g.C + m * g.ldc + n) = v;
}

}  // namespace fp
