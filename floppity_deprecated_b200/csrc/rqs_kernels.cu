// rqs_kernels.cu -- fused rational-quadratic-spline coupling kernels (forward / inverse / backward).
//
// Replaces nflows PiecewiseRationalQuadraticCouplingTransform._piecewise_cdf +
// unconstrained_rational_quadratic_spline + searchsorted (reached from
// /root/reference/modules.py:63 via sbi build_nsf; ~60 eager ATen ops + boolean-mask
// gather/scatter per layer upstream) with ONE kernel per direction:
//   - persistent CTAs, each streaming tiles of `rows` consecutive samples;
//   - the [rows, d_tr, 3K-1] conditioner output of a tile is one contiguous byte range, fetched
//     by the TMA engine as a 1-D bulk copy (cp.async.bulk) into a 2-stage shared-memory ring;
//   - one thread per (sample, transformed feature): softmax/softplus parameterisation, bin search,
//     RQ evaluation, all in registers, params read from smem at odd stride (conflict free);
//   - the per-sample log-det sum and the layer's LU-linear (y = M(x+pre)+post, D<=64) are fused in
//     a second phase that works on the tile's theta rows held in shared memory.
// HBM traffic per sample per layer = params + theta in + theta out + logdet.
#include "common.cuh"
#include "rqs_math.cuh"
#include "prof.h"
#include <math.h>
#include <stdlib.h>
#include <algorithm>

namespace fp {

constexpr int RQS_MAX_STAGES = 2;
// ring depth of the spline kernels (1 or 2), FLOPPITY_RQS_STAGES: with one stage a CTA holds half the shared memory, so
// twice as many CTAs (and rows in flight through the arithmetic) fit an SM and other CTAs' loads cover a CTA's wait
static int rqs_stages() {
  static int n = 0;
  if (n == 0) { const char* e = getenv("FLOPPITY_RQS_STAGES"); n = (e && e[0] == '2') ? 2 : ((e && e[0] == '1') ? 1 : 2); }
  return n;
}

struct RqsArgs {
  const float* theta_in;
  const float* params;
  float* theta_out;
  float* theta_sp;      // optional: post-spline / pre-LU rows (saved for the LU backward)
  float* logdet;        // accumulated
  int32_t* bin_idx;     // optional
  int32_t* status;      // optional
  int64_t R;
  int D, d_tr, parity, rows, stages;
  int ldp;              // row stride of params in floats (>= d_tr*P; padded to a multiple of 4 by the engine)
  RqsConsts c;
  const float* luM;     // optional [D, D]
  const float* luPre;   // optional [D] added before the matvec
  const float* luPost;  // optional [D] added after
  const float* lu_logdet;  // optional device scalar added to every row's logdet (LULinear logabsdet)
  float pre_sign;       // luPre is multiplied by this (-1 for the inverse direction: x - b)
};

// One stage of the ring = everything a tile needs, fetched by the TMA engine on ONE mbarrier:
//   [ params: rows*d_tr*P floats | theta rows: rows*D floats | logdet: rows floats ]
// The transformed values overwrite the theta rows in place; nothing is read from global by the SM's load path.
struct RqsSmem {
  float* stage[RQS_MAX_STAGES];
  float* ladbuf;
  float* luM;
  uint64_t* bars;
  int stage_floats, th_off, ld_off;
};

__device__ __forceinline__ RqsSmem carve(unsigned char* base, int tile_floats, int rows, int D, int d_tr, bool lu, int RQS_STAGES) {
  RqsSmem s;
  s.th_off = tile_floats;                       // rows % 4 == 0 -> every region starts 16-byte aligned
  s.ld_off = tile_floats + rows * D;
  s.stage_floats = tile_floats + rows * D + rows;
  float* f = reinterpret_cast<float*>(base);
  for (int i = 0; i < RQS_STAGES; ++i) { s.stage[i] = f; f += s.stage_floats; }
  s.ladbuf = f; f += rows * d_tr;
  s.luM = f; if (lu) f += D * D + D;           // M^T, then the (signed) pre-bias
  uintptr_t p = (reinterpret_cast<uintptr_t>(f) + 7) & ~uintptr_t(7);
  s.bars = reinterpret_cast<uint64_t*>(p);
  return s;
}

static size_t rqs_smem_bytes(int tile_floats, int rows, int D, int d_tr, bool lu) {
  const int RQS_STAGES = rqs_stages();
  size_t f = (size_t)RQS_STAGES * (tile_floats + rows * D + rows) + (size_t)rows * d_tr + (lu ? (size_t)D * D + D : 0);
  return f * 4 + 8 + RQS_STAGES * 8 + 16;
}

// Issue the loads of one tile into a stage: bulk TMA when byte counts and addresses allow it, otherwise a
// cooperative plain copy (the final ragged tile, or unaligned caller buffers).  Must be called by ALL threads.
__device__ __forceinline__ void load_chunk(float* dst, const float* src, int n_floats, uint64_t* bar, bool bulk) {
  if (bulk) {
    if (threadIdx.x == 0) bulk_g2s(dst, src, (uint32_t)n_floats * 4u, bar);
  } else {
    for (int i = threadIdx.x; i < n_floats; i += blockDim.x) dst[i] = src[i];
  }
}
__device__ __forceinline__ bool bulk_ok(const void* p, int n_floats) {
  return ((reinterpret_cast<uintptr_t>(p) & 15) == 0) && ((n_floats & 3) == 0);
}
__device__ __forceinline__ void load_tile(const RqsArgs& a, const RqsSmem& s, int st, int64_t tile, int row_floats) {
  int64_t r0 = tile * a.rows;
  int rows_here = (int)min((int64_t)a.rows, a.R - r0);
  const float* ps = a.params + r0 * row_floats;
  const float* ts = a.theta_in + r0 * a.D;
  const float* ls = a.logdet ? a.logdet + r0 : nullptr;
  const int np = rows_here * row_floats, nt = rows_here * a.D, nl = ls ? rows_here : 0;
  const bool bp = bulk_ok(ps, np), bt = bulk_ok(ts, nt), bl = ls && bulk_ok(ls, nl);
  uint32_t tx = (bp ? np : 0) + (bt ? nt : 0) + (bl ? nl : 0);
  uint64_t* bar = &s.bars[st];
  if (threadIdx.x == 0) {
    if (tx) mbar_expect_tx(bar, tx * 4u);
    else mbar_arrive(bar);                       // nothing asynchronous: complete the phase by hand
  }
  float* base = s.stage[st];
  load_chunk(base, ps, np, bar, bp);
  load_chunk(base + s.th_off, ts, nt, bar, bt);
  if (ls) load_chunk(base + s.ld_off, ls, nl, bar, bl);
}

template <bool INVERSE, int KT>
__global__ void __launch_bounds__(512) k_rqs_apply(RqsArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int row_floats = a.ldp;
  const int tile_floats = a.rows * row_floats;
  const bool lu = a.luM != nullptr;
  const int RQS_STAGES = a.stages;
  RqsSmem s = carve(smem_raw, tile_floats, a.rows, a.D, a.d_tr, lu, RQS_STAGES);
  const int64_t ntiles = (a.R + a.rows - 1) / a.rows;
  const int D = a.D, d_tr = a.d_tr;
  const float inv_dtr = 1.f / (float)d_tr, inv_D = 1.f / (float)D;

  if (threadIdx.x == 0) {
    for (int i = 0; i < RQS_STAGES; ++i) mbar_init(&s.bars[i], 1);
    mbar_fence_init();
  }
  // M is kept TRANSPOSED in shared memory: in phase 2 consecutive threads own consecutive output features f of a row,
  // so for a fixed k they read consecutive words M^T[k][f] (conflict-free) while x[k] is a broadcast; with the
  // row-major copy the stride between threads was D words (4-way bank conflicts at D = 40)
  if (lu) for (int i = threadIdx.x; i < D * D; i += blockDim.x) s.luM[(i % D) * D + i / D] = a.luM[i];
  float* const spre = s.luM + D * D;
  if (lu) for (int i = threadIdx.x; i < D; i += blockDim.x) spre[i] = a.luPre ? a.pre_sign * a.luPre[i] : 0.f;
  const float lu_ld = a.lu_logdet ? *a.lu_logdet : 0.f;
  __syncthreads();

  for (int st = 0; st < RQS_STAGES; ++st) {          // prologue: fill the ring
    int64_t tile = (int64_t)blockIdx.x + (int64_t)st * gridDim.x;
    if (tile < ntiles) load_tile(a, s, st, tile, row_floats);
  }

  int it = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int st = it % RQS_STAGES;
    const uint32_t parity = (uint32_t)(it / RQS_STAGES) & 1u;
    const int64_t r0 = tile * a.rows;
    const int rows_here = (int)min((int64_t)a.rows, a.R - r0);
    float* P = s.stage[st];
    float* rowbuf = P + s.th_off;
    float* ldbuf = P + s.ld_off;

    mbar_wait(&s.bars[st], parity);
    __syncthreads();                                 // (also orders the cooperative-copy fallback)

    // phase 1: one (sample, transformed feature) per thread; y overwrites x in the staged theta rows
    for (int e = threadIdx.x; e < rows_here * d_tr; e += blockDim.x) {
      int row = (int)(((float)e + 0.5f) * inv_dtr);  // e / d_tr, exact for e < 2^22
      int j = e - row * d_tr;
      int f = 2 * j + a.parity;
      float x = rowbuf[row * D + f];
      float y, lad; int bin; bool nd = false;
      float* pe = P + row * row_floats + j * a.c.P;
      if (KT > 0) {
        constexpr int KK = KT > 0 ? KT : 2;
        if (INVERSE) rqs_inverse_elem_t<KK>(pe, a.c, x, y, lad, bin, nd);
        else rqs_forward_elem_t<KK>(pe, a.c, x, y, lad, bin);
      } else {
        if (INVERSE) rqs_inverse_elem(pe, a.c, x, y, lad, bin, nd);
        else rqs_forward_elem(pe, a.c, x, y, lad, bin);
      }
      rowbuf[row * D + f] = y;
      s.ladbuf[e] = lad;
      if (a.bin_idx) a.bin_idx[(r0 + row) * d_tr + j] = bin;
      if (a.status) {
        if (nd) atomicAdd(&a.status[1], 1);
        if (!(fabsf(y) <= 3.0e38f) || !(fabsf(lad) <= 3.0e38f)) atomicAdd(&a.status[0], 1);
      }
    }
    __syncthreads();

    // phase 2: write rows (optionally through the LU-linear) and the per-sample log-det
    for (int i = threadIdx.x; i < rows_here * D; i += blockDim.x) {
      int row = (int)(((float)i + 0.5f) * inv_D);
      int f = i - row * D;
      float v = rowbuf[i];
      if (a.theta_sp) a.theta_sp[r0 * D + i] = v;
      if (lu) {
        const float* xr = rowbuf + row * D;
        const float* m = s.luM + f;              // column f of M^T = row f of M
        float acc = 0.f;
        if (a.luPre) {
          for (int k = 0; k < D; ++k) acc = fmaf(m[k * D], xr[k] + spre[k], acc);
        } else {
          for (int k = 0; k < D; ++k) acc = fmaf(m[k * D], xr[k], acc);
        }
        v = acc + (a.luPost ? a.luPost[f] : 0.f);
      }
      a.theta_out[r0 * D + i] = v;
    }
    if (a.logdet) {
      for (int row = threadIdx.x; row < rows_here; row += blockDim.x) {
        float sum = 0.f;
        for (int j = 0; j < d_tr; ++j) sum += s.ladbuf[row * d_tr + j];
        a.logdet[r0 + row] = ldbuf[row] + sum + lu_ld;
      }
    }
    __syncthreads();
    // stage `st` is fully consumed: refill it with the tile RQS_STAGES iterations ahead
    int64_t nxt = tile + (int64_t)RQS_STAGES * gridDim.x;
    if (nxt < ntiles) load_tile(a, s, st, nxt, row_floats);
  }
}

// ---- backward ----------------------------------------------------------------------------------
struct RqsBwdArgs {
  const float* theta_in;     // [R, D] layer input (pre-spline)
  const float* params;       // [R, d_tr, P]
  const float* d_theta_out;  // [R, D] grad wrt post-spline rows
  const float* d_logdet;     // [R]
  float* d_params;           // [R, d_tr, P]
  float* d_theta_in;         // [R, D]
  int64_t R;
  int D, d_tr, parity, rows, stages;
  int ldp;                   // row stride of params / d_params in floats
  RqsConsts c;
  // optional fused LULinear backward (y = M x + b after the spline): d_theta_out is then the gradient wrt y
  const float* luM;          // [D, D]  (NULL: no LU, d_theta_out is the gradient wrt the spline output)
  const float* theta_sp;     // [R, D]  spline output x saved by the forward
  float* dM;                 // [D, D] += dy^T x
  float* dbias;              // [D]    += colsum(dy)
};

template <int KT>
__global__ void __launch_bounds__(512) k_rqs_backward(RqsBwdArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int row_floats = a.ldp;
  const int tile_floats = a.rows * row_floats;
  const int RQS_STAGES = a.stages;
  float* stage[RQS_MAX_STAGES];
  {
    float* f = reinterpret_cast<float*>(smem_raw);
    for (int i = 0; i < RQS_STAGES; ++i) { stage[i] = f; f += tile_floats; }
  }
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (((size_t)RQS_STAGES * tile_floats * 4 + 7) & ~size_t(7)));
  const int64_t ntiles = (a.R + a.rows - 1) / a.rows;
  const int D = a.D, d_tr = a.d_tr;
  const bool lu = a.luM != nullptr;
  // LU scratch behind the barriers: M, dM accumulator, bias accumulator, dy tile, x tile, dx tile
  float* luM = reinterpret_cast<float*>(bars + RQS_STAGES);
  float* accM = luM + D * D;
  float* accb = accM + D * D;
  float* dyb = accb + D;
  float* xsb = dyb + a.rows * D;
  float* dxb = xsb + a.rows * D;

  if (threadIdx.x == 0) {
    for (int i = 0; i < RQS_STAGES; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  if (lu) {
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) { luM[i] = a.luM[i]; accM[i] = 0.f; }
    for (int i = threadIdx.x; i < D; i += blockDim.x) accb[i] = 0.f;
  }
  __syncthreads();

  auto load_params = [&](int st, int64_t tile) {
    int64_t r0 = tile * a.rows;
    int rows_here = (int)min((int64_t)a.rows, a.R - r0);
    const float* ps = a.params + r0 * row_floats;
    const int np = rows_here * row_floats;
    const bool bp = bulk_ok(ps, np);
    if (threadIdx.x == 0) {
      if (bp) mbar_expect_tx(&bars[st], (uint32_t)np * 4u);
      else mbar_arrive(&bars[st]);
    }
    load_chunk(stage[st], ps, np, &bars[st], bp);
  };
  for (int st = 0; st < RQS_STAGES; ++st) {
    int64_t tile = (int64_t)blockIdx.x + (int64_t)st * gridDim.x;
    if (tile < ntiles) load_params(st, tile);
  }

  int it = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int st = it % RQS_STAGES;
    const uint32_t parity = (uint32_t)(it / RQS_STAGES) & 1u;
    const int64_t r0 = tile * a.rows;
    const int rows_here = (int)min((int64_t)a.rows, a.R - r0);
    if (lu) {
      // LULinear backward on this tile: dx = M^T dy (feeds the spline backward), dM += dy^T x, dbias += colsum(dy)
      const int nrd = rows_here * D;
      for (int i = threadIdx.x; i < nrd; i += blockDim.x) {
        dyb[i] = a.d_theta_out[r0 * D + i];
        xsb[i] = a.theta_sp[r0 * D + i];
      }
      __syncthreads();
      for (int i = threadIdx.x; i < nrd; i += blockDim.x) {
        const int row = i / D, k = i - row * D;
        const float* dy = dyb + row * D;
        float acc = 0.f;
        for (int f = 0; f < D; ++f) acc = fmaf(luM[f * D + k], dy[f], acc);
        dxb[i] = acc;
      }
      for (int p = threadIdx.x; p < D * D; p += blockDim.x) {      // each thread owns fixed (f, k) entries
        const int f = p / D, k = p - f * D;
        float acc = 0.f;
        for (int row = 0; row < rows_here; ++row) acc = fmaf(dyb[row * D + f], xsb[row * D + k], acc);
        accM[p] += acc;
      }
      for (int f = threadIdx.x; f < D; f += blockDim.x) {
        float acc = 0.f;
        for (int row = 0; row < rows_here; ++row) acc += dyb[row * D + f];
        accb[f] += acc;
      }
    }
    mbar_wait(&bars[st], parity);
    __syncthreads();

    float* P = stage[st];
    for (int e = threadIdx.x; e < rows_here * d_tr; e += blockDim.x) {
      int row = e / d_tr, j = e - row * d_tr;
      int f = 2 * j + a.parity;
      int64_t gr = r0 + row;
      float* pe = P + row * row_floats + j * a.c.P;
      float x = a.theta_in[gr * D + f];
      float gy = lu ? dxb[row * D + f] : a.d_theta_out[gr * D + f];
      float gl = a.d_logdet[gr];
      // in place: every parameter is read before its gradient overwrites it
      float gx;
      if (KT > 0) {
        constexpr int KK = KT > 0 ? KT : 2;
        gx = rqs_backward_elem_t<KK>(pe, a.c, x, gy, gl);
      } else {
        gx = rqs_backward_elem(pe, a.c, x, gy, gl);
      }
      a.d_theta_in[gr * D + f] = gx;
    }
    // identity features: gradient passes straight through the coupling
    for (int i = threadIdx.x; i < rows_here * D; i += blockDim.x) {
      int f = i % D;
      if ((f & 1) != a.parity) a.d_theta_in[r0 * D + i] = lu ? dxb[i] : a.d_theta_out[r0 * D + i];
    }
    // write the gradient tile: bulk store when 16B-granular, else plain
    uint32_t bytes = (uint32_t)rows_here * row_floats * 4u;
    float* dst = a.d_params + r0 * row_floats;
    if (bulk_ok(dst, rows_here * row_floats)) {
      fence_proxy_async_smem();
      __syncthreads();
      if (threadIdx.x == 0) {
        bulk_s2g(dst, P, bytes);
        bulk_commit();
        bulk_wait_read<0>();   // smem may be refilled once the store has read it
      }
      __syncthreads();
    } else {
      __syncthreads();
      for (int i = threadIdx.x; i < rows_here * row_floats; i += blockDim.x) dst[i] = P[i];
      __syncthreads();
    }
    int64_t nxt = tile + (int64_t)RQS_STAGES * gridDim.x;
    if (nxt < ntiles) load_params(st, nxt);
  }
  if (threadIdx.x == 0) bulk_wait_all<0>();
  if (lu) {
    __syncthreads();
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) atomicAdd(&a.dM[i], accM[i]);
    for (int i = threadIdx.x; i < D; i += blockDim.x) atomicAdd(&a.dbias[i], accb[i]);
  }
}

// ---- host side ------------------------------------------------------------------------------------
RqsConsts make_rqs_consts(const fp_nsf_config* cfg) {
  RqsConsts c;
  c.K = cfg->K; c.P = 3 * cfg->K - 1;
  c.tb = cfg->tail_bound; c.span = (float)(2.0 * (double)cfg->tail_bound);
  c.mbw = cfg->min_bin_width; c.mbh = cfg->min_bin_height; c.md = cfg->min_derivative;
  c.mixw = (float)(1.0 - (double)cfg->min_bin_width * cfg->K);
  c.mixh = (float)(1.0 - (double)cfg->min_bin_height * cfg->K);
  c.sqrtH = (float)sqrt((double)cfg->H);
  c.inv_sqrtH = 1.0f / c.sqrtH;
  c.dconst = (float)log(exp(1.0 - (double)cfg->min_derivative) - 1.0);
  return c;
}

// rows per tile: multiple of 4 (keeps every full tile 16-byte granular for the bulk copy), sized to ~20 KB per
// stage so that 4-5 CTAs stay resident per SM (the kernels are issue-bound: occupancy hides the MUFU/LDS latency)
static int pick_rows(int row_floats, int d_tr) {
  static int kb = 0;
  if (kb == 0) { const char* e = getenv("FLOPPITY_RQS_TILE_KB"); kb = e ? atoi(e) : 20; if (kb < 2 || kb > 96) kb = 20; }
  int rows = (kb * 1024) / (row_floats * 4);
  rows = rows / 4 * 4;
  if (rows < 4) rows = 4;
  if (rows > 128) rows = 128;
  while (rows > 4 && rows * d_tr > 512) rows -= 4;   // one element per thread, at most 512 threads
  return rows;
}

static int pick_threads(int rows, int d_tr) {
  int t = (rows * d_tr + 31) / 32 * 32;
  if (t > 512) t = 512;
  if (t < 64) t = 64;
  return t;
}

int launch_rqs_apply(bool inverse, const fp_nsf_config* cfg, int parity, const float* theta_in, const float* params,
                     float* theta_out, float* theta_sp, float* logdet, int32_t* bin_idx, int64_t R, int32_t* status,
                     const float* luM, const float* luPre, float pre_sign, const float* luPost,
                     const float* lu_logdet, int ldp, cudaStream_t stream) {
  if (R <= 0) return 0;
  if (cfg->K < 2 || cfg->K > 64 || cfg->D < 2 || cfg->D > 64) return FP_E_BADARG;
  RqsArgs a{};
  a.theta_in = theta_in; a.params = params; a.theta_out = theta_out; a.theta_sp = theta_sp;
  a.logdet = logdet; a.bin_idx = bin_idx; a.status = status; a.R = R;
  a.D = cfg->D; a.parity = parity & 1;
  a.d_tr = (a.parity == 0) ? (cfg->D + 1) / 2 : cfg->D / 2;
  a.c = make_rqs_consts(cfg);
  a.luM = luM; a.luPre = luPre; a.luPost = luPost; a.lu_logdet = lu_logdet; a.pre_sign = pre_sign;
  if (ldp > 0 && ldp < a.d_tr * a.c.P) return FP_E_BADARG;
  a.ldp = ldp > 0 ? ldp : a.d_tr * a.c.P;
  int row_floats = a.ldp;
  a.rows = pick_rows(row_floats, a.d_tr);
  a.stages = rqs_stages();
  int threads = pick_threads(a.rows, a.d_tr);
  size_t smem = rqs_smem_bytes(a.rows * row_floats, a.rows, a.D, a.d_tr, luM != nullptr);
  void (*kern)(RqsArgs);
#define FP_RQS_PICK(KV) (inverse ? k_rqs_apply<true, KV> : k_rqs_apply<false, KV>)
  switch (cfg->K) {
    case 8: kern = FP_RQS_PICK(8); break;
    case 10: kern = FP_RQS_PICK(10); break;
    case 16: kern = FP_RQS_PICK(16); break;
    default: kern = FP_RQS_PICK(0); break;
  }
#undef FP_RQS_PICK
  FP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  FP_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
  if (occ < 1) occ = 1;
  int64_t ntiles = (R + a.rows - 1) / a.rows;
  int grid = (int)std::min<int64_t>(ntiles, (int64_t)num_sms() * occ);
  // algorithmic bytes: params + theta in + theta out (transformed features) + logdet
  ProfScope ps(inverse ? TAG_RQS_INV : TAG_RQS_FWD, 0.0, stream, 4.0 * (double)R * ((double)(a.d_tr * a.c.P) + 2.0 * a.d_tr + 1.0));
  kern<<<grid, threads, smem, stream>>>(a);
  FP_LAUNCH_CHECK();
  return 0;
}

int launch_rqs_backward(const fp_nsf_config* cfg, int parity, const float* theta_in, const float* params,
                        const float* d_theta_out, const float* d_logdet, float* d_params, float* d_theta_in,
                        int64_t R, int ldp, cudaStream_t stream, const float* luM, const float* theta_sp, float* dM,
                        float* dbias) {
  if (R <= 0) return 0;
  if (cfg->K < 2 || cfg->K > 64 || cfg->D < 2 || cfg->D > 64) return FP_E_BADARG;
  RqsBwdArgs a{};
  a.theta_in = theta_in; a.params = params; a.d_theta_out = d_theta_out; a.d_logdet = d_logdet;
  a.d_params = d_params; a.d_theta_in = d_theta_in; a.R = R;
  if (luM && (!theta_sp || !dM || !dbias)) return FP_E_BADARG;
  a.luM = luM; a.theta_sp = theta_sp; a.dM = dM; a.dbias = dbias;
  a.D = cfg->D; a.parity = parity & 1;
  a.d_tr = (a.parity == 0) ? (cfg->D + 1) / 2 : cfg->D / 2;
  a.c = make_rqs_consts(cfg);
  if (ldp > 0 && ldp < a.d_tr * a.c.P) return FP_E_BADARG;
  a.ldp = ldp > 0 ? ldp : a.d_tr * a.c.P;
  int row_floats = a.ldp;
  a.rows = pick_rows(row_floats, a.d_tr);
  a.stages = rqs_stages();
  const int RQS_STAGES = a.stages;
  int threads = pick_threads(a.rows, a.d_tr);
  size_t smem = (size_t)RQS_STAGES * a.rows * row_floats * 4 + 8 + RQS_STAGES * 8 + 16;
  smem += ((size_t)2 * a.D * a.D + a.D + (size_t)3 * a.rows * a.D) * 4;      // LU scratch (see the kernel)
  void (*kern)(RqsBwdArgs);
  switch (cfg->K) {
    case 8: kern = k_rqs_backward<8>; break;
    case 10: kern = k_rqs_backward<10>; break;
    case 16: kern = k_rqs_backward<16>; break;
    default: kern = k_rqs_backward<0>; break;
  }
  FP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  FP_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
  if (occ < 1) occ = 1;
  int64_t ntiles = (R + a.rows - 1) / a.rows;
  int grid = (int)std::min<int64_t>(ntiles, (int64_t)num_sms() * occ);
  ProfScope ps(TAG_RQS_BWD, 0.0, stream, 4.0 * (double)R * (2.0 * (a.d_tr * a.c.P) + 3.0 * a.d_tr + 1.0));
  kern<<<grid, threads, smem, stream>>>(a);
  FP_LAUNCH_CHECK();
  return 0;
}

}  // namespace fp

extern "C" {

int fp_rqs_forward(const fp_nsf_config* cfg, int32_t parity, const float* theta_in, const float* params,
                   float* theta_out, float* logdet, int32_t* bin_idx, int64_t R, int32_t* status, void* stream) {
  return fp::launch_rqs_apply(false, cfg, parity, theta_in, params, theta_out, nullptr, logdet, bin_idx, R, status,
                              nullptr, nullptr, 1.f, nullptr, nullptr, 0, (cudaStream_t)stream);
}

int fp_rqs_inverse(const fp_nsf_config* cfg, int32_t parity, const float* theta_in, const float* params,
                   float* theta_out, float* logdet, int32_t* bin_idx, int64_t R, int32_t* status, void* stream) {
  return fp::launch_rqs_apply(true, cfg, parity, theta_in, params, theta_out, nullptr, logdet, bin_idx, R, status,
                              nullptr, nullptr, 1.f, nullptr, nullptr, 0, (cudaStream_t)stream);
}

int fp_rqs_backward(const fp_nsf_config* cfg, int32_t parity, const float* theta_in, const float* params,
                    const float* d_theta_out, const float* d_logdet, float* d_params, float* d_theta_in,
                    int64_t R, void* stream) {
  return fp::launch_rqs_backward(cfg, parity, theta_in, params, d_theta_out, d_logdet, d_params, d_theta_in, R, 0,
                                 (cudaStream_t)stream, nullptr, nullptr, nullptr, nullptr);
}

}  // extern "C"
