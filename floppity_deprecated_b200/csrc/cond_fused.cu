// cond_fused.cu -- the whole ResidualNet conditioner of one coupling layer as ONE persistent tcgen05 kernel (hidden 128).
//
// Stands for nflows ResidualNet.forward (initial layer, Bk x ResidualBlock with context GLU, final layer), reached
// from /root/reference/modules.py:63 via sbi build_nsf; upstream runs it as ~10 eager ops per block
// (SURVEY.md 2.3 iii-iv), gemm_tc.cu runs it as 2*Bk+2 GEMM launches with every activation making a round trip
// through HBM/L2.  Here a CTA takes a tile of 128 flow rows through the whole chain
//     h0 = theta W0x^T + ctx0 ;  per block: t = relu(h) W1^T + b1 ; u = relu(t) W2^T + b2 ; h += u * sigmoid(ctx_k)
//     params = h Wf^T + bf            (sigmoid(ctx_k) arrives precomputed, once per context row: k_gate_sigmoid)
// with the activations never leaving the SM:
//   * the A operand of every GEMM lives in TENSOR MEMORY (tcgen05.mma with A from TMEM): the epilogue threads own one
//     row each, apply bias / context / GLU / relu, split the result into TF32 hi + lo (3xTF32, fp32-accurate) and
//     write it back with tcgen05.st as the next GEMM's operand -- no shared-memory round trip, no swizzle;
//   * the residual stream h stays in the epilogue threads' registers;
//   * weights (pre-split hi / lo, fp_nsf_prepare) stream from L2 through a TMA ring that runs ahead across GEMM and
//     tile boundaries;
//   * the only global traffic is theta in, the spline parameters out and -- in training mode -- the saved activations
//     (h_k, t_k, u_k) out, staged in shared memory and written by TMA tensor stores.
// TMEM map (512 columns): [0,128) hi*hi accumulator, [128,256) cross-term accumulator, [256,384) operand hi,
// [384,512) operand lo.  (tcgen05.mma kind::tf32 reads 32 bytes of every A row per instruction, so an MMA costs
// ~75 cycles for any N <= 128: narrower N units do not pay -- every GEMM of the chain is issued at N = 128.)
#include "common.cuh"
#include "prof.h"
#include "layout.h"
#include "tc_common.cuh"
#include "gemm_epilogue.cuh"
#include <stdlib.h>
#include <algorithm>

namespace fp {

constexpr int CF_BM = 128;
constexpr int CF_WSTAGES = 3;
constexpr int CF_TILE_BYTES = 128 * 128;                // one [128 rows x 32 floats] SWIZZLE_128B box
constexpr int CF_WSTAGE_BYTES = 2 * CF_TILE_BYTES;      // hi + lo
constexpr int CF_SLOT_BYTES = 2 * CF_TILE_BYTES;        // staging slot: 64 columns of 128 rows
constexpr int CF_NSLOTS = 4;                            // two per column half
constexpr int CF_THREADS = 320;                         // TMA producer, MMA issuer, 8 epilogue warps
constexpr int CF_MAXBK = 4;
constexpr int CF_SB_BLOCKS = 2;                         // blocks whose biases live in shared memory
constexpr int CF_SBIAS_BYTES = CF_SB_BLOCKS * 2 * 128 * 4;
constexpr int CF_SMEM_BYTES = CF_WSTAGES * CF_WSTAGE_BYTES + CF_NSLOTS * CF_SLOT_BYTES + CF_SBIAS_BYTES + 256 /*barriers*/;
constexpr uint32_t CF_T_ACC = 0, CF_T_AHI = 256, CF_T_ALO = 384;

struct CondFwdMaps {
  CUtensorMap w0h, w0l;       // W0x hi / lo   [128, D]
  CUtensorMap wh, wl;         // blocks: rows k*258 = W1_k, k*258+129 = W2_k   [Bk*258, 128]
  CUtensorMap wfh, wfl;       // Wf hi / lo    [dP, 128]
  CUtensorMap prm;            // spline parameters out [R, dP]
  CUtensorMap h[CF_MAXBK + 1], t[CF_MAXBK], u[CF_MAXBK];   // saved activations out [R, 128] (training)
};
struct CondFwdParams {
  const float* theta; int32_t D;
  const float* ctx; int64_t ldg; int32_t ctx_div;     // ctx_all + column of this layer's slot 0; slot k at +128 k
  const float* b1[CF_MAXBK]; const float* b2[CF_MAXBK]; const float* bf;
  int64_t R; int32_t Bk; int32_t dP; int32_t training;
  long long* trace;      // debug: SM-clock stamps of CTA 0 (fp_cf_set_trace)
  int32_t* status;       // optional status words; [2] counts operand pieces beyond fp16's range (3xFP16 mode)
  DropKey dk[CF_MAXBK];  // dropout mask keys of the blocks (training with dropout_p > 0; nflows ResidualBlock.dropout)
  int32_t drop_on;
};

// D[tmem] (+)= A[tmem] * B[smem desc], kind::tf32 (A: lane = row, one 32-bit column per k)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
        "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
        "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15]))
      : "memory");
}
// same with fp16 operands (two k per 32-bit TMEM column, K = 16 per instruction)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// One column half (4 warps, 128 threads) of the epilogue group stages 64-column pieces of a [128 x *] result in its
// own two shared-memory slots and hands them to the TMA store engine.
struct Stager {
  uint32_t slots;       // shared address of this half's two slots
  uint32_t op;          // store ops issued so far by this half
  int bar_id;           // named barrier of this half
  bool issuer;          // the one thread that talks to the TMA engine
  uint32_t row_off;     // this thread's row inside the tile * 128 bytes
  uint32_t swz;         // row & 7
  __device__ __forceinline__ uint32_t begin() {
    if (issuer) bulk_wait_read<1>();          // the op that used this slot two ops ago has been read out
    named_bar_sync(bar_id, 128);
    return slots + (op & 1u) * (uint32_t)CF_SLOT_BYTES;
  }
  // both slots at once (a phase with two results): every earlier store of this half has been read out
  __device__ __forceinline__ void begin2(uint32_t& s0, uint32_t& s1) {
    if (issuer) bulk_wait_read<0>();
    named_bar_sync(bar_id, 128);
    s0 = slots + (op & 1u) * (uint32_t)CF_SLOT_BYTES;
    s1 = slots + ((op + 1u) & 1u) * (uint32_t)CF_SLOT_BYTES;
  }
  // 16 consecutive floats of this thread's row: columns [piece*16, piece*16+16) of 32-column box `box`
  __device__ __forceinline__ void put16(uint32_t slot, int box, int piece, const float (&v)[16]) const {
    const uint32_t base = slot + (uint32_t)box * (uint32_t)CF_TILE_BYTES + row_off;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sts128(base + ((((uint32_t)(piece * 4 + j)) ^ swz) << 4), make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
  }
  // two 64-column results (both slots) in one hand-over
  __device__ __forceinline__ void end2(uint32_t s0, const CUtensorMap* m0, uint32_t s1, const CUtensorMap* m1, int col0, int row0) {
    fence_proxy_async_smem();
    named_bar_sync(bar_id, 128);
    if (issuer) {
      for (int b = 0; b < 2; ++b) tma_store_2d(m0, s0 + (uint32_t)b * (uint32_t)CF_TILE_BYTES, col0 + 32 * b, row0);
      bulk_commit();
      for (int b = 0; b < 2; ++b) tma_store_2d(m1, s1 + (uint32_t)b * (uint32_t)CF_TILE_BYTES, col0 + 32 * b, row0);
      bulk_commit();
    }
    op += 2;
  }
  __device__ __forceinline__ void end(uint32_t slot, const CUtensorMap* map, int col0, int row0, int nbox) {
    fence_proxy_async_smem();
    named_bar_sync(bar_id, 128);
    if (issuer) {
      for (int b = 0; b < nbox; ++b) tma_store_2d(map, slot + (uint32_t)b * (uint32_t)CF_TILE_BYTES, col0 + 32 * b, row0);
      bulk_commit();
    }
    ++op;
  }
};

// 16 operand values of this thread's row, K columns kcol .. kcol+15: split into the two operand parts and written to
// tensor memory.  TF32: one 32-bit column per k (hi | lo regions).  FP16: two k per column, residual scaled by 2^11.
template <bool F16>
__device__ __forceinline__ void put_operand16(uint32_t t_hi, uint32_t t_lo, uint32_t kcol, const float (&w)[16],
                                              int32_t* status) {
  if (F16) {
    // the conversions saturate at +-65504: count pieces that reach the edge of fp16's range so the host can say so
    // (z-scored inputs and O(1..10) activations never do; the host then re-runs with FLOPPITY_FUSED_FP16=0)
    float mx = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) mx = fmaxf(mx, fabsf(w[j]));
    if (mx > 6.0e4f && status) atomicAdd(status + 2, 1);
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split_f16x2(w[2 * j], w[2 * j + 1], hi[j], lo[j]);
    tmem_st_8(t_hi + (kcol >> 1), hi);
    tmem_st_8(t_lo + (kcol >> 1), lo);
  } else {
    float hi[16], lo[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { hi[j] = tf32_round(w[j]); lo[j] = w[j] - hi[j]; }
    tmem_st_16(t_hi + kcol, hi);
    tmem_st_16(t_lo + kcol, lo);
  }
}

#define CF_TRACE(slot) do { if (p.trace && blockIdx.x == 0 && (slot) < 512) p.trace[(slot)] = clock64(); } while (0)

template <bool F16, bool DROP>
__global__ void __launch_bounds__(CF_THREADS, 1)
k_cond_fwd(const __grid_constant__ CondFwdMaps maps, const CondFwdParams p) {
  constexpr int KB_COLS = F16 ? 64 : 32;        // contraction columns per weight stage (one 128-byte swizzle row)
  constexpr int NKB = 128 / KB_COLS;            // stages per hidden-128 contraction
  constexpr float XS = F16 ? FP16_LO_UNSCALE : 1.f;   // the cross-term accumulator carries the 2^11 of the fp16 residual
  extern __shared__ __align__(1024) uint8_t cf_smem[];     // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = cf_smem;
  const uint32_t smem_a = smem_u32(smem);
  const uint32_t slots_a = smem_a + CF_WSTAGES * CF_WSTAGE_BYTES;
  float* sbias = reinterpret_cast<float*>(smem + CF_WSTAGES * CF_WSTAGE_BYTES + CF_NSLOTS * CF_SLOT_BYTES);
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + CF_WSTAGES * CF_WSTAGE_BYTES + CF_NSLOTS * CF_SLOT_BYTES + CF_SBIAS_BYTES);
  uint64_t* w_empty = w_full + CF_WSTAGES;
  uint64_t* a_ready = w_empty + CF_WSTAGES;     // [1]
  uint64_t* acc_full = a_ready + 1;             // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_tiles = (p.R + CF_BM - 1) / CF_BM;
  const int Bk = p.Bk;
  const int nF = (p.dP + 127) / 128;            // 128-column units of the final layer
  const int nkb0 = F16 ? 1 : (p.D + 31) / 32;   // weight stages of the initial layer (K = D <= 64)
  const int npc0 = (p.D + 31) / 32;             // 16-column theta pieces per thread (1 or 2)

  if (threadIdx.x == 0) {
    for (int s = 0; s < CF_WSTAGES; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    mbar_init(a_ready, 8);
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 8); }
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w0h) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.wh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.wl) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.wfh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.prm) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // weights (fp_nsf_prepare), theta and the context projection come from earlier kernels in the stream
  if ((smem_a & 1023u) != 0u) __trap();     // the swizzled tiles rely on the declared alignment
  for (int e = threadIdx.x; e < CF_SB_BLOCKS * 256; e += CF_THREADS) {
    const int k = e >> 8, r = e & 255;
    sbias[e] = (k < Bk) ? ((r < 128) ? p.b1[k][r] : p.b2[k][r - 128]) : 0.f;
  }
  __syncthreads();

  if (warp == 0) {
    // ===== weight producer: the same sequence of k-blocks for every tile, running ahead through the ring =====
    if (lane == 0) {
      uint32_t it = 0;
      auto stage = [&](const CUtensorMap* mh, const CUtensorMap* ml, int col, int row, uint32_t bytes) {
        const int s = it % CF_WSTAGES;
        const uint32_t ph = (it / CF_WSTAGES) & 1u;
        mbar_wait(&w_empty[s], ph ^ 1u);
        uint8_t* sb = smem + s * CF_WSTAGE_BYTES;
        mbar_expect_tx(&w_full[s], bytes);
        tma_load_2d(sb, mh, col, row, &w_full[s]);
        tma_load_2d(sb + CF_TILE_BYTES, ml, col, row, &w_full[s]);
        ++it;
      };
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < nkb0; ++kb) stage(&maps.w0h, &maps.w0l, kb * KB_COLS, 0, 2 * CF_TILE_BYTES);
        for (int k = 0; k < Bk; ++k)
          for (int g = 0; g < 2; ++g)
            for (int kb = 0; kb < NKB; ++kb) stage(&maps.wh, &maps.wl, kb * KB_COLS, F16 ? (2 * k + g) * 128 : k * 258 + g * 129, 2 * CF_TILE_BYTES);
        for (int j = 0; j < nF; ++j)
          for (int kb = 0; kb < NKB; ++kb) stage(&maps.wfh, &maps.wfl, kb * KB_COLS, j * 128, 2 * CF_TILE_BYTES);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc128 = F16 ? make_idesc_f16(128, 128) : make_idesc_tf32(128, 128);
      const uint32_t t_ahi = tmem_base + CF_T_AHI, t_alo = tmem_base + CF_T_ALO;
      uint32_t it = 0, n_ar = 0, n_ae = 0;
      int pend = 0;          // final-layer units whose read-out has not been awaited yet
      // one k-block (ring stage) of a GEMM unit: nk k-steps of 8, operand columns from acol
      auto kblock = [&](uint32_t d_hh, uint32_t d_x, uint32_t acol, int nk, uint32_t idesc, bool first) {
        const int s = it % CF_WSTAGES;
        const uint32_t ph = (it / CF_WSTAGES) & 1u;
        mbar_wait(&w_full[s], ph);
        tcgen05_fence_after();
        uint8_t* sb = smem + s * CF_WSTAGE_BYTES;
        const uint64_t b_hi = make_kmajor_desc(sb), b_lo = make_kmajor_desc(sb + CF_TILE_BYTES);
        for (int k = 0; k < nk; ++k) {
          const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);
          const uint32_t acc = (!first || k > 0) ? 1u : 0u;
          // one k-step = 32 bytes of every operand row: 8 tf32 or 16 fp16 contraction indices = 8 operand columns
          if (F16) {
            umma_f16_ts(d_hh, t_ahi + acol + 8u * k, b_hi + adv, idesc, acc);
            umma_f16_ts(d_x, t_alo + acol + 8u * k, b_hi + adv, idesc, acc);
            umma_f16_ts(d_x, t_ahi + acol + 8u * k, b_lo + adv, idesc, 1u);
          } else {
            umma_tf32_ts(d_hh, t_ahi + acol + 8u * k, b_hi + adv, idesc, acc);
            umma_tf32_ts(d_x, t_alo + acol + 8u * k, b_hi + adv, idesc, acc);
            umma_tf32_ts(d_x, t_ahi + acol + 8u * k, b_lo + adv, idesc, 1u);
          }
        }
        umma_commit(&w_empty[s]);
        ++it;
      };
      auto wait_operand = [&]() {
        mbar_wait(a_ready, n_ar & 1u);
        ++n_ar;
        tcgen05_fence_after();
      };
      auto wait_drained = [&]() {
        mbar_wait(&acc_empty[0], n_ae & 1u);
        ++n_ae;
        --pend;
        tcgen05_fence_after();
      };
      const uint32_t d_hh = tmem_base + CF_T_ACC, d_x = d_hh + 128;
      int ts = 0;     // trace slot (MMA: 0..255)
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        wait_operand();                                   // theta tile split into the operand columns
        while (pend > 0) wait_drained();                  // (already true by program order: keeps the phase count in step)
        CF_TRACE(ts); ++ts;
        for (int kb = 0; kb < nkb0; ++kb)
          kblock(d_hh, d_x, 32u * kb, F16 ? (p.D + 15) / 16 : min(4, (p.D - 32 * kb + 7) / 8), idesc128, kb == 0);
        umma_commit(&acc_full[0]);
        CF_TRACE(ts); ++ts;
        for (int u = 0; u < 2 * Bk; ++u) {
          wait_operand();
          CF_TRACE(ts); ++ts;
          for (int kb = 0; kb < NKB; ++kb) kblock(d_hh, d_x, 32u * kb, 4, idesc128, kb == 0);
          umma_commit(&acc_full[0]);
          CF_TRACE(ts); ++ts;
        }
        wait_operand();                                   // final hidden state
        CF_TRACE(ts); ++ts;
        for (int j = 0; j < nF; ++j) {
          if (pend > 0) wait_drained();                   // the read-out of the previous unit has drained the accumulators
          for (int kb = 0; kb < NKB; ++kb) kblock(d_hh, d_x, 32u * kb, 4, idesc128, kb == 0);
          umma_commit(&acc_full[0]);
          ++pend;
          CF_TRACE(ts); ++ts;
        }
      }
    }
  } else {
    // ===== epilogue warps 2..9: TMEM lane quarter q (= warp % 4), column half hf; one row per thread =====
    // Every global operand of a phase (context / gate rows) is requested one 16-column piece ahead -- the first piece
    // before the accumulator is even complete -- and the biases sit in shared memory, so the serial chain
    // MMA -> read-out -> operand write-back has no exposed L2 round trips.
    const int ew = warp - 2;
    const int q = warp & 3;
    const int hf = ew >> 2;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const uint32_t t_acc = tmem_base + lane_sel + CF_T_ACC;
    const uint32_t t_ahi = tmem_base + lane_sel + CF_T_AHI, t_alo = tmem_base + lane_sel + CF_T_ALO;
    const bool training = p.training != 0;
    Stager st;
    st.slots = slots_a + (uint32_t)(2 * hf) * (uint32_t)CF_SLOT_BYTES;
    st.op = 0;
    st.bar_id = 1 + hf;
    st.issuer = ((ew & 3) == 0) && lane == 0;
    st.row_off = (uint32_t)(q * 32 + lane) * 128u;
    st.swz = (uint32_t)(lane & 7);
    const uint32_t cdiv = p.ctx_div <= 0 ? 0xFFFFFFFFu : (uint32_t)p.ctx_div;
    uint32_t n_af[2] = {0u, 0u};
    int es = 256;    // trace slot (epilogue warp 2: 256..511)
    auto wait_acc = [&](int b) {
      mbar_wait(&acc_full[b], n_af[b] & 1u);
      ++n_af[b];
      tcgen05_fence_after();
      if (threadIdx.x == 64) { CF_TRACE(es); ++es; }
    };
    auto operand_done = [&]() {
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
      if (threadIdx.x == 64) { CF_TRACE(es); ++es; }
    };
    auto ldg16 = [](const float* src, float4 (&g)[4]) {
#pragma unroll
      for (int j = 0; j < 4; ++j) g[j] = __ldg(reinterpret_cast<const float4*>(src + 4 * j));
    };
    // biases of the first CF_SB_BLOCKS blocks: [k][b1 | b2][128] in shared memory (broadcast reads)
    const uint32_t sbias_a = smem_u32(sbias) + (uint32_t)hf * 256u;
    float hres[4][16];      // this row's residual stream, columns hf*64 .. hf*64+63

    const float* ctx_row = nullptr;
    float4 aux[4], nxt[4];
    // theta tile -> operand columns: this thread's half of the 32 (D <= 32) or 64 (D <= 64) padded columns; also points
    // ctx_row at the tile's context row and requests its first piece.  Runs for the first tile before the loop and for
    // every further tile right after the previous tile's last accumulator has been drained (all MMAs that read the old
    // operand are complete by then), so the next tile's first GEMM overlaps the previous tile's last store.
    auto theta_operand = [&](int64_t tile) {
      const int64_t m = tile * CF_BM + q * 32 + lane;
      const int64_t mc = m < p.R ? m : p.R - 1;
      ctx_row = p.ctx + (int64_t)((uint32_t)mc / cdiv) * p.ldg + hf * 64;
      const int npc = npc0;                   // 16-column pieces per thread: 1 or 2
      ldg16(ctx_row, aux);                    // first context piece of the initial layer, ahead of its accumulator
      prefetch_l1(ctx_row + 32);
      for (int pc = 0; pc < npc; ++pc) {
        float w[16];
        const int c0 = hf * 16 * npc + pc * 16;
        const float* tp = p.theta + mc * p.D + c0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c0 + 4 * j < p.D) v = __ldg(reinterpret_cast<const float4*>(tp + 4 * j));
          w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
        }
        put_operand16<F16>(t_ahi, t_alo, (uint32_t)c0, w, p.status);
      }
      operand_done();
    };
    if ((int64_t)blockIdx.x < n_tiles) theta_operand(blockIdx.x);

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t row0 = tile * CF_BM;

      {  // ---- initial layer: h0 = acc + ctx slot 0
        uint32_t slot = 0;
        if (training) slot = st.begin();      // slot hand-shake while the MMAs run; the store itself after the hand-over
        wait_acc(0);
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {
          float v[16], x[16];
          const uint32_t col = (uint32_t)(hf * 64 + pc * 16);
          if (pc < 3) ldg16(ctx_row + (pc + 1) * 16, nxt);
          else ldg16(ctx_row + 128, nxt);     // first gate piece of block 0
          tmem_ld_2x16(t_acc + col, t_acc + 128u + col, v, x);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            v[4 * j] += XS * x[4 * j] + aux[j].x; v[4 * j + 1] += XS * x[4 * j + 1] + aux[j].y;
            v[4 * j + 2] += XS * x[4 * j + 2] + aux[j].z; v[4 * j + 3] += XS * x[4 * j + 3] + aux[j].w;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) { hres[pc][j] = v[j]; x[j] = fmaxf(v[j], 0.f); }
          if (training) st.put16(slot, pc >> 1, pc & 1, v);
          put_operand16<F16>(t_ahi, t_alo, col, x, p.status);
#pragma unroll
          for (int j = 0; j < 4; ++j) aux[j] = nxt[j];
        }
        operand_done();
        if (training) st.end(slot, &maps.h[0], hf * 64, (int)row0, 2);
        prefetch_l1(ctx_row + 128);           // gate rows of block 0 (used two phases from now)
        prefetch_l1(ctx_row + 128 + 32);
      }

      for (int k = 0; k < Bk; ++k) {
        const bool sb = k < CF_SB_BLOCKS;
        {  // ---- t = acc + b1 ; operand relu(t)            (aux keeps holding the first gate piece of this block)
          uint32_t slot = 0;
          if (training) slot = st.begin();
          wait_acc(0);
          const float* bp = p.b1[k] + hf * 64;
          const uint32_t sbp = sbias_a + (uint32_t)k * 1024u;
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            float v[16], x[16];
            const uint32_t col = (uint32_t)(hf * 64 + pc * 16);
            tmem_ld_2x16(t_acc + col, t_acc + 128u + col, v, x);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 c = sb ? lds128(sbp + (uint32_t)(pc * 64 + j * 16)) : __ldg(reinterpret_cast<const float4*>(bp + pc * 16 + 4 * j));
              v[4 * j] += XS * x[4 * j] + c.x; v[4 * j + 1] += XS * x[4 * j + 1] + c.y;
              v[4 * j + 2] += XS * x[4 * j + 2] + c.z; v[4 * j + 3] += XS * x[4 * j + 3] + c.w;
            }
            if (DROP) {
              // t' = dropout(t): element (row, col) of the [R, 128] activation, pairs of columns share one hash word
              // (gemm_epilogue.cuh drop_word); what is saved and what feeds the next GEMM both come from t'
              const uint64_t w0 = ((uint64_t)(row0 + q * 32 + lane) * 128u + col) >> 1;
              const float ds = p.dk[k].scale;
              const uint32_t thr = p.dk[k].thr;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint32_t w = drop_word(p.dk[k], w0 + (uint64_t)j);
                v[2 * j] = (w & 0xffffu) >= thr ? v[2 * j] * ds : 0.f;
                v[2 * j + 1] = (w >> 16) >= thr ? v[2 * j + 1] * ds : 0.f;
              }
            }
            if (training) st.put16(slot, pc >> 1, pc & 1, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = fmaxf(v[j], 0.f);
            put_operand16<F16>(t_ahi, t_alo, col, x, p.status);
          }
          operand_done();
          if (training) st.end(slot, &maps.t[k], hf * 64, (int)row0, 2);
        }
        {  // ---- u = acc + b2 ; h += u * gate (slot 1+k of ctx_all holds sigmoid(gate), k_gate_sigmoid) ; operand relu(h)
           //      (the final layer takes h itself)
          uint32_t slot_u = 0, slot_h = 0;
          if (training) st.begin2(slot_u, slot_h);
          wait_acc(0);
          const float* bp = p.b2[k] + hf * 64;
          const uint32_t sbp = sbias_a + (uint32_t)k * 1024u + 512u;
          const float* gp = ctx_row + 128 * (1 + k);
          const bool last = (k == Bk - 1);
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            float v[16], x[16];
            const uint32_t col = (uint32_t)(hf * 64 + pc * 16);
            if (pc < 3) ldg16(gp + (pc + 1) * 16, nxt);
            else if (!last) ldg16(gp + 128, nxt);       // first gate piece of the next block
            tmem_ld_2x16(t_acc + col, t_acc + 128u + col, v, x);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 c = sb ? lds128(sbp + (uint32_t)(pc * 64 + j * 16)) : __ldg(reinterpret_cast<const float4*>(bp + pc * 16 + 4 * j));
              v[4 * j] += XS * x[4 * j] + c.x; v[4 * j + 1] += XS * x[4 * j + 1] + c.y;
              v[4 * j + 2] += XS * x[4 * j + 2] + c.z; v[4 * j + 3] += XS * x[4 * j + 3] + c.w;
            }
            if (training) st.put16(slot_u, pc >> 1, pc & 1, v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              hres[pc][4 * j] = fmaf(v[4 * j], aux[j].x, hres[pc][4 * j]);
              hres[pc][4 * j + 1] = fmaf(v[4 * j + 1], aux[j].y, hres[pc][4 * j + 1]);
              hres[pc][4 * j + 2] = fmaf(v[4 * j + 2], aux[j].z, hres[pc][4 * j + 2]);
              hres[pc][4 * j + 3] = fmaf(v[4 * j + 3], aux[j].w, hres[pc][4 * j + 3]);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = last ? hres[pc][j] : fmaxf(hres[pc][j], 0.f);
            put_operand16<F16>(t_ahi, t_alo, col, x, p.status);
#pragma unroll
            for (int j = 0; j < 4; ++j) aux[j] = nxt[j];
          }
          operand_done();
          if (training) {
#pragma unroll
            for (int pc = 0; pc < 4; ++pc) st.put16(slot_h, pc >> 1, pc & 1, hres[pc]);
            st.end2(slot_u, &maps.u[k], slot_h, &maps.h[k + 1], hf * 64, (int)row0);
          }
          if (!last) { prefetch_l1(gp + 128); prefetch_l1(gp + 128 + 32); }    // gate rows of the next block
        }
      }

      // ---- final layer: 128-column units.  The residual registers are dead by now, so a unit's 64 columns are drained
      //      into registers at once and the accumulators handed back: bias, staging and the TMA store of unit j run
      //      under the MMAs of unit j+1.
      for (int j = 0; j < nF; ++j) {
        wait_acc(0);
        float o[4][16];
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {
          float x[16];
          const uint32_t col = (uint32_t)(hf * 64 + pc * 16);
          tmem_ld_2x16(t_acc + col, t_acc + 128u + col, o[pc], x);
#pragma unroll
          for (int e = 0; e < 16; ++e) o[pc][e] += XS * x[e];
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[0]);
        if (j == nF - 1 && tile + (int64_t)gridDim.x < n_tiles) theta_operand(tile + gridDim.x);
        const int n0 = j * 128 + hf * 64;
        if (n0 + 64 <= p.dP) {
#pragma unroll
          for (int pc = 0; pc < 4; ++pc)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float4 c = __ldg(reinterpret_cast<const float4*>(p.bf + n0 + pc * 16 + 4 * e));
              o[pc][4 * e] += c.x; o[pc][4 * e + 1] += c.y; o[pc][4 * e + 2] += c.z; o[pc][4 * e + 3] += c.w;
            }
        } else {
#pragma unroll
          for (int pc = 0; pc < 4; ++pc)
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int n = n0 + pc * 16 + e;
              o[pc][e] += (n < p.dP) ? __ldg(p.bf + n) : 0.f;
            }
        }
        const uint32_t slot = st.begin();
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) st.put16(slot, pc >> 1, pc & 1, o[pc]);
        st.end(slot, &maps.prm, n0, (int)row0, (n0 < p.dP ? 1 : 0) + (n0 + 32 < p.dP ? 1 : 0));
        if (threadIdx.x == 64) { CF_TRACE(es); ++es; }
      }
    }
    if (st.issuer) bulk_wait_all<0>();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static long long* g_cf_trace = nullptr;

bool fused_enabled();      // flow_engine.cu (fp_set_mode)
uint64_t layer_seed(uint64_t seed, int layer, int block);
// Operand format of the fused kernel: 3xFP16 (default) or 3xTF32 (FLOPPITY_FUSED_FP16=0 / fp_set_mode).  Both carry 22
// mantissa bits through three tensor-core products; fp16 operands take half the tensor-core time, half the TMEM columns
// and half the weight bytes.  fp16's exponent range (6e-5 .. 65504 before precision degrades / saturates) covers what
// flows through a conditioner's forward pass -- z-scored parameters, O(1..10) activations, O(1) weights; the gradient
// GEMMs, whose operands span many more orders of magnitude, stay on 3xTF32.
bool cond_fused_f16();

// fp16 hi / scaled-residual copies of one layer's matrices in the fused kernel's own layout (layout.h f16*)
__global__ void k_prep_f16(const float* __restrict__ params, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, fp_nsf_config cfg) {
  const Layout L(&cfg);
  const int i = blockIdx.y;
  const int H = L.H, D = L.D, dP = L.dP(i);
  const int64_t nW = (int64_t)2 * L.Bk * H * H, nF = L.f16WfRows(i) * H, n = L.f16LayerHalves(i);
  const int64_t base = L.f16LayerOff(i);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    float v = 0.f;
    if (e < nW) {
      const int m = (int)(e / ((int64_t)H * H));
      const int64_t r = e - (int64_t)m * H * H;
      v = params[((m & 1) ? L.w2(i, m >> 1) : L.w1(i, m >> 1)) + r];
    } else if (e < nW + nF) {
      const int64_t r = e - nW;
      v = (r / H < dP) ? params[L.wf(i) + r] : 0.f;
    } else {
      const int64_t r = e - nW - nF;
      const int row = (int)(r / 64), col = (int)(r % 64);
      v = (col < D) ? params[L.w0x(i) + (int64_t)row * D + col] : 0.f;
    }
    uint32_t h2, l2;
    split_f16x2(v, 0.f, h2, l2);
    hi[base + e] = (uint16_t)(h2 & 0xffffu);
    lo[base + e] = (uint16_t)(l2 & 0xffffu);
  }
}
bool cond_fwd_supported(const Layout& L, const float* prepared);
int launch_cond_prepare(const fp_nsf_config* cfg, const float* params, float* prepared, cudaStream_t stream) {
  Layout L(cfg);
  if (!cond_fused_f16() || !cond_fwd_supported(L, prepared)) return 0;
  ProfScope ps(TAG_OTHER, 0.0, stream);
  dim3 grid(64, L.T);
  k_prep_f16<<<grid, 256, 0, stream>>>(params, reinterpret_cast<uint16_t*>(prepared + L.prepF16Hi()),
                                       reinterpret_cast<uint16_t*>(prepared + L.prepF16Lo()), *cfg);
  FP_LAUNCH_CHECK();
  return 0;
}

// Shapes the fused kernel covers (a property of the flow, identical for every layer and every pass)
bool cond_fwd_supported(const Layout& L, const float* prepared) {
  if (!tc_enabled() || !fused_enabled() || !prepared) return false;
  return L.H == 128 && L.D <= 64 && (L.D & 3) == 0 && L.Bk >= 1 && L.Bk <= CF_MAXBK;
}

// Conditioner forward of layer i as one launch.  h/t/u: saved-activation buffers ([R,128] each) when training, else unused.
// Returns FP_E_NOTREADY when the shape is outside what the fused kernel covers (the caller then runs the GEMM chain).
int launch_cond_fwd(const fp_nsf_config* cfg, const Layout& L, const float* params, const float* prepared, int i,
                    const float* theta_i, const float* ctx_all, int64_t R, int ctx_div, float* const* h, float* const* t,
                    float* const* u, float* prm, bool training, float drop_p, uint64_t drop_seed, int32_t* status,
                    cudaStream_t stream) {
  if (!cond_fwd_supported(L, prepared)) return FP_E_NOTREADY;
  if (R < 1 || R > 2147483647LL - 256) return FP_E_TOOLARGE;
  if ((reinterpret_cast<uintptr_t>(ctx_all) & 15) || (reinterpret_cast<uintptr_t>(theta_i) & 15)) return FP_E_BADARG;
  const int dP = L.dP(i);
  CondFwdMaps maps;
  const bool f16 = cond_fused_f16();
  bool ok = true;
  if (f16) {
    const uint16_t* hi = reinterpret_cast<const uint16_t*>(prepared + L.prepF16Hi());
    const uint16_t* lo = reinterpret_cast<const uint16_t*>(prepared + L.prepF16Lo());
    ok = ok && get_tensor_map_f16(hi + L.f16W0(i), 128, 64, 64, &maps.w0h);
    ok = ok && get_tensor_map_f16(lo + L.f16W0(i), 128, 64, 64, &maps.w0l);
    ok = ok && get_tensor_map_f16(hi + L.f16W(i, 0, 0), (int64_t)2 * L.Bk * 128, 128, 128, &maps.wh);
    ok = ok && get_tensor_map_f16(lo + L.f16W(i, 0, 0), (int64_t)2 * L.Bk * 128, 128, 128, &maps.wl);
    ok = ok && get_tensor_map_f16(hi + L.f16Wf(i), dP, 128, 128, &maps.wfh);
    ok = ok && get_tensor_map_f16(lo + L.f16Wf(i), dP, 128, 128, &maps.wfl);
  } else {
    const float* hi = prepared + L.prepHi();
    const float* lo = prepared + L.prepLo();
    ok = ok && get_tensor_map(hi + L.w0x(i), 128, L.D, L.D, &maps.w0h);
    ok = ok && get_tensor_map(lo + L.w0x(i), 128, L.D, L.D, &maps.w0l);
    ok = ok && get_tensor_map(hi + L.w1(i, 0), (int64_t)L.Bk * 258, 128, 128, &maps.wh);
    ok = ok && get_tensor_map(lo + L.w1(i, 0), (int64_t)L.Bk * 258, 128, 128, &maps.wl);
    ok = ok && get_tensor_map(hi + L.wf(i), dP, 128, 128, &maps.wfh);
    ok = ok && get_tensor_map(lo + L.wf(i), dP, 128, 128, &maps.wfl);
  }
  ok = ok && get_tensor_map(prm, R, dP, L.ldP(i), &maps.prm);
  for (int k = 0; k <= CF_MAXBK; ++k) maps.h[k] = maps.prm;
  for (int k = 0; k < CF_MAXBK; ++k) { maps.t[k] = maps.prm; maps.u[k] = maps.prm; }
  if (training) {
    for (int k = 0; k <= L.Bk && ok; ++k) ok = get_tensor_map(h[k], R, 128, 128, &maps.h[k]);
    for (int k = 0; k < L.Bk && ok; ++k) ok = get_tensor_map(t[k], R, 128, 128, &maps.t[k]) && get_tensor_map(u[k], R, 128, 128, &maps.u[k]);
  }
  if (!ok) return FP_E_NOTREADY;
  CondFwdParams p{};
  p.theta = theta_i; p.D = L.D;
  p.ctx = ctx_all + L.ctxCol(i, 0); p.ldg = L.Nctx; p.ctx_div = ctx_div;
  for (int k = 0; k < L.Bk; ++k) { p.b1[k] = params + L.b1(i, k); p.b2[k] = params + L.b2(i, k); }
  p.bf = params + L.bf(i);
  p.R = R; p.Bk = L.Bk; p.dP = dP; p.training = training ? 1 : 0;
  p.trace = g_cf_trace;
  p.status = status;
  p.drop_on = (training && drop_p > 0.f) ? 1 : 0;
  for (int k = 0; k < L.Bk; ++k) p.dk[k] = make_drop_key(layer_seed(drop_seed, i, k), drop_p);

  static bool attr_set = false;
  if (!attr_set) {
    FP_CUDA_OK(cudaFuncSetAttribute(k_cond_fwd<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF_SMEM_BYTES));
    FP_CUDA_OK(cudaFuncSetAttribute(k_cond_fwd<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF_SMEM_BYTES));
    FP_CUDA_OK(cudaFuncSetAttribute(k_cond_fwd<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF_SMEM_BYTES));
    FP_CUDA_OK(cudaFuncSetAttribute(k_cond_fwd<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF_SMEM_BYTES));
    attr_set = true;
  }
  const int64_t tiles = (R + CF_BM - 1) / CF_BM;
  const int grid = (int)std::min<int64_t>(tiles, num_sms());
  const double flops = 2.0 * (double)R * ((double)L.D * 128 + (double)L.Bk * 2 * 128 * 128 + 128.0 * dP);
  const double bytes = 4.0 * ((double)R * (L.D + dP + (training ? (3.0 * L.Bk + 1) * 128 : 0.0)) +
                              (double)(128 * L.D + L.Bk * 2 * 128 * 128 + dP * 128) * 2);
  ProfScope ps(TAG_COND_FWD, flops, stream, bytes);
  static int pdl_on = -1;
  if (pdl_on < 0) { const char* e = getenv("FLOPPITY_PDL"); pdl_on = (e && e[0] == '0') ? 0 : 1; }
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)grid, 1, 1);
  lc.blockDim = dim3(CF_THREADS, 1, 1);
  lc.dynamicSmemBytes = CF_SMEM_BYTES;
  lc.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = pdl_on ? 1 : 0;
  if (p.drop_on) {
    if (f16) FP_CUDA_OK(cudaLaunchKernelEx(&lc, k_cond_fwd<true, true>, maps, p));
    else FP_CUDA_OK(cudaLaunchKernelEx(&lc, k_cond_fwd<false, true>, maps, p));
  } else {
    if (f16) FP_CUDA_OK(cudaLaunchKernelEx(&lc, k_cond_fwd<true, false>, maps, p));
    else FP_CUDA_OK(cudaLaunchKernelEx(&lc, k_cond_fwd<false, false>, maps, p));
  }
  FP_LAUNCH_CHECK();
  return 0;
}

}  // namespace fp

extern "C" {
// debug: device buffer of >= 512 int64 receiving SM-clock stamps of CTA 0 of the fused conditioner (NULL = off)
int fp_cf_set_trace(void* buf) { fp::g_cf_trace = reinterpret_cast<long long*>(buf); return 0; }
}
