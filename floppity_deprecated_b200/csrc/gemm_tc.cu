// gemm_tc.cu -- conditioner GEMMs on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), fp32-accurate.
//
// Same contract as the SIMT kernel (gemm_simt.cu): C[M,N] = pro(A)[M,K] * B[N,K]^T with the fused ResidualNet
// epilogues (nflows ResidualNet reached from /root/reference/modules.py:63).  The reference computes these layers in
// fp32; a single TF32/bf16 pass would miss the 1e-4 parity bar, so every product is evaluated as a three-product split
//     a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo                                              (error ~2^-21 |a||b|)
// accumulated in fp32 TMEM accumulators (hi*hi and the cross terms separately, see the kernel comment), in one of two
// operand formats:
//   3xFP16 (default, fp_set_gemm_format): x_hi = fp16(x), x_lo = fp16((x - x_hi) * 2^11); 64 contraction indices per 64 KB
//          stage, converted in place; gradient operands scaled by one power of two per backward pass; range guard.
//   3xTF32: x_hi = tf32(x), x_lo = x - x_hi; 32 contraction indices per stage.
//
// Persistent warp-specialised CTA (448 threads, 576 for 3xFP16; 128x128 output tiles; K in blocks of one 128-byte swizzle row):
//   warp 0      TMA producer: cp.async.bulk.tensor loads of the raw fp32 A tile and the pre-split B_hi / B_lo tiles
//               (weights are split once per optimiser step) into a 3-stage shared-memory ring (SWIZZLE_128B).
//   warp 1      MMA issuer: one elected thread issues 3 products x 4 k-steps of tcgen05.mma (M128 N128, K8 tf32 / K16 f16)
//               per stage from shared-memory descriptors; tcgen05.commit releases the stage / signals the epilogue.
//   warps 2..   splitter (4 warps, 8 for 3xFP16): relu prologue + hi/lo split of the A tile IN PLACE, fence.proxy.async,
//               arrive on the stage's "ready" mbarrier.
//   last 8      epilogue: tcgen05.ld TMEM -> registers -> shared (32-column slab) -> coalesced float4 stores with
//               bias / context-add / GLU-gate+residual / relu-mask applied; overlaps the next tile's mainloop
//               (two accumulator sets in TMEM).
#include "common.cuh"
#include "prof.h"
#include "gemm_epilogue.cuh"
#include "tc_common.cuh"
#include "layout.h"
#include <cuda.h>
#include <map>
#include <mutex>
#include <tuple>
#include <stdlib.h>
#include <algorithm>

namespace fp {

int launch_gemm(const fp_gemm_args* a, cudaStream_t stream);

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 32, TC_STAGES = 3;
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;          // 16 KB: 128 rows x 128 B
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;         // A(raw->hi), A_lo, B_hi, B_lo
constexpr int TC_THREADS = 448;                           // producer, MMA, 4 splitter warps, 8 epilogue warps
constexpr int TC_CS_WARP_BYTES = 32 * 128;                // epilogue staging: per warp 32 rows x 32 columns, XOR-swizzled
constexpr int TC_CS_BYTES = 8 * TC_CS_WARP_BYTES;         // eight epilogue warps
constexpr int TC_KCHUNK_BLOCKS = 32;                      // K chunk (1024) for long contractions, see k_gemm_tc_nt
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + TC_CS_BYTES + 1024 /*align*/ + 256 /*barriers*/;

#define FP_TRACE(slot) do { if (trace && blockIdx.x == 0) trace[(slot)] = clock64(); } while (0)

// Persistent kernel: CTA c processes output tiles c, c+grid, ... ; every role walks the same tile sequence.
// TMEM (512 columns) holds two accumulator sets {hi*hi, cross terms} so the epilogue of tile i overlaps the
// mainloop of tile i+1.  Tensor-core accumulation truncates, so its error grows with the number of chained
// MMAs: the big hi*hi products get their own accumulator (the small cross terms cannot disturb it) and
// contractions longer than 1024 are cut into chunks whose partial sums are added in fp32 by the epilogue
// (only for the plain/bias epilogue, which is all the long-K context projection needs).
// F16: 16-bit operand pieces (3xFP16, see tc_common.cuh): a stage holds 64 contraction indices -- the raw fp32 A tile as two
// 16 KB boxes, converted IN PLACE into the fp16 hi tile (+0) and the scaled-residual lo tile (+16 KB), and the pre-split
// fp16 B_hi / B_lo tiles (16 KB each): the same 64 KB stage carries twice the contraction depth, the MMAs of a k-block
// take half the tensor-pipe time, and the TMA moves 32 instead of 48 KB per 32 contraction indices.
constexpr int TC16_THREADS = TC_THREADS + 128;            // 3xFP16 variant: eight converting warps
// FMT: 0 = 3xTF32, 1 = 3xFP16 with A converted in the kernel, 2 = 3xFP16 with A PRE-SPLIT by the caller (fp16 hi / lo copies
// of an operand many tiles share, e.g. the context batch of the context projection: mapA / mapAl are its fp16 maps, the
// converting warps idle and the MMA warp waits for the TMA directly).
template <int EPI, bool PAIR, int FMT>
__global__ void __launch_bounds__(FMT == 1 ? TC16_THREADS : TC_THREADS, 1)
k_gemm_tc_nt(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAl,
             const __grid_constant__ CUtensorMap mapBh,
             const __grid_constant__ CUtensorMap mapBl, const fp_gemm_args g, int nkb_total, int nchunks, int vec_all,
             long long* trace, int dbg, int gfast) {
  // PAIR: the two CTAs of a cluster compute a 256 x 128 output block with cta_group::2 MMAs (rank 0 issues them): each CTA
  // loads its own 128 rows of A and only HALF of the B tile (64 rows), so a stage is 48 KB instead of 64 KB -- a
  // four-deep ring in the same 192 KB and a third less L2 -> shared-memory traffic per flop.  The single-CTA ring is
  // latency-bound on long contractions (tools/tc_trace_c3.py: ~3300 cycles around the ring for ~810 cycles of MMA work
  // per k-block, three stages).
  constexpr int BROWS = PAIR ? TC_BN / 2 : TC_BN;             // rows of the B tile this CTA holds
  constexpr int B_TILE = BROWS * TC_BK * 4;
  constexpr int STAGE = 2 * TC_TILE_BYTES + 2 * B_TILE;       // A (raw -> hi), A_lo, B_hi, B_lo
  constexpr int STAGES = TC_STAGES * TC_STAGE_BYTES / STAGE;
  constexpr bool F16 = FMT != 0, A16 = FMT == 2;
  constexpr int NSPLITW = FMT == 1 ? 8 : 4;                   // converting warps 2 .. 2 + NSPLITW - 1, then eight epilogue warps
  constexpr int EPIW0 = 2 + NSPLITW;
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_a = smem_u32(smem);                                  // shared-space address of the ring
  const uint32_t cs_a = smem_a + TC_STAGES * TC_STAGE_BYTES;              // epilogue staging slabs
  uint64_t* full_raw = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES + TC_CS_BYTES);
  uint64_t* ready = full_raw + STAGES;
  uint64_t* empty = ready + STAGES;
  uint64_t* acc_full = empty + STAGES;      // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  pdl_launch_dependents();      // the next kernel in the stream may start its own prologue as SMs free up
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int64_t cta_id = PAIR ? (int64_t)(blockIdx.x >> 1) : (int64_t)blockIdx.x;       // the pair walks ONE tile sequence
  const int64_t cta_step = PAIR ? (int64_t)(gridDim.x >> 1) : (int64_t)gridDim.x;
  constexpr int ROWS_M = PAIR ? 2 * TC_BM : TC_BM;               // output rows of one work item
  const int tiles_n = (g.N + TC_BN - 1) / TC_BN;
  const int64_t tiles_m = (g.M + ROWS_M - 1) / ROWS_M;
  const int64_t n_super = tiles_m * tiles_n;
  const int kb_per_chunk = (nkb_total + nchunks - 1) / nchunks;
  // rasterisation: concurrently running CTAs should share tiles of the LARGER operand stream, i.e. walk the
  // dimension of the smaller operand fastest (activations [M,K] vs weights [N,K]); the other one stays in L2
  // `gfast` tiles of the fast dimension form a GROUP that sweeps the whole slow dimension before the next group starts:
  // when neither operand fits the 126 MB L2 (c3's context projection: A 151 MB, B hi+lo 614 MB) the un-grouped order
  // re-read ALL of A from DRAM for every tile of N -- 47 GB of DRAM reads for 3.4 GB of algorithmic bytes (ncu,
  // profiles/r2n_ctx_gemm_ncu_full_summary.csv).  A group keeps its slice of the fast operand L2-resident.
  const bool m_fast = g.M <= (int64_t)g.N;
  const int64_t tiles_f = m_fast ? tiles_m : (int64_t)tiles_n, tiles_s = m_fast ? (int64_t)tiles_n : tiles_m;
  const int64_t gf = (gfast > 0 && gfast < tiles_f) ? gfast : tiles_f;
  auto tile_fs = [&](int64_t st, int64_t& f, int64_t& sl) {
    const int64_t per = gf * tiles_s;
    const int64_t gi = st / per, idx = st - gi * per;
    const int64_t gs = min(gf, tiles_f - gi * gf);
    f = gi * gf + idx % gs;
    sl = idx / gs;
  };
  auto tile_m0 = [&](int64_t st) -> int64_t { int64_t f, sl; tile_fs(st, f, sl); return (m_fast ? f : sl) * ROWS_M + (int64_t)rank * TC_BM; };
  auto tile_n0 = [&](int64_t st) -> int { int64_t f, sl; tile_fs(st, f, sl); return (int)(m_fast ? sl : f) * TC_BN; };

  if (threadIdx.x == 0) {
    FP_TRACE(500);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_raw[s], 1); mbar_init(&ready[s], PAIR ? 8 : NSPLITW); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], PAIR ? 16 : 8); }
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBl) : "memory");
  }
  if (warp == 1) { if (PAIR) tmem_alloc_pair(tmem_slot, 512); else tmem_alloc(tmem_slot, 512); }
  tcgen05_fence_before();
  if (PAIR) cluster_sync_all();      // both CTAs' barriers are initialised before either one signals the other's
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: everything above overlapped the previous kernel's tail; its results are only
  // touched below this point
  pdl_wait();
  if (threadIdx.x == 0) FP_TRACE(501);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t it = 0;   // global k-block counter -> ring slot / phase
      for (int64_t st = cta_id; st < n_super; st += cta_step) {
        const int64_t m0 = tile_m0(st);
        const int n0 = tile_n0(st) + (int)rank * BROWS;      // this CTA's rows of the B tile
        for (int kb = 0; kb < nkb_total; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          uint8_t* sb = smem + s * STAGE;
          if (A16) {
            mbar_expect_tx(&full_raw[s], 4 * TC_TILE_BYTES);
            tma_load_2d(sb, &mapA, kb * 64, (int)m0, &full_raw[s]);
            tma_load_2d(sb + TC_TILE_BYTES, &mapAl, kb * 64, (int)m0, &full_raw[s]);
            tma_load_2d(sb + 2 * TC_TILE_BYTES, &mapBh, kb * 64, n0, &full_raw[s]);
            tma_load_2d(sb + 3 * TC_TILE_BYTES, &mapBl, kb * 64, n0, &full_raw[s]);
          } else if (F16) {
            mbar_expect_tx(&full_raw[s], 4 * TC_TILE_BYTES);
            tma_load_2d(sb, &mapA, kb * 64, (int)m0, &full_raw[s]);
            tma_load_2d(sb + TC_TILE_BYTES, &mapA, kb * 64 + 32, (int)m0, &full_raw[s]);
            tma_load_2d(sb + 2 * TC_TILE_BYTES, &mapBh, kb * 64, n0, &full_raw[s]);
            tma_load_2d(sb + 3 * TC_TILE_BYTES, &mapBl, kb * 64, n0, &full_raw[s]);
          } else {
          mbar_expect_tx(&full_raw[s], TC_TILE_BYTES + ((dbg & 8) ? 1 : 2) * B_TILE);
          tma_load_2d(sb, &mapA, kb * TC_BK, (int)m0, &full_raw[s]);
          tma_load_2d(sb + 2 * TC_TILE_BYTES, &mapBh, kb * TC_BK, n0, &full_raw[s]);
          if (!(dbg & 8)) tma_load_2d(sb + 2 * TC_TILE_BYTES + B_TILE, &mapBl, kb * TC_BK, n0, &full_raw[s]);
          }
          if (it < 60) FP_TRACE(it * 8 + 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = F16 ? make_idesc_f16(ROWS_M, TC_BN) : make_idesc_tf32(ROWS_M, TC_BN);
      uint32_t it = 0, tile_it = 0;
      for (int64_t st = cta_id; st < n_super; st += cta_step) {
        for (int kc = 0; kc < nchunks; ++kc, ++tile_it) {
          const int buf = tile_it & 1;
          const uint32_t aph = (tile_it >> 1) & 1u;
          if (PAIR) mbar_wait_cluster(&acc_empty[buf], aph ^ 1u);
          else mbar_wait(&acc_empty[buf], aph ^ 1u);        // epilogue has drained this accumulator set
          tcgen05_fence_after();
          const uint32_t t_hh = tmem_base + (uint32_t)(buf * 256);
          const uint32_t t_x = t_hh + 128;
          const int kb0 = kc * kb_per_chunk;
          const int kb1 = min(nkb_total, kb0 + kb_per_chunk);
          for (int kb = kb0; kb < kb1; ++kb, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1u;
            if (PAIR) mbar_wait_cluster(&ready[s], ph);      // both CTAs' A tiles converted, both halves of B landed
            else mbar_wait(A16 ? &full_raw[s] : &ready[s], ph);
            tcgen05_fence_after();
            if (it < 60) FP_TRACE(it * 8 + 3);
            uint8_t* sb = smem + s * STAGE;
            const uint64_t a_hi = make_kmajor_desc(sb), a_lo = make_kmajor_desc(sb + TC_TILE_BYTES);
            const uint64_t b_hi = make_kmajor_desc(sb + 2 * TC_TILE_BYTES), b_lo = make_kmajor_desc(sb + 2 * TC_TILE_BYTES + B_TILE);
#pragma unroll
            for (int k = 0; k < TC_BK / 8; ++k) {
              const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);      // 32 bytes per UMMA_K inside the swizzle row
              const uint32_t acc = (kb > kb0 || k > 0) ? 1u : 0u;
              if (F16) {          // 32 bytes per k-step = 16 fp16 contraction indices
                umma_f16(t_hh, a_hi + adv, b_hi + adv, idesc, acc);
                if (!(dbg & 64)) umma_f16(t_x, a_lo + adv, b_hi + adv, idesc, acc);
                if (!(dbg & (4 | 64))) umma_f16(t_x, a_hi + adv, b_lo + adv, idesc, 1u);
              } else if (PAIR) {
                umma_tf32_pair(t_hh, a_hi + adv, b_hi + adv, idesc, acc);
                if (!(dbg & 64)) umma_tf32_pair(t_x, a_lo + adv, b_hi + adv, idesc, acc);
                if (!(dbg & (4 | 64))) umma_tf32_pair(t_x, a_hi + adv, b_lo + adv, idesc, 1u);
              } else {
                umma_tf32(t_hh, a_hi + adv, b_hi + adv, idesc, acc);
                if (!(dbg & 64)) umma_tf32(t_x, a_lo + adv, b_hi + adv, idesc, acc);      // 64: single TF32 pass (reduced precision)
                if (!(dbg & (4 | 64))) umma_tf32(t_x, a_hi + adv, b_lo + adv, idesc, 1u);
              }
            }
            if (PAIR) umma_commit_pair(&empty[s]);           // the ring slot is reusable in BOTH CTAs
            else umma_commit(&empty[s]);   // ring slot reusable once these MMAs have read it
            if (it < 60) FP_TRACE(it * 8 + 4);
          }
          if (PAIR) umma_commit_pair(&acc_full[buf]);
          else umma_commit(&acc_full[buf]);     // this accumulator set is complete
        }
      }
    }
  } else if (warp < EPIW0) {
    // ===== splitter warps: relu prologue + hi/lo split of the A tile, in place =====
    const int tid_s = threadIdx.x - 64;       // 0..127
    const bool relu_a = (g.flags & FP_G_RELU_A) != 0;
    const float a_s = (F16 && g.a_scale) ? __ldg(g.a_scale) : 1.f;      // power-of-two operand scale (gradient operands)
    const bool scaled = a_s != 1.f;
    float amax = 0.f;                                                     // fp16 range guard
    const int pf_fl = (EPI >= 0 ? EPI : g.flags) & (FP_G_GLU_RES | FP_G_RELUMASK | FP_G_ACCUM | FP_G_ADD_CTX);
    const int pf_bytes = (pf_fl && vec_all) ? TC_BN * 4 : 0;
    uint32_t it = 0;
    for (int64_t st = cta_id; st < n_super && !A16; st += cta_step) {
      for (int kb = 0; kb < nkb_total; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1u;
        if (kb == 0 && pf_bytes > 0) {
          // L2 prefetch of this tile's epilogue operands (one row segment per thread): the epilogue warps can only
          // keep ~4 KB of loads in flight each, so at HBM latency they are bandwidth-starved; at L2-hit latency not
          const int64_t m = tile_m0(st) + tid_s;
          const int n0p = tile_n0(st);
          const uint32_t nb = (uint32_t)min(pf_bytes, (g.N - n0p) * 4) & ~15u;
          if (m < g.M && nb > 0) {
            if (pf_fl & FP_G_GLU_RES) l2_prefetch(g.Rsd + m * g.ldr + n0p, nb);
            if ((pf_fl & (FP_G_GLU_RES | FP_G_ADD_CTX)) && (g.ldg & 3) == 0)      // gate / context slot of this row's context row
              l2_prefetch(g.G + (g.ctx_div > 0 ? m / g.ctx_div : (int64_t)0) * g.ldg + n0p, nb);
            if (pf_fl & FP_G_RELUMASK) l2_prefetch(g.Msk + m * g.ldm + n0p, nb);
            if (pf_fl & FP_G_ACCUM) l2_prefetch(g.C + m * g.ldc + n0p, nb);
          }
        }
        mbar_wait(&full_raw[s], ph);
        if (tid_s == 0 && it < 60) FP_TRACE(it * 8 + 1);
        const uint32_t aA = smem_a + (uint32_t)(s * STAGE) + (uint32_t)tid_s * 16u;
        float4 v[TC_TILE_BYTES / 16 / 128];
        if (F16) {
          // (row, 8-index chunk) items; a quarter-warp takes chunks 0-3 of an even row and chunks 4-7 of the next (or the
          // other way round): its eight 16-byte reads and its eight 16-byte writes each cover all 32 banks exactly once
          const uint32_t sA = smem_a + (uint32_t)(s * STAGE);
          const int q8 = tid_s >> 3, u = tid_s & 7;
          float4 raw[8];
          if (!(dbg & 1)) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = (i >> 1) * 64 + 2 * q8 + (u >> 2);
            const int cp = (u & 3) + 4 * ((u >> 2) ^ (i & 1));
            const uint32_t rowa = sA + (uint32_t)(cp >> 2) * (uint32_t)TC_TILE_BYTES + (uint32_t)r * 128u;
            const uint32_t cc = 2u * (uint32_t)(cp & 3), sw = (uint32_t)(r & 7);
            raw[2 * i] = lds128(rowa + ((cc ^ sw) << 4));
            raw[2 * i + 1] = lds128(rowa + (((cc + 1u) ^ sw) << 4));
          }
          }
          // hi / lo row r overwrite exactly the raw bytes of row r (both boxes), and a row pair belongs to ONE quarter-warp
          // over its two passes: the in-place hazard is warp-local
          __syncwarp();
          if (!(dbg & 1)) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = (i >> 1) * 64 + 2 * q8 + (u >> 2);
            const int cp = (u & 3) + 4 * ((u >> 2) ^ (i & 1));
            float4 a = raw[2 * i], b = raw[2 * i + 1];
            if (scaled) { a.x *= a_s; a.y *= a_s; a.z *= a_s; a.w *= a_s; b.x *= a_s; b.y *= a_s; b.z *= a_s; b.w *= a_s; }
            if (relu_a) {
              a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
              b.x = fmaxf(b.x, 0.f); b.y = fmaxf(b.y, 0.f); b.z = fmaxf(b.z, 0.f); b.w = fmaxf(b.w, 0.f);
            }
            amax = fmaxf(amax, fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))),
                                     fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w)))));
            uint4 hi, lo;
            split_f16x2(a.x, a.y, hi.x, lo.x); split_f16x2(a.z, a.w, hi.y, lo.y);
            split_f16x2(b.x, b.y, hi.z, lo.z); split_f16x2(b.z, b.w, hi.w, lo.w);
            const uint32_t dst = sA + (uint32_t)r * 128u + (((uint32_t)cp ^ (uint32_t)(r & 7)) << 4);
            sts128u(dst, hi);
            sts128u(dst + (uint32_t)TC_TILE_BYTES, lo);
          }
          }
        } else if (!(dbg & 1)) {
#pragma unroll
        for (int i = 0; i < TC_TILE_BYTES / 16 / 128; ++i) v[i] = lds128(aA + (uint32_t)i * 2048u);
#pragma unroll
        for (int i = 0; i < TC_TILE_BYTES / 16 / 128; ++i) {
          float4 w = v[i];
          if (relu_a) { w.x = fmaxf(w.x, 0.f); w.y = fmaxf(w.y, 0.f); w.z = fmaxf(w.z, 0.f); w.w = fmaxf(w.w, 0.f); }
          const float4 h = make_float4(tf32_round(w.x), tf32_round(w.y), tf32_round(w.z), tf32_round(w.w));
          sts128(aA + (uint32_t)i * 2048u, h);
          sts128(aA + (uint32_t)i * 2048u + TC_TILE_BYTES, make_float4(w.x - h.x, w.y - h.y, w.z - h.z, w.w - h.w));
        }
        }
        fence_proxy_async_smem();             // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_leader(&ready[s]); else mbar_arrive(&ready[s]); }
        if (tid_s == 0 && it < 60) FP_TRACE(it * 8 + 2);
      }
    }
    if (F16 && g.status && !(amax <= 6.0e4f)) atomicAdd(g.status + 2, 1);      // also counts NaN operands
  } else {
    // ===== epilogue warps 6..13: two warps per TMEM lane quarter, each owning one 64-column half =====
    // Warp-private: 32 accumulator rows x 32 columns go TMEM -> registers -> this warp's XOR-swizzled slab ->
    // registers with 8 lanes per row, so every global access of a warp covers four full 128-byte row segments.
    const int ew = warp - EPIW0;               // 0..7
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int half = ew >> 2;                  // column half
    const uint32_t slabw = cs_a + (uint32_t)ew * (uint32_t)TC_CS_WARP_BYTES;
    const uint32_t st_base = slabw + (uint32_t)lane * 128u;
    const uint32_t sw_st = (uint32_t)(lane & 7);
    const int rsub = lane >> 3;                // row within a group of four
    const int cch = lane & 7;                  // float4 chunk within the 32-column slab row
    const EpiArgs ea = make_epi(g);
    const bool single = (dbg & 64) != 0;       // single TF32 pass: the cross-term accumulator is never written
    // 3xFP16: the cross terms carry the 2^11 of the scaled residual; the operand scale is undone on the sum
    const float x_s = F16 ? FP16_LO_UNSCALE : 1.f;
    const float o_s = (F16 && g.a_scale) ? 1.f / __ldg(g.a_scale) : 1.f;
    const int fl_rt = g.flags & FP_EPI_MASK;
    const bool need_ctx = ((EPI >= 0 ? EPI : fl_rt) & (FP_G_ADD_CTX | FP_G_GLU_RES)) != 0;
    const uint32_t cdiv = ea.ctx_div <= 0 ? 0xFFFFFFFFu : (uint32_t)ea.ctx_div;
    // slab read addresses of this thread's eight rows (row r = rsub + 4 i, so r & 7 only depends on i & 1)
    const uint32_t ld_even = slabw + (uint32_t)rsub * 128u + ((uint32_t)(cch ^ rsub) << 4);
    const uint32_t ld_odd = slabw + (uint32_t)(rsub + 4) * 128u + ((uint32_t)(cch ^ (rsub + 4)) << 4);
    uint32_t tile_it = 0;
    for (int64_t st = cta_id; st < n_super; st += cta_step) {
      const int64_t m0 = tile_m0(st);
      const int n0 = tile_n0(st);
      const int64_t mrow = m0 + q * 32 + rsub;                              // first of this thread's 8 rows
      const bool full = (m0 + TC_BM <= ea.M);
      uint32_t cr8[8];                                                      // context rows of the 8 rows
      if (need_ctx) {
        uint32_t cr = (uint32_t)mrow / cdiv, rem = (uint32_t)mrow - cr * cdiv;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          cr8[i] = cr;
          rem += 4u;
          while (rem >= cdiv) { rem -= cdiv; ++cr; }
        }
      }
      for (int kc = 0; kc < nchunks; ++kc, ++tile_it) {
        const int buf = tile_it & 1;
        const uint32_t aph = (tile_it >> 1) & 1u;
        mbar_wait(&acc_full[buf], aph);
        tcgen05_fence_after();
        if (threadIdx.x == 192 && tile_it < 8) FP_TRACE(480 + tile_it * 2);
        const uint32_t t_hh = tmem_base + (uint32_t)(buf * 256) + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64);
#pragma unroll 1
        for (int chunk = 0; chunk < 2; ++chunk) {
          const int n = n0 + half * 64 + chunk * 32 + cch * 4;
          const bool vec = vec_all && (n + 3 < ea.N);
          const bool fast = full && vec && kc == 0;
          // the auxiliary operands of the first four rows are requested before the TMEM read-out
          EpiAux aux[4];
          float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (fast) {
            if ((EPI >= 0 ? EPI : fl_rt) & FP_G_BIAS) bias4 = __ldg(reinterpret_cast<const float4*>(ea.bias + n));
#pragma unroll
            for (int i = 0; i < 4; ++i) epi_load4<EPI>(ea, fl_rt, mrow + 4 * i, need_ctx ? (int64_t)cr8[i] : 0, n, aux[i]);
          }
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            float v[16], x[16];
            const uint32_t tc = t_hh + (uint32_t)(chunk * 32 + p * 16);
            tmem_ld_2x16(tc, tc + 128u, v, x);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts128(st_base + ((((uint32_t)(p * 4 + j)) ^ sw_st) << 4),
                     single ? make_float4(v[4 * j] * o_s, v[4 * j + 1] * o_s, v[4 * j + 2] * o_s, v[4 * j + 3] * o_s)
                     : F16  ? make_float4(fmaf(x[4 * j], x_s, v[4 * j]) * o_s, fmaf(x[4 * j + 1], x_s, v[4 * j + 1]) * o_s,
                                          fmaf(x[4 * j + 2], x_s, v[4 * j + 2]) * o_s, fmaf(x[4 * j + 3], x_s, v[4 * j + 3]) * o_s)
                            : make_float4(v[4 * j] + x[4 * j], v[4 * j + 1] + x[4 * j + 1], v[4 * j + 2] + x[4 * j + 2],
                                          v[4 * j + 3] + x[4 * j + 3]));
          }
          if (chunk == 1) {                       // all TMEM reads of this set are done: hand it back to the MMA warp
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) { if (PAIR) mbar_arrive_leader(&acc_empty[buf]); else mbar_arrive(&acc_empty[buf]); }
          }
          __syncwarp();
          if (dbg & 2) {
          } else if (fast) {
            // straight-line code, four independent rows at a time (the epilogue warps are few, so instruction-level
            // parallelism across rows is what hides the MUFU / shared / global latencies)
            float4 o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = lds128(((i & 1) ? ld_odd : ld_even) + (uint32_t)(i >> 1) * 1024u);
            EpiAux aux2[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) epi_load4<EPI>(ea, fl_rt, mrow + 16 + 4 * i, need_ctx ? (int64_t)cr8[4 + i] : 0, n, aux2[i]);
#pragma unroll
            for (int i = 0; i < 4; ++i) epi_store4<EPI>(ea, fl_rt, mrow + 4 * i, n, bias4, aux[i], o[i]);
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = lds128(((i & 1) ? ld_odd : ld_even) + (uint32_t)(2 + (i >> 1)) * 1024u);
#pragma unroll
            for (int i = 0; i < 4; ++i) epi_store4<EPI>(ea, fl_rt, mrow + 16 + 4 * i, n, bias4, aux2[i], o[i]);
          } else if (n < ea.N) {
            // ragged tiles, unaligned operands and the later chunks of a long contraction: run-time flags, row by row
            const int fl_s = (EPI >= 0 ? EPI : fl_rt);
            const int fl_g = (kc == 0) ? fl_s : ((fl_s & ~FP_G_BIAS) | FP_G_ACCUM);
            if (vec && (fl_g & FP_G_BIAS)) bias4 = __ldg(reinterpret_cast<const float4*>(ea.bias + n));
#pragma unroll 1
            for (int i = 0; i < 8; ++i) {
              const int r = rsub + 4 * i;
              const int64_t m = mrow + 4 * i;
              if (m < ea.M) {
                const float4 o = lds128(slabw + (uint32_t)r * 128u + ((uint32_t)(cch ^ (r & 7)) << 4));
                if (vec) {
                  EpiAux a1;
                  epi_load4<-1>(ea, fl_g, m, need_ctx ? (int64_t)cr8[i] : 0, n, a1);
                  epi_store4<-1>(ea, fl_g, m, n, bias4, a1, o);
                } else {
                  gemm_epilogue_store_slow(ea, fl_g, m, n, o, min(4, ea.N - n));
                }
              }
            }
          }
          __syncwarp();
        }
        if (threadIdx.x == 192 && tile_it < 8) FP_TRACE(480 + tile_it * 2 + 1);
      }
    }
  }
  tcgen05_fence_before();
  if (PAIR) cluster_sync_all();      // the leader's MMAs read the partner's shared memory until the last commit
  else __syncthreads();
  if (threadIdx.x == 0) FP_TRACE(502);
  if (warp == 1) { if (PAIR) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
}

// Stage geometry of the weight-gradient kernel.  (Measured: 16-row stages in a six-deep ring are SLOWER than 32-row
// stages in a three-deep one -- 15.1 vs 14.2 us for [40960 x 128]^T [40960 x 128] -- twice the TMA requests and barrier
// hand-offs outweigh the extra loads in flight; gpurun_out/tn_ablate*.log.)
constexpr int TN_BK = 32, TN_STAGES = 3;
constexpr int TN_TILE_BYTES = TC_BM * TN_BK * 4;           // 8 KB: 16 batch rows x 128 columns
constexpr int TN_STAGE_BYTES = 4 * TN_TILE_BYTES;          // A(raw->hi), A_lo, B(raw->hi), B_lo
static_assert(TN_STAGES * TN_STAGE_BYTES == TC_STAGES * TC_STAGE_BYTES, "the TN ring reuses the NT kernel's shared-memory budget");
static_assert(3 * TN_STAGES * 8 + 2 * 2 * 8 + 8 <= 256, "barrier block");

// ---------------------------------------------------------------------------------------------------------
// TN variant (weight gradients): C[M,N] += sum_r A[r,M] * B[r,N], both operands row-major over the batch index r
// (MN-major for the tensor core), contraction over r split across CTAs, partial tiles added with vector atomics.
// Work item = (output tile, k-split); every CTA walks items blockIdx.x, +gridDim.x, ...  Both operands are raw fp32
// activations, so the splitter warps split BOTH (relu prologue on B for the relu'd activation operand).
// F16: 3xFP16 operand pieces (see k_gemm_tc_nt): a stage holds 64 batch rows -- the raw fp32 tiles of A and B (four boxes of
// 32 columns x 64 rows each, 32 KB per operand) are converted IN PLACE into MN-major fp16 hi (first 16 KB) and scaled-residual
// lo (second 16 KB) tiles of two 64-column atoms; the four output rows of a batch row (hi / lo of two atoms) occupy exactly
// the bytes of its four raw rows, and a warp converts whole row pairs, so the in-place hazard is warp-local.
// B16 (with F16): the B operand (the activations every M-tile shares) was PRE-SPLIT by the caller into fp16 hi / lo copies:
// the TMA writes its two 64-column atoms straight into the MN-major operand layout (mapB = hi, mapBl = lo) and only A is
// converted.
template <bool F16, bool B16>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_gemm_tc_tn(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
             const __grid_constant__ CUtensorMap mapBl, const fp_gemm_args g,
             int nkb_total, int nsplit, int vec_all, int dbg, int gm) {
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_a = smem_u32(smem);                                  // shared-space address of the ring
  const uint32_t cs_a = smem_a + TN_STAGES * TN_STAGE_BYTES;              // epilogue staging slabs
  uint64_t* full_raw = reinterpret_cast<uint64_t*>(smem + TN_STAGES * TN_STAGE_BYTES + TC_CS_BYTES);
  uint64_t* ready = full_raw + TN_STAGES;
  uint64_t* empty = ready + TN_STAGES;
  uint64_t* acc_full = empty + TN_STAGES;      // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  pdl_launch_dependents();      // the next kernel in the stream may start its own prologue as SMs free up
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = (g.N + TC_BN - 1) / TC_BN;
  const int tiles_m = (int)((g.M + TC_BM - 1) / TC_BM);
  const int64_t n_items = (int64_t)tiles_m * tiles_n * nsplit;
  const int kb_per_split = (nkb_total + nsplit - 1) / nsplit;
  constexpr uint32_t GROUP_BYTES = TN_BK * 128;   // one 32-float MN group: TN_BK k-rows x 128 B

  if (threadIdx.x == 0) {
    for (int s = 0; s < TN_STAGES; ++s) { mbar_init(&full_raw[s], 1); mbar_init(&ready[s], 8); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: everything above overlapped the previous kernel's tail; its results are only
  // touched below this point
  pdl_wait();

  // Item order.  gm == 0 (operands that fit the L2): the splits of one output tile run side by side, the tile stays hot.
  // gm > 0 (c3's context weight gradient: A 2.9 GB, B 151 MB, C 307 MB): groups of gm M-tiles; inside a group the batch
  // range is the SLOW index, so all concurrently running CTAs read the same ~1000 batch rows (the 8 MB slab of B and the
  // group's slice of A stay in L2 while every tile of the group uses them) and the group's part of C stays in L2 across
  // the ranges.  The split-minor order re-read all of B from DRAM for every M-tile: 45 GB of reads for 3.7 GB.
  auto item_decode = [&](int64_t w, int& m0, int& n0, int& kb0, int& kb1) {
    int sp; int64_t t;
    if (gm > 0) {
      const int64_t per = (int64_t)gm * tiles_n * nsplit;
      const int64_t gi = w / per, r = w - gi * per;
      const int64_t gs = min((int64_t)gm, (int64_t)tiles_m - gi * gm);
      sp = (int)(r / (gs * tiles_n));
      const int64_t ti = r - (int64_t)sp * gs * tiles_n;
      t = (gi * gm + ti / tiles_n) * tiles_n + ti % tiles_n;
    } else {
      sp = (int)(w % nsplit);
      t = w / nsplit;
    }
    m0 = (int)(t / tiles_n) * TC_BM;
    n0 = (int)(t % tiles_n) * TC_BN;
    kb0 = sp * kb_per_split;
    kb1 = min(nkb_total, kb0 + kb_per_split);
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t w = blockIdx.x; w < n_items; w += gridDim.x) {
        int m0, n0, kb0, kb1; item_decode(w, m0, n0, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % TN_STAGES;
          const uint32_t ph = (it / TN_STAGES) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          uint8_t* sb = smem + s * TN_STAGE_BYTES;
          if (F16) {
            mbar_expect_tx(&full_raw[s], 4 * TN_TILE_BYTES);
#pragma unroll
            for (int gidx = 0; gidx < 4; ++gidx) {     // four boxes of 32 columns x 64 batch rows per operand
              tma_load_2d(sb + gidx * 8192, &mapA, m0 + gidx * 32, kb * 64, &full_raw[s]);
              if (!B16) tma_load_2d(sb + 2 * TN_TILE_BYTES + gidx * 8192, &mapB, n0 + gidx * 32, kb * 64, &full_raw[s]);
            }
            if (B16) {
#pragma unroll
              for (int at = 0; at < 2; ++at) {         // two atoms of 64 columns x 64 batch rows, hi and lo
                tma_load_2d(sb + 2 * TN_TILE_BYTES + at * 8192, &mapB, n0 + at * 64, kb * 64, &full_raw[s]);
                tma_load_2d(sb + 3 * TN_TILE_BYTES + at * 8192, &mapBl, n0 + at * 64, kb * 64, &full_raw[s]);
              }
            }
          } else {
          mbar_expect_tx(&full_raw[s], 2 * TN_TILE_BYTES);
#pragma unroll
          for (int gidx = 0; gidx < 4; ++gidx) {     // four 32-float MN groups per operand
            tma_load_2d(sb + gidx * GROUP_BYTES, &mapA, m0 + gidx * 32, kb * TN_BK, &full_raw[s]);
            tma_load_2d(sb + 2 * TN_TILE_BYTES + gidx * GROUP_BYTES, &mapB, n0 + gidx * 32, kb * TN_BK, &full_raw[s]);
          }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = F16 ? make_idesc_f16(TC_BM, TC_BN, 1, 1) : make_idesc_tf32(TC_BM, TC_BN, 1, 1);
      uint32_t it = 0, tile_it = 0;
      for (int64_t w = blockIdx.x; w < n_items; w += gridDim.x, ++tile_it) {
        int m0, n0, kb0, kb1; item_decode(w, m0, n0, kb0, kb1);
        if (kb0 >= kb1) { --tile_it; continue; }
        const int buf = tile_it & 1;
        const uint32_t aph = (tile_it >> 1) & 1u;
        mbar_wait(&acc_empty[buf], aph ^ 1u);
        tcgen05_fence_after();
        const uint32_t t_hh = tmem_base + (uint32_t)(buf * 256);
        const uint32_t t_x = t_hh + 128;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % TN_STAGES;
          const uint32_t ph = (it / TN_STAGES) & 1u;
          mbar_wait(&ready[s], ph);
          tcgen05_fence_after();
          uint8_t* sb = smem + s * TN_STAGE_BYTES;
          if (F16) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {                          // 16 batch rows per MMA = two 8-row groups of 1024 B
              const uint32_t koff = (uint32_t)k * 2048u;
              const uint64_t a_hi = make_mnmajor_desc16(sb + koff, 8192u);
              const uint64_t a_lo = make_mnmajor_desc16(sb + TN_TILE_BYTES + koff, 8192u);
              const uint64_t b_hi = make_mnmajor_desc16(sb + 2 * TN_TILE_BYTES + koff, 8192u);
              const uint64_t b_lo = make_mnmajor_desc16(sb + 3 * TN_TILE_BYTES + koff, 8192u);
              const uint32_t acc = (kb > kb0 || k > 0) ? 1u : 0u;
              umma_f16(t_hh, a_hi, b_hi, idesc, acc);
              if (!(dbg & 64)) umma_f16(t_x, a_lo, b_hi, idesc, acc);
              if (!(dbg & (4 | 64))) umma_f16(t_x, a_hi, b_lo, idesc, 1u);
            }
          } else
#pragma unroll
          for (int k = 0; k < TN_BK / 8; ++k) {
            const uint32_t koff = (uint32_t)k * 1024u;             // 8 k-rows of 128 B
            const uint64_t a_hi = make_mnmajor_desc(sb + koff, GROUP_BYTES);
            const uint64_t a_lo = make_mnmajor_desc(sb + TN_TILE_BYTES + koff, GROUP_BYTES);
            const uint64_t b_hi = make_mnmajor_desc(sb + 2 * TN_TILE_BYTES + koff, GROUP_BYTES);
            const uint64_t b_lo = make_mnmajor_desc(sb + 3 * TN_TILE_BYTES + koff, GROUP_BYTES);
            const uint32_t acc = (kb > kb0 || k > 0) ? 1u : 0u;
            umma_tf32(t_hh, a_hi, b_hi, idesc, acc);
            if (!(dbg & 64)) umma_tf32(t_x, a_lo, b_hi, idesc, acc);
            if (!(dbg & (4 | 64))) umma_tf32(t_x, a_hi, b_lo, idesc, 1u);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else if (warp < 10) {
    // splitter, eight warps (both operands are raw activations: twice the conversion work of the NT kernel's k-block; the
    // epilogue runs once per output tile and makes do with four warps): raw A at +0 -> hi in place, lo at +TILE; raw B at
    // +2*TILE -> hi in place, lo at +3*TILE.  Measured (tools/c3_gemm_ablate.py): ablating the conversion altogether makes
    // the kernel 23..31 % faster, but doubling the converting warps from four to eight changes nothing -- what the
    // conversion costs is its LATENCY inside the three-stage ring (TMA -> convert -> MMA -> free), not its throughput.
    const int tid_s = threadIdx.x - 64;       // 0..255
    const bool relu_b = (g.flags & FP_G_RELU_B) != 0;
    // FP_G_COLSUM_A: the bias gradient colsum(A) rides along.  This thread's eight 16-byte chunks of the raw A tile
    // are chunk (tid_s & 7) of k-row (tid_s >> 3) of successive MN groups (one group per pass); the 32-byte-atom swizzle XORs the
    // 32-byte chunk index with (k-row & 3), which does not depend on i, so the chunk is the same four logical
    // columns for every i & 1 and every k-block: one float4 accumulator per MN group.
    const bool colsum = (g.flags & FP_G_COLSUM_A) != 0;
    float* const cs_out = g.U;
    uint32_t it = 0;
    if (F16) {
      // One warp converts one PAIR of batch rows of one operand per pass (2 x 128 columns = 32 items of 8 columns): lane ->
      // atom a (64 columns), row of the pair, 8-column chunk cp; a quarter-warp takes chunks 0-3 of one row and 4-7 of the
      // other, so its eight 16-byte reads and writes cover all 32 banks once.  Lanes l and l ^ 12 own the same columns.
      const int sw_w = warp - 2;                                   // 0..7
      const int qw = lane >> 3, u = lane & 7;
      const int at = qw >> 1, rr = u >> 2;                         // atom, row within the pair
      const int cp = (u & 3) + 4 * (rr ^ (qw & 1));                // 8-column chunk within the atom
      const uint32_t src_off = (uint32_t)(2 * at + (cp >> 2)) * 8192u;      // raw box of these 8 columns
      const uint32_t cc = 2u * (uint32_t)(cp & 3);
      const float a_s = g.a_scale ? __ldg(g.a_scale) : 1.f;
      const bool scaled = a_s != 1.f;
      float amax = 0.f;
      for (int64_t w = blockIdx.x; w < n_items; w += gridDim.x) {
        int m0, n0, kb0, kb1; item_decode(w, m0, n0, kb0, kb1);
        float cs8[8];
        const bool do_cs = colsum && n0 == 0;                      // the tiles of the first column block carry the bias gradient
#pragma unroll
        for (int j = 0; j < 8; ++j) cs8[j] = 0.f;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % TN_STAGES;
          const uint32_t ph = (it / TN_STAGES) & 1u;
          mbar_wait(&full_raw[s], ph);
          const uint32_t sS = smem_a + (uint32_t)(s * TN_STAGE_BYTES);
          // pass p = 0..7: operand (p >> 2), row pair 4 * sw_w + (p & 3) + ... : 32 row pairs per operand over 8 warps
          constexpr int NP = B16 ? 4 : 8;
          float4 raw[2 * NP];
          if (!(dbg & 1)) {
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            const int row = 2 * (8 * (p & 3) + sw_w) + rr;
            const uint32_t ra = sS + (uint32_t)(p >> 2) * (uint32_t)(2 * TN_TILE_BYTES) + src_off + (uint32_t)row * 128u;
            const uint32_t sw = (uint32_t)(row & 7);
            raw[2 * p] = lds128(ra + ((cc ^ sw) << 4));
            raw[2 * p + 1] = lds128(ra + (((cc + 1u) ^ sw) << 4));
          }
          }
          __syncwarp();
          if (!(dbg & 1)) {
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            const int row = 2 * (8 * (p & 3) + sw_w) + rr;
            float4 a = raw[2 * p], b = raw[2 * p + 1];
            if (p < 4) {
              if (scaled) { a.x *= a_s; a.y *= a_s; a.z *= a_s; a.w *= a_s; b.x *= a_s; b.y *= a_s; b.z *= a_s; b.w *= a_s; }
              if (do_cs) {
                cs8[0] += a.x; cs8[1] += a.y; cs8[2] += a.z; cs8[3] += a.w; cs8[4] += b.x; cs8[5] += b.y; cs8[6] += b.z; cs8[7] += b.w;
              }
            } else if (relu_b) {
              a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
              b.x = fmaxf(b.x, 0.f); b.y = fmaxf(b.y, 0.f); b.z = fmaxf(b.z, 0.f); b.w = fmaxf(b.w, 0.f);
            }
            amax = fmaxf(amax, fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))),
                                     fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w)))));
            uint4 hi, lo;
            split_f16x2(a.x, a.y, hi.x, lo.x); split_f16x2(a.z, a.w, hi.y, lo.y);
            split_f16x2(b.x, b.y, hi.z, lo.z); split_f16x2(b.z, b.w, hi.w, lo.w);
            const uint32_t dst = sS + (uint32_t)(p >> 2) * (uint32_t)(2 * TN_TILE_BYTES) + (uint32_t)at * 8192u + (uint32_t)row * 128u +
                                 (((uint32_t)cp ^ (uint32_t)(row & 7)) << 4);
            sts128u(dst, hi);
            sts128u(dst + (uint32_t)TN_TILE_BYTES, lo);
          }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ready[s]);
        }
        if (do_cs && kb0 < kb1) {
          const float inv = 1.f / a_s;
#pragma unroll
          for (int j = 0; j < 8; ++j) cs8[j] = (cs8[j] + __shfl_xor_sync(0xffffffffu, cs8[j], 12)) * inv;
          if (!(qw & 1)) {                                          // one lane of each (l, l ^ 12) pair
            const int m = m0 + 64 * at + 8 * cp;
            if (vec_all && m + 7 < (int)g.M) {
              red_add_f32x4(cs_out + m, make_float4(cs8[0], cs8[1], cs8[2], cs8[3]));
              red_add_f32x4(cs_out + m + 4, make_float4(cs8[4], cs8[5], cs8[6], cs8[7]));
            } else {
              for (int j = 0; j < 8; ++j) if (m + j < (int)g.M) red_add_f32(cs_out + m + j, cs8[j]);
            }
          }
        }
      }
      if (g.status && !(amax <= 6.0e4f)) atomicAdd(g.status + 2, 1);
    } else
    for (int64_t w = blockIdx.x; w < n_items; w += gridDim.x) {
      int m0, n0, kb0, kb1; item_decode(w, m0, n0, kb0, kb1);
      float4 cs[4];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) cs[q4] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % TN_STAGES;
        const uint32_t ph = (it / TN_STAGES) & 1u;
        mbar_wait(&full_raw[s], ph);
        const uint32_t aA = smem_a + (uint32_t)(s * TN_STAGE_BYTES) + (uint32_t)tid_s * 16u;
        if (!(dbg & 1))
#pragma unroll
        for (int i = 0; i < TN_TILE_BYTES / 16 / 256; ++i) {
          const uint32_t off = aA + (uint32_t)i * 4096u;
          const float4 v = lds128(off);
          float4 u = lds128(off + 2 * TN_TILE_BYTES);
          constexpr int IPG = (int)(GROUP_BYTES / 4096u);      // iterations per 32-float MN group
          static_assert(IPG == 1, "one pass of the 256 splitter threads covers one 32-float MN group");
          cs[i / IPG].x += v.x; cs[i / IPG].y += v.y; cs[i / IPG].z += v.z; cs[i / IPG].w += v.w;
          const float4 h = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
          sts128(off, h);
          sts128(off + TN_TILE_BYTES, make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w));
          if (relu_b) { u.x = fmaxf(u.x, 0.f); u.y = fmaxf(u.y, 0.f); u.z = fmaxf(u.z, 0.f); u.w = fmaxf(u.w, 0.f); }
          const float4 hb = make_float4(tf32_round(u.x), tf32_round(u.y), tf32_round(u.z), tf32_round(u.w));
          sts128(off + 2 * TN_TILE_BYTES, hb);
          sts128(off + 3 * TN_TILE_BYTES, make_float4(u.x - hb.x, u.y - hb.y, u.z - hb.z, u.w - hb.w));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[s]);
      }
      if (colsum && n0 == 0 && kb0 < kb1) {
        // lanes l, l^10, l^20, l^30 hold the same logical columns (k-rows differing in their low two bits)
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          float4 c = cs[q4];
#pragma unroll
          for (int sh = 10; sh <= 20; sh += 10) {
            c.x += __shfl_xor_sync(0xffffffffu, c.x, sh); c.y += __shfl_xor_sync(0xffffffffu, c.y, sh);
            c.z += __shfl_xor_sync(0xffffffffu, c.z, sh); c.w += __shfl_xor_sync(0xffffffffu, c.w, sh);
          }
          if (lane < 8) {               // k-row & 3 == 0: physical chunk == logical chunk
            const int m = m0 + q4 * 32 + lane * 4;
            if (vec_all && m + 3 < (int)g.M) red_add_f32x4(cs_out + m, c);
            else {
              const float e[4] = {c.x, c.y, c.z, c.w};
              for (int j = 0; j < 4; ++j) if (m + j < (int)g.M) red_add_f32(cs_out + m + j, e[j]);
            }
          }
        }
      }
    }
  } else {
    // epilogue: accumulate the partial tile into C with vector reductions (warp-private staging, full 128-byte
    // row segments per access -- same scheme as the NT kernel)
    const int ew = warp - 10;                  // 0..3: one warp per TMEM lane quarter, both column halves in turn
    const int q = warp & 3;
    const uint32_t slabw = cs_a + (uint32_t)ew * (uint32_t)TC_CS_WARP_BYTES;
    const uint32_t st_base = slabw + (uint32_t)lane * 128u;
    const uint32_t sw_st = (uint32_t)(lane & 7);
    const int rsub = lane >> 3, cch = lane & 7;
    float* const gC = g.C; const int64_t g_ldc = g.ldc, gM = g.M; const int gN = g.N;
    const bool single = (dbg & 64) != 0;
    const float x_s = F16 ? FP16_LO_UNSCALE : 1.f;
    const float o_s = (F16 && g.a_scale) ? 1.f / __ldg(g.a_scale) : 1.f;
    uint32_t tile_it = 0;
    for (int64_t w = blockIdx.x; w < n_items; w += gridDim.x, ++tile_it) {
      int m0, n0, kb0, kb1; item_decode(w, m0, n0, kb0, kb1);
      if (kb0 >= kb1) { --tile_it; continue; }
      const int buf = tile_it & 1;
      const uint32_t aph = (tile_it >> 1) & 1u;
      mbar_wait(&acc_full[buf], aph);
      tcgen05_fence_after();
      const uint32_t t_hh0 = tmem_base + (uint32_t)(buf * 256) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c4 = 0; c4 < 4; ++c4) {
        const int half = c4 >> 1, chunk = c4 & 1;
        const uint32_t t_hh = t_hh0 + (uint32_t)(half * 64);
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          float v[16], x[16];
          const uint32_t tc = t_hh + (uint32_t)(chunk * 32 + p * 16);
          tmem_ld_2x16(tc, tc + 128u, v, x);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128(st_base + ((((uint32_t)(p * 4 + j)) ^ sw_st) << 4),
                   single ? make_float4(v[4 * j] * o_s, v[4 * j + 1] * o_s, v[4 * j + 2] * o_s, v[4 * j + 3] * o_s)
                   : F16  ? make_float4(fmaf(x[4 * j], x_s, v[4 * j]) * o_s, fmaf(x[4 * j + 1], x_s, v[4 * j + 1]) * o_s,
                                        fmaf(x[4 * j + 2], x_s, v[4 * j + 2]) * o_s, fmaf(x[4 * j + 3], x_s, v[4 * j + 3]) * o_s)
                          : make_float4(v[4 * j] + x[4 * j], v[4 * j + 1] + x[4 * j + 1], v[4 * j + 2] + x[4 * j + 2],
                                        v[4 * j + 3] + x[4 * j + 3]));
        }
        if (c4 == 3) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        __syncwarp();
        const int n = n0 + half * 64 + chunk * 32 + cch * 4;
        const bool vec = vec_all && (n + 3 < gN);
        const int64_t mrow = (int64_t)m0 + q * 32 + rsub;
        if (dbg & 2) {
        } else if (vec && m0 + TC_BM <= gM) {            // full tile: straight-line, four rows in flight
          float* cp = gC + mrow * g_ldc + n;
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            float4 o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = rsub + 4 * (4 * b + i);
              o[i] = lds128(slabw + (uint32_t)r * 128u + ((uint32_t)(cch ^ (r & 7)) << 4));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) red_add_f32x4(cp + (int64_t)(4 * (4 * b + i)) * g_ldc, o[i]);
          }
        } else if (n < gN) {
#pragma unroll 1
          for (int i = 0; i < 8; ++i) {
            const int r = rsub + 4 * i;
            const int64_t m = mrow + 4 * i;
            if (m < gM) {
              const float4 o = lds128(slabw + (uint32_t)r * 128u + ((uint32_t)(cch ^ (r & 7)) << 4));
              float* cp = gC + m * g_ldc + n;
              if (vec) {
                red_add_f32x4(cp, o);
              } else {
                const float e[4] = {o.x, o.y, o.z, o.w};
                for (int j = 0; j < 4 && n + j < gN; ++j) red_add_f32(cp + j, e[j]);
              }
            }
          }
        }
        __syncwarp();
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---- weight / activation split helper: hi = tf32(x), lo = x - hi ---------------------------------------
__global__ void k_split_tf32(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = x[i];
    float h = tf32_round(v);
    hi[i] = h;
    lo[i] = v - h;
  }
}

// fp16 pieces for the 3xFP16 GEMMs: hi = fp16(x), lo = fp16((x - hi) * 2^11)
__global__ void k_split_f16(const float* __restrict__ x, uint32_t* __restrict__ hi2, uint32_t* __restrict__ lo2, int64_t n,
                            int32_t* __restrict__ status) {
  const int64_t np = n >> 1;
  float amax = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 v = *reinterpret_cast<const float2*>(x + 2 * i);
    uint32_t h, l;
    split_f16x2(v.x, v.y, h, l);
    hi2[i] = h; lo2[i] = l;
    amax = fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.y)));
  }
  if (status && !(amax <= 6.0e4f)) atomicAdd(status + 2, 1);
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    uint32_t h, l;
    split_f16x2(x[n - 1], 0.f, h, l);
    reinterpret_cast<uint16_t*>(hi2)[n - 1] = (uint16_t)(h & 0xFFFFu);
    reinterpret_cast<uint16_t*>(lo2)[n - 1] = (uint16_t)(l & 0xFFFFu);
  }
}

// fp16 pieces of the K-major weight matrices the forward GEMMs read: blockIdx.y = layer * (2 Bk + 2) + {2b: W1_b, 2b+1: W2_b,
// 2Bk: W0x, 2Bk+1: Wf}; the copy of the tensor at float offset woff starts at half index 2 * woff (layout.h prepG16Hi)
__global__ void k_split_f16_weights(const float* __restrict__ params, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                    fp_nsf_config cfg) {
  const Layout L(&cfg);
  const int per = 2 * L.Bk + 2;
  const int layer = blockIdx.y / per, which = blockIdx.y % per;
  int64_t woff, n;
  if (which == per - 1) { woff = L.wf(layer); n = (int64_t)L.dP(layer) * L.H; }
  else if (which == per - 2) { woff = L.w0x(layer); n = (int64_t)L.H * L.D; }
  else { woff = (which & 1) ? L.w2(layer, which >> 1) : L.w1(layer, which >> 1); n = (int64_t)L.H * L.H; }
  if (which == per - 2 && (L.D & 7) != 0) {
    // W0x [H, D] with D not a multiple of 8: rows padded to ld16(D) halves (zeros), when the region has room (D >= 4)
    if (L.D < 4) return;
    const int ld = Layout::ld16(L.D);
    const int64_t np = (int64_t)L.H * ld;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (int64_t)gridDim.x * blockDim.x) {
      const int r = (int)(i / ld), c = (int)(i % ld);
      uint32_t h, l;
      split_f16x2(c < L.D ? params[woff + (int64_t)r * L.D + c] : 0.f, 0.f, h, l);
      hi[2 * woff + i] = (uint16_t)(h & 0xFFFFu);
      lo[2 * woff + i] = (uint16_t)(l & 0xFFFFu);
    }
    return;
  }
  uint32_t* hi2 = reinterpret_cast<uint32_t*>(hi + 2 * woff);
  uint32_t* lo2 = reinterpret_cast<uint32_t*>(lo + 2 * woff);
  const float* x = params + woff;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; 2 * i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float a = x[2 * i], b = (2 * i + 1 < n) ? x[2 * i + 1] : 0.f;
    uint32_t h, l;
    split_f16x2(a, b, h, l);
    hi2[i] = h; lo2[i] = l;
  }
}

// transpose + split of every weight matrix the dgrad GEMMs need, one launch: blockIdx.y = tensor id
// (layer * (2*Bk+2) + {0: W0x, 1+2b: W1_b, 2+2b: W2_b, 2Bk+1: Wf}), blockIdx.x = 32x32 tile
// F16: fp16 hi / scaled-residual lo at half index 2 * dst (see layout.h prepGT16Hi), leading dimension rounded up to 8
template <bool F16>
__global__ void k_transpose_split_weights(const float* __restrict__ params, float* __restrict__ t_hi,
                                          float* __restrict__ t_lo, fp_nsf_config cfg) {
  __shared__ float tile[32][33];
  const Layout L(&cfg);
  const int per = 2 * L.Bk + 2;
  const int layer = blockIdx.y / per, which = blockIdx.y % per;
  int rows, cols, ld_dst; int64_t src, dst;
  if (which == 0) { rows = L.H; cols = L.D; src = L.w0x(layer); dst = L.tW0(layer); ld_dst = L.H; }
  else if (which == per - 1) { rows = L.dP(layer); cols = L.H; src = L.wf(layer); dst = L.tWf(layer); ld_dst = L.ldP(layer); }
  else {
    const int b = (which - 1) / 2;
    rows = L.H; cols = L.H; ld_dst = L.H;
    if ((which - 1) % 2 == 0) { src = L.w1(layer, b); dst = L.tW1(layer, b); }
    else { src = L.w2(layer, b); dst = L.tW2(layer, b); }
  }
  const int tiles_c = (cols + 31) / 32, tiles_r = (rows + 31) / 32;
  if ((int)blockIdx.x >= tiles_c * tiles_r) return;
  const int r0 = (blockIdx.x / tiles_c) * 32, c0 = (blockIdx.x % tiles_c) * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? params[src + (int64_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, r = r0 + threadIdx.x;      // dst[c][r]
    if (c < cols && r < rows) {
      float v = tile[threadIdx.x][i];
      if (F16) {
        uint32_t h2, l2;
        split_f16x2(v, 0.f, h2, l2);
        const int64_t o = 2 * dst + (int64_t)c * Layout::ld16(ld_dst) + r;
        reinterpret_cast<uint16_t*>(t_hi)[o] = (uint16_t)(h2 & 0xFFFFu);
        reinterpret_cast<uint16_t*>(t_lo)[o] = (uint16_t)(l2 & 0xFFFFu);
      } else {
      float h = tf32_round(v);
      t_hi[dst + (int64_t)c * ld_dst + r] = h;
      t_lo[dst + (int64_t)c * ld_dst + r] = v - h;
      }
    }
  }
}

// ---- host side: tensor maps (cached) ---------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 row-major [rows, cols] with leading dimension ld (floats); box = 32 cols x box_rows rows, SWIZZLE_128B
bool get_tensor_map(const float* base, int64_t rows, int64_t cols, int64_t ld, CUtensorMap* out, int box_rows, bool atom32) {
  typedef std::tuple<const void*, int64_t, int64_t, int64_t, int, bool> Key;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  Key key(base, rows, cols, ld, box_rows, atom32);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  if (cache.size() > 4096) cache.clear();
  cache[key] = m;
  *out = m;
  return true;
}

// 2-D fp16 row-major [rows, cols] with leading dimension ld (halves); box = 64 columns x box_rows rows, SWIZZLE_128B
bool get_tensor_map_f16(const void* base, int64_t rows, int64_t cols, int64_t ld, CUtensorMap* out, int box_rows) {
  typedef std::tuple<const void*, int64_t, int64_t, int64_t, int> Key;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  Key key(base, rows, cols, ld, box_rows);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  if (cache.size() > 4096) cache.clear();
  cache[key] = m;
  *out = m;
  return true;
}

static long long* g_tc_trace = nullptr;
static int g_tc_dbg = 0;     // debug switches (fp_tc_set_debug): 1 no split, 2 no epilogue stores, 4 two products, 8 no B_lo, 32 pairs
bool tc_single_pass();        // flow_engine.cu (fp_set_mode): reduced-precision variant, ONE TF32 product instead of three

// Launch with programmatic stream serialisation: the kernel may begin (barrier init, TMEM allocation, tensor-map
// prefetch) while the previous kernel in the stream drains; it calls griddepcontrol.wait before touching memory.
// FLOPPITY_PDL=0 falls back to plain stream order.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl_n(void (*kernel)(KArgs...), int grid, int threads, cudaStream_t stream, int cluster, Args... args);
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, cudaStream_t stream, int cluster, Args... args) {
  return launch_pdl_n(kernel, grid, TC_THREADS, stream, cluster, args...);
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl_n(void (*kernel)(KArgs...), int grid, int threads, cudaStream_t stream, int cluster, Args... args) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("FLOPPITY_PDL"); on = (e && e[0] == '0') ? 0 : 1; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3((unsigned)threads, 1, 1);
  cfg.dynamicSmemBytes = TC_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster > 1) {           // CTA pairs: the two CTAs of a cluster land on the two SMs of one TPC
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (on) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Can this GEMM go to the tcgen05 NT kernel?  (both operands K-major, TMA-compatible strides, no dropout/split-K)
static bool tc_supported(const fp_gemm_args* a) {
  if (!a->a_kmajor || !a->b_kmajor) return false;
  if ((a->lda & 3) || (a->ldb & 3)) return false;
  if ((reinterpret_cast<uintptr_t>(a->A) & 15) || (reinterpret_cast<uintptr_t>(a->B) & 15)) return false;
  if (a->flags & (FP_G_RELU_B | FP_G_ATOMIC)) return false;
  if (a->split_k > 1) return false;
  if (a->M < 1 || a->N < 1 || a->K < 1 || a->M > 2147483647LL) return false;
  return true;
}

// bytes of the operand slice a tile group keeps L2-resident (FLOPPITY_L2_GROUP_MB, default 40 of the 126 MB)
static double l2_group_bytes() {
  static double v = 0.0;
  if (v == 0.0) { const char* e = getenv("FLOPPITY_L2_GROUP_MB"); v = 1e6 * ((e && atof(e) > 0.0) ? atof(e) : 40.0); }
  return v;
}

// B_hi / B_lo: the pre-split weight matrix, same shape and ld as a->B
int launch_gemm_tc(const fp_gemm_args* a, const float* B_hi, const float* B_lo, cudaStream_t stream) {
  if (!tc_enabled() || !tc_supported(a) || !B_hi || !B_lo) return FP_E_NOTREADY;
  if ((reinterpret_cast<uintptr_t>(B_hi) & 15) || (reinterpret_cast<uintptr_t>(B_lo) & 15)) return FP_E_NOTREADY;
  // CTA pairs (cta_group::2, see the kernel comment) are OFF by default: measured on c3's context projection
  // (tools/c3_gemm_ablate.py) the pair kernel is correct but slower (25.7 ms against 18.2 ms) -- one issuing thread now
  // paces two SMs through cluster-scope hand-shakes, which costs more than the deeper ring and the halved B traffic buy.
  // FLOPPITY_PAIR=1 or fp_tc_set_debug(32) selects it (the parity suite runs it once).
  static int pair_on = -1;
  if (pair_on < 0) { const char* e = getenv("FLOPPITY_PAIR"); pair_on = (e && e[0] == '1') ? 1 : 0; }
  const bool pair = (pair_on || (g_tc_dbg & 32)) && a->K >= 256 && a->M >= 2 * TC_BM;
  CUtensorMap mA, mBh, mBl;
  if (!get_tensor_map(a->A, a->M, a->K, a->lda, &mA)) return FP_E_NOTREADY;
  if (!get_tensor_map(B_hi, a->N, a->K, a->ldb, &mBh, pair ? TC_BN / 2 : TC_BN)) return FP_E_NOTREADY;
  if (!get_tensor_map(B_lo, a->N, a->K, a->ldb, &mBl, pair ? TC_BN / 2 : TC_BN)) return FP_E_NOTREADY;
  const int nkb = (int)((a->K + TC_BK - 1) / TC_BK);
  // long contractions are chunked (see the kernel comment); only the plain / bias epilogue can be accumulated
  const bool plain = (a->flags & ~(FP_G_RELU_A | FP_G_BIAS | FP_G_TF32X3)) == 0;
  const int nchunks = (plain && nkb > TC_KCHUNK_BLOCKS) ? (nkb + TC_KCHUNK_BLOCKS - 1) / TC_KCHUNK_BLOCKS : 1;
  // L2-aware tile order (see the kernel): group the fast dimension so that a group's slice of its operand is ~40 MB
  int gfast = 0;
  {
    const bool m_fast = a->M <= (int64_t)a->N;
    const double fast_bytes = m_fast ? 4.0 * (double)a->M * (double)a->K : 8.0 * (double)a->N * (double)a->K;
    const double tile_bytes = (m_fast ? 4.0 * (pair ? 2 : 1) * TC_BM : 8.0 * TC_BN) * (double)a->K;
    if (fast_bytes > 64e6 && !(g_tc_dbg & 128)) gfast = (int)std::max(1.0, l2_group_bytes() / tile_bytes);
  }
  const int64_t tiles = ((a->M + (pair ? 2 : 1) * TC_BM - 1) / ((pair ? 2 : 1) * TC_BM)) * ((a->N + TC_BN - 1) / TC_BN);
  const int grid = pair ? 2 * (int)std::min<int64_t>(tiles, num_sms() / 2) : (int)std::min<int64_t>(tiles, num_sms());
  ProfScope ps(TAG_GEMM_TC_NT, 2.0 * (double)a->M * (double)a->N * (double)a->K, stream, gemm_algo_bytes(a));
  // float4 epilogue needs every pointer the flags touch to be 16-byte aligned with a leading dimension % 4 == 0
  auto ok = [](const void* p, int64_t ld) { return p == nullptr || (((reinterpret_cast<uintptr_t>(p) & 15) == 0) && ((ld & 3) == 0)); };
  int vec_all = ok(a->C, a->ldc) ? 1 : 0;
  if ((a->flags & FP_G_BIAS) && !ok(a->bias, 4)) vec_all = 0;
  if ((a->flags & (FP_G_ADD_CTX | FP_G_GLU_RES)) && !ok(a->G, a->ldg)) vec_all = 0;
  if ((a->flags & FP_G_GLU_RES) && (!ok(a->Rsd, a->ldr) || !ok(a->U, a->ldu))) vec_all = 0;
  if ((a->flags & FP_G_RELUMASK) && !ok(a->Msk, a->ldm)) vec_all = 0;
  // one instantiation per epilogue the engine uses (compile-time flags), run-time flags for anything else
#define FP_TC_NT_CASE(E)                                                                                              \
  case (E): {                                                                                                         \
    static bool attr_set = false;                                                                                     \
    if (!attr_set) {                                                                                                  \
      FP_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc_nt<(E), false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES)); \
      FP_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc_nt<(E), true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES)); \
      attr_set = true;                                                                                                \
    }                                                                                                                 \
    if (pair)                                                                                                         \
      FP_CUDA_OK(launch_pdl(k_gemm_tc_nt<(E), true, 0>, grid, stream, 2, mA, mA, mBh, mBl, *a, nkb, nchunks, vec_all, g_tc_trace, g_tc_dbg | (tc_single_pass() ? 64 : 0), gfast)); \
    else                                                                                                              \
      FP_CUDA_OK(launch_pdl(k_gemm_tc_nt<(E), false, 0>, grid, stream, 1, mA, mA, mBh, mBl, *a, nkb, nchunks, vec_all, g_tc_trace, g_tc_dbg | (tc_single_pass() ? 64 : 0), gfast)); \
    break;                                                                                                            \
  }
  int epi = a->flags & FP_EPI_MASK;
  switch (epi) {
    case 0: case FP_G_BIAS: case FP_G_ADD_CTX: case FP_G_BIAS | FP_G_GLU_RES: case FP_G_RELUMASK:
    case FP_G_RELUMASK | FP_G_ACCUM: case FP_G_ACCUM: case FP_G_BIAS | FP_G_DROP_OUT: break;
    default: epi = -1;
  }
  switch (epi) {
    FP_TC_NT_CASE(0)
    FP_TC_NT_CASE(FP_G_BIAS)
    FP_TC_NT_CASE(FP_G_ADD_CTX)
    FP_TC_NT_CASE(FP_G_BIAS | FP_G_GLU_RES)
    FP_TC_NT_CASE(FP_G_RELUMASK)
    FP_TC_NT_CASE(FP_G_RELUMASK | FP_G_ACCUM)
    FP_TC_NT_CASE(FP_G_ACCUM)
    FP_TC_NT_CASE(FP_G_BIAS | FP_G_DROP_OUT)
    FP_TC_NT_CASE(-1)
  }
#undef FP_TC_NT_CASE
  FP_LAUNCH_CHECK();
  return 0;
}

// 3xFP16 variant: B_hi16 / B_lo16 = fp16 pieces of B [N, K] with leading dimension a->ldb in HALVES
int launch_gemm_tc16(const fp_gemm_args* a, const void* B_hi16, const void* B_lo16, cudaStream_t stream) {
  if (!tc_enabled() || !a->a_kmajor || !a->b_kmajor || !B_hi16 || !B_lo16) return FP_E_NOTREADY;
  if ((a->lda & 3) || (a->ldb & 7) || (reinterpret_cast<uintptr_t>(a->A) & 15)) return FP_E_NOTREADY;
  if ((reinterpret_cast<uintptr_t>(B_hi16) & 15) || (reinterpret_cast<uintptr_t>(B_lo16) & 15)) return FP_E_NOTREADY;
  if (a->flags & (FP_G_RELU_B | FP_G_ATOMIC)) return FP_E_NOTREADY;
  if (a->split_k > 1 || a->M < 1 || a->N < 1 || a->K < 1 || a->M > 2147483647LL) return FP_E_NOTREADY;
  CUtensorMap mA, mAl, mBh, mBl;
  // A pre-split by the caller (plain / bias epilogues, no prologue, no operand scale): no conversion in the kernel
  const bool a16 = a->pre_hi16 && a->pre_lo16 && !(a->flags & ~(FP_G_BIAS | FP_G_TF32X3)) && !a->a_scale && !(a->pre_ld & 7) &&
                   !(reinterpret_cast<uintptr_t>(a->pre_hi16) & 15) && !(reinterpret_cast<uintptr_t>(a->pre_lo16) & 15);
  if (a16) {
    if (!get_tensor_map_f16(a->pre_hi16, a->M, a->K, a->pre_ld, &mA)) return FP_E_NOTREADY;
    if (!get_tensor_map_f16(a->pre_lo16, a->M, a->K, a->pre_ld, &mAl)) return FP_E_NOTREADY;
  } else if (!get_tensor_map(a->A, a->M, a->K, a->lda, &mA)) return FP_E_NOTREADY;
  if (!get_tensor_map_f16(B_hi16, a->N, a->K, a->ldb, &mBh)) return FP_E_NOTREADY;
  if (!get_tensor_map_f16(B_lo16, a->N, a->K, a->ldb, &mBl)) return FP_E_NOTREADY;
  const int nkb = (int)((a->K + 63) / 64);
  constexpr int kchunk = TC_KCHUNK_BLOCKS / 2;       // the same 1024 contraction indices per accumulation chain
  const bool plain = (a->flags & ~(FP_G_RELU_A | FP_G_BIAS | FP_G_TF32X3)) == 0;
  const int nchunks = (plain && nkb > kchunk) ? (nkb + kchunk - 1) / kchunk : 1;
  int gfast = 0;
  {
    const bool m_fast = a->M <= (int64_t)a->N;
    const double fast_bytes = 4.0 * (double)(m_fast ? a->M : (int64_t)a->N) * (double)a->K;      // fp32 A rows / fp16 hi + lo B rows
    const double tile_bytes = 4.0 * TC_BM * (double)a->K;
    if (fast_bytes > 64e6 && !(g_tc_dbg & 128)) gfast = (int)std::max(1.0, l2_group_bytes() / tile_bytes);
  }
  const int64_t tiles = ((a->M + TC_BM - 1) / TC_BM) * ((a->N + TC_BN - 1) / TC_BN);
  const int grid = (int)std::min<int64_t>(tiles, num_sms());
  ProfScope ps(TAG_GEMM_TC_NT, 2.0 * (double)a->M * (double)a->N * (double)a->K, stream, gemm_algo_bytes(a));
  auto ok = [](const void* p, int64_t ld) { return p == nullptr || (((reinterpret_cast<uintptr_t>(p) & 15) == 0) && ((ld & 3) == 0)); };
  int vec_all = ok(a->C, a->ldc) ? 1 : 0;
  if ((a->flags & FP_G_BIAS) && !ok(a->bias, 4)) vec_all = 0;
  if ((a->flags & (FP_G_ADD_CTX | FP_G_GLU_RES)) && !ok(a->G, a->ldg)) vec_all = 0;
  if ((a->flags & FP_G_GLU_RES) && (!ok(a->Rsd, a->ldr) || !ok(a->U, a->ldu))) vec_all = 0;
  if ((a->flags & FP_G_RELUMASK) && !ok(a->Msk, a->ldm)) vec_all = 0;
#define FP_TC_NT16_CASE(E)                                                                                            \
  case (E): {                                                                                                         \
    static bool attr_set = false;                                                                                     \
    if (!attr_set) {                                                                                                  \
      FP_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc_nt<(E), false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES)); \
      attr_set = true;                                                                                                \
    }                                                                                                                 \
    FP_CUDA_OK(launch_pdl_n(k_gemm_tc_nt<(E), false, 1>, grid, TC16_THREADS, stream, 1, mA, mA, mBh, mBl, *a, nkb, nchunks, vec_all, g_tc_trace, g_tc_dbg | (tc_single_pass() ? 64 : 0), gfast)); \
    break;                                                                                                            \
  }
  int epi = a->flags & FP_EPI_MASK;
  if (a16) {
    static bool attr_set = false;
    if (!attr_set) {
      FP_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc_nt<0, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
      FP_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc_nt<FP_G_BIAS, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
      attr_set = true;
    }
    const int dbg = g_tc_dbg | (tc_single_pass() ? 64 : 0);
    if (epi == FP_G_BIAS)
      FP_CUDA_OK(launch_pdl(k_gemm_tc_nt<FP_G_BIAS, false, 2>, grid, stream, 1, mA, mAl, mBh, mBl, *a, nkb, nchunks, vec_all, g_tc_trace, dbg, gfast));
    else
      FP_CUDA_OK(launch_pdl(k_gemm_tc_nt<0, false, 2>, grid, stream, 1, mA, mAl, mBh, mBl, *a, nkb, nchunks, vec_all, g_tc_trace, dbg, gfast));
    FP_LAUNCH_CHECK();
    return 0;
  }
  switch (epi) {
    case 0: case FP_G_BIAS: case FP_G_ADD_CTX: case FP_G_BIAS | FP_G_GLU_RES: case FP_G_RELUMASK:
    case FP_G_RELUMASK | FP_G_ACCUM: case FP_G_ACCUM: case FP_G_BIAS | FP_G_DROP_OUT: break;
    default: epi = -1;
  }
  switch (epi) {
    FP_TC_NT16_CASE(0)
    FP_TC_NT16_CASE(FP_G_BIAS)
    FP_TC_NT16_CASE(FP_G_ADD_CTX)
    FP_TC_NT16_CASE(FP_G_BIAS | FP_G_GLU_RES)
    FP_TC_NT16_CASE(FP_G_RELUMASK)
    FP_TC_NT16_CASE(FP_G_RELUMASK | FP_G_ACCUM)
    FP_TC_NT16_CASE(FP_G_ACCUM)
    FP_TC_NT16_CASE(FP_G_BIAS | FP_G_DROP_OUT)
    FP_TC_NT16_CASE(-1)
  }
#undef FP_TC_NT16_CASE
  FP_LAUNCH_CHECK();
  return 0;
}

int launch_split_f16(const float* x, void* hi16, void* lo16, int64_t n, cudaStream_t stream, int32_t* status) {
  if (n <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(x) & 7) || (reinterpret_cast<uintptr_t>(hi16) & 3) || (reinterpret_cast<uintptr_t>(lo16) & 3)) return FP_E_BADARG;
  int blocks = (int)std::min<int64_t>((n / 2 + 255) / 256 + 1, 8 * num_sms());
  ProfScope ps(TAG_OTHER, 0.0, stream);
  k_split_f16<<<blocks, 256, 0, stream>>>(x, reinterpret_cast<uint32_t*>(hi16), reinterpret_cast<uint32_t*>(lo16), n, status);
  FP_LAUNCH_CHECK();
  return 0;
}

// wgrad: C[M,N] += A[R,M]^T * B[R,N]   (a_kmajor == 0, b_kmajor == 0, FP_G_ATOMIC [, FP_G_RELU_B])
static int launch_gemm_tc_tn_fmt(const fp_gemm_args* a, cudaStream_t stream, bool f16);
int launch_gemm_tc_tn(const fp_gemm_args* a, cudaStream_t stream) { return launch_gemm_tc_tn_fmt(a, stream, false); }
int launch_gemm_tc16_tn(const fp_gemm_args* a, cudaStream_t stream) { return launch_gemm_tc_tn_fmt(a, stream, true); }
static int launch_gemm_tc_tn_fmt(const fp_gemm_args* a, cudaStream_t stream, bool f16) {
  const int BK = f16 ? 64 : TN_BK;          // batch rows per stage
  if (!tc_enabled()) return FP_E_NOTREADY;
  if (a->a_kmajor || a->b_kmajor) return FP_E_NOTREADY;
  if ((a->flags & ~(FP_G_ATOMIC | FP_G_RELU_B | FP_G_COLSUM_A)) != 0 || !(a->flags & FP_G_ATOMIC)) return FP_E_NOTREADY;
  if ((a->flags & FP_G_COLSUM_A) && !a->U) return FP_E_BADARG;
  if ((a->lda & 3) || (a->ldb & 3)) return FP_E_NOTREADY;
  if ((reinterpret_cast<uintptr_t>(a->A) & 15) || (reinterpret_cast<uintptr_t>(a->B) & 15)) return FP_E_NOTREADY;
  if (a->M < 1 || a->N < 1 || a->K < 1 || a->M > 65536) return FP_E_NOTREADY;
  CUtensorMap mA, mB;
  // TF32: MN-major tiles written by the TMA itself (32-byte-atom swizzle); 3xFP16: raw boxes of 64 rows, plain 128-byte swizzle
  // B pre-split by the caller (fp16 hi / lo copies of the activations, no relu prologue): boxes of 64 halves x 64 rows
  const bool b16 = f16 && a->pre_hi16 && a->pre_lo16 && !(a->flags & FP_G_RELU_B) && !(a->pre_ld & 7) &&
                   !(reinterpret_cast<uintptr_t>(a->pre_hi16) & 15) && !(reinterpret_cast<uintptr_t>(a->pre_lo16) & 15);
  CUtensorMap mBl;
  if (!get_tensor_map(a->A, a->K, a->M, a->lda, &mA, BK, !f16)) return FP_E_NOTREADY;   // [R rows, M cols]
  if (b16) {
    if (!get_tensor_map_f16(a->pre_hi16, a->K, a->N, a->pre_ld, &mB, 64)) return FP_E_NOTREADY;
    if (!get_tensor_map_f16(a->pre_lo16, a->K, a->N, a->pre_ld, &mBl, 64)) return FP_E_NOTREADY;
  } else if (!get_tensor_map(a->B, a->K, a->N, a->ldb, &mB, BK, !f16)) return FP_E_NOTREADY;   // [R rows, N cols]
  static bool attr_set = false;
  if (!attr_set) {
    FP_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc_tn<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    FP_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc_tn<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    FP_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc_tn<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    attr_set = true;
  }
  const int nkb = (int)((a->K + BK - 1) / BK);
  const int64_t tiles = ((a->M + TC_BM - 1) / TC_BM) * ((a->N + TC_BN - 1) / TC_BN);
  int nsplit = (int)std::max<int64_t>(1, (num_sms() + tiles - 1) / tiles);
  // A few output tiles (the final layer's [290 x 128] is three): give every CTA ALL tiles of one batch range
  // (items c, c + nsplit, ... of CTA c share their k-range), so the operand the tiles have in common is re-read from
  // L2 microseconds later instead of from HBM by three CTAs that drift apart (221 MB of DRAM reads for 159 MB).
  if (tiles > 1 && tiles <= 8) nsplit = num_sms();
  nsplit = std::min(nsplit, std::max(1, nkb * BK / 128));   // at least 128 batch rows per split
  const int kchunk = TC_KCHUNK_BLOCKS * TC_BK / BK;  // bound the truncating accumulation chain (1024 batch rows)
  nsplit = std::max(nsplit, (nkb + kchunk - 1) / kchunk);
  // Whole waves: with a handful of output tiles (c3's 512 x 512 blocks: 16 tiles x 19 splits = 304 items on 148 SMs =
  // 2.05 waves) the last round runs almost empty.  Take the split count in [nsplit, 2 nsplit] that fills its rounds best.
  if (tiles * nsplit > num_sms() && tiles < num_sms()) {
    int best = nsplit; double best_eff = 0.0;
    for (int c = nsplit; c <= 2 * nsplit && c * 4 <= nkb; ++c) {
      const int64_t items = tiles * c;
      const double eff = (double)items / (double)(((items + num_sms() - 1) / num_sms()) * num_sms());
      if (eff > best_eff + 0.02) { best_eff = eff; best = c; }
    }
    nsplit = best;
  }
  const int kbps = (nkb + nsplit - 1) / nsplit;
  nsplit = (nkb + kbps - 1) / kbps;                        // no empty splits
  const int grid = (int)std::min<int64_t>(tiles * nsplit, num_sms());
  // L2-aware item order for operands that do not fit the L2 (see the kernel): groups of gm M-tiles whose part of C is ~40 MB
  int gm = 0;
  if (4.0 * (double)a->K * ((double)a->M + (double)a->N) > 96e6 && tiles > 2 * num_sms() && !(g_tc_dbg & 128))
    gm = (int)std::max(1.0, l2_group_bytes() / (4.0 * TC_BM * (double)a->N));
  int vec_all = (((reinterpret_cast<uintptr_t>(a->C) & 15) == 0) && ((a->ldc & 3) == 0)) ? 1 : 0;
  if ((a->flags & FP_G_COLSUM_A) && (reinterpret_cast<uintptr_t>(a->U) & 15)) vec_all = 0;
  ProfScope ps(TAG_GEMM_TC_TN, 2.0 * (double)a->M * (double)a->N * (double)a->K, stream, gemm_algo_bytes(a));
  if (b16)
    FP_CUDA_OK(launch_pdl(k_gemm_tc_tn<true, true>, grid, stream, 1, mA, mB, mBl, *a, nkb, nsplit, vec_all, g_tc_dbg | (tc_single_pass() ? 64 : 0), gm));
  else if (f16)
    FP_CUDA_OK(launch_pdl(k_gemm_tc_tn<true, false>, grid, stream, 1, mA, mB, mB, *a, nkb, nsplit, vec_all, g_tc_dbg | (tc_single_pass() ? 64 : 0), gm));
  else
    FP_CUDA_OK(launch_pdl(k_gemm_tc_tn<false, false>, grid, stream, 1, mA, mB, mB, *a, nkb, nsplit, vec_all, g_tc_dbg | (tc_single_pass() ? 64 : 0), gm));
  FP_LAUNCH_CHECK();
  return 0;
}

int launch_split_tf32(const float* x, float* hi, float* lo, int64_t n, cudaStream_t stream) {
  if (n <= 0) return 0;
  int blocks = (int)std::min<int64_t>((n + 255) / 256, 8 * num_sms());
  ProfScope ps(TAG_OTHER, 0.0, stream);
  k_split_tf32<<<blocks, 256, 0, stream>>>(x, hi, lo, n);
  FP_LAUNCH_CHECK();
  return 0;
}

int launch_transpose_split_weights(const fp_nsf_config* cfg, const float* params, float* t_hi, float* t_lo, cudaStream_t stream) {
  Layout L(cfg);
  int maxdim_r = std::max(L.H, L.dPmax()), maxdim_c = std::max(L.H, L.D);
  int tiles = ((maxdim_r + 31) / 32) * ((maxdim_c + 31) / 32);
  dim3 grid(tiles, L.T * (2 * L.Bk + 2));
  ProfScope ps(TAG_OTHER, 0.0, stream);
  k_transpose_split_weights<false><<<grid, dim3(32, 8), 0, stream>>>(params, t_hi, t_lo, *cfg);
  FP_LAUNCH_CHECK();
  return 0;
}

// fp16 operand copies for the 3xFP16 GEMMs, once per optimiser step (after launch_split_tf32 / launch_transpose_split_weights)
int launch_prepare_gemm16(const fp_nsf_config* cfg, const float* params, float* prepared, cudaStream_t stream) {
  Layout L(cfg);
  uint16_t* hi = reinterpret_cast<uint16_t*>(prepared + L.prepG16Hi());
  uint16_t* lo = reinterpret_cast<uint16_t*>(prepared + L.prepG16Lo());
  { int r = launch_split_f16(params + L.ctx_w, hi + 2 * L.ctx_w, lo + 2 * L.ctx_w, L.Nctx * L.C, stream, nullptr); if (r) return r; }
  ProfScope ps(TAG_OTHER, 0.0, stream);
  {
    const int64_t nmax = (int64_t)std::max(L.H, L.dPmax()) * L.H;
    dim3 grid((unsigned)std::min<int64_t>((nmax / 2 + 255) / 256, 64), (unsigned)(L.T * (2 * L.Bk + 2)));
    k_split_f16_weights<<<grid, 256, 0, stream>>>(params, hi, lo, *cfg);
    FP_LAUNCH_CHECK();
  }
  {
    int maxdim_r = std::max(L.H, L.dPmax()), maxdim_c = std::max(L.H, L.D);
    int tiles = ((maxdim_r + 31) / 32) * ((maxdim_c + 31) / 32);
    dim3 grid(tiles, L.T * (2 * L.Bk + 2));
    k_transpose_split_weights<true><<<grid, dim3(32, 8), 0, stream>>>(params, prepared + L.prepGT16Hi(), prepared + L.prepGT16Lo(), *cfg);
    FP_LAUNCH_CHECK();
  }
  return 0;
}

// router used by the engine for the context projection (weights split by fp_nsf_prepare)
int launch_gemm_ctx(const fp_gemm_args* a, cudaStream_t stream) { return launch_gemm(a, stream); }

}  // namespace fp

extern "C" {
// debug: device buffer of >= 512 int64 receiving SM-clock stamps of CTA 0 (NULL switches tracing off)
int fp_tc_set_trace(void* buf) { fp::g_tc_trace = reinterpret_cast<long long*>(buf); return 0; }
// debug: ablation switches for kernel timing experiments (results are WRONG when non-zero); 0 = normal
int fp_tc_set_debug(int32_t bits) { fp::g_tc_dbg = bits; return 0; }
// Standalone entry: C = pro(A) * B^T (+ epilogue) on the tensor cores; B_hi/B_lo from fp_split_tf32.
// Returns FP_E_NOTREADY when the shape/stride/flags are not supported by the tcgen05 kernel.
int fp_gemm_tc(const fp_gemm_args* args, const float* B_hi, const float* B_lo, void* stream) {
  if (!args) return FP_E_BADARG;
  return fp::launch_gemm_tc(args, B_hi, B_lo, (cudaStream_t)stream);
}
// Weight-gradient form on the tensor cores: C[M,N] += A[K,M]^T * B[K,N] (a_kmajor = b_kmajor = 0, FP_G_ATOMIC).
int fp_gemm_tc_tn(const fp_gemm_args* args, void* stream) {
  if (!args) return FP_E_BADARG;
  return fp::launch_gemm_tc_tn(args, (cudaStream_t)stream);
}
int fp_gemm_tc16(const fp_gemm_args* args, const void* B_hi16, const void* B_lo16, void* stream) {
  if (!args) return FP_E_BADARG;
  return fp::launch_gemm_tc16(args, B_hi16, B_lo16, (cudaStream_t)stream);
}
int fp_gemm_tc16_tn(const fp_gemm_args* args, void* stream) {
  if (!args) return FP_E_BADARG;
  return fp::launch_gemm_tc16_tn(args, (cudaStream_t)stream);
}
int fp_split_f16(const float* x, void* hi16, void* lo16, int64_t n, void* stream) {
  return fp::launch_split_f16(x, hi16, lo16, n, (cudaStream_t)stream, nullptr);
}
int fp_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream) {
  return fp::launch_split_tf32(x, hi, lo, n, (cudaStream_t)stream);
}
}
