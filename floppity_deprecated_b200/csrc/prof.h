// prof.h -- launch accounting: a global launch counter (always on) and, when enabled by bench.py, CUDA-event
// timing of every launch grouped BY KERNEL, together with the ALGORITHMIC work (flops and HBM bytes) each launch
// was asked to do.  Events are recorded on the launching stream.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fp {
enum ProfTag {
  TAG_GEMM_SIMT = 0,       // k_gemm (exact fp32 SIMT GEMM)
  TAG_GEMM_TC_NT = 1,      // k_gemm_tc_nt (tcgen05 forward / dgrad / context projection)
  TAG_GEMM_TC_TN = 2,      // k_gemm_tc_tn (tcgen05 weight gradients)
  TAG_COND_FWD = 3,        // k_cond_fwd (fused tcgen05 conditioner, hidden 128)
  TAG_RQS_FWD = 4,         // k_rqs_apply<forward>
  TAG_RQS_INV = 5,         // k_rqs_apply<inverse>
  TAG_RQS_BWD = 6,         // k_rqs_backward
  TAG_COND_SMALL_FWD = 7,  // k_cond_small_fwd (shared-memory-resident conditioner, hidden <= 64)
  TAG_COND_SMALL_BWD = 8,  // k_cond_small_bwd
  TAG_EMBED = 9,           // convolutional embedding kernels
  TAG_SMALL_WGRAD = 10,    // k_small_wgrad (all weight gradients of a hidden <= 64 conditioner layer)
  TAG_OTHER = 11,
  NTAGS = 12
};

struct ProfScope {
  // flops / bytes: the algorithmic work of the launch (either may be 0)
  ProfScope(int tag, double flops, cudaStream_t s, double bytes = 0.0);
  ~ProfScope();
  cudaStream_t s_;
  int slot_;
};
}  // namespace fp
