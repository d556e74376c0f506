// prof.h -- launch accounting: a global launch counter (always on) and, when enabled by bench.py, CUDA-event
// timing of every launch grouped by kernel class, together with the ALGORITHMIC work (bytes or flops) each
// launch was asked to do.  Events are recorded on the launching stream.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fp {
enum ProfTag { TAG_GEMM_SIMT = 0, TAG_GEMM_TC = 1, TAG_RQS_FWD = 2, TAG_RQS_INV = 3, TAG_RQS_BWD = 4, TAG_OTHER = 5, NTAGS = 6 };

struct ProfScope {
  // work = algorithmic flops (GEMM classes) or bytes (HBM-bound classes); bytes = algorithmic HBM bytes of a GEMM
  ProfScope(int tag, double work, cudaStream_t s, double bytes = 0.0);
  ~ProfScope();
  cudaStream_t s_;
  int slot_;
};
}  // namespace fp
