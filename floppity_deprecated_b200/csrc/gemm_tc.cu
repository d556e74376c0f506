// gemm_tc.cu -- router for the context-projection GEMM.  (tcgen05 path lands here; SIMT for now.)
#include "common.cuh"
namespace fp {
int launch_gemm(const fp_gemm_args* a, float drop_p, uint64_t drop_seed, int drop_width, cudaStream_t stream);
int launch_gemm_ctx(const fp_gemm_args* a, cudaStream_t stream) { return launch_gemm(a, 0.f, 0, 0, stream); }
}  // namespace fp
