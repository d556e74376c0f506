#!/usr/bin/env bash
# Builds libfloppity_b200.so (sm_100a only) next to the Python package.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libfloppity_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr)
OBJS=()
for f in rqs_kernels gemm_simt gemm_tc cond_fused misc_kernels retrieval_kernels flow_engine prof; do
  "${NVCC}" "${FLAGS[@]}" -c "${HERE}/${f}.cu" -o "${HERE}/${f}.o" 2> "${HERE}/${f}.ptxas.log" || { cat "${HERE}/${f}.ptxas.log"; exit 1; }
  OBJS+=("${HERE}/${f}.o")
done
"${NVCC}" -shared -o "${OUT}" "${OBJS[@]}" -lcudart
echo "built ${OUT}"
