#!/usr/bin/env bash
# Builds libfloppity_b200.so (sm_100a only) next to the Python package.  Translation units compile in parallel.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libfloppity_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr)
UNITS=(rqs_kernels gemm_simt gemm_tc cond_fused cond_small misc_kernels retrieval_kernels embed_conv flow_engine prof)
OBJS=()
PIDS=()
for f in "${UNITS[@]}"; do
  [ -f "${HERE}/${f}.cu" ] || continue
  "${NVCC}" "${FLAGS[@]}" -c "${HERE}/${f}.cu" -o "${HERE}/${f}.o" 2> "${HERE}/${f}.ptxas.log" &
  PIDS+=("$!:${f}")
  OBJS+=("${HERE}/${f}.o")
done
FAIL=0
for pf in "${PIDS[@]}"; do
  if ! wait "${pf%%:*}"; then cat "${HERE}/${pf##*:}.ptxas.log"; FAIL=1; fi
done
[ "${FAIL}" = 0 ] || exit 1
"${NVCC}" -shared -o "${OUT}" "${OBJS[@]}" -lcudart
echo "built ${OUT}"
