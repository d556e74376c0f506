#!/usr/bin/env bash
# One-GPU evidence run: parity tests, bench lines for every config, c5 sweep, ncu launch list + full capture.
# usage (under gpurun): bash tools/gpu_round.sh <tag>
set -uo pipefail
TAG="${1:-r1x}"
OUT="gpurun_out/${TAG}"
mkdir -p "${OUT}"
python -m pytest tests -m gpu -x -q > "${OUT}/gpu_tests.log" 2>&1; echo "pytest rc=$?" | tee -a "${OUT}/gpu_tests.log"
tail -3 "${OUT}/gpu_tests.log"
python bench.py --steps 20 --warmup 5 > "${OUT}/bench_c2.json" 2> "${OUT}/bench_c2.err"; echo "bench c2 rc=$?"
python bench.py --steps 20 --warmup 5 --graph 1 --no-cpu --no-extras > "${OUT}/bench_c2_graph.json" 2> "${OUT}/bench_c2_graph.err"; echo "bench c2 graph rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > "${OUT}/bench_ref.json" 2> "${OUT}/bench_ref.err"; echo "bench ref rc=$?"
for c in c1 c4 c3; do
  timeout 600 python bench.py --config $c --steps 6 --warmup 3 --no-cpu > "${OUT}/bench_${c}.json" 2> "${OUT}/bench_${c}.err"; echo "bench $c rc=$?"
done
timeout 900 python tools/sweep_c5.py > "${OUT}/sweep_c5.jsonl" 2> "${OUT}/sweep_c5.err"; echo "sweep rc=$?"
timeout 600 python tools/kbench.py rqs > "${OUT}/kbench_rqs.log" 2>&1; cp gpurun_out/kbench.json "${OUT}/kbench_rqs.json" 2>/dev/null; echo "kbench rc=$?"
timeout 300 python tools/cf_trace.py > "${OUT}/cond_fused_trace.log" 2>&1; echo "cf trace rc=$?"
tail -2 "${OUT}"/bench_*.err | tail -30
# ncu: launch list of one step + full capture of the dominant GEMM (same command ran plain right before)
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-extras"
$CMD > "${OUT}/plain.log" 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 450 --csv --log-file "${OUT}/launches.csv" $CMD > "${OUT}/ncu_list.log" 2>&1
echo "ncu list rc=$?"
$CMD > "${OUT}/plain2.log" 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_cond_fwd" -s 32 -c 2 -o "${OUT}/prof_cond_fwd" $CMD > "${OUT}/ncu_full_cf.log" 2>&1
echo "ncu full (fused conditioner) rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tc_tn|k_rqs_backward" -s 80 -c 6 -o "${OUT}/prof_top" $CMD > "${OUT}/ncu_full.log" 2>&1
echo "ncu full (wgrad, spline backward) rc=$?"
ls -la "${OUT}"
