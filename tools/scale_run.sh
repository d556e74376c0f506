#!/usr/bin/env bash
# N-GPU scaling evidence (run under `gpurun --gpus N`): training bench + c5 sweep, one rank per GPU over NCCL.
set -uo pipefail
N="${1:-2}"; TAG="${2:-r1x}"
OUT="gpurun_out/${TAG}"; mkdir -p "${OUT}"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node ${N} --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus ${N} --steps 20 --warmup 5 > "${OUT}/bench_n${N}.json" 2> "${OUT}/bench_n${N}.err"; echo "bench rc=$?"
grep "\[bench\]" "${OUT}/bench_n${N}.err" | head -2
$TR --master-port 29512 tools/sweep_c5.py --max-exp 8 > "${OUT}/sweep_c5_n${N}.jsonl" 2> "${OUT}/sweep_c5_n${N}.err"; echo "sweep rc=$?"
cut -c1-220 "${OUT}/sweep_c5_n${N}.jsonl"
