#!/usr/bin/env bash
# flow + embedding parity, fused-kernel trace summary, headline bench (short)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_flow.py tests/test_embedding.py -m gpu -x -q 2>&1 | grep -E "^E  |passed|failed" | cut -c1-200 | head -4
timeout 100 python tools/cf_trace.py 2>&1 | grep -E "whole|last stamp"
timeout 240 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/b_check.json 2> gpurun_out/b_check.err
tail -1 gpurun_out/b_check.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/b_check.json") if l.startswith("{")][-1])
print(round(d["value"]), round(d["ms_per_step"],3), {k:round(d[k]) for k in d if "per_s" in k and k!="ms_per_step"})
PY
