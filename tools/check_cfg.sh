#!/usr/bin/env bash
# flow-level GPU parity tests + a short bench of one config:  bash tools/check_cfg.sh c4
CFG="${1:-c4}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_flow.py -m gpu -x -q 2>&1 | grep -E "^E  |passed|failed" | cut -c1-200 | head -5
timeout 240 python bench.py --config "$CFG" --steps 6 --warmup 3 --no-cpu > "gpurun_out/b_${CFG}_check.json" 2> "gpurun_out/b_${CFG}_check.err"
tail -1 "gpurun_out/b_${CFG}_check.err"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/b_${CFG}_check.json") if l.startswith("{")][-1])
print(round(d["value"]), round(d["ms_per_step"],2), {k:round(d[k]) for k in d if "per_s" in k and k!="ms_per_step"})
for k,v in d["kernel_classes"].items(): print("  ",k,{a:round(b,3) for a,b in v.items()})
PY
