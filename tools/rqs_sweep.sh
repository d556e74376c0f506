#!/usr/bin/env bash
# spline-kernel geometry sweep inside the real training step (class averages from bench.py's event pass)
mkdir -p gpurun_out
for st in 2 1; do for kb in 10 20 40; do
  FLOPPITY_RQS_STAGES=$st FLOPPITY_RQS_TILE_KB=$kb python bench.py --steps 8 --warmup 3 --no-cpu --no-extras > gpurun_out/rs_${st}_${kb}.json 2>/dev/null
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/rs_${st}_${kb}.json") if l.startswith("{")][-1])
k=d["kernel_classes"]
print("stages $st tile ${kb}KB: step %.3f ms  rqs_fwd %.1f us (%.0f%% hbm)  rqs_bwd %.1f us (%.0f%% hbm)" % (d["ms_per_step"], k["rqs_forward"]["avg_us"], 100*k["rqs_forward"]["frac_of_hbm_peak"], k["rqs_backward"]["avg_us"], 100*k["rqs_backward"]["frac_of_hbm_peak"]))
PY
done; done
