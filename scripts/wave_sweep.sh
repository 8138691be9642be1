#!/bin/bash
# usage: scripts/wave_sweep.sh <frames per wave>...   (GPU box: headline bench with CZB_WAVE_FRAMES / CZB_BUDGET_GB=48)
for w in "$@"; do
  CZB_WAVE_FRAMES=$w CZB_BUDGET_GB=48 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline 2>/dev/null > gpurun_out/wave_$w.json
  python - "$w" <<'PY'
import json, sys
w = sys.argv[1]
d = json.loads(open(f"gpurun_out/wave_{w}.json").read().strip().splitlines()[-1])
print("wave", w, round(d["value"], 1), "GB/s", round(d["ms_per_step"], 1), "ms", {k: round(v, 1) for k, v in d["roofline"]["kernel_ms_per_step"].items()})
PY
done
