#!/bin/bash
# usage: scripts/co_exec.sh <carveout percent>...   (GPU box: k_exec's preferred shared-memory carveout, kernel ms per 131072 frames)
for c in "$@"; do
  CZB_EXEC_CARVEOUT=$c python bench.py --frames 131072 --steps 3 --warmup 1 --distinct 512 --no-e2e --no-cpu-baseline > gpurun_out/co_exec.json 2>/dev/null
  python - "$c" <<'PY'
import json, sys
d = json.load(open("gpurun_out/co_exec.json")); k = d["roofline"]["kernel_ms_per_step"]
print("carveout", sys.argv[1], "exec=%.2f fse=%.2f huff=%.2f" % (k["exec"], k["fse"], k["huff"]))
PY
done
