#!/bin/bash
# usage: scripts/ab2.sh "<flags>" tag   -- isolated kernel times (no wave overlap)
CZB_NVCC_FLAGS="$1" python cairo_zstd_b200/build.py --force > /dev/null 2>&1
python bench.py --frames 131072 --steps 3 --warmup 1 --distinct 512 --no-e2e --no-cpu-baseline > gpurun_out/ab_$2.json 2> gpurun_out/ab_$2.err || { tail -3 gpurun_out/ab_$2.err; exit 1; }
python - <<PY
import json
d=json.load(open("gpurun_out/ab_$2.json"))
k=d["roofline"]["kernel_ms_per_step"]
print("$2", "flags=[$1]", "GB/s=%.1f"%d["value"], " ".join(f"{n}={k[n]:.2f}" for n in ("huff","fse","exec")))
PY
