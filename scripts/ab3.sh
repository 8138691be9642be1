#!/bin/bash
# usage: scripts/ab3.sh "<nvcc -D flags>" <tag> [ENV=VAL ...]   (GPU box: rebuild with flags, short bench with extra env, print kernel ms)
set -e
flags="$1"; tag="$2"; shift 2
CZB_NVCC_FLAGS="$flags" python cairo_zstd_b200/build.py --force > /dev/null 2>&1
env "$@" python bench.py --frames 262144 --steps 3 --warmup 1 --distinct 512 --no-e2e --no-cpu-baseline > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err || { tail -3 gpurun_out/ab_$tag.err; exit 1; }
python - <<PY
import json
d=json.load(open("gpurun_out/ab_$tag.json"))
k=d["roofline"]["kernel_ms_per_step"]
print("$tag", "flags=[$flags] env=[$*]", "GB/s=%.1f"%d["value"], "ms/step=%.1f"%d["ms_per_step"], " ".join(f"{n}={k[n]:.1f}" for n in ("huff","fse","exec")))
PY
