#!/bin/bash
# usage: scripts/ab5.sh "<nvcc -D flags | ->" <tag> [ENV=VAL ...]   (GPU box: rebuild with flags, 3-wave bench, print kernel ms)
flags="$1"; tag="$2"; shift 2
if [ "$flags" != "-" ]; then CZB_NVCC_FLAGS="$flags" python cairo_zstd_b200/build.py --force > /dev/null 2>&1 || { echo "$tag build failed"; exit 1; }; fi
env "$@" python bench.py --frames ${AB_FRAMES:-393216} --steps 3 --warmup 1 --distinct 512 --no-e2e --no-cpu-baseline > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err || { tail -3 gpurun_out/ab_$tag.err; exit 1; }
python - <<PY
import json
d=json.load(open("gpurun_out/ab_$tag.json"))
k=d["roofline"]["kernel_ms_per_step"]
print("$tag", "flags=[$flags] env=[$*]", "GB/s=%.1f"%d["value"], "ms=%.2f"%d["ms_per_step"], " ".join(f"{n}={k.get(n,0):.1f}" for n in ("huff","fse","exec")), flush=True)
PY
