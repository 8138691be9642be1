#!/bin/bash
# usage: scripts/ab4.sh "<nvcc -D flags>" <tag> <config: 3|4|5> [ENV=VAL ...]  (GPU box: rebuild, run one parity config of perf_configs.py)
flags="$1"; tag="$2"; cfg="$3"; shift 3
if [ "$flags" != "-" ]; then CZB_NVCC_FLAGS="$flags" python cairo_zstd_b200/build.py --force > /dev/null 2>&1; fi  # "-": keep the library as it is
echo -n "$tag [$flags] [$*] "; env "$@" timeout 300 python scripts/perf_configs.py $cfg 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config'][:8], '%.1f ms %.1f GB/s' % (d['ms'], d['GBps']), d.get('kernel_ms'))"
