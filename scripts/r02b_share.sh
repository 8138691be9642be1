#!/bin/bash
# GPU box: which frames of a mixed batch get a CTA (CZB_BIG_SHARE: frames holding >= 1/share of the wave's compressed bytes)
for s in 2048 4096 8192 16384 32768; do echo "share $s"; CZB_BIG_SHARE=$s python scripts/perf_configs.py 5 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['GBps'],1), d['kernel_ms'])"; done
