#!/bin/bash
# GPU box: parity tests, then the large-frame configurations (4: long window, 5: mixed sizes) of the current build and of variants (tag=path)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02b_pytest.log
for c in 4 5; do python scripts/perf_configs.py $c 2>/dev/null | tail -1 | cut -c1-330; done
for v in "$@"; do for c in 4 5; do echo "variant ${v%%=*}"; CZB_LIB=$PWD/"${v#*=}" python scripts/perf_configs.py $c 2>/dev/null | tail -1 | cut -c1-330; done; done
