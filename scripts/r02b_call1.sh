#!/bin/bash
# GPU box: parity tests, then A/B of the current build against variant builds given as arguments (tag=path)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02b_pytest.log
scripts/ab5.sh - cur
for v in "$@"; do scripts/ab5.sh - "${v%%=*}" CZB_LIB=$PWD/"${v#*=}"; done
