#!/bin/bash
# GPU box: A/B runs given as arguments (tag=path), then one full ncu capture of k_fse (source view) on a 65 536-frame wave
scripts/ab5.sh - cur
for v in "$@"; do scripts/ab5.sh - "${v%%=*}" CZB_LIB=$PWD/"${v#*=}"; done
S="python bench.py --frames 65536 --steps 1 --warmup 1 --distinct 512 --no-e2e --no-cpu-baseline --no-other-configs"
ncu --set full --clock-control none --import-source on -k 'regex:k_fse' -s 1 -c 1 -f -o gpurun_out/r02b_fse $S > gpurun_out/r02b_ncu_fse.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -1
