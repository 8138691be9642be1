#!/bin/bash
# GPU box, final evidence of round 2 (second session): the default bench line, then the ncu launch list and one full capture
# of the decode kernels.  Every profiled command first runs plainly and must exit 0.
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02b_pytest.log
python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err || { tail -5 gpurun_out/r02b_bench.err; exit 1; }
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-other-configs"
$B > gpurun_out/r02b_ncu_plain.json 2> gpurun_out/r02b_ncu_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02b_launches.csv $B > gpurun_out/r02b_ncu_launches.log 2>&1
S="python bench.py --frames 65536 --steps 1 --warmup 1 --distinct 512 --no-e2e --no-cpu-baseline --no-other-configs"
$S > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k 'regex:k_fse|k_huff|k_exec|k_xxh64' -s 6 -c 6 -f -o gpurun_out/r02b_full $S > gpurun_out/r02b_ncu_full.log 2>&1
ls -la gpurun_out/r02b*
