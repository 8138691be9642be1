#!/bin/bash
# 2 x B200: the bench line at N = 2 (config 2 weak scaling, config5_sharded strong scaling) with the final build
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_n2.json 2> gpurun_out/r02b_bench_n2.err; echo "rc=$?"; tail -c 400 gpurun_out/r02b_bench_n2.err
python - <<PY
import json
for l in open("gpurun_out/r02b_bench_n2.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["n_gpus"], round(d["value"],1), d["ms_per_step"], d.get("config5_sharded",{}).get("value"), d.get("config5_sharded",{}).get("ms_per_rank"), d["e2e"]["value"] if d.get("e2e") else None)
PY
