#!/bin/bash
# 8 x B200: the bench line at N = 8 (config 2 weak scaling, config5_sharded strong scaling) with the final build
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_n8.json 2> gpurun_out/r02b_bench_n8.err; echo "rc=$?"; tail -c 300 gpurun_out/r02b_bench_n8.err
python - <<PY
import json
for l in open("gpurun_out/r02b_bench_n8.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["n_gpus"], round(d["value"],1), d["ms_per_step"], d.get("config5_sharded",{}).get("value"), d.get("config5_sharded",{}).get("imbalance"), d["e2e"]["value"] if d.get("e2e") else None, d["e2e"].get("frac_of_link") if d.get("e2e") else None)
PY
