#!/bin/bash
# flow executor parameter sweep on configs 4 and 5 (GPU box)
run() { flags="$1"; shift; scripts/ab4.sh "$flags" sweep "$@" 2>&1 | tail -1 | cut -c1-260; }
run "" 4
run - 5
run - 5 CZB_BIG_SHARE=2048
run - 5 CZB_BIG_SHARE=1024
run - 5 CZB_BIG_SHARE=512
run "-DCZB_FLOW_WIN_LOG=16 -DCZB_FLOW_SLICE=4096" 4
run - 5
run - 5 CZB_BIG_SHARE=2048
run - 5 CZB_BIG_SHARE=1024
run - 3
run "-DCZB_FLOW_WIN_LOG=16 -DCZB_FLOW_SLICE=4096 -DCZB_FLOW_WARPS=32 -DCZB_FLOW_MIN_CTAS=1" 4
run - 5 CZB_BIG_SHARE=2048
run "-DCZB_FLOW_WIN_LOG=16 -DCZB_FLOW_SLICE=8192 -DCZB_FLOW_WARPS=16" 4
run - 5 CZB_BIG_SHARE=2048
