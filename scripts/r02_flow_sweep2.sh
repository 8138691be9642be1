#!/bin/bash
run() { flags="$1"; shift; scripts/ab4.sh "$flags" sweep "$@" 2>&1 | tail -1 | cut -c1-200; }
run "" 4
run "-DCZB_FLOW_WARPS=32 -DCZB_FLOW_MIN_CTAS=1" 4
run "-DCZB_FLOW_WARPS=24 -DCZB_FLOW_MIN_CTAS=1" 4
run "-DCZB_FLOW_WARPS=12 -DCZB_FLOW_MIN_CTAS=2" 4
run "-DCZB_FLOW_WIN_LOG=17 -DCZB_FLOW_SLICE=8192 -DCZB_FLOW_WARPS=32 -DCZB_FLOW_MIN_CTAS=1" 4
