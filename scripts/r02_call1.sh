#!/bin/bash
# round 2, GPU call 1: overlap A/B on the round-1 kernels, then compute-sanitizer over the executor tests
scripts/ab5.sh "" base
scripts/ab5.sh - overlap7 CZB_OVERLAP=1
scripts/ab5.sh "-DEXEC_CTAS_PER_SM=6 -DEXEC_MIN_CTAS=6" exec6
scripts/ab5.sh - overlap6 CZB_OVERLAP=1
scripts/ab5.sh "-DEXEC_CTAS_PER_SM=5 -DEXEC_MIN_CTAS=5" exec5
scripts/ab5.sh - overlap5 CZB_OVERLAP=1
CZB_NVCC_FLAGS="" python cairo_zstd_b200/build.py --force > /dev/null 2>&1
scripts/r02_sanitize.sh r02a
