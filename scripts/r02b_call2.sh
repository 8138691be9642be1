#!/bin/bash
# GPU box: quick A/B of the current build (and optional variants given as arguments: tag=path)
scripts/ab5.sh - cur
for v in "$@"; do scripts/ab5.sh - "${v%%=*}" CZB_LIB=$PWD/"${v#*=}"; done
