#!/bin/bash
# GPU box: run the tests named on the command line
timeout 900 python -m pytest "$@" -x -q 2>&1 | tail -15
