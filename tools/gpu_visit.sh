#!/bin/bash
# One development visit: a subset of the parity tests, then A/B runs of environment switches.
# Usage: gpurun -- bash tools/gpu_visit.sh "<pytest -k expr>" "VAR=a" "VAR=b VAR2=c" ...
mkdir -p gpurun_out
K="$1"; shift
timeout -s KILL 500 python -m pytest tests -x -q -m gpu ${K:+-k "$K"} > gpurun_out/pytest_visit.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_visit.log
tail -8 gpurun_out/pytest_visit.log
bash tools/gpu_ab.sh "$@" 2>&1 | tee gpurun_out/ab_visit.log
