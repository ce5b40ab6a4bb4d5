#!/bin/bash
# run selected GPU tests with a hard time limit.  usage: gpurun -- bash tools/gpu_test.sh <tag> <seconds> <pytest args...>
TAG=$1; LIMIT=$2; shift 2
mkdir -p gpurun_out
timeout -s KILL $LIMIT python -m pytest "$@" -m gpu -q -rs > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_${TAG}.log
