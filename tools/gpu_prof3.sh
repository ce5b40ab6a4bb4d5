#!/bin/bash
# ncu --set full of selected kernels of one full-size chunk.  Usage: gpurun -- bash tools/gpu_prof3.sh <tag> <kernel regex> <skip> <count> [ENV=..]
TAG=${1:-x}; KRE=${2:-tc_gemm}; SKIP=${3:-0}; CNT=${4:-13}; shift 4
mkdir -p gpurun_out
env "$@" timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k "regex:${KRE}" -s ${SKIP} -c ${CNT} -f -o gpurun_out/prof_${TAG} \
   python bench.py --batch 4096 --steps 1 --warmup 3 --no-cpu-baseline --max-rows 32768 > gpurun_out/ncu_full_${TAG}.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out | tail -4
