#!/bin/bash
# one ncu --set full capture of one kernel.  usage: gpurun -- bash tools/gpu_ncu1.sh <tag> <kernel regex> <skip> [bench args...]
TAG=$1; KREGEX=$2; SKIP=$3; shift 3
mkdir -p gpurun_out
timeout -s KILL 240 ncu --set full --clock-control none --import-source on -k "regex:$KREGEX" -s $SKIP -c 1 -f -o gpurun_out/prof_${TAG} \
   python bench.py "$@" --no-cpu-baseline > gpurun_out/ncu_${TAG}.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_${TAG}.log
