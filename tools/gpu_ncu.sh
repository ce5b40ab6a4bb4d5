#!/bin/bash
# one ncu --set full capture of one kernel of a bench step.  usage: gpurun -- bash tools/gpu_ncu.sh <tag> <kernel regex> [skip]
TAG=$1; KREGEX=$2; SKIP=${3:-4}
mkdir -p gpurun_out
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k "regex:$KREGEX" -s $SKIP -c 1 -f -o gpurun_out/prof_${TAG} \
   python bench.py --batch 4096 --steps 1 --warmup 3 --quick --no-cpu-baseline > gpurun_out/ncu_${TAG}.log 2>&1; echo "ncu rc=$?"
