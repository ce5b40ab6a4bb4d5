#!/bin/bash
# Pipeline event traces of one block kernel (CTA 0).  Usage: gpurun -- bash tools/gpu_trace.sh <block idx> [<block idx> ...]
mkdir -p gpurun_out
KWS_PROFILE_BUILD=1 python -m speech_recognition_b200.build --force > gpurun_out/profile_build.log 2>&1 || { tail gpurun_out/profile_build.log; exit 1; }
for B in "$@"; do
  for KO in 0 31; do
    KWS_TRACE=$B KWS_KNOCKOUT=$KO timeout -s KILL 200 python bench.py --batch 4096 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/trace_${B}_${KO}.err
    grep "kws trace" gpurun_out/trace_${B}_${KO}.err
    mv gpurun_out/trace.bin gpurun_out/trace_${B}_${KO}.bin
  done
done
ls -la gpurun_out/*.bin
