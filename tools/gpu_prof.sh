#!/bin/bash
# ncu evidence: launch list of one bench step + --set full of every kernel of one step (after 3 warm-up steps).
# One step at --batch 4096 --max-rows 32768 = 14 launches: augment, stft_mel, conv1_block1 (fused), 10 x tc_gemm, head.
# Usage: gpurun -- bash tools/gpu_prof.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
   python bench.py --batch 4096 --steps 1 --warmup 3 --no-cpu-baseline --max-rows 32768 > gpurun_out/ncu_launch_${TAG}.log 2>&1; echo "ncu launches exit $?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k "regex:augment|stft_mel|conv1_block1|tc_gemm|head_kernel" -s 42 -c 14 -f -o gpurun_out/prof_${TAG} \
   python bench.py --batch 4096 --steps 1 --warmup 3 --no-cpu-baseline --max-rows 32768 > gpurun_out/ncu_full_${TAG}.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out | tail -6
