#!/bin/bash
# bench + ncu evidence.  Usage: gpurun -- bash tools/gpu_bench.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench exit $?"
cat gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_${TAG}.csv \
   python bench.py --batch 1024 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1; echo "ncu launches exit $?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 26 -c 12 -f -o gpurun_out/prof_tc_${TAG} \
   python bench.py --batch 1024 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out | tail -12
