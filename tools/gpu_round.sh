#!/bin/bash
# One GPU-box visit: tests, smoke, quick perf.  Usage: gpurun -- bash tools/gpu_round.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -8 gpurun_out/smoke.log
timeout -s KILL 300 python tools/quick_perf.py tc 4096 > gpurun_out/perf_tc.log 2>&1
cat gpurun_out/perf_tc.log
