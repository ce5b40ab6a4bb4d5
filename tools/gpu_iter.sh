#!/bin/bash
# One development visit: TC parity tests, then forward timing at several chunk sizes.
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests -x -q -m gpu -k "tc or host_entry" > gpurun_out/pytest_tc.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_tc.log
tail -15 gpurun_out/pytest_tc.log
for mr in ${MRS:-512 2048 8192}; do
  timeout -s KILL 200 python tools/quick_perf.py tc 4096 $mr 2>&1 | grep -E "forward|Error|error" | tee -a gpurun_out/perf_iter.log
done
