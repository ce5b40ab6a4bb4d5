#!/bin/bash
# compute-sanitizer memcheck over the kernels that are new in r02 (small cases; hard time limits)
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL 280 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest "$@" -m gpu -q -x > gpurun_out/sanitize_$name.log 2>&1; echo "$name rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" gpurun_out/sanitize_$name.log | head -5; }
run stretch tests/test_gpu_dropin.py -k "time_stretch"
run head_fp32 tests/test_gpu_parity.py -k "test_forward_fp32 or test_convert_32_to_12 or test_vote_tie"
run tc tests/test_gpu_parity.py -k "test_forward_tc or fused_and_unfused"
run families tests/test_gpu_parity.py -k "time_sliced"
run hostpipe tests/test_gpu_dropin.py -k "pcm16_and_staging"
