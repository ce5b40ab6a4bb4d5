#!/bin/bash
# compute-sanitizer memcheck over the kernels rewritten at the end of r02 (front end, head prologue); small cases, hard time limits
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL 120 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest "$@" -m gpu -q -x > gpurun_out/sanitize_$name.log 2>&1; echo "== $name rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" gpurun_out/sanitize_$name.log | head -5; }
run frontend tests/test_gpu_parity.py -k "test_features_tc or contrib_audio or audio_converter"
run head_tc tests/test_gpu_parity.py -k "test_forward_tc and not layers"
