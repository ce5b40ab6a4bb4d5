#!/bin/bash
# quick GPU visit: the tensor-core parity tests, one bench line, optionally an ncu capture of one kernel.
# A kernel that dead-locks must not burn the GPU budget: a 150 s canary runs first and everything else is skipped if it hangs.
# usage: gpurun -- bash tools/gpu_quick.sh <tag> [kernel regex for ncu] [skip count]
TAG=${1:-q}; KREGEX=${2:-}; SKIP=${3:-4}
mkdir -p gpurun_out
timeout -s KILL 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_and_unfused or forward_tc" > gpurun_out/canary_${TAG}.log 2>&1; rc=$?
echo "canary rc=$rc"; tail -3 gpurun_out/canary_${TAG}.log
if [ $rc -ne 0 ]; then tail -30 gpurun_out/canary_${TAG}.log; echo "canary failed: skipping the rest"; exit 1; fi
KWS_AGREEMENT_CLIPS=16384 timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_agreement.py -m gpu -q --timeout 300 -x > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_${TAG}.log
timeout -s KILL 120 python bench.py --steps 10 --quick --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_${TAG}.err
KWS_FUSE_V1=1 timeout -s KILL 120 python bench.py --steps 10 --quick --no-cpu-baseline > gpurun_out/bench_${TAG}_v1.json 2> gpurun_out/bench_${TAG}_v1.err; echo "bench v1 rc=$?"
if [ -n "$KREGEX" ]; then
timeout -s KILL 240 ncu --set full --clock-control none --import-source on -k "regex:$KREGEX" -s $SKIP -c 1 -f -o gpurun_out/prof_${TAG} \
   python bench.py --batch 4096 --steps 1 --warmup 3 --quick --no-cpu-baseline > gpurun_out/ncu_${TAG}.log 2>&1; echo "ncu rc=$?"
fi
python - <<PY
import json
for t in ("${TAG}", "${TAG}_v1"):
  try:
    d=json.load(open("gpurun_out/bench_%s.json" % t))
    print(t, round(d["value"]), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["frac"],3), d.get("block_ms_per_step"), {k:round(v["ms_per_step"],3) for k,v in d.get("kernel_classes",{}).items()})
  except Exception as e: print(t, "ERR", e)
PY
