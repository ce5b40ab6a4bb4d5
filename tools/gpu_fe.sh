#!/bin/bash
# front-end visit: the front-end parity tests, config 2 (front end only) and the default bench line.
# usage: gpurun -- bash tools/gpu_fe.sh <tag>
TAG=${1:-fe}
mkdir -p gpurun_out
timeout -s KILL 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "features or contrib or audio_converter or host_entry" > gpurun_out/pytest_${TAG}.log 2>&1; rc=$?
echo "pytest rc=$rc"; tail -4 gpurun_out/pytest_${TAG}.log
if [ $rc -ne 0 ]; then tail -40 gpurun_out/pytest_${TAG}.log; exit 1; fi
timeout -s KILL 100 python bench.py --config 2 --steps 10 --no-cpu-baseline > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_${TAG}_c2.err; echo "config 2 rc=$?"
timeout -s KILL 120 python bench.py --steps 10 --quick --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_${TAG}.err
python - <<PY
import json
for t in ("${TAG}_c2", "${TAG}"):
  try:
    d=json.load(open("gpurun_out/bench_%s.json" % t))
    print(t, round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "roof", d.get("roofline",{}).get("frac"), {k:round(v["ms_per_step"],3) for k,v in d.get("kernel_classes",{}).items()})
  except Exception as e: print(t, "ERR", e)
PY
