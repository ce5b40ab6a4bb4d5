#!/bin/bash
# configs 4 and 5 only.  usage: gpurun -- bash tools/gpu_c45.sh <tag>
TAG=${1:-c45}
mkdir -p gpurun_out
for c in 4 5; do timeout -s KILL 200 python bench.py --config $c --steps 10 > gpurun_out/bench_${TAG}_c$c.json 2> gpurun_out/bench_${TAG}_c$c.err; echo "config $c rc=$?"; done
python - <<PY
import json
for n in ("bench_${TAG}_c4","bench_${TAG}_c5"):
    d=json.load(open(f"gpurun_out/{n}.json")); print(n, round(d["value"]), "e2e", round(d["e2e"]["value"]), d["config"].get("host_chunk_clips"))
PY
