#!/bin/bash
mkdir -p gpurun_out
for CFG in "$@"; do
  env $CFG timeout -s KILL 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab.json 2>gpurun_out/ab.err
  python - "$CFG" <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/ab.json")); print(sys.argv[1],"value",round(d["value"]),"e2e",round(d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[1],"failed",e); print(open("gpurun_out/ab.err").read()[-800:])
PY
done
