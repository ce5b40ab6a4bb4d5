#!/bin/bash
# A/B of an environment switch on the bench's per-block times.  Usage: gpurun -- bash tools/gpu_ab.sh "VAR=a" "VAR=b" ...
mkdir -p gpurun_out
for CFG in "$@"; do
  env $CFG timeout -s KILL 200 python bench.py --steps ${AB_STEPS:-5} --warmup 3 --no-cpu-baseline ${AB_ARGS} > gpurun_out/ab.json 2>gpurun_out/ab.err
  python - "$CFG" <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/ab.json"))
    print(sys.argv[1],"value",round(d["value"]),"ms/step",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"]),{k:round(v["ms_per_step"],3) for k,v in d["kernel_classes"].items()})
    print("   blocks",d.get("block_ms_per_step"))
except Exception as e:
    print(sys.argv[1],"failed",e); print(open("gpurun_out/ab.err").read()[-1500:])
PY
done
