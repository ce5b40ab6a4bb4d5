#!/bin/bash
# chunk-size sweep of the forward
mkdir -p gpurun_out
for MR in ${@:-8192 16384 32768}; do
  timeout -s KILL 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --max-rows $MR > gpurun_out/sweep_$MR.json 2>gpurun_out/sweep_$MR.err
  python - $MR <<'PY'
import json,sys
mr=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/sweep_{mr}.json"))
    print(mr,"value",round(d["value"]),"ms/step",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"]),{k:round(v["ms_per_step"],3) for k,v in d["kernel_classes"].items()})
except Exception as e:
    print(mr,"failed",e); print(open(f"gpurun_out/sweep_{mr}.err").read()[-800:])
PY
done
