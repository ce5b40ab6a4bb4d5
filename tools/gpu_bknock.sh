#!/bin/bash
# knockout runs of the depthwise+pointwise block kernels (profiling build, results are wrong on purpose)
mkdir -p gpurun_out
export KWS_LIBKWS=$PWD/speech_recognition_b200/libkws_prof.so
for k in 0 32 30 62; do
  KWS_KNOCKOUT=$k timeout -s KILL 100 python bench.py --steps 5 --quick --no-cpu-baseline --batch 8192 > gpurun_out/bknock_$k.json 2> gpurun_out/bknock_$k.err; echo "knock $k rc=$?"
done
python - <<'PY'
import json
print("bits: 1 no TMA stores, 2 no FIR, 4 no MMAs, 8 no epilogue math, 16 no raw loads, 32 no weight loads")
for k in (0,32,30,62):
    try:
        d=json.load(open(f"gpurun_out/bknock_{k}.json")); print("knock", k, [round(x,3) for x in d["block_ms_per_step"]])
    except Exception as e: print(k, "ERR", e)
PY
