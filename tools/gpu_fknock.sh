#!/bin/bash
# knockout runs of the fused conv1d_1 + block 1 kernel (profiling build, results are wrong on purpose): which role bounds it?
mkdir -p gpurun_out
export KWS_LIBKWS=$PWD/speech_recognition_b200/libkws_prof.so
for k in 0 1 2 4 8 3 9 11; do
  KWS_FKNOCK=$k timeout -s KILL 100 python bench.py --steps 5 --quick --no-cpu-baseline --batch 8192 > gpurun_out/fknock_$k.json 2> gpurun_out/fknock_$k.err; echo "knock $k rc=$?"
done
python - <<'PY'
import json
for k in (0,1,2,4,8,3,9,11):
    try:
        d=json.load(open(f"gpurun_out/fknock_{k}.json")); print("knock", k, "slice_conv1 ms/step", round(d["kernel_classes"]["slice_conv1"]["ms_per_step"],3))
    except Exception as e: print(k, "ERR", e)
PY
