#!/bin/bash
# bench line under a few launch-configuration knobs (A/B aid)
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout -s KILL 150 python bench.py --steps 10 --quick --no-cpu-baseline $EXTRA > gpurun_out/knob_$name.json 2> gpurun_out/knob_$name.err; echo "$name rc=$?"; }
EXTRA="--max-rows 65536" run rows64k A=1
EXTRA="--max-rows 131072" run rows128k A=1
EXTRA="--max-rows 65536" run rows64k_chunk8k KWS_HOST_CHUNK=8192
EXTRA="--max-rows 131072" run rows128k_chunk8k KWS_HOST_CHUNK=8192
EXTRA="--max-rows 65536" run rows64k_chunk2k KWS_HOST_CHUNK=2048
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/knob_*.json")):
    try:
        d=json.load(open(f)); print(f.split("knob_")[1][:-5], round(d["value"]), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), d.get("block_ms_per_step"), {k:round(v["ms_per_step"],3) for k,v in d.get("kernel_classes",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
