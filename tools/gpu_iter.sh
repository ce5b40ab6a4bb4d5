#!/bin/bash
# One development visit: parity tests, then a bench line with the per-class breakdown.
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -x -q -m gpu ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/pytest_iter.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_iter.log
tail -25 gpurun_out/pytest_iter.log
timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --max-rows ${MR:-8192} > gpurun_out/bench_iter.json 2>gpurun_out/bench_iter.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_iter.json"))
    print("value",round(d["value"]),"ms/step",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"]))
    print({k:round(v["ms_per_step"],3) for k,v in d["kernel_classes"].items()}); print("blocks",d.get("block_ms_per_step"))
    print("roofline",d["roofline"]["achieved"],d["roofline"]["frac"])
except Exception as e:
    print("bench failed",e); print(open("gpurun_out/bench_iter.err").read()[-2000:])
PY
