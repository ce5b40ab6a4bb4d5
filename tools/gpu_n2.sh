#!/bin/bash
# 2-GPU visit (gpurun --gpus 2): the NCCL shard test and the N = 2 bench line.  usage: gpurun --gpus 2 -- bash tools/gpu_n2.sh <tag>
TAG=${1:-n2}
mkdir -p gpurun_out
timeout -s KILL 70 python -m pytest tests/test_gpu_sharded.py -m gpu -q -rs > gpurun_out/pytest_sharded_${TAG}.log 2>&1; echo "sharded test rc=$?"; tail -3 gpurun_out/pytest_sharded_${TAG}.log
port=$((29500 + RANDOM % 2000))
timeout -s KILL 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 10 --warmup 3 --quick --no-cpu-baseline > gpurun_out/scale_${TAG}_n2.json 2> gpurun_out/scale_${TAG}_n2.err; echo "bench n2 rc=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_${TAG}_n2.json")); print("n2 value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]))
except Exception as e: print("ERR", e)
PY
