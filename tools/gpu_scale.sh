#!/bin/bash
# multi-GPU visit (gpurun --gpus 8): NCCL shard test, the 1/2/4/8 scaling curve of the default bench line, the fixed job and
# configs 4 / 5 on 8 GPUs.  usage: gpurun --gpus 8 -- bash tools/gpu_scale.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${TAG}.txt 2>&1
run() {  # run <n> <out> <bench args...>
  local n=$1 out=$2; shift 2
  local port=$((29500 + RANDOM % 2000))
  if [ "$n" = "1" ]; then timeout -s KILL 200 python bench.py --gpus 1 "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err
  else timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err; fi
  echo "$out rc=$?"
}
timeout -s KILL 300 python -m pytest tests/test_gpu_sharded.py -m gpu -q -rs > gpurun_out/pytest_sharded_${TAG}.log 2>&1; echo "sharded test rc=$?"; tail -3 gpurun_out/pytest_sharded_${TAG}.log
for n in 1 2 4 8; do run $n scale_${TAG}_n$n --steps 10 --warmup 3 --quick --no-cpu-baseline; done
run 8 job_${TAG}_n8 --config 3 --job --steps 5 --no-cpu-baseline
run 8 c4_${TAG}_n8 --config 4 --steps 10 --no-cpu-baseline
run 8 c5_${TAG}_n8 --config 5 --batch 16384 --steps 10 --no-cpu-baseline
python - <<PY
import json
for n in ("scale_${TAG}_n1","scale_${TAG}_n2","scale_${TAG}_n4","scale_${TAG}_n8","job_${TAG}_n8","c4_${TAG}_n8","c5_${TAG}_n8"):
    try:
        d=json.load(open(f"gpurun_out/{n}.json"))
        print(n, "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), d["scaling"])
    except Exception as e: print(n, "ERR", e)
PY
