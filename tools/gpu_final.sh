#!/bin/bash
# round-end evidence: every GPU test, smoke, the bench line of every config, the ncu launch list of the bench command and one
# ncu --set full capture of every kernel of a step (last: it is the longest and the least urgent).
# usage: gpurun -- bash tools/gpu_final.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out; rm -f gpurun_out/agreement.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi_${TAG}.txt 2>&1
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 400 -rs > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_${TAG}.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_${TAG}.log
timeout -s KILL 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
timeout -s KILL 200 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; echo "reference rc=$?"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
   python bench.py --batch 4096 --steps 2 --warmup 3 --quick --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1; echo "ncu launches rc=$?"
for c in 2 4 5; do timeout -s KILL 300 python bench.py --config $c --steps 10 > gpurun_out/bench_${TAG}_c$c.json 2> gpurun_out/bench_${TAG}_c$c.err; echo "config $c rc=$?"; done
timeout -s KILL 200 python bench.py --config 3 --job --steps 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_job.json 2> gpurun_out/bench_${TAG}_job.err; echo "job rc=$?"
python - <<PY
import json
for n in ("bench_${TAG}","bench_${TAG}_reference","bench_${TAG}_c2","bench_${TAG}_c4","bench_${TAG}_c5","bench_${TAG}_job"):
    try:
        d=json.load(open(f"gpurun_out/{n}.json"))
        print(n, round(d["value"]), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "roof", d.get("roofline",{}).get("frac"), d.get("block_ms_per_step"), {k:round(v["ms_per_step"],3) for k,v in d.get("kernel_classes",{}).items()})
    except Exception as e: print(n, "ERR", e)
PY
if [ "${2:-}" != "nofull" ]; then
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k "regex:augment|stft_mel|conv1_block1|tc_gemm|head_kernel" -s 42 -c 14 -f -o gpurun_out/prof_${TAG} \
   python bench.py --batch 4096 --steps 1 --warmup 3 --quick --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1; echo "ncu full rc=$?"
fi
