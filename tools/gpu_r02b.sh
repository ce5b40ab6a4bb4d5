#!/bin/bash
# round 2, second GPU visit: transposed fused conv1d_1 + block 1 kernel -- parity tests, A/B against the r01 form, ncu
mkdir -p gpurun_out
rm -f gpurun_out/agreement.jsonl
KWS_AGREEMENT_CLIPS=32768 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_agreement.py -m gpu -q --timeout 600 -x > gpurun_out/pytest_gpu_b.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu_b.log
timeout 300 python bench.py --steps 10 --quick --no-cpu-baseline > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r02b.err
KWS_FUSE_V1=1 timeout 300 python bench.py --steps 10 --quick --no-cpu-baseline > gpurun_out/bench_r02b_v1.json 2> gpurun_out/bench_r02b_v1.err; echo "v1 rc=$?"
for c in 4 5; do timeout 600 python bench.py --config $c --steps 10 > gpurun_out/bench_r02b_c$c.json 2> gpurun_out/bench_r02b_c$c.err; echo "config $c rc=$?"; tail -c 300 gpurun_out/bench_r02b_c$c.err; done
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:conv1_block1" -s 4 -c 1 -f -o gpurun_out/prof_r02b_fused \
   python bench.py --batch 4096 --steps 1 --warmup 3 --quick --no-cpu-baseline > gpurun_out/ncu_r02b.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
for n in ("bench_r02b","bench_r02b_v1","bench_r02b_c4","bench_r02b_c5"):
    try:
        d=json.load(open(f"gpurun_out/{n}.json"))
        print(n, round(d["value"]), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["frac"],3), d.get("block_ms_per_step"), {k:round(v["ms_per_step"],3) for k,v in d.get("kernel_classes",{}).items()})
        for r in d.get("sweep",[]): print("   ", r)
    except Exception as e: print(n, "ERR", e)
PY
