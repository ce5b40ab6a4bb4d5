#!/bin/bash
# round 2, first GPU visit: all GPU tests (new agreement / drop-in tests included), every bench config, FIR A/B
mkdir -p gpurun_out; cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
rm -f gpurun_out/agreement.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nvidia-smi topo -m >> gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -rs > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r02a.err
for c in 2 4 5; do timeout 600 python bench.py --config $c --steps 10 > gpurun_out/bench_r02a_c$c.json 2> gpurun_out/bench_r02a_c$c.err; echo "config $c rc=$?"; tail -c 300 gpurun_out/bench_r02a_c$c.err; done
timeout 600 python bench.py --config 3 --job --steps 3 --no-cpu-baseline > gpurun_out/bench_r02a_job.json 2> gpurun_out/bench_r02a_job.err; echo "job rc=$?"
# A/B: the r01 packed-half FIR against the fp32-accumulated one
cp speech_recognition_b200/libkws.so /tmp/libkws_fp32fir.so
KWS_FIR_FP16=1 timeout 600 python -m speech_recognition_b200.build --force > gpurun_out/build_fir16.log 2>&1
timeout 300 python bench.py --steps 10 --quick --no-cpu-baseline > gpurun_out/bench_r02a_fir16.json 2> gpurun_out/bench_r02a_fir16.err; echo "fir16 rc=$?"
cp /tmp/libkws_fp32fir.so speech_recognition_b200/libkws.so
timeout 300 python bench.py --steps 10 --quick --no-cpu-baseline > gpurun_out/bench_r02a_fir32.json 2> gpurun_out/bench_r02a_fir32.err; echo "fir32 rc=$?"
python - <<'PY'
import json
for n in ("bench_r02a","bench_r02a_fir16","bench_r02a_fir32","bench_r02a_c2","bench_r02a_c4","bench_r02a_c5","bench_r02a_job"):
    try:
        d=json.load(open(f"gpurun_out/{n}.json"))
        print(n, round(d["value"]), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["frac"],3), d.get("block_ms_per_step"), {k:round(v["ms_per_step"],3) for k,v in d.get("kernel_classes",{}).items()})
        for k,v in d.get("e2e_variants",{}).items(): print("   ", k, round(v["value"]))
        for r in d.get("sweep",[]): print("   ", r)
    except Exception as e: print(n, "ERR", e)
PY
