#!/bin/bash
# Round-end measurements on one B200 box: full GPU test suite, both bench arms the way the driver runs them, the
# batched-probing study and BASELINE configs 1, 3, 4.
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -q --durations=8) > gpurun_out/r2_final_tests.log 2>&1; tail -14 gpurun_out/r2_final_tests.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench_driver_args.json 2> gpurun_out/r2_final_bench_driver_args.err
timeout 600 python bench.py > gpurun_out/r2_final_bench_1gpu.json 2> gpurun_out/r2_final_bench_1gpu.err
timeout 600 python bench.py --blur fir --no-cpu --steps 300 > gpurun_out/r2_final_bench_1gpu_fir.json 2> gpurun_out/r2_final_bench_1gpu_fir.err
timeout 600 python scripts/batched_probe.py --images 12 --out gpurun_out/r2_batched_probe.json > gpurun_out/r2_batched_probe.log 2>&1
timeout 600 python scripts/run_configs.py --only=cfg1 --only=cfg3 --only=cfg4 > gpurun_out/r2_configs_1gpu.json 2> gpurun_out/r2_configs_1gpu.err
timeout 600 python scripts/corpus_bench.py --gpus 1 --count 256 --cpu-count 0 --pinned-ab 256 --out gpurun_out/r2_decode_handoff.json > gpurun_out/r2_decode_handoff.log 2>&1
python - <<PY
import json
for f in ("r2_final_bench_reference","r2_final_bench_driver_args","r2_final_bench_1gpu","r2_final_bench_1gpu_fir"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d.get("ms_per_step"), d.get("e2e",{}).get("value"), d.get("roofline",{}).get("frac"), d.get("cpu_baseline",{}).get("value"))
    except Exception as e: print(f, "ERR", e)
try:
    b=json.load(open("gpurun_out/r2_batched_probe.json")); print(b["mean_ms_per_image"], b["encodes_per_image"], b["wasted_share"])
    h=json.load(open("gpurun_out/r2_decode_handoff.json")); print(h["decode_handoff"])
except Exception as e: print("ERR", e)
PY
