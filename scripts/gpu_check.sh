#!/bin/bash
# quick GPU regression: the fast parity tests, then one bench line (summary printed)
timeout 900 python -m pytest tests -m "gpu and not slow" -q 2>&1 | tail -6
timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu "$@" > gpurun_out/r2_bench_check.json 2> gpurun_out/r2_bench_check.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_check.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "sync", d["sync_calls"]["ms_per_step"], "two", d["two_callers"]["ms_per_step"], "cached", d["cached_source"]["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["frac"], d["kernel_ms"])
PY
tail -3 gpurun_out/r2_bench_check.err
