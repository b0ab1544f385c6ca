python bench.py --steps 300 --warmup 5 --no-cpu > gpurun_out/r2_bench3a.json 2> gpurun_out/r2_bench3a.err
python bench.py --steps 300 --warmup 5 --no-cpu --source-rows first-score > gpurun_out/r2_bench3b.json 2> gpurun_out/r2_bench3b.err
python - <<PY
import json
for f in ("a","b"):
    d=json.load(open(f"gpurun_out/r2_bench3{f}.json"))
    print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "sync", d["sync_calls"]["ms_per_step"], "two", d["two_callers"]["ms_per_step"], "cached", d["cached_source"]["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["frac"], d["kernel_ms"])
PY
tail -3 gpurun_out/r2_bench3a.err
