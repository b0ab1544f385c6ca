python bench.py --no-cpu > gpurun_out/bench_r1w.json 2> gpurun_out/bench_r1w.err; tail -3 gpurun_out/bench_r1w.err; python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_r1w.json')); print(d['value'], d['ms_per_step'], d['e2e'])
PY
