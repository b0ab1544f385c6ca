python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python scripts/profile_step.py --steps 6 | tail -3
python bench.py --no-cpu > gpurun_out/bench_r1v.json 2> gpurun_out/bench_r1v.err; python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_r1v.json')); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['kernel_ms'], d['cached_source'], d['other_blur'], d['roofline']['whole_step'], d['e2e']['value'])
PY
