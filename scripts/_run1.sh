python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python scripts/profile_step.py --steps 6 | tail -2
python bench.py --no-cpu > gpurun_out/bench_r1u.json 2> gpurun_out/bench_r1u.err; cut -c1-330 gpurun_out/bench_r1u.json; python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_r1u.json')); print(d['kernel_ms'], d['cached_source'], d['other_blur'], d['roofline']['whole_step'])
PY
