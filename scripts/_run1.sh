python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python scripts/profile_step.py --steps 6 | tail -2
python scripts/rows_variants.py
python bench.py --no-cpu > gpurun_out/bench_r1t.json 2> gpurun_out/bench_r1t.err; cat gpurun_out/bench_r1t.json
