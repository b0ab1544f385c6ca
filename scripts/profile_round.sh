#!/bin/bash
# ncu evidence for the round: each command first runs WITHOUT ncu and must exit 0.
#  1. launch list (gpu__time_duration.sum) of the bench command itself and of profile_step.py (the same step, no reporting)
#  2. one --set full capture of profile_step.py with source import, read back here with scripts/ncu_summary.py / ncu_hot.py
set -e
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_prof_bench_plain.json 2> gpurun_out/r2_prof_bench_plain.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_final_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_prof_bench_ncu.log 2>&1 || echo "bench launch list failed"
timeout 600 python scripts/profile_step.py --steps 3 > gpurun_out/r2_prof_step_plain.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_final_launches.csv \
    python scripts/profile_step.py --steps 3 > gpurun_out/r2_prof_step_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -f -o gpurun_out/r2_final_prof \
    python scripts/profile_step.py --steps 2 > gpurun_out/r2_prof_full.log 2>&1
ls -la gpurun_out/r2_final_prof.ncu-rep
tail -3 gpurun_out/r2_prof_step_plain.log
