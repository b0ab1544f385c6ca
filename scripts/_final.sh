python bench.py > gpurun_out/r1_final_bench_1gpu.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
python bench.py --blur fir --no-cpu > gpurun_out/r1_final_bench_1gpu_fir.json 2>> gpurun_out/bench.err
python bench.py --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches.log 2>&1
