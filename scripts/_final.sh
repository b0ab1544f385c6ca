python bench.py > gpurun_out/r1_final_bench_1gpu.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
python bench.py --impl reference > gpurun_out/r1_final_bench_reference.json 2>> gpurun_out/bench.err
