python scripts/run_configs.py > gpurun_out/r1_configs_1gpu.json 2> gpurun_out/configs.err; tail -5 gpurun_out/configs.err
cp gpurun_out/corpus_1gpu.csv gpurun_out/r1_corpus_1gpu.csv 2>/dev/null
