#!/bin/bash
# One gpurun call per GPU count: what the box gives pinned uploads, the bench (torchrun, one rank per GPU) and the
# config-5 corpus sweep.  Usage (on the GPU box): scripts/multi_gpu_round.sh N [corpus images] [cpu-arm images]
N=${1:-2}; COUNT=${2:-2000}; CPUC=${3:-192}
mkdir -p gpurun_out
nproc > gpurun_out/r2_box_${N}gpu.txt; nvidia-smi -L >> gpurun_out/r2_box_${N}gpu.txt
nvidia-smi topo -m >> gpurun_out/r2_box_${N}gpu.txt 2>&1
scripts/ubench/h2d_bw --ms 400 > gpurun_out/r2_h2d_bw_${N}gpu.txt 2>&1
if [ "$N" -gt 1 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
else
  python bench.py --gpus 1 --steps 200 --warmup 5 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
fi
python scripts/corpus_bench.py --gpus $N --count $COUNT --cpu-count $CPUC --out gpurun_out/r2_cfg5_${N}gpu.json \
    > gpurun_out/r2_cfg5_${N}gpu.log 2>&1
tail -3 gpurun_out/r2_h2d_bw_${N}gpu.txt
python - <<PY
import json
b=json.load(open("gpurun_out/r2_bench_${N}gpu.json"))
print("bench N=${N}: value", b["value"], "e2e", b["e2e"]["value"], "h2d GB/s per rank", b["e2e"].get("h2d_gbs"), b["e2e"].get("cpu_binding"))
c=json.load(open("gpurun_out/r2_cfg5_${N}gpu.json"))
print("cfg5 N=${N}:", c["gpu_arm"]["encodes_per_s"], "enc/s on", c["host_cores"], "cores; cpu arm", c.get("cpu_arm",{}).get("encodes_per_s"), "digest", c["gpu_arm"]["csv_digest_image_bytes_passes"])
PY
