#!/bin/bash
# 8-GPU bench lines ($1 = tag), launched the way the driver launches them: weak (4096 per GPU), strong cfg2 (4096 in all),
# strong cfg5 (8192 x 512 atoms in all).  gpurun --gpus 8 -- 'bash scripts/gpu_bench_8.sh r02c'
set -u
TAG=${1:-r02}
N=${GPUS:-8}
mkdir -p gpurun_out
run() { out=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" 2> gpurun_out/$out.err | grep '^{' | tail -1 > gpurun_out/$out.json; echo "$out: $(head -c 200 gpurun_out/$out.json)"; }
run bench_${TAG}_cfg2_${N}gpu --steps 20 --warmup 5 --no-cpu-baseline
run bench_${TAG}_cfg2_strong_${N}gpu --steps 20 --warmup 5 --strong --no-cpu-baseline
run bench_${TAG}_cfg5_strong_${N}gpu --workload cfg5 --steps 5 --warmup 3 --strong --no-cpu-baseline
