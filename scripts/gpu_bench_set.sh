#!/bin/bash
# the single-GPU bench lines committed under profiles/ ($1 = tag): headline, the other BASELINE configurations, the
# reference's own job shapes (32 / 64 samples), graph replay vs direct launches, a complete 1000-step trajectory
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
run() { out=$1; shift; timeout 900 python bench.py "$@" 2> gpurun_out/$out.err | tail -1 > gpurun_out/$out.json; echo "$out: $(head -c 160 gpurun_out/$out.json)"; }
run bench_${TAG}_cfg2_4096 --steps 20 --warmup 5
run bench_${TAG}_reference_arm --impl reference --steps 5 --warmup 2
for w in cfg1 cfg3 cfg4 cfg5; do run bench_${TAG}_$w --workload $w --steps 5 --warmup 3 --no-cpu-baseline; done
for n in 32 64; do run bench_${TAG}_cfg2_$n --samples $n --steps 50 --warmup 5 --no-cpu-baseline; done
B200MD_PC_DIRECT=1 run bench_${TAG}_cfg2_32_direct --samples 32 --steps 50 --warmup 5 --no-cpu-baseline
run bench_${TAG}_cfg1_64x100 --workload cfg1 --samples 64 --steps 50 --warmup 5 --no-cpu-baseline
run bench_${TAG}_full_trajectory_1024 --samples 1024 --full-trajectory
