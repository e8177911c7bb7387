#!/bin/bash
# ncu evidence for the bench command (run via gpurun, 1 GPU).  $1 = tag for the output names.
set -u
TAG=${1:-r01}
ARGS="--steps 2 --warmup 3 --samples ${SAMPLES:-512} --no-cpu-baseline"
mkdir -p gpurun_out
python bench.py $ARGS > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py $ARGS > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list exit $?"
python bench.py $ARGS > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KERNEL_REGEX:-egnn_edge} -s 8 -c 2 -f \
    -o gpurun_out/prof_${TAG} python bench.py $ARGS > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out | tail -12
