#!/bin/bash
# ncu evidence for the bench command (run via gpurun, 1 GPU).  $1 = tag for the output names.
# launch list of the whole bench + one --set full capture each of the edge and the node tensor-memory kernels
set -u
TAG=${1:-r02}
ARGS="--steps 2 --warmup 3 --samples ${SAMPLES:-4096} --no-cpu-baseline"
mkdir -p gpurun_out
python bench.py $ARGS > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py $ARGS > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list exit $?"
for K in egnn_edge_ts egnn_node_ts; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 12 -c 2 -f \
      -o gpurun_out/prof_${TAG}_$K python bench.py $ARGS > gpurun_out/ncu_full_${TAG}_$K.log 2>&1
  echo "full capture $K exit $?"
done
ls -la gpurun_out | tail -8
