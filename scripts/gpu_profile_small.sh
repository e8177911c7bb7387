#!/bin/bash
# ncu --set full of the memory-side kernels (graph / score / update), one launch each, from the standalone microbench
# (sizes beyond L2) and -- for the internal EGNN graph kernels -- from the bench.  $1 = tag.
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
python scripts/kernel_rooflines.py > gpurun_out/kernel_rooflines_${TAG}.json 2> gpurun_out/kernel_rooflines_${TAG}.err || tail -5 gpurun_out/kernel_rooflines_${TAG}.err
for K in pc_coordinates_update_v4 atom_types_update wrapped_gaussian_score wrap_kernel_v4 radius_graph_count radius_graph_fill; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -f -o gpurun_out/prof_${TAG}_small_$K \
      python scripts/kernel_rooflines.py > gpurun_out/ncu_small_${TAG}_$K.log 2>&1
  echo "$K exit $?"
done
ncu --set full --clock-control none --import-source on -k regex:egnn_graph -s 9 -c 3 -f -o gpurun_out/prof_${TAG}_small_egnn_graph \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_small_${TAG}_egnn_graph.log 2>&1
echo "egnn_graph exit $?"
ls -la gpurun_out/*small* | tail -12
