#!/bin/bash
# cycle traces of the edge kernel for the library variants scratch_libs/t_*.so (built with -DB200MD_TS_TRACE)
set -u
mkdir -p gpurun_out
LIB=diffusion_for_multi_scale_molecular_dynamics_b200/csrc/libb200md.so
cp $LIB /tmp/lib_orig.so
for v in scratch_libs/t_*.so; do
  name=$(basename $v .so)
  cp $v $LIB
  timeout 300 python scripts/ts_trace.py ${SAMPLES:-4096} > gpurun_out/ts_trace_$name.txt 2> gpurun_out/ts_trace_$name.err
  echo "$name: $(wc -l < gpurun_out/ts_trace_$name.txt) records"
done
cp /tmp/lib_orig.so $LIB
