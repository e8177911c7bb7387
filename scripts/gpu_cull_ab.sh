#!/bin/bash
# image culling on / off (scratch_libs/cull_*.so): parity tests of the graph / statistics kernels + standalone timings
set -u
mkdir -p gpurun_out
LIB=diffusion_for_multi_scale_molecular_dynamics_b200/csrc/libb200md.so
cp $LIB /tmp/lib_orig.so
for v in scratch_libs/cull_*.so; do
  name=$(basename $v .so)
  cp $v $LIB
  timeout 600 python -m pytest tests -q -m gpu -k "adjacency or edges or pair_distances or stillinger or silicon or force_field or small_boxes or cutoff" > gpurun_out/pytest_$name.log 2>&1
  echo "$name: $(tail -1 gpurun_out/pytest_$name.log)"
  timeout 600 python scripts/kernel_rooflines.py > gpurun_out/kernel_rooflines_$name.json 2> gpurun_out/kernel_rooflines_$name.err
  python - $name <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/kernel_rooflines_{sys.argv[1]}.json"))
    for k in d["kernels"]:
        print(f'  {k["ms"]:9.3f} ms  {k["achieved_GBps"]:8.1f} GB/s  {k["kernel"][:100]}')
except Exception as exc:
    print("failed", exc, open(f"gpurun_out/kernel_rooflines_{sys.argv[1]}.err").read()[-1500:])
PY
done
cp /tmp/lib_orig.so $LIB
