#!/bin/bash
# A/B of prebuilt library variants in scratch_libs/*.so: parity (edge kernel tests) + short bench for each
set -u
mkdir -p gpurun_out
LIB=diffusion_for_multi_scale_molecular_dynamics_b200/csrc/libb200md.so
cp $LIB /tmp/lib_orig.so
for v in scratch_libs/*.so; do
  name=$(basename $v .so)
  cp $v $LIB
  timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "production_width or tcgen05_edge or range_guard" > gpurun_out/pytest_$name.log 2>&1
  echo "$name: $(tail -1 gpurun_out/pytest_$name.log)"
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --kernels f16x3ts ${BENCH_ARGS:-} > gpurun_out/bench_$name.log 2> gpurun_out/bench_$name.err
  python - "$name" <<'PY'
import json, sys
k = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_{k}.log").read().strip().splitlines()[-1])
    print(k, round(d["value"], 2), "structures/s", round(d["ms_per_step"], 1), "ms/step; edge launch", round(d["roofline"]["avg_launch_ms"], 2), "ms; node", round(d["roofline"]["class_ms_per_step"]["node"], 2), "ms/step", d["clocks"])
except Exception as exc:
    print(k, "bench failed", exc)
PY
done
cp /tmp/lib_orig.so $LIB
