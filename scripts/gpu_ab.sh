#!/bin/bash
# A/B of prebuilt library variants in scratch_libs/*.so: parity (edge / node kernel tests) + short bench for each
set -u
mkdir -p gpurun_out
LIB=diffusion_for_multi_scale_molecular_dynamics_b200/csrc/libb200md.so
cp $LIB /tmp/lib_orig.so
for round in ${ROUNDS:-1}; do
for v in scratch_libs/*.so; do
  name=$(basename $v .so)
  cp $v $LIB
  if [ "$round" = "1" ]; then
    timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "production or tcgen05_edge or range_guard or short_segments" > gpurun_out/pytest_$name.log 2>&1
    echo "$name: $(tail -1 gpurun_out/pytest_$name.log)"
  fi
  timeout 300 python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/bench_${name}_$round.log 2> gpurun_out/bench_${name}_$round.err
  python - "$name" "$round" <<'PY'
import json, sys
k, r = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(f"gpurun_out/bench_{k}_{r}.log").read().strip().splitlines()[-1])
    print(k, r, round(d["value"], 2), "structures/s", round(d["ms_per_step"], 1), "ms/step; e2e", round(d["e2e"]["value"], 2), "edge launch", round(d["roofline"]["avg_launch_ms"], 2), "ms; node", round(d["roofline"]["class_ms_per_step"]["node"], 2), "ms/step", d["clocks"])
except Exception as exc:
    print(k, r, "bench failed", exc, open(f"gpurun_out/bench_{k}_{r}.err").read()[-800:])
PY
done
done
cp /tmp/lib_orig.so $LIB
