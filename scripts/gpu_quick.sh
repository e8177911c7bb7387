#!/bin/bash
# quick GPU loop: EGNN parity tests, then a short bench per kernel implementation in $KERNELS (default f16x3ts)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu ${PYTEST_ARGS:-} > gpurun_out/pytest_quick.log 2>&1
tail -3 gpurun_out/pytest_quick.log
for k in ${KERNELS:-f16x3ts}; do
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --kernels $k ${BENCH_ARGS:-} > gpurun_out/bench_$k.log 2> gpurun_out/bench_$k.err
  python - "$k" <<'PY'
import json, sys
k = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_{k}.log").read().strip().splitlines()[-1])
    print(k, round(d["value"], 2), "structures/s", round(d["ms_per_step"], 1), "ms/step; edge launch", round(d["roofline"]["avg_launch_ms"], 2), "ms",
          {a: round(b, 2) for a, b in d["roofline"]["class_ms_per_step"].items()}, d["clocks"])
except Exception as exc:
    print(k, "bench failed", exc)
    print(open(f"gpurun_out/bench_{k}.err").read()[-1500:])
PY
done
