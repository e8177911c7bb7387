#!/bin/bash
# Run on the GPU box (via gpurun): build, GPU parity tests, smoke, a short bench.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py ${BENCH_ARGS:---steps 3 --warmup 3} > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -2 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
