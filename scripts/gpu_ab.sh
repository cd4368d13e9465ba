#!/bin/bash
# A/B trip: kernel tests, then bench.py under each value of an env knob, then the launch list and the full GPU suite.
# Usage: bash scripts/gpu_ab.sh VAR val1 val2 ...     (e.g. AOTB_CONV_TILING wide narrow)
mkdir -p gpurun_out
VAR=$1; shift
echo "== kernel tests"
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_kernels.log
for v in "$@"; do
  echo "== bench $VAR=$v"
  env $VAR=$v timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_${VAR}_$v.log | cut -c1-420
done
echo "== ncu launch list ($VAR=$1)"
env $VAR=$1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 1200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 12 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
