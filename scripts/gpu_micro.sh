#!/bin/bash
# Kernel tests, conv microbench, bench, launch list, full GPU suite.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_kernels.log
timeout 600 python scripts/conv_microbench.py > gpurun_out/conv_microbench.log 2>&1; tail -3 gpurun_out/conv_microbench.log | cut -c1-300
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_full.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 1200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 12 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
