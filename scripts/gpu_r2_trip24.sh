#!/bin/bash
# Round 2 trip 24: (1) the test order that failed in trips 21 / 22, now with diagnostics; (2) the whole suite in the driver's order;
# (3) the default bench.
mkdir -p gpurun_out
T0=$SECONDS
note() { echo "== $1 (t=$((SECONDS-T0))s)"; }
note "1. failing order"
timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_engine.py -m gpu -q --tb=short > gpurun_out/t24_order.txt 2>&1; grep -E "passed|failed|AssertionError|differ" gpurun_out/t24_order.txt | cut -c1-1200 | head -20
note "2. whole GPU suite"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final2_pytest_gpu.txt 2>&1; echo "exit $?" >> gpurun_out/final2_pytest_gpu.txt; tail -5 gpurun_out/final2_pytest_gpu.txt
note "3. bench default"
timeout 500 python bench.py > gpurun_out/final2_bench_default.json 2> gpurun_out/final2_bench_default.err; cut -c1-160 gpurun_out/final2_bench_default.json; tail -2 gpurun_out/final2_bench_default.err
note "done"
