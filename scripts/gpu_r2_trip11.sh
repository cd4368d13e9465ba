#!/bin/bash
mkdir -p gpurun_out
T0=$SECONDS
note() { echo "== $1 (t=$((SECONDS-T0))s)"; }
note "1. whole GPU suite"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t11_pytest_gpu.txt 2>&1; echo "exit $?" >> gpurun_out/t11_pytest_gpu.txt; tail -6 gpurun_out/t11_pytest_gpu.txt
note "2. smoke"
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
note "3. bench, driver-style default + steps 20"
timeout 500 python bench.py > gpurun_out/t11_bench_default.json 2> gpurun_out/t11_bench_default.err; cut -c1-200 gpurun_out/t11_bench_default.json; tail -2 gpurun_out/t11_bench_default.err
timeout 500 python bench.py --steps 20 --warmup 3 > gpurun_out/t11_bench_steps20.json 2> gpurun_out/t11_bench_steps20.err; cut -c1-200 gpurun_out/t11_bench_steps20.json; tail -2 gpurun_out/t11_bench_steps20.err
note "4. reference arm"
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 2>&1 | tail -1 | cut -c1-300
note "done"
