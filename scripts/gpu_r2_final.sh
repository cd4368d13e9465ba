#!/bin/bash
# Round 2, final verification of the committed tree (one B200): whole GPU suite, smoke(), the driver's bench commands.
mkdir -p gpurun_out
T0=$SECONDS
note() { echo "== $1 (t=$((SECONDS-T0))s)"; }
note "1. whole GPU suite"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_pytest_gpu.txt 2>&1; echo "exit $?" >> gpurun_out/final_pytest_gpu.txt; tail -5 gpurun_out/final_pytest_gpu.txt
note "2. smoke"
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
note "3. bench default (99 steps) and the driver's --steps 20"
timeout 500 python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; cut -c1-160 gpurun_out/final_bench_default.json; tail -2 gpurun_out/final_bench_default.err
timeout 500 python bench.py --steps 20 --warmup 3 > gpurun_out/final_bench_steps20.json 2> gpurun_out/final_bench_steps20.err; cut -c1-160 gpurun_out/final_bench_steps20.json; tail -2 gpurun_out/final_bench_steps20.err
note "4. launch shares of the default bench"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 1500 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 12 --warmup 3 --skip-cpu-baseline --cfg4-frames 0 --no-full-clip > gpurun_out/final_under_ncu.log 2>&1; python scripts/launch_shares.py gpurun_out/final_launches.csv 2>/dev/null | head -14
note "done"
