#!/bin/bash
# GPU trip 17 (the last ~100 s of round-1 GPU time): two-group LT softmax layout parity + bench, the multi-engine /
# new-object golden, same-box default bench, then (if time is left) one ncu --set full capture of the LT kernel.
mkdir -p gpurun_out
T0=$SECONDS
AOTB_TEST_GROUPS=1 timeout 40 python -m pytest tests/test_gpu_tc.py tests/test_gpu_zevents.py -q -m gpu -k "groups or new_objects" > gpurun_out/t17_tests.txt 2>&1; echo "exit $? at $((SECONDS-T0))s" >> gpurun_out/t17_tests.txt
tail -5 gpurun_out/t17_tests.txt
AOTB_LT_VARIANT=groups timeout 30 python bench.py --skip-cpu-baseline > gpurun_out/t17_bench_groups.json 2> gpurun_out/t17_bench_groups.err; echo "groups bench exit $? at $((SECONDS-T0))s"
grep -o '"value": [0-9.]*\|avg_launch_us": [0-9.]*\|"frac": [0-9.]*' gpurun_out/t17_bench_groups.json | head -4
timeout 30 python bench.py --skip-cpu-baseline > gpurun_out/t17_bench_tile.json 2> gpurun_out/t17_bench_tile.err; echo "tile bench exit $? at $((SECONDS-T0))s"
grep -o '"value": [0-9.]*\|avg_launch_us": [0-9.]*\|"frac": [0-9.]*' gpurun_out/t17_bench_tile.json | head -4
AOTB_LT_VARIANT=groups AOTB_GRAPHS=0 timeout 45 ncu --set full --clock-control none --import-source on -k regex:lt_attn_tc_kernel -s 100 -c 1 -o gpurun_out/t17_prof_lt_groups python bench.py --steps 40 --warmup 3 --skip-cpu-baseline > gpurun_out/t17_prof.log 2>&1; echo "ncu exit $? at $((SECONDS-T0))s"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
