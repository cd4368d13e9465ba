#!/bin/bash
# GPU trip 18: does polling (instead of suspend-hint sleeping) on the softmax <-> MMA mbarrier chain shorten the LT kernel?
mkdir -p gpurun_out
T0=$SECONDS
AOTB_LT_SPIN=1 timeout 25 python bench.py --skip-cpu-baseline > gpurun_out/t18_bench_tile_spin.json 2> gpurun_out/t18_bench_tile_spin.err; echo "tile+spin exit $? at $((SECONDS-T0))s"
grep -o '"value": [0-9.]*\|avg_launch_us": [0-9.]*\|"frac": [0-9.]*' gpurun_out/t18_bench_tile_spin.json | head -4
AOTB_LT_SPIN=1 AOTB_LT_VARIANT=groups timeout 25 python bench.py --skip-cpu-baseline > gpurun_out/t18_bench_groups_spin.json 2> gpurun_out/t18_bench_groups_spin.err; echo "groups+spin exit $? at $((SECONDS-T0))s"
grep -o '"value": [0-9.]*\|avg_launch_us": [0-9.]*\|"frac": [0-9.]*' gpurun_out/t18_bench_groups_spin.json | head -4
AOTB_LT_SPIN=1 AOTB_TEST_GROUPS=1 timeout 30 python -m pytest tests/test_gpu_tc.py -q -m gpu -k "lt_attention" > gpurun_out/t18_tests.txt 2>&1; echo "exit $? at $((SECONDS-T0))s" >> gpurun_out/t18_tests.txt
tail -3 gpurun_out/t18_tests.txt
