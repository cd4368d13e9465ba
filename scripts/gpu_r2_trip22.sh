#!/bin/bash
# Round 2 trip 22: one mbarrier arrival per WARP (was per thread) in the attention and conv kernels.
mkdir -p gpurun_out
T0=$SECONDS
note() { echo "== $1 (t=$((SECONDS-T0))s)"; }
note "1. LT microbench, all layouts (before: tile 241.9 us / 2398 clk per tile at m=20, pair 2381)"
timeout 300 python scripts/lt_microbench.py --variants tile,groups,ahead,pair --frames 1,5,20 --json gpurun_out/t22_lt_microbench.json 2>&1 | tail -16
note "2. conv microbench (before: profiles/r02_trip17_conv_microbench.json)"
timeout 200 python scripts/conv_microbench.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: continue
    print(r['shape'], r['us_mode0'], r['phases_us_mode0'])
"
note "3. parity: tensor-core kernels, engines, full geometry"
timeout 700 python -m pytest tests/test_gpu_tc.py tests/test_gpu_zz_deaot_gemm.py tests/test_gpu_engine.py tests/test_gpu_full_geometry.py tests/test_gpu_conv_chain.py -m gpu -q -x 2>&1 | tail -4
note "4. bench"
timeout 300 python bench.py --skip-cpu-baseline --cfg4-frames 200 > gpurun_out/t22_bench.json 2> gpurun_out/t22_bench.err; python -c "
import json; d=json.load(open('gpurun_out/t22_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_us'], d['roofline_conv']['encoder']['ms'], d['cfg4']['value'])"; tail -2 gpurun_out/t22_bench.err
note "done"
