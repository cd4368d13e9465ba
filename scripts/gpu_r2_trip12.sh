#!/bin/bash
mkdir -p gpurun_out
T0=$SECONDS
note() { echo "== $1 (t=$((SECONDS-T0))s)"; }
note "1. tests touched by the LT diet / GN stats / merge changes"
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py tests/test_gpu_engine.py tests/test_gpu_full_geometry.py tests/test_gpu_zz_deaot_gemm.py -m gpu -q -x 2>&1 | tail -4
note "2. LT microbench (tile with the packed-fp32 softmax)"
timeout 120 python scripts/lt_microbench.py --variants tile,pair --frames 1,5,10,20 --json gpurun_out/t12_lt_microbench.json 2>&1 | tail -8
note "3. bench (99 frames)"
timeout 300 python bench.py --skip-cpu-baseline --cfg4-frames 0 > gpurun_out/t12_bench.json 2> gpurun_out/t12_bench.err; python -c "
import json; d=json.load(open('gpurun_out/t12_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_us'], d['gpu_launches'])"; tail -2 gpurun_out/t12_bench.err
note "4. launch shares"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 1500 --csv --log-file gpurun_out/t12_launches.csv python bench.py --steps 12 --warmup 3 --skip-cpu-baseline --cfg4-frames 0 --no-full-clip > gpurun_out/t12_under_ncu.log 2>&1; python scripts/launch_shares.py gpurun_out/t12_launches.csv 2>/dev/null | head -16
note "done"
