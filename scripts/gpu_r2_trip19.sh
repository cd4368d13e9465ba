#!/bin/bash
mkdir -p gpurun_out
T0=$SECONDS
note() { echo "== $1 (t=$((SECONDS-T0))s)"; }
note "1. groupnorm + engines"
timeout 500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_engine.py tests/test_gpu_window.py -m gpu -q -x -k "groupnorm or golden or graph or swin" 2>&1 | tail -3
note "2. bench cfg2 + launch shares"
timeout 300 python bench.py --skip-cpu-baseline --cfg4-frames 0 > gpurun_out/t19_bench.json 2> gpurun_out/t19_bench.err; python -c "
import json; d=json.load(open('gpurun_out/t19_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'])"; tail -2 gpurun_out/t19_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 1500 --csv --log-file gpurun_out/t19_launches.csv python bench.py --steps 12 --warmup 3 --skip-cpu-baseline --cfg4-frames 0 --no-full-clip > gpurun_out/t19_under_ncu.log 2>&1; python scripts/launch_shares.py gpurun_out/t19_launches.csv 2>/dev/null | head -14
note "done"
