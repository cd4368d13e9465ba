#!/bin/bash
mkdir -p gpurun_out
T0=$SECONDS
note() { echo "== $1 (t=$((SECONDS-T0))s)"; }
note "1. local attention tests + engines"
timeout 500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_engine.py tests/test_gpu_full_geometry.py -m gpu -q -x -k "local or golden or oracle" 2>&1 | tail -3
note "2. bench cfg2 + launch shares"
timeout 300 python bench.py --skip-cpu-baseline --cfg4-frames 0 > gpurun_out/t18_bench.json 2> gpurun_out/t18_bench.err; python -c "
import json; d=json.load(open('gpurun_out/t18_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'])"; tail -2 gpurun_out/t18_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 1500 --csv --log-file gpurun_out/t18_launches.csv python bench.py --steps 12 --warmup 3 --skip-cpu-baseline --cfg4-frames 0 --no-full-clip > gpurun_out/t18_under_ncu.log 2>&1; python scripts/launch_shares.py gpurun_out/t18_launches.csv 2>/dev/null | head -8
note "3. bench cfg3 model"
timeout 300 python bench.py --model r50_deaotl --skip-cpu-baseline --cfg4-frames 0 > gpurun_out/t18_bench_deaotl.json 2> gpurun_out/t18_bench_deaotl.err; python -c "
import json; d=json.load(open('gpurun_out/t18_bench_deaotl.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'])"; tail -2 gpurun_out/t18_bench_deaotl.err
note "done"
