#!/bin/bash
mkdir -p gpurun_out
T0=$SECONDS
note() { echo "== $1 (t=$((SECONDS-T0))s)"; }
note "1. DeAOT tiled short-term kernel"
timeout 120 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "local" 2>&1 | tail -3
note "2. DeAOT engines vs goldens"
timeout 300 python -m pytest tests/test_gpu_engine.py tests/test_gpu_full_geometry.py tests/test_gpu_zz_deaot_gemm.py tests/test_gpu_zevents.py tests/test_gpu_window.py -m gpu -q -x -k "deaot or gated or events or skip" 2>&1 | tail -3
note "3. bench r50_deaotl (99 frames)"
timeout 300 python bench.py --model r50_deaotl --skip-cpu-baseline --cfg4-frames 0 > gpurun_out/t16_bench_deaotl.json 2> gpurun_out/t16_bench_deaotl.err; python -c "
import json; d=json.load(open('gpurun_out/t16_bench_deaotl.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['avg_launch_us'])"; tail -2 gpurun_out/t16_bench_deaotl.err
note "4. launch shares"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1500 --csv --log-file gpurun_out/t16_launches_deaotl.csv python bench.py --model r50_deaotl --steps 12 --warmup 3 --skip-cpu-baseline --cfg4-frames 0 --no-full-clip > gpurun_out/t16_under_ncu.log 2>&1; python scripts/launch_shares.py gpurun_out/t16_launches_deaotl.csv 2>/dev/null | head -12
note "done"
