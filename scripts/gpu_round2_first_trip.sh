#!/bin/bash
# First GPU trip of round 2 (one B200, ~12 min): validates everything that was built after round 1's GPU budget ran out and
# measures the candidates.  Every step has its own timeout and writes into gpurun_out/ as it goes.
#   gpurun --timeout 900 -- 'bash scripts/gpu_round2_first_trip.sh'
mkdir -p gpurun_out
T0=$SECONDS
note() { echo "== $1 (t=$((SECONDS-T0))s)"; }
note "1. whole GPU suite (includes the DeAOT GEMM-path kernels, events / skip goldens)"
timeout 420 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.txt 2>&1; echo "exit $?" >> gpurun_out/r2_pytest_gpu.txt; tail -4 gpurun_out/r2_pytest_gpu.txt
note "2. 'ahead' LT layout parity (three score buffers; never run on a GPU before) -- short timeout: a protocol bug would hang"
AOTB_TEST_VARIANTS=ahead timeout 60 python -m pytest tests/test_gpu_tc.py -m gpu -q -k "layouts and ahead" > gpurun_out/r2_pytest_ahead.txt 2>&1; echo "exit $?" >> gpurun_out/r2_pytest_ahead.txt; tail -4 gpurun_out/r2_pytest_ahead.txt
note "2b. fused DeAOT long-term attention kernel (gp_attn_tc.cu; never run on a GPU before) -- short timeout"
AOTB_TEST_VARIANTS=gp_tc timeout 60 python -m pytest tests/test_gpu_zz_deaot_gemm.py -m gpu -q -k "fused" > gpurun_out/r2_pytest_gp_tc.txt 2>&1; echo "exit $?" >> gpurun_out/r2_pytest_gp_tc.txt; tail -4 gpurun_out/r2_pytest_gp_tc.txt
note "3. LT microbench: tile vs ahead at 1 / 5 / 10 / 20 memory frames"
timeout 60 python scripts/lt_microbench.py --variants tile,ahead --frames 1,5,10,20 --json gpurun_out/r2_lt_microbench.json 2>&1 | tail -9
note "4. bench cfg2: default, then AOTB_LT_VARIANT=ahead"
timeout 120 python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err; cut -c1-220 gpurun_out/r2_bench_full.json
AOTB_LT_VARIANT=ahead timeout 60 python bench.py --skip-cpu-baseline > gpurun_out/r2_bench_full_ahead.json 2> gpurun_out/r2_bench_full_ahead.err; grep -o '"value": [0-9.]*\|avg_launch_us": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2_bench_full_ahead.json | head -4
note "5. bench cfg3 model (r50_deaotl): SIMT long-term attention vs GEMM path"
timeout 120 python bench.py --model r50_deaotl --skip-cpu-baseline > gpurun_out/r2_bench_deaotl_simt.json 2> gpurun_out/r2_bench_deaotl_simt.err; grep -o '"value": [0-9.]*\|avg_launch_us": [0-9.]*' gpurun_out/r2_bench_deaotl_simt.json | head -3
AOTB_DEAOT_LT=gemm timeout 120 python bench.py --model r50_deaotl --skip-cpu-baseline > gpurun_out/r2_bench_deaotl_gemm.json 2> gpurun_out/r2_bench_deaotl_gemm.err; grep -o '"value": [0-9.]*\|avg_launch_us": [0-9.]*' gpurun_out/r2_bench_deaotl_gemm.json | head -3; tail -2 gpurun_out/r2_bench_deaotl_gemm.err
AOTB_DEAOT_LT=tc timeout 120 python bench.py --model r50_deaotl --skip-cpu-baseline > gpurun_out/r2_bench_deaotl_tc.json 2> gpurun_out/r2_bench_deaotl_tc.err; grep -o '"value": [0-9.]*\|avg_launch_us": [0-9.]*' gpurun_out/r2_bench_deaotl_tc.json | head -3; tail -2 gpurun_out/r2_bench_deaotl_tc.err
note "6. ncu --set full of a LONG LT launch (20 memory frames), default layout"
timeout 90 ncu --set full --clock-control none --import-source on -k regex:lt_attn_tc -s 6 -c 1 -o gpurun_out/r2_prof_lt_m20 python scripts/lt_microbench.py --variants tile --frames 20 --reps 2 > gpurun_out/r2_prof_lt_m20.log 2>&1; ls -la gpurun_out/*.ncu-rep 2>/dev/null
note "7. ncu launch list of the default bench (kernel shares of the step)"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 1500 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 12 --warmup 3 --skip-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1; python scripts/launch_shares.py gpurun_out/r2_launches.csv 2>/dev/null | head -14
note "done"
