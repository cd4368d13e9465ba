#!/bin/bash
# Round 2, trip 3 (one B200): knock-out study of the 'pair' LT kernel (which resource sets the tile time), ncu of the pair
# kernel, the driver-style bench line (steps 20 + full_clip + cfg4 500 frames + baselines).
mkdir -p gpurun_out
T0=$SECONDS
note() { echo "== $1 (t=$((SECONDS-T0))s)"; }
note "1. knock-outs (1 no ex2, 2 no PV MMAs, 4 no P write-back, 8 no max exchange, 16 no S MMAs after the first three, 32 no hi/lo)"
for k in 0 1 2 4 8 16 32 3 18 19 27 63; do
  echo "knock $k: $(AOTB_LT_KNOCK=$k timeout 60 python scripts/lt_microbench.py --variants pair --frames 20 --reps 10 2>&1 | tail -2 | tr '\n' ' ' | cut -c1-330)"
done 2>&1 | tee gpurun_out/t3_knock.txt
note "2. ncu --set full: pair kernel, 20 memory frames"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:lt_attn_pair -s 3 -c 1 -o gpurun_out/t3_prof_lt_m20_pair python scripts/lt_microbench.py --variants pair --frames 20 --reps 2 > gpurun_out/t3_prof_lt_m20_pair.log 2>&1; tail -2 gpurun_out/t3_prof_lt_m20_pair.log
note "3. bench cfg2 (driver-style)"
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/t3_bench_steps20.json 2> gpurun_out/t3_bench_steps20.err; cut -c1-300 gpurun_out/t3_bench_steps20.json; tail -3 gpurun_out/t3_bench_steps20.err
note "4. pytest: tc + zevents (multi-engine streams) + new op kernels"
timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_zevents.py tests/test_gpu_zz_deaot_gemm.py tests/test_gpu_window.py -m gpu -q > gpurun_out/t3_pytest.txt 2>&1; echo "exit $?" >> gpurun_out/t3_pytest.txt; tail -4 gpurun_out/t3_pytest.txt
timeout 100 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "aggregation" >> gpurun_out/t3_pytest.txt 2>&1; tail -2 gpurun_out/t3_pytest.txt
note "done"
