#!/bin/bash
mkdir -p gpurun_out
T0=$SECONDS
note() { echo "== $1 (t=$((SECONDS-T0))s)"; }
note "0. pair (4 softmax warps, 1 thread / row) parity"
timeout 120 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "layouts and pair" 2>&1 | tail -3
note "1. tile vs pair microbench"
timeout 120 python scripts/lt_microbench.py --variants tile,pair --frames 1,5,10,20 --json gpurun_out/t3b_lt_microbench.json 2>&1 | tail -8
for sp in 3 4 6 8; do timeout 60 python scripts/lt_microbench.py --variants pair --frames 10,20 --splits $sp 2>&1 | tail -2; done
note "2. knock-outs (1 no ex2, 2 no PV MMAs, 4 no P write-back, 16 no S MMAs after the first three, 32 no hi/lo)"
for k in 0 1 2 4 16 32 3 18 19 23 55; do
  echo "knock $k: $(AOTB_LT_KNOCK=$k timeout 60 python scripts/lt_microbench.py --variants pair --frames 20 --reps 10 2>&1 | tail -2 | tr '\n' ' ' | cut -c1-330)"
done 2>&1 | tee gpurun_out/t3b_knock.txt
note "3. ncu pair"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:lt_attn_pair -s 3 -c 1 -o gpurun_out/t3b_prof_lt_m20_pair python scripts/lt_microbench.py --variants pair --frames 20 --reps 2 > gpurun_out/t3b_prof_lt_m20_pair.log 2>&1; tail -2 gpurun_out/t3b_prof_lt_m20_pair.log
note "4. bench cfg2 (driver-style)"
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/t3_bench_steps20.json 2> gpurun_out/t3_bench_steps20.err; cut -c1-300 gpurun_out/t3_bench_steps20.json; tail -3 gpurun_out/t3_bench_steps20.err
note "done"
