#!/bin/bash
# Round 2, trip 2 (one B200, ~12 min): full-geometry parity vs the real reference's goldens, new bench line (full_clip, conv
# roofline, eager-GPU baseline), DeAOT on the fused kernel (incl. self-attention), ncu source-level captures of the LT kernel
# at a 20-frame bank and of two conv shapes, launch lists (cfg2, cfg3, smoke).
mkdir -p gpurun_out
T0=$SECONDS
note() { echo "== $1 (t=$((SECONDS-T0))s)"; }
note "0. 'pair' LT layout (first execution): parity cases under a short timeout, then the microbench"
timeout 120 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "layouts and pair" > gpurun_out/t2_pytest_pair.txt 2>&1; echo "exit $?" >> gpurun_out/t2_pytest_pair.txt; tail -4 gpurun_out/t2_pytest_pair.txt
timeout 120 python scripts/lt_microbench.py --variants tile,pair --frames 1,5,10,20 --json gpurun_out/t2_lt_microbench.json 2>&1 | tail -9
for sp in 3 5 8; do timeout 60 python scripts/lt_microbench.py --variants pair --frames 10,20 --splits $sp 2>&1 | tail -2; done
note "1. GPU suite"
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/t2_pytest_gpu.txt 2>&1; echo "exit $?" >> gpurun_out/t2_pytest_gpu.txt; tail -5 gpurun_out/t2_pytest_gpu.txt
note "2. bench cfg2 (driver-style: --steps 20 -> also the full_clip sub-record)"
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/t2_bench_steps20.json 2> gpurun_out/t2_bench_steps20.err; cut -c1-300 gpurun_out/t2_bench_steps20.json; tail -3 gpurun_out/t2_bench_steps20.err
note "3. bench cfg3 model (r50_deaotl, fused tcgen05 long-term + self attention)"
timeout 200 python bench.py --model r50_deaotl --skip-cpu-baseline > gpurun_out/t2_bench_deaotl.json 2> gpurun_out/t2_bench_deaotl.err; cut -c1-200 gpurun_out/t2_bench_deaotl.json; tail -3 gpurun_out/t2_bench_deaotl.err
note "4. ncu --set full: LT kernel, 20 memory frames (3 warm-up launches skipped)"
for v in tile pair ahead; do timeout 120 ncu --set full --clock-control none --import-source on -k regex:lt_attn_tc -s 3 -c 1 -o gpurun_out/t2_prof_lt_m20_$v python scripts/lt_microbench.py --variants $v --frames 20 --reps 2 > gpurun_out/t2_prof_lt_m20_$v.log 2>&1; tail -2 gpurun_out/t2_prof_lt_m20_$v.log; done
note "5. ncu --set full: conv shapes l3 1x1 256->1024 (idx 9), l1 1x1 64->256 (idx 2), l2 3x3 (idx 5)"
for i in 9 2 5; do
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 3 -c 1 -o gpurun_out/t2_prof_conv_$i python scripts/conv_one.py $i --res > gpurun_out/t2_prof_conv_$i.log 2>&1; tail -1 gpurun_out/t2_prof_conv_$i.log
done
note "6. launch lists: cfg3 bench and smoke()"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1500 --csv --log-file gpurun_out/t2_launches_deaotl.csv python bench.py --model r50_deaotl --steps 12 --warmup 3 --skip-cpu-baseline --no-full-clip > gpurun_out/t2_bench_deaotl_under_ncu.log 2>&1; python scripts/launch_shares.py gpurun_out/t2_launches_deaotl.csv 2>/dev/null | head -16
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/t2_launches_smoke.csv python __graft_entry__.py smoke > gpurun_out/t2_smoke_under_ncu.log 2>&1; tail -2 gpurun_out/t2_smoke_under_ncu.log; python scripts/launch_shares.py gpurun_out/t2_launches_smoke.csv 2>/dev/null | head -12
note "7. conv microbench (phase stamps per shape)"
timeout 120 python scripts/conv_microbench.py > gpurun_out/t2_conv_microbench.log 2>&1; tail -3 gpurun_out/t2_conv_microbench.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
note "done"
