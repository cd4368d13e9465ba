#!/bin/bash
# One GPU trip: parity tests, smoke, bench, and the ncu launch list of the bench command.
# Usage (from the repo root, under gpurun): bash scripts/gpu_check.sh [quick|tc|impls]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
if [ "$1" == "tc" ]; then
  echo "== tensor-core kernel tests first"
  timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -s 2>&1 | tail -40 | tee gpurun_out/pytest_tc.log
fi
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; tail -25 gpurun_out/pytest_gpu.log; grep -a -E "dlogit|sharp" gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
if [ "$1" == "quick" ]; then
  timeout 600 python bench.py --steps 30 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_short.log
  exit 0
fi
echo "== bench full clip (defaults)"
timeout 900 python bench.py 2>&1 | tail -2 | tee gpurun_out/bench_full.log
if [ "$1" == "impls" ]; then
  for impl in simt tc_fast; do
    echo "== bench full clip AOTB_LT_IMPL=$impl"
    AOTB_LT_IMPL=$impl timeout 900 python bench.py 2>&1 | tail -2 | tee gpurun_out/bench_full_$impl.log
  done
  echo "== bench full clip AOTB_CONV_IMPL=simt"
  AOTB_CONV_IMPL=simt timeout 900 python bench.py 2>&1 | tail -2 | tee gpurun_out/bench_full_convsimt.log
fi
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 1200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 12 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-200
if [ "$2" == "prof" ] || [ "$1" == "prof" ]; then
  echo "== ncu --set full: long-term attention kernel (late frame, large bank) and conv kernel"
  AOTB_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:lt_attn_tc_kernel -s 345 -c 1 -o gpurun_out/prof_lt python bench.py --steps 99 --warmup 3 > gpurun_out/prof_lt.log 2>&1
  AOTB_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2000 -c 4 -o gpurun_out/prof_conv python bench.py --steps 20 --warmup 3 > gpurun_out/prof_conv.log 2>&1
  ls -la gpurun_out/*.ncu-rep
fi
if [ "$2" == "pdl" ] || [ "$3" == "pdl" ]; then
  echo "== programmatic dependent launch: parity tests + bench with AOTB_PDL=1"
  AOTB_PDL=1 timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_tc.py -m gpu -q > gpurun_out/pytest_pdl.log 2>&1; tail -3 gpurun_out/pytest_pdl.log
  AOTB_PDL=1 timeout 900 python bench.py 2>&1 | tail -2 | tee gpurun_out/bench_full_pdl.log
fi
