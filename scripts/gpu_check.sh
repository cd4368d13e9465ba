#!/bin/bash
# One GPU trip: parity tests, smoke, a short bench, and the ncu launch list of the bench command.
# Usage (from the repo root, under gpurun): bash scripts/gpu_check.sh [quick]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== pytest -m gpu" 
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench (short)"
timeout 600 python bench.py --steps 30 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_short.log
if [ "$1" != "quick" ]; then
echo "== bench (full clip)"
timeout 900 python bench.py 2>&1 | tail -3 | tee gpurun_out/bench_full.log
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 12 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log
fi
