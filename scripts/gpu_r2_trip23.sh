#!/bin/bash
# Round 2 trip 23: test_cuda_graph_replay_matches_eager failed when tests/test_gpu_tc.py ran before it (trips 21, 22) -- find out why.
mkdir -p gpurun_out
T0=$SECONDS
note() { echo "== $1 (t=$((SECONDS-T0))s)"; }
note "1. the test alone"
timeout 200 python -m pytest tests/test_gpu_engine.py -m gpu -q -x -k graph_replay --tb=short 2>&1 | tail -15
note "2. poisoned allocator: NaN / big / zero"
for k in nan big zero; do timeout 200 python scripts/poison_check.py $k 2>&1 | tail -12; done
note "3. the failing order, traceback kept"
timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_engine.py -m gpu -q -x -k "lt_attention or conv2d or linear_tc or graph_replay" --tb=short > gpurun_out/t23_order.txt 2>&1; tail -30 gpurun_out/t23_order.txt
note "done"
