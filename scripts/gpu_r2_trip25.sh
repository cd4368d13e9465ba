#!/bin/bash
# Round 2 trip 25: the test order that exposed the per-video position table (trips 21, 22, 24), after the fix.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_engine.py -m gpu -q --tb=short > gpurun_out/t25_order.txt 2>&1; grep -E "passed|failed|AssertionError|differ" gpurun_out/t25_order.txt | cut -c1-600 | head
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -1
