#!/bin/bash
# Round 2, trip 5 (1 GPU): first execution of the persistent conv-chain kernel (short timeouts: a protocol bug traps or hangs).
mkdir -p gpurun_out
T0=$SECONDS
note() { echo "== $1 (t=$((SECONDS-T0))s)"; }
note "1. conv chain vs per-layer kernels (small)"
timeout 90 python -m pytest tests/test_gpu_conv_chain.py -m gpu -q -x -k "per_layer and 31" 2>&1 | tail -6
note "2. conv chain, layer-1 size"
timeout 90 python -m pytest tests/test_gpu_conv_chain.py -m gpu -q -x -k "per_layer and 121" 2>&1 | tail -6
note "3. engine with the chain vs goldens"
timeout 200 python -m pytest tests/test_gpu_conv_chain.py -m gpu -q -x -k "engine" 2>&1 | tail -6
note "4. bench: chain off / on (99 frames, no extras)"
timeout 200 python bench.py --skip-cpu-baseline --cfg4-frames 0 > gpurun_out/t5_bench_nochain.json 2> gpurun_out/t5_bench_nochain.err; python -c "
import json; d=json.load(open('gpurun_out/t5_bench_nochain.json')); print('chain off', d['value'], d['e2e']['value'], d['roofline_conv']['encoder'])"
AOTB_CONV_CHAIN=1 timeout 200 python bench.py --skip-cpu-baseline --cfg4-frames 0 > gpurun_out/t5_bench_chain.json 2> gpurun_out/t5_bench_chain.err; python -c "
import json; d=json.load(open('gpurun_out/t5_bench_chain.json')); print('chain on ', d['value'], d['e2e']['value'], d['roofline_conv']['encoder'])"; tail -3 gpurun_out/t5_bench_chain.err
note "done"
