#!/bin/bash
mkdir -p gpurun_out
echo "== chain tests"; timeout 200 python -m pytest tests/test_gpu_conv_chain.py -m gpu -q -x 2>&1 | tail -3
echo "== timeline"; AOTB_CHAIN_PROF=1 AOTB_CONV_CHAIN=1 timeout 120 python scripts/chain_profile.py 2>&1 | tail -50 | tee gpurun_out/t8_chain_profile.txt
echo "== bench chain on"; AOTB_CONV_CHAIN=1 timeout 200 python bench.py --skip-cpu-baseline --cfg4-frames 0 > gpurun_out/t8_bench_chain.json 2> gpurun_out/t8_bench_chain.err; python -c "
import json; d=json.load(open('gpurun_out/t8_bench_chain.json')); print('chain on ', d['value'], d['e2e']['value'], d['roofline_conv']['encoder'])"; tail -3 gpurun_out/t8_bench_chain.err
