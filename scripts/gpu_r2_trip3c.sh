#!/bin/bash
mkdir -p gpurun_out
echo "== knock-outs of the default 'tile' LT kernel at a 20-frame bank (1 no ex2, 2 no PV MMAs, 4 no P write-back, 8 no TMEM score read, 16 no S MMAs)"
for k in 0 1 2 4 8 16 18 9 3 10 26 27 31; do
  echo "knock $k: $(AOTB_LT_KNOCK=$k timeout 60 python scripts/lt_microbench.py --variants tile --frames 20 --reps 10 2>&1 | tail -1 | cut -c1-200)"
done 2>&1 | tee gpurun_out/t3c_knock_tile.txt
echo "== pair (8 softmax warps) parity"
timeout 120 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "layouts and pair" 2>&1 | tail -2
echo "== new op tests"
timeout 120 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "preprocess or mask_writer or aggregation" 2>&1 | tail -2
