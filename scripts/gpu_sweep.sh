#!/bin/bash
mkdir -p gpurun_out
cd scripts && timeout 900 python conv_sweep.py > ../gpurun_out/conv_sweep.log 2>&1; cd ..; tail -3 gpurun_out/conv_sweep.log | cut -c1-400
mv scripts/gpurun_out/conv_sweep.json gpurun_out/ 2>/dev/null
for v in 0 1; do
  echo "== bench AOTB_PDL=$v"
  AOTB_PDL=$v timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_pdl_$v.log | cut -c1-200
done
