#!/bin/bash
# ncu --set full captures (one GPU, eager launches): self-attention + long-term attention + local attention of one
# late frame, and four conv launches.  Reports come back in gpurun_out/ and are summarised under profiles/.
mkdir -p gpurun_out
AOTB_GRAPHS=0 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'lt_attn_tc_kernel|local_attn_tile' -s 540 -c 3 -f -o gpurun_out/prof_attn python bench.py --steps 99 --warmup 3 > gpurun_out/prof_attn.log 2>&1
tail -2 gpurun_out/prof_attn.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
