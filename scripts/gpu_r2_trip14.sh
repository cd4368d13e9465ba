#!/bin/bash
mkdir -p gpurun_out
for k in 0 1 2 4 8 3 11 15; do
  echo "== AOTB_CHAIN_KNOCK=$k"
  AOTB_CHAIN_KNOCK=$k AOTB_CHAIN_PROF=1 AOTB_CONV_CHAIN=1 timeout 100 python scripts/chain_profile.py 2>&1 | grep -E "chain time|^  0 |^  1 |^  3 |^ 12 |^ 28 |^ 29 " | cut -c1-150
done 2>&1 | tee gpurun_out/t14_chain_knock.txt
