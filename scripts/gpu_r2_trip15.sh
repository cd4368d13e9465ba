#!/bin/bash
mkdir -p gpurun_out
for cfg in "8 4 120" "4 8 300" "2 8 600" "3 12 2000" "1 16 4000"; do
  set -- $cfg
  echo "== MINCHUNKS=$1 MAXSPLIT=$2 ITEMS=$3"
  AOTB_CHAIN_MINCHUNKS=$1 AOTB_CHAIN_MAXSPLIT=$2 AOTB_CHAIN_ITEMS=$3 AOTB_CHAIN_PROF=1 AOTB_CONV_CHAIN=1 timeout 100 python scripts/chain_profile.py 2>&1 | grep -E "chain time|Error|error" | cut -c1-150
done 2>&1 | tee gpurun_out/t15_chain_split_sweep.txt
