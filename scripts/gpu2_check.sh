#!/bin/bash
# Two-GPU trip (gpurun --gpus 2): bench under torchrun (video-level DP), sharded-bank check, reference arm under torchrun.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== sharded long-term bank (split-KV over 2 ranks)"
timeout 600 $TR --master-port 29541 scripts/test_sharded_2gpu.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -12 | tee gpurun_out/sharded_2gpu.log
echo "== bench --gpus 2"
timeout 900 $TR --master-port 29542 bench.py --gpus 2 --steps 99 --warmup 3 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -2 | tee gpurun_out/bench_2gpu.log
echo "== bench --gpus 1 (same box)"
timeout 900 python bench.py --gpus 1 --steps 99 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_1gpu_samebox.log
echo "== reference arm under torchrun"
timeout 600 $TR --master-port 29543 bench.py --impl reference --gpus 2 --steps 6 --warmup 1 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -1 | cut -c1-400 | tee gpurun_out/bench_ref_2gpu.log
