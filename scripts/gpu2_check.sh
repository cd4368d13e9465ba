#!/bin/bash
# Two-GPU trip (gpurun --gpus 2): sharded long-term bank on NCCL (graphs on / off) and over peer memory, the bench under
# torchrun (video-level DP + the cfg4 sharded sub-record), the same cfg4 clip on one GPU of the same box.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
F='^W\|^\*\*\*\|OMP_NUM'
echo "== sharded long-term bank (split-KV over 2 ranks): NCCL exchange, then peer-memory exchange"
timeout 300 $TR --master-port 29541 scripts/test_sharded_2gpu.py 2>&1 | grep -v "$F" | tail -14 | tee gpurun_out/g2_sharded_2gpu.log
echo "== same with graphs off (eager exchange)"
AOTB_SHARD_GRAPHS=0 AOTB_TEST_P2P=0 timeout 200 $TR --master-port 29544 scripts/test_sharded_2gpu.py 2>&1 | grep -v "$F" | tail -6 | tee gpurun_out/g2_sharded_2gpu_nograph.log
echo "== bench --gpus 2 (cfg2 video-DP + cfg4 sharded sub-record, 200 frames)"
timeout 600 $TR --master-port 29542 bench.py --gpus 2 --steps 99 --warmup 3 --cfg4-frames 200 --skip-cpu-baseline 2>&1 | grep -v "$F" | tail -2 | tee gpurun_out/g2_bench_2gpu.json
echo "== bench --gpus 1 on the same box (cfg4 unsharded, 200 frames)"
timeout 600 python bench.py --gpus 1 --steps 99 --warmup 3 --cfg4-frames 200 --skip-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/g2_bench_1gpu.json
echo "== cfg4 with the peer-memory exchange"
AOTB_SHARD_XCHG=p2p timeout 400 $TR --master-port 29545 bench.py --gpus 2 --steps 20 --warmup 3 --no-full-clip --cfg4-frames 200 --skip-cpu-baseline 2>&1 | grep -v "$F" | tail -2 | tee gpurun_out/g2_bench_2gpu_p2p.json
echo "== reference arm under torchrun"
timeout 300 $TR --master-port 29543 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | grep -v "$F" | tail -1 | cut -c1-400 | tee gpurun_out/g2_bench_ref_2gpu.log
