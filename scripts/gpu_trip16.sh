#!/bin/bash
# GPU trip 16 (last ~10 GPU-minutes of round 1): highest-value checks first, every step under its own timeout and
# writing into gpurun_out/ as it goes.
#   1. Swin-B encoder path: window-attention / patch-merge kernels, encoder vs oracle, SwinB-AOTL/DeAOTL goldens
#   2. software-pipelined LT softmax (AOTB_LT_PIPE): bit-identical to the serial variant
#   3. bench (cfg2) with the pipelined variant     4. short SwinB-AOTL bench at 592x1040
mkdir -p gpurun_out
T0=$SECONDS
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/t16_smi.txt 2>&1
echo "== 1. Swin tests"
timeout 240 python -m pytest tests/test_gpu_window.py -q -m gpu > gpurun_out/t16_window.txt 2>&1; echo "exit $? at $((SECONDS-T0))s" >> gpurun_out/t16_window.txt
tail -4 gpurun_out/t16_window.txt
echo "== 2. pipelined LT softmax parity"
AOTB_TEST_PIPE=1 timeout 120 python -m pytest tests/test_gpu_tc.py -q -m gpu -k pipelined > gpurun_out/t16_pipe.txt 2>&1; echo "exit $? at $((SECONDS-T0))s" >> gpurun_out/t16_pipe.txt
tail -4 gpurun_out/t16_pipe.txt
echo "== 3. bench cfg2, AOTB_LT_PIPE=1"
AOTB_LT_PIPE=1 timeout 170 python bench.py --skip-cpu-baseline > gpurun_out/t16_bench_pipe.json 2> gpurun_out/t16_bench_pipe.err; echo "exit $? at $((SECONDS-T0))s"
cut -c1-400 gpurun_out/t16_bench_pipe.json; grep -o '"roofline.*avg_launch_us[^,]*' gpurun_out/t16_bench_pipe.json | cut -c1-300
echo "== 4. bench swinb_aotl (592x1040), 30 frames"
timeout 170 python bench.py --model swinb_aotl --steps 30 --skip-cpu-baseline > gpurun_out/t16_bench_swin.json 2> gpurun_out/t16_bench_swin.err; echo "exit $? at $((SECONDS-T0))s"
cut -c1-300 gpurun_out/t16_bench_swin.json; tail -3 gpurun_out/t16_bench_swin.err
echo "== 5. engine tests with the pipelined variant (if time is left)"
AOTB_LT_PIPE=1 timeout 200 python -m pytest tests/test_gpu_engine.py -q -m gpu -k "golden or growth" > gpurun_out/t16_engine_pipe.txt 2>&1; echo "exit $? at $((SECONDS-T0))s" >> gpurun_out/t16_engine_pipe.txt
tail -3 gpurun_out/t16_engine_pipe.txt
