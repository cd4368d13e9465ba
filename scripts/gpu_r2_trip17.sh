#!/bin/bash
mkdir -p gpurun_out
T0=$SECONDS
note() { echo "== $1 (t=$((SECONDS-T0))s)"; }
note "1. conv tests + engines (finish with prefetched bias / residual)"
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py tests/test_gpu_engine.py tests/test_gpu_conv_chain.py tests/test_gpu_window.py -m gpu -q -x -k "conv or linear or engine or golden or swin or chain" 2>&1 | tail -3
note "2. conv microbench"
timeout 120 python scripts/conv_microbench.py > gpurun_out/t17_conv_microbench.log 2>&1; python - <<'PY'
import json
for r in json.load(open('gpurun_out/conv_microbench.json')):
    print(f'{r["shape"]:24s} us={r["us_mode0"]:6.2f}', {k: v for k, v in r["phases_us_mode0"].items() if k in ("staged", "finish start", "finish stored", "exit")})
PY
note "3. bench (99 frames)"
timeout 300 python bench.py --skip-cpu-baseline --cfg4-frames 0 > gpurun_out/t17_bench.json 2> gpurun_out/t17_bench.err; python -c "
import json; d=json.load(open('gpurun_out/t17_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline_conv']['encoder'])"; tail -2 gpurun_out/t17_bench.err
note "done"
