#!/bin/bash
# Round 2 trip 21: bulk-copy finish of the conv kernel, A/B against the per-thread-store finish.
mkdir -p gpurun_out
T0=$SECONDS
note() { echo "== $1 (t=$((SECONDS-T0))s)"; }
note "1. per-shape A/B (bit equality, graph timing, finish stamps)"
timeout 300 python scripts/conv_bulk_ab.py 2>&1 | cut -c1-420
note "2. parity tests with the bulk finish as the engine's conv path"
AOTB_CONV_TILING=bulk timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_engine.py tests/test_gpu_full_geometry.py -m gpu -q -x 2>&1 | tail -4
note "3. bench A/B (99-frame clip)"
for t in model bulk; do
  AOTB_CONV_TILING=$t timeout 300 python bench.py --skip-cpu-baseline --cfg4-frames 0 > gpurun_out/t21_bench_$t.json 2> gpurun_out/t21_bench_$t.err
  python - "$t" <<'PY'
import json, sys
t = sys.argv[1]
r = json.loads(open(f"gpurun_out/t21_bench_{t}.json").read().strip().splitlines()[-1])
print(t, "value", r["value"], "e2e", r["e2e"]["value"], "encoder ms", r["roofline_conv"]["encoder"]["ms"], "conv ms/frame", r["roofline_conv"]["ms_per_frame"])
PY
done
note "done"
