#!/usr/bin/env python
"""Timeline of the persistent conv-chain kernel on the R50 encoder of a 480p frame (AOTB_CHAIN_PROF=1): per layer the first /
last publish time, the time work items spent waiting for their inputs, and the busy fraction of the SMs.
    AOTB_CHAIN_PROF=1 AOTB_CONV_CHAIN=1 python scripts/chain_profile.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aot_benchmark_b200 import EngineConfig, build_engine, build_vos_model, ops  # noqa: E402

assert ops.CONV_CHAIN and os.environ.get("AOTB_CHAIN_PROF") == "1"
cfg = EngineConfig("p", "r50_aotl")
torch.manual_seed(0)
model = build_vos_model(cfg.MODEL_VOS, cfg).cuda().eval()
eng = build_engine(cfg.MODEL_ENGINE, phase="eval", aot_model=model, gpu_id=0)
img = torch.randn(1, 3, 481, 849, device="cuda")
mask = torch.zeros(1, 1, 481, 849, device="cuda")
mask[:, :, 100:200, 100:300] = 1
with torch.no_grad():
    eng.add_reference_frame(img, mask, obj_nums=[1], frame_step=0)
    e0 = eng.aot_engines[0]
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(4):
        e0._encode(img, st)
    torch.cuda.synchronize()
chain = e0._enc._chain[1]
prof = chain.profile().double()
tiles, table = ops.conv_chain_dump(chain.layers)
t0 = prof[:, 0].min()
prof = (prof - t0) / 1e3                       # us
total = prof[:, 3].max().item()
print(f"work items {len(tiles)}, layers {len(table)}, chain time {total:.1f} us")
ncta = min(148, len(tiles))
busy = 0.0
for i, t in enumerate(tiles):
    busy += (prof[i, 3] - prof[i, 1]).item()
print(f"sum of (published - inputs complete) over work items = {busy:.0f} us-SM -> {busy / (ncta * total) * 100:.1f} % of {ncta} SMs x chain time")
print("layer  M      K-chunks BN S items | first start  last publish | avg wait-for-inputs  avg inputs->acc  avg acc->publish")
for li, row in enumerate(table):
    idx = [i for i, t in enumerate(tiles) if t[0] == li]
    p = prof[idx]
    print(f"{li:3d} {row[0]:6d} {row[9]:6d} {row[1]:4d} {row[7]} {len(idx):5d} | {p[:, 0].min():9.1f} {p[:, 3].max():12.1f} | "
          f"{(p[:, 1] - p[:, 0]).mean():12.2f} {(p[:, 2] - p[:, 1]).mean():14.2f} {(p[:, 3] - p[:, 2]).mean():14.2f}")
