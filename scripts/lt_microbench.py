"""Long-term attention kernel microbenchmark (B200): time aotb_lt_attn_tc_f16x2 alone at fixed bank sizes.

    python scripts/lt_microbench.py [--variants tile,groups,ahead] [--frames 1,5,10,20] [--n 1674] [--json out.json]

For every (variant, memory frames m) it packs random Q / K / V (cfg2 shape: N = 1674 queries, Tk = N*m keys, 8 heads x 32),
checks the output against the default layout, and times `reps` back-to-back launches with CUDA events on the launching
stream after a warm-up (the 126 MB L2 holds the packed bank of small m; m >= 10 streams from HBM).  Reports us per launch,
algorithmic TFLOP/s (4*N*Tk*C) and cycles per 128x128 score tile per SM -- the number the layouts are designed against
(~1024 = MUFU / TMEM-read floor, 896 = tensor floor of the exact mode).  For an ncu capture of a LONG launch:

    ncu --set full --clock-control none --import-source on -k regex:lt_attn_tc -s 6 -c 1 -o gpurun_out/lt_m20 \\
        python scripts/lt_microbench.py --variants tile --frames 20 --reps 2
"""
import argparse
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aot_benchmark_b200 import ops  # noqa: E402
from aot_benchmark_b200.engine import lt_splits  # noqa: E402

H, D = 8, 32


def pack(x, cap, div=1.0):
    dst = torch.zeros(H, cap, 64, dtype=torch.float16, device=x.device)
    ops.tc_pack_rows(x, dst, 0, div)
    return dst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="tile,groups")
    ap.add_argument("--frames", default="1,5,10,20")
    ap.add_argument("--n", type=int, default=1674)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--splits", type=int, default=0, help="0 = engine policy (lt_splits)")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    N = a.n
    g = torch.Generator().manual_seed(0)
    Q = (torch.randn(N, H * D, generator=g) * 3).to(dev)
    ncap = ((N + 255) // 256) * 256
    Qp = pack(Q, ncap, math.sqrt(D))
    try:
        sm_clock = torch.cuda.clock_rate() * 1e6            # MHz (NVML) -> Hz, sampled once; B200 boost is 1965 MHz
    except Exception:
        sm_clock = 1.965e9
    rows = []
    for m in [int(x) for x in a.frames.split(",")]:
        Tk = N * m
        K = torch.randn(Tk, H * D, generator=g).to(dev)
        V = torch.randn(Tk, H * D, generator=g).to(dev)
        kcap = ((Tk + 127) // 128) * 128 + 128
        Kp, Vp = pack(K, kcap), pack(V, kcap)
        ref = None
        for v in a.variants.split(","):
            splits = a.splits or lt_splits(N, H, Tk, variant=v)
            part = None
            if splits > 1:
                part = (torch.empty(splits, N, H * D, device=dev), torch.empty(splits, H, N, device=dev),
                        torch.empty(splits, H, N, device=dev))
            O = torch.empty(N, H * D, device=dev)
            run = lambda: ops.lt_attention_tc(Qp, Kp, Vp, N, Tk, O=O, splits=splits, exact=True, part=part, variant=v,
                                              merge=False)
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps):
                run()
            e1.record()
            torch.cuda.synchronize()
            us = 1e3 * e0.elapsed_time(e1) / a.reps
            ops.lt_attention_tc(Qp, Kp, Vp, N, Tk, O=O, splits=splits, exact=True, part=part, variant=v)   # merged
            torch.cuda.synchronize()
            if ref is None:
                ref = O.clone()
            err = (O - ref).abs().max().item()
            tiles = ((N + 127) // 128) * H * ((Tk + 127) // 128)      # 128 x 128 score tiles of the launch
            clk_per_tile = us * 1e-6 * sm_clock / (tiles / 148)
            row = {"variant": v, "frames": m, "Tk": Tk, "splits": splits, "us": round(us, 2),
                   "tflops": round(4.0 * N * Tk * H * D / us / 1e6, 1), "clk_per_tile_per_sm": round(clk_per_tile),
                   "max_abs_diff_vs_first": err}
            rows.append(row)
            print(row, flush=True)
    if a.json:
        json.dump(rows, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
