#!/usr/bin/env python
"""Sweep the N tile x split-K factor of the tensor-core conv on the shapes of one R50-AOTL 480p frame (graph-replayed,
L2-warm, CUDA events) -- the data behind the tile policy in aotb_conv2d_nhwc_tc.  GPU only; writes
gpurun_out/conv_sweep.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aot_benchmark_b200 import ops  # noqa: E402
from aot_benchmark_b200._lib import lib  # noqa: E402
from conv_microbench import SHAPES  # noqa: E402

REP = 20
EXTRA = [("dec 3x3 256->256 @8x", 61, 107, 256, 256, 3, 1, 1), ("dec 3x3 256->256 @4x", 121, 213, 256, 256, 3, 1, 1),
         ("l2 ds 1x1s2 256->512", 121, 213, 256, 512, 1, 2, 0), ("l3 ds 1x1s2 512->1024", 61, 107, 512, 1024, 1, 2, 0),
         ("l3 3x3s2 256->256", 61, 107, 256, 256, 3, 2, 1), ("l2 3x3s2 128->128", 121, 213, 128, 128, 3, 2, 1)]


def time_graph(fn):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(2):
            fn()
        st.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for _ in range(REP):
                fn()
        gr.replay()
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(5):
            gr.replay()
        e1.record(st)
        st.synchronize()
    return e0.elapsed_time(e1) * 1000 / (5 * REP)


def main():
    d = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    results = []
    for name, H, W, Cin, Cout, K, s, p in SHAPES + EXTRA:
        x = torch.randn(1, H, W, Cin, generator=g).to(d)
        w = (torch.randn(K * K * Cin, Cout, generator=g) / (K * K * Cin) ** 0.5).to(d)
        wh, wl = ops.split_fp16(w)
        b = torch.randn(Cout, generator=g).to(d)
        Ho, Wo = (H + 2 * p - K) // s + 1, (W + 2 * p - K) // s + 1
        out = torch.empty(1, Ho, Wo, Cout, device=d)
        nchunks = (K * K * Cin + 63) // 64
        mt = (Ho * Wo + 127) // 128
        row = {"shape": name, "M": Ho * Wo, "K": K * K * Cin, "N": Cout, "us": {}}
        fn = lambda: ops.conv2d_tc(x, wh, wl, b, out, KH=K, KW=K, stride=s, pad=p, act=1)  # noqa: E731
        lib().aotb_set_conv_tiling(0)
        row["us"]["policy"] = round(time_graph(fn), 2)
        for bi, BN in ((1, 64), (2, 128), (3, 256)):
            if Cout % BN:
                continue
            for S in (1, 2, 4, 8):
                ctas = mt * (Cout // BN) * S
                if S > nchunks or (S > 1 and ctas > 320):
                    continue
                lib().aotb_set_conv_tiling((bi << 4) | (S << 8))
                row["us"][f"bn{BN}_s{S}"] = round(time_graph(fn), 2)
        lib().aotb_set_conv_tiling(0)
        best = min(row["us"], key=row["us"].get)
        row["best"] = best
        results.append(row)
        print(json.dumps(row), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(results, open("gpurun_out/conv_sweep.json", "w"), indent=1)


if __name__ == "__main__":
    main()
