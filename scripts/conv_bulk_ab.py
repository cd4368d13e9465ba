#!/usr/bin/env python
"""A/B of the conv finish: per-thread float4 stores (tiling mask 0) against the bulk-copy finish (mask 8).

For every shape of scripts/conv_microbench.SHAPES, with and without a residual: bit-equality of the two outputs, a CUDA graph
of REP back-to-back launches timed with CUDA events, and the median per-CTA 'finish start' -> 'finish stored' stamps.
GPU only; writes gpurun_out/conv_bulk_ab.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from aot_benchmark_b200 import ops  # noqa: E402
from aot_benchmark_b200._lib import lib  # noqa: E402
from conv_microbench import SHAPES, NAMES  # noqa: E402

REP = 20


def timed(fn, st):
    with torch.cuda.stream(st):
        for _ in range(3):
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
    return round(e0.elapsed_time(e1) * 1000 / (5 * REP), 2)


def main():
    d = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    ws = ops._tc_workspace(d)
    st = torch.cuda.Stream()
    results = []
    for name, H, W, Cin, Cout, K, s, p in SHAPES:
        x = torch.randn(1, H, W, Cin, generator=g).to(d)
        w = (torch.randn(K * K * Cin, Cout, generator=g) / (K * K * Cin) ** 0.5).to(d)
        wh, wl = ops.split_fp16(w)
        b = torch.randn(Cout, generator=g).to(d)
        Ho, Wo = (H + 2 * p - K) // s + 1, (W + 2 * p - K) // s + 1
        res = torch.randn(1, Ho, Wo, Cout, generator=g).to(d)
        row = {"shape": name, "M": Ho * Wo, "K": K * K * Cin, "N": Cout}
        for tag, r in (("", None), ("_res", res)):
            outs = {}
            for mode in (0, 8):
                lib().aotb_set_conv_tiling(mode)
                out = torch.full((1, Ho, Wo, Cout), float("nan"), device=d)
                fn = lambda: ops.conv2d_tc(x, wh, wl, b, out, res=r, KH=K, KW=K, stride=s, pad=p, act=1)  # noqa: E731
                row[f"us_mode{mode}{tag}"] = timed(fn, st)
                outs[mode] = out.clone()
                lib().aotb_set_conv_tiling(4 | mode)
                ws.zero_()
                fn()
                torch.cuda.synchronize()
                st8 = ws.view(torch.int64)[: 12 * 4096].view(-1, 12).cpu()
                st8 = st8[st8[:, 7] != 0][:, :10]
                rel = (st8[:, 1:] - st8[:, :1]).double() / 1965.0
                rel = torch.where(st8[:, 1:] != 0, rel, torch.full_like(rel, float("nan")))
                med = dict(zip(NAMES, torch.nanmedian(rel, dim=0).values.tolist()))
                row[f"finish_us_mode{mode}{tag}"] = round(med["finish stored"] - med["finish start"], 2)
            row[f"bit_equal{tag}"] = bool(torch.equal(outs[0], outs[8]))
        lib().aotb_set_conv_tiling(0)
        results.append(row)
        print(json.dumps(row), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(results, open("gpurun_out/conv_bulk_ab.json", "w"), indent=1)


if __name__ == "__main__":
    main()
