#!/usr/bin/env python
"""One tensor-core conv shape, launched a few times (for `ncu --set full -k regex:conv_tc -s 3 -c 1`).
    python scripts/conv_one.py <index into scripts/conv_microbench.SHAPES> [--res]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from aot_benchmark_b200 import ops  # noqa: E402
from conv_microbench import SHAPES  # noqa: E402

name, H, W, Cin, Cout, K, s, p = SHAPES[int(sys.argv[1])]
d = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x = torch.randn(1, H, W, Cin, generator=g).to(d)
w = (torch.randn(K * K * Cin, Cout, generator=g) / (K * K * Cin) ** 0.5).to(d)
wh, wl = ops.split_fp16(w)
b = torch.randn(Cout, generator=g).to(d)
Ho, Wo = (H + 2 * p - K) // s + 1, (W + 2 * p - K) // s + 1
out = torch.empty(1, Ho, Wo, Cout, device=d)
res = torch.randn(1, Ho, Wo, Cout, generator=g).to(d) if "--res" in sys.argv else None
for _ in range(6):
    ops.conv2d_tc(x, wh, wl, b, out, res=res, KH=K, KW=K, stride=s, pad=p, act=1)
torch.cuda.synchronize()
print(name, "ok")
