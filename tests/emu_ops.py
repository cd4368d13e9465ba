"""TEST INFRASTRUCTURE ONLY: torch-CPU emulations of a few C-ABI entry points' *contracts* (include/aotb200.h).

There is no GPU in the build container, so the host-side orchestration of a new path (weight packing in plan.py,
buffer wiring and call order in engine.py) is checked here by swapping these emulations in for `aot_benchmark_b200.ops`
and comparing the result with the oracle.  Each emulation follows the index arithmetic of the CUDA kernel it stands
for (not the oracle's formulation), so the test also cross-checks that arithmetic against the reference semantics.
Nothing under aot_benchmark_b200/ imports this module; the product has no CPU path.
"""
import math

import torch
import torch.nn.functional as F


def _act(x, act):
    if act == 1:
        return F.relu(x)
    if act == 2:
        return F.gelu(x)
    if act == 3:
        return F.silu(x)
    if act == 4:
        return F.relu6(x)
    return x


def image_to_nhwc4(img, out, stream=None):
    out.zero_()
    out[..., :3] = img.permute(0, 2, 3, 1)
    return out


def conv2d(x, w, bias, out, res=None, KH=1, KW=1, stride=1, pad=0, dil=1, act=0, stream=None):
    Cin, Cout = x.shape[3], w.shape[1]
    wt = w.view(KH, KW, Cin, Cout).permute(3, 2, 0, 1)
    y = F.conv2d(x.permute(0, 3, 1, 2), wt, bias, stride, pad, dil).permute(0, 2, 3, 1)
    if res is not None:
        y = y + res
    out.copy_(_act(y, act))
    return out


def linear(x, wt, bias, out, res=None, act=0, stream=None):
    y = x @ wt
    if bias is not None:
        y = y + bias
    if res is not None:
        y = y + res
    out.copy_(_act(y, act))
    return out


def layernorm(x, gamma, beta, out, add=None, out2=None, stream=None):
    y = F.layer_norm(x, (x.shape[1],), gamma, beta, 1e-5)
    out.copy_(y)
    if out2 is not None:
        out2.copy_(y + add)
    return out


def window_attention(qkv, qkv_bias, rel_bias, out, H, W, heads, shift, window=7, stream=None):
    """Literal restatement of window_attn_kernel's addressing (csrc/window_attn.cu)."""
    WS, D = window, 32
    C = out.shape[1]
    assert C == heads * D and qkv.shape == (H * W, 3 * C)
    Hp, Wp = -(-H // WS) * WS, -(-W // WS) * WS
    T = WS * WS
    ty, tx = torch.arange(T) // WS, torch.arange(T) % WS
    scale = 0.17677669529663687
    for wy in range(Hp // WS):
        for wx in range(Wp // WS):
            ys, xs = wy * WS + ty, wx * WS + tx
            y, x = (ys + shift) % Hp, (xs + shift) % Wp
            valid = (y < H) & (x < W)
            src = torch.where(valid, y * W + x, torch.zeros_like(y))
            rows = torch.where(valid[:, None], qkv[src], qkv_bias[None, :].expand(T, -1))
            if shift > 0:
                ry = torch.where(ys < Hp - WS, 0, torch.where(ys < Hp - shift, 1, 2))
                rx = torch.where(xs < Wp - WS, 0, torch.where(xs < Wp - shift, 1, 2))
            else:
                ry = rx = torch.zeros(T, dtype=torch.long)
            reg = ry * 3 + rx
            mask = torch.where(reg[:, None] != reg[None, :], -100.0, 0.0)
            for h in range(heads):
                q = rows[:, h * D:(h + 1) * D] * scale
                k = rows[:, C + h * D:C + (h + 1) * D]
                v = rows[:, 2 * C + h * D:2 * C + (h + 1) * D]
                s = q @ k.t() + rel_bias[h] + mask
                o = torch.softmax(s, dim=-1) @ v
                out[src[valid], h * D:(h + 1) * D] = o[valid]
    return out


def patch_merge(x, out, H, W, stream=None):
    C = x.shape[1]
    H2, W2 = (H + 1) // 2, (W + 1) // 2
    out.zero_()
    o = out.view(H2, W2, 4, C)
    xm = x.view(H, W, C)
    for q in range(4):
        dy, dx = q & 1, q >> 1
        sub = xm[dy::2, dx::2]
        o[:sub.shape[0], :sub.shape[1], q] = sub
    return out


EMULATED = ("image_to_nhwc4", "conv2d", "linear", "layernorm", "window_attention", "patch_merge")


def install(monkeypatch, ops_module):
    g = globals()
    for name in EMULATED:
        monkeypatch.setattr(ops_module, name, g[name])
