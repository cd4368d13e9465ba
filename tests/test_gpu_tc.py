"""GPU: the tcgen05 long-term attention kernel (lt_attn_tc.cu) vs the fp64 oracle and vs the fp32 SIMT
kernel, in exact (fp16x2 split) and fast modes, with KV splits and a device-resident key count."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

H, D = 8, 32


def _pack(x, cap, div=1.0):
    from aot_benchmark_b200 import ops
    dst = torch.zeros(H, cap, 64, dtype=torch.float16, device=x.device)
    ops.tc_pack_rows(x, dst, 0, div)
    return dst


def _ref(Q, K, V):
    from oracle import aot_oracle as O
    return O.multihead_attention(Q.double().cpu().unsqueeze(1), K.double().cpu().unsqueeze(1),
                                 V.double().cpu().unsqueeze(1), H)[:, 0]


def test_pack_rows_layout():
    d = torch.device("cuda:0")
    x = torch.randn(50, 256, device=d) * 3
    p = _pack(x, 64, div=math.sqrt(32.0))
    # true IEEE division like the CPU reference (torch's CUDA kernel multiplies by a reciprocal instead)
    xs = (x.cpu() / math.sqrt(32.0)).to(d).view(50, H, D).permute(1, 0, 2)
    hi = xs.half()
    lo = (xs - hi.float()).half()
    assert torch.equal(p[:, :50, :32], hi) and torch.equal(p[:, :50, 32:], lo)
    assert p[:, 50:].abs().max().item() == 0
    assert ((hi.float() + lo.float()) - xs).abs().max().item() < 1e-6


@pytest.mark.parametrize("N,Tk,qs,exact,tol", [
    (128, 128, 1.0, True, 3e-5),
    (300, 700, 1.0, True, 3e-5),
    (1674, 5022, 4.0, True, 1e-4),      # |S| up to ~100: 2^-22 relative operand error ~ 2e-5 on S
    (300, 700, 1.0, False, 5e-3),
    (1674, 3348, 2.0, False, 2e-2),
])
def test_lt_attention_tc(N, Tk, qs, exact, tol):
    from aot_benchmark_b200 import ops
    d = torch.device("cuda:0")
    g = torch.Generator().manual_seed(N + Tk)
    Q = (torch.randn(N, 256, generator=g) * qs).to(d)
    K = torch.randn(Tk, 256, generator=g).to(d)
    V = torch.randn(Tk, 256, generator=g).to(d)
    ref = _ref(Q, K, V)
    ncap = ((N + 255) // 256) * 256
    kcap = ((Tk + 127) // 128) * 128 + 128
    Qp, Kp, Vp = _pack(Q, ncap, math.sqrt(32.0)), _pack(K, kcap), _pack(V, kcap)
    O = torch.full((N, 256), float("nan"), device=d)
    dbg = torch.zeros(128 * 128 + 128 * 64, device=d)
    ops.lt_attention_tc(Qp, Kp, Vp, N, Tk, O=O, exact=exact, dbg=dbg)
    torch.cuda.synchronize()
    err = (O.cpu().double() - ref).abs().max().item()
    if not (err < tol):
        # diagnostics: raw scores of tile (q 0..127, keys 0..127, head 0) and the un-normalised output
        S = dbg[:128 * 128].view(128, 128).cpu().double()
        nq, nk = min(N, 128), min(Tk, 128)
        Sref = (Q[:nq, :32].double().cpu() / math.sqrt(32.0)) @ K[:nk, :32].double().cpu().t()
        print("S err", (S[:nq, :nk] - Sref).abs().max().item(), "S ref max", Sref.abs().max().item())
        print("S[0,:8]", S[0, :8].tolist(), "ref", Sref[0, :8].tolist())
        Od = dbg[128 * 128:].view(128, 64).cpu()
        print("O' row0", Od[0, :8].tolist(), Od[0, 32:40].tolist())
    assert err < tol, f"max |dO| = {err}"


def test_lt_attention_tc_matches_simt_and_splits():
    from aot_benchmark_b200 import ops
    d = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    N, Tk = 1674, 1674 * 4 + 77
    Q = (torch.randn(N, 256, generator=g) * 3).to(d)
    K = torch.randn(Tk, 256, generator=g).to(d)
    V = torch.randn(Tk, 256, generator=g).to(d)
    simt = torch.empty(N, 256, device=d)
    ops.attention(Q, K, V, simt, H, D, D)
    ncap, kcap = 1792, ((Tk + 127) // 128) * 128 + 256
    Qp, Kp, Vp = _pack(Q, ncap, math.sqrt(32.0)), _pack(K, kcap), _pack(V, kcap)
    O1 = torch.empty(N, 256, device=d)
    ops.lt_attention_tc(Qp, Kp, Vp, N, Tk, O=O1, exact=True)
    assert (O1 - simt).abs().max().item() < 1e-4
    for splits in (2, 5):
        part = (torch.empty(splits, N, 256, device=d), torch.empty(splits, H, N, device=d),
                torch.empty(splits, H, N, device=d))
        O2 = torch.empty(N, 256, device=d)
        ops.lt_attention_tc(Qp, Kp, Vp, N, Tk, O=O2, splits=splits, exact=True, part=part)
        assert (O2 - simt).abs().max().item() < 1e-4
    tk_dev = torch.tensor([Tk], dtype=torch.int32, device=d)
    O3 = torch.empty(N, 256, device=d)
    ops.lt_attention_tc(Qp, Kp, Vp, N, 1, O=O3, Tk_dev=tk_dev, exact=True)
    assert torch.equal(O3, O1)
    # more splits than tiles: empty splits must contribute nothing
    part = (torch.empty(8, 200, 256, device=d), torch.empty(8, H, 200, device=d), torch.empty(8, H, 200, device=d))
    O4 = torch.empty(200, 256, device=d)
    ops.lt_attention_tc(Qp, Kp, Vp, 200, 300, O=O4, splits=8, exact=True, part=part)
    ref = torch.empty(200, 256, device=d)
    ops.attention(Q[:200], K[:300], V[:300], ref, H, D, D)
    assert (O4 - ref).abs().max().item() < 1e-4


@pytest.mark.parametrize("variant", ["groups", "ahead", "pair"])
@pytest.mark.parametrize("N,Tk,splits,exact", [(128, 128, 1, True), (300, 700, 1, True), (1674, 5022 + 77, 1, True),
                                               (1674, 1674 * 7, 5, True), (200, 300, 8, True), (300, 700, 1, False),
                                               (1674, 1674 * 3 + 5, 3, False)])
def test_lt_attention_tc_layouts(N, Tk, splits, exact, variant):
    """The alternative softmax layouts ("groups": 2 threads per row, one TMEM read per tile; "ahead": three score
    buffers, TMEM read under the ex2 pass; "pair": two co-resident CTAs per SM, 64-key tiles, packed-fp32 softmax with a
    truncating hi / lo split of P) compute the same maxima and the same P as the default one-tile layout; only
    the association of the row sums may differ: outputs agree to ~1e-6 and all match the fp64 oracle."""
    from aot_benchmark_b200 import ops
    d = torch.device("cuda:0")
    g = torch.Generator().manual_seed(N * 7 + Tk)
    Q = (torch.randn(N, 256, generator=g) * 3).to(d)
    K = torch.randn(Tk, 256, generator=g).to(d)
    V = torch.randn(Tk, 256, generator=g).to(d)
    ncap = ((N + 255) // 256) * 256
    kcap = ((Tk + 127) // 128) * 128 + 128
    Qp, Kp, Vp = _pack(Q, ncap, math.sqrt(32.0)), _pack(K, kcap), _pack(V, kcap)
    outs = []
    for v in ("tile", variant):
        part = None
        if splits > 1:
            part = (torch.zeros(splits, N, 256, device=d), torch.zeros(splits, H, N, device=d),
                    torch.zeros(splits, H, N, device=d))
        O = torch.full((N, 256), float("nan"), device=d)
        ops.lt_attention_tc(Qp, Kp, Vp, N, Tk, O=O, splits=splits, exact=exact, part=part, variant=v)
        torch.cuda.synchronize()
        outs.append((O, part))
    assert torch.isfinite(outs[1][0]).all()
    # exact mode: same P to fp32 rounding.  fast mode keeps one fp16 pass of P, whose rounding depends on the running row
    # maximum and therefore on the key tiling: layouts with other tiles differ at the 1e-4 level (both within 5e-2 of fp64)
    assert (outs[0][0] - outs[1][0]).abs().max().item() < (2e-5 if exact else 1e-3)
    if splits > 1 and variant != "pair":                        # ("pair" cuts the key range into 64-key tiles: other split bounds)
        assert torch.equal(outs[0][1][1], outs[1][1][1])          # per-split row maxima are identical
    ref = _ref(Q, K, V)
    assert (outs[1][0].cpu().double() - ref).abs().max().item() < (2e-4 if exact else 5e-2)


def _pack_w(w):  # [Cout,Cin,KH,KW] -> fp32 [K, Cout] (k = (ky,kx,ci))
    co, ci, kh, kw = w.shape
    return w.permute(2, 3, 1, 0).reshape(kh * kw * ci, co).contiguous()


@pytest.mark.parametrize("cfg", [
    # B, H, W, Cin, Cout, K, stride, pad, res, act
    (1, 37, 53, 64, 64, 1, 1, 0, False, 1),
    (1, 121, 213, 64, 256, 1, 1, 0, True, 1),      # layer1 conv3 + residual, BN=256/128 path
    (1, 31, 54, 256, 256, 3, 1, 1, False, 1),      # layer3 3x3
    (1, 61, 107, 128, 128, 3, 2, 1, False, 1),     # strided 3x3
    (1, 61, 107, 256, 512, 1, 2, 0, False, 0),     # strided 1x1 downsample
    (1, 31, 54, 1024, 256, 1, 1, 0, False, 0),     # projector, K = 1024
    (2, 20, 24, 128, 192, 3, 1, 1, True, 0),       # batch 2, Cout = 192 (BN = 64 only)
    (1, 1674, 1, 512, 256, 1, 1, 0, True, 0),      # linear with in-place style residual
    (1, 65, 97, 4, 64, 7, 2, 3, False, 1),         # 7x7 stem on the zero-padded 4-channel image (K = 196 -> 256)
])
def test_conv2d_tc(cfg):
    from aot_benchmark_b200 import ops
    import torch.nn.functional as F
    B, Hh, Ww, Cin, Cout, K, s, p, use_res, act = cfg
    d = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, Cin, Hh, Ww, generator=g) * 2
    w = torch.randn(Cout, Cin, K, K, generator=g) / math.sqrt(Cin * K * K)
    b = torch.randn(Cout, generator=g)
    ref = F.conv2d(x.double(), w.double(), b.double(), s, p)
    res = torch.randn(ref.shape, generator=g) if use_res else None
    if use_res:
        ref = ref + res.double()
    if act == 1:
        ref = F.relu(ref)
    wk = _pack_w(w).to(d)
    wh, wl = ops.split_fp16(wk)
    xg = x.permute(0, 2, 3, 1).contiguous().to(d)
    out = torch.full((B, ref.shape[2], ref.shape[3], Cout), float("nan"), device=d)
    rg = res.permute(0, 2, 3, 1).contiguous().to(d) if use_res else None
    from aot_benchmark_b200._lib import lib
    scale = ref.abs().max().item()
    try:
        for mode in (1, 0):      # narrow tiles, then the default wide tiles + split-K clusters
            assert lib().aotb_set_conv_tiling(mode) == 0
            out.fill_(float("nan"))
            ops.conv2d_tc(xg, wh, wl, b.to(d), out, res=rg, KH=K, KW=K, stride=s, pad=p, act=act)
            torch.cuda.synchronize()
            err = (out.permute(0, 3, 1, 2).cpu().double() - ref).abs().max().item()
            # tensor-core fp32 accumulation over K up to 2304 terms: ~5e-6 relative (fp32 CUDA cores: ~1e-6)
            assert err < 1e-5 * max(scale, 1.0) + 1e-5, f"tiling {mode}: err {err} scale {scale}"
            # split-K sums in cluster-rank order: bit-identical from run to run
            out_b = torch.full_like(out, float("nan"))
            ops.conv2d_tc(xg, wh, wl, b.to(d), out_b, res=rg, KH=K, KW=K, stride=s, pad=p, act=act)
            assert torch.equal(out, out_b)
    finally:
        lib().aotb_set_conv_tiling(0)
    # and the fp32 CUDA-core kernel on the same problem agrees
    out2 = torch.empty_like(out)
    old = ops.CONV_IMPL
    ops.CONV_IMPL = "simt"
    try:
        ops.conv2d(xg, wk, b.to(d), out2, res=rg, KH=K, KW=K, stride=s, pad=p, act=act)
    finally:
        ops.CONV_IMPL = old
    assert (out - out2).abs().max().item() < 2e-5 * max(scale, 1.0)


def test_linear_tc_channel_slices_inplace():
    from aot_benchmark_b200 import ops
    import torch.nn.functional as F
    d = torch.device("cuda:0")
    g = torch.Generator().manual_seed(2)
    big = torch.randn(300, 1024, generator=g)
    w = torch.randn(256, 512, generator=g) / 22          # nn.Linear weight [out, in]
    b = torch.randn(256, generator=g)
    y = torch.randn(300, 512, generator=g)
    ref = y[:, 256:].double() + F.linear(big[:, 512:].double(), w.double(), b.double())
    wk = w.t().contiguous().to(d)
    ops.register_tc_weights(wk, *ops.split_fp16(wk))
    yg, bg = y.to(d), big.to(d)
    ops.linear(bg[:, 512:], wk, b.to(d), yg[:, 256:], res=yg[:, 256:])
    assert (yg[:, 256:].cpu().double() - ref).abs().max().item() < 5e-5
    assert torch.equal(yg[:, :256].cpu(), y[:, :256])
