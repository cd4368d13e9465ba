"""GPU: every C-ABI kernel against a CPU fp32 restatement (torch functional ops / the oracle)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def _pack_conv(w):  # [Cout,Cin,KH,KW] -> [KH*KW*Cin, Cout]
    co, ci, kh, kw = w.shape
    return w.permute(2, 3, 1, 0).reshape(kh * kw * ci, co).contiguous()


@pytest.mark.parametrize("cfg", [
    # B, H, W, Cin, Cout, K, stride, pad, res, act
    (1, 37, 53, 64, 64, 1, 1, 0, False, 1),
    (1, 37, 53, 64, 256, 1, 1, 0, True, 1),
    (1, 31, 54, 256, 256, 3, 1, 1, False, 1),
    (1, 61, 107, 128, 128, 3, 2, 1, False, 1),
    (1, 65, 97, 3, 64, 7, 2, 3, False, 1),       # stem: scalar A path
    (1, 40, 50, 128, 11, 1, 1, 0, False, 0),     # conv_out: scalar B path
    (1, 33, 49, 11, 256, 17, 16, 8, False, 0),   # dense ID bank path (scalar A)
    (2, 20, 24, 32, 96, 3, 1, 1, True, 0),
    (1, 16, 16, 24, 144, 1, 1, 0, False, 4),     # K % 16 != 0, ReLU6 (mobilenet)
    (1, 121, 213, 256, 128, 1, 1, 0, True, 0),   # big M
])
def test_conv2d(cfg):
    from aot_benchmark_b200 import ops
    B, H, W, Cin, Cout, K, s, p, use_res, act = cfg
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, K, K, generator=g) / math.sqrt(Cin * K * K)
    b = torch.randn(Cout, generator=g)
    ref = F.conv2d(x, w, b, s, p)
    res = torch.randn_like(ref) if use_res else None
    if use_res:
        ref = ref + res
    ref = {0: lambda t: t, 1: F.relu, 4: F.relu6}[act](ref)
    d = _dev()
    xg = x.permute(0, 2, 3, 1).contiguous().to(d)
    out = torch.empty(B, ref.shape[2], ref.shape[3], Cout, device=d)
    rg = res.permute(0, 2, 3, 1).contiguous().to(d) if use_res else None
    ops.conv2d(xg, _pack_conv(w).to(d), b.to(d), out, res=rg, KH=K, KW=K, stride=s, pad=p, act=act)
    assert _rel(out.permute(0, 3, 1, 2), ref) < 2e-5


def test_conv2d_channel_slices():
    """ld > C on input and output (writing into / reading from a wider buffer)."""
    from aot_benchmark_b200 import ops
    d = _dev()
    g = torch.Generator().manual_seed(2)
    big_in = torch.randn(1, 10, 12, 96, generator=g)
    w = torch.randn(48, 32, 1, 1, generator=g) * 0.1
    b = torch.randn(48, generator=g)
    big_out = torch.zeros(1, 10, 12, 128)
    ref = F.conv2d(big_in[..., 32:64].permute(0, 3, 1, 2), w, b).permute(0, 2, 3, 1)
    gi, go = big_in.to(d), big_out.to(d)
    ops.conv2d(gi[..., 32:64], _pack_conv(w).to(d), b.to(d), go[..., 64:112])
    assert _rel(go[..., 64:112], ref) < 2e-5
    assert go[..., :64].abs().max().item() == 0 and go[..., 112:].abs().max().item() == 0


def test_linear_inplace_residual():
    from aot_benchmark_b200 import ops
    d = _dev()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1674, 512, generator=g)
    w = torch.randn(256, 512, generator=g) / 22
    b = torch.randn(256, generator=g)
    y = torch.randn(1674, 256, generator=g)
    ref = y + F.linear(x, w, b)
    yg = y.to(d)
    ops.linear(x.to(d), w.t().contiguous().to(d), b.to(d), yg, res=yg)
    assert _rel(yg, ref) < 2e-5
    ref2 = F.silu(F.linear(x, w, b))
    o = torch.empty(1674, 256, device=d)
    ops.linear(x.to(d), w.t().contiguous().to(d), b.to(d), o, act=3)
    assert _rel(o, ref2) < 2e-5


def test_transposes_maxpool():
    from aot_benchmark_b200 import ops
    d = _dev()
    x = torch.randn(1, 3, 45, 67)
    o = torch.empty(1, 45, 67, 3, device=d)
    ops.nchw_to_nhwc(x.to(d), o)
    assert torch.equal(o.cpu(), x.permute(0, 2, 3, 1))
    back = torch.empty(1, 3, 45, 67, device=d)
    ops.nhwc_to_nchw(o, back)
    assert torch.equal(back.cpu(), x)
    y = torch.randn(1, 64, 41, 59)
    ref = F.max_pool2d(y, 3, 2, 1)
    mo = torch.empty(1, ref.shape[2], ref.shape[3], 64, device=d)
    ops.maxpool3x3s2(y.permute(0, 2, 3, 1).contiguous().to(d), mo)
    assert torch.equal(mo.permute(0, 3, 1, 2).cpu(), ref)


@pytest.mark.parametrize("K,stride,dil,act", [(5, 1, 1, 0), (3, 2, 1, 4), (3, 1, 2, 4)])
def test_dwconv(K, stride, dil, act):
    from aot_benchmark_b200 import ops
    d = _dev()
    g = torch.Generator().manual_seed(4)
    C = 96
    x = torch.randn(1, C, 23, 31, generator=g)
    w = torch.randn(C, 1, K, K, generator=g) * 0.2
    b = torch.randn(C, generator=g) if act else None
    pad = (K - 1) // 2 * dil
    ref = F.conv2d(x, w, b, stride, pad, dil, C)
    if act == 4:
        ref = F.relu6(ref)
    out = torch.empty(1, ref.shape[2], ref.shape[3], C, device=d)
    ops.dwconv(x.permute(0, 2, 3, 1).contiguous().to(d), w.permute(2, 3, 1, 0).reshape(K * K, C).contiguous().to(d),
               b.to(d) if b is not None else None, out, K=K, stride=stride, pad=pad, dil=dil, act=act)
    assert _rel(out.permute(0, 3, 1, 2), ref) < 1e-5


@pytest.mark.parametrize("align", [True, False])
def test_bilinear(align):
    from aot_benchmark_b200 import ops
    d = _dev()
    x = torch.randn(1, 32, 31, 54)
    ref = F.interpolate(x, size=(61, 107), mode="bilinear", align_corners=align)
    out = torch.empty(1, 61, 107, 32, device=d)
    ops.bilinear(x.permute(0, 2, 3, 1).contiguous().to(d), out, align)
    assert (out.permute(0, 3, 1, 2).cpu() - ref).abs().max().item() < 2e-6


def test_layernorm_with_pos():
    from aot_benchmark_b200 import ops
    d = _dev()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1000, 256, generator=g) * 3 + 1
    ga, be = torch.randn(256, generator=g), torch.randn(256, generator=g)
    pos = torch.randn(1000, 256, generator=g)
    ref = F.layer_norm(x, (256,), ga, be)
    o1 = torch.empty(1000, 256, device=d)
    o2 = torch.empty(1000, 512, device=d)
    ops.layernorm(x.to(d), ga.to(d), be.to(d), o1, add=pos.to(d), out2=o2[:, 256:])
    assert (o1.cpu() - ref).abs().max().item() < 1e-5
    assert (o2[:, 256:].cpu() - (ref + pos)).abs().max().item() < 1e-5


@pytest.mark.parametrize("P,C,G,act", [(1674, 1024, 32, 2), (6527, 128, 8, 1), (1674, 512, 2, 0), (25773, 128, 8, 1)])
def test_groupnorm(P, C, G, act):
    from aot_benchmark_b200 import ops
    d = _dev()
    g = torch.Generator().manual_seed(6)
    x = torch.randn(1, P, C, generator=g) * 2 + 0.5
    ga, be = torch.randn(C, generator=g), torch.randn(C, generator=g)
    ref = F.group_norm(x.permute(0, 2, 1), G, ga, be).permute(0, 2, 1)
    ref = {0: lambda t: t, 1: F.relu, 2: F.gelu}[act](ref)
    xg = x.to(d)
    ws = ops.groupnorm_workspace(1, G, d)
    ops.groupnorm(xg, ga.to(d), be.to(d), xg, G, act, ws)
    assert (xg.cpu() - ref).abs().max().item() < 2e-5


def test_eltwise():
    from aot_benchmark_b200 import ops
    d = _dev()
    a, b = torch.randn(50, 64), torch.randn(50, 64)
    ag, bg = a.to(d), b.to(d)
    o = torch.zeros(50, 128, device=d)
    ops.eltwise(ops.EW_SILU_MUL, ag, bg, o[:, 64:])
    assert (o[:, 64:].cpu() - F.silu(a) * b).abs().max().item() < 1e-6
    ops.eltwise(ops.EW_FILL, None, None, o[:, :64], scalar=1.0)
    assert torch.equal(o[:, :64].cpu(), torch.ones(50, 64))
    ops.eltwise(ops.EW_ADD, ag, bg, o[:, :64])
    assert torch.equal(o[:, :64].cpu(), a + b)


@pytest.mark.parametrize("H,dq,dv,N,Tk,qs", [(8, 32, 32, 100, 333, 1.0), (8, 32, 32, 1674, 5022, 6.0),
                                             (1, 128, 1024, 70, 150, 1.0), (1, 128, 1024, 300, 900, 4.0)])
def test_attention(H, dq, dv, N, Tk, qs):
    from aot_benchmark_b200 import ops
    from oracle import aot_oracle as O
    d = _dev()
    g = torch.Generator().manual_seed(7)
    Q = torch.randn(N, 1, H * dq, generator=g) * qs
    K = torch.randn(Tk, 1, H * dq, generator=g)
    V = torch.randn(Tk, 1, H * dv, generator=g)
    ref = O.multihead_attention(Q.double(), K.double(), V.double(), H, d_att=dq)[:, 0]
    o = torch.empty(N, H * dv, device=d)
    Kg = torch.zeros(Tk + 77, H * dq, device=d)
    Kg[:Tk] = K[:, 0].to(d)
    Vg = torch.zeros(Tk + 77, H * dv, device=d)
    Vg[:Tk] = V[:, 0].to(d)
    ops.attention(Q[:, 0].to(d), Kg, Vg, o, H, dq, dv, Tk=Tk)
    assert (o.cpu().double() - ref).abs().max().item() < 2e-5
    tk_dev = torch.tensor([Tk], dtype=torch.int32, device=d)
    o2 = torch.empty_like(o)
    ops.attention(Q[:, 0].to(d), Kg, Vg, o2, H, dq, dv, Tk=1, Tk_dev=tk_dev)
    assert torch.equal(o, o2)


def test_attention_split_kv_merge():
    """Split-KV partials merged with the exact LSE merge == unsharded attention (cfg4 row e)."""
    from aot_benchmark_b200 import ops
    d = _dev()
    g = torch.Generator().manual_seed(8)
    H, dd, N, Tk, R = 8, 32, 257, 1000, 3
    Q = (torch.randn(N, H * dd, generator=g) * 5).to(d)
    K = torch.randn(Tk, H * dd, generator=g).to(d)
    V = torch.randn(Tk, H * dd, generator=g).to(d)
    full = torch.empty(N, H * dd, device=d)
    ops.attention(Q, K, V, full, H, dd, dd)
    bounds = [0, 300, 650, Tk]
    Op = torch.empty(R, N, H * dd, device=d)
    Mp = torch.empty(R, H, N, device=d)
    Lp = torch.empty(R, H, N, device=d)
    for r in range(R):
        ops.attention(Q, K[bounds[r]:bounds[r + 1]], V[bounds[r]:bounds[r + 1]], Op[r], H, dd, dd, Mout=Mp[r], Lout=Lp[r])
    merged = torch.empty_like(full)
    ops.attn_merge(Op, Mp, Lp, merged, H, dd)
    assert (merged - full).abs().max().item() < 1e-5


def test_local_attention_aot():
    from aot_benchmark_b200 import ops
    from oracle import aot_oracle as O
    d = _dev()
    g = torch.Generator().manual_seed(9)
    H, dd, h, w = 8, 32, 13, 22
    q = torch.randn(1, 256, h, w, generator=g) * 2
    k = torch.randn(1, 256, h, w, generator=g)
    v = torch.randn(1, 256, h, w, generator=g)
    rkw = torch.randn(1800, 32, 1, 1, generator=g) * 0.2
    rkb = torch.randn(1800, generator=g) * 0.1
    rv = torch.randn(8, 32, 225, generator=g) * 0.3
    ref = O.local_attention(q.double(), k.double(), v.double(), rkw.double(), rkb.double(), rv.double(), H)[:, 0]
    tok = lambda t: t[0].permute(1, 2, 0).reshape(h * w, -1).contiguous().to(d)
    out = torch.empty(h * w, 256, device=d)
    ops.local_attention(tok(q), tok(k), tok(v), rkw.view(1800, 32).contiguous().to(d), rkb.to(d), rv.to(d), out,
                        h, w, H, dd, dd)
    assert (out.cpu().double() - ref).abs().max().item() < 2e-5


def test_local_attention_deaot():
    from aot_benchmark_b200 import ops
    from oracle import aot_oracle as O
    d = _dev()
    g = torch.Generator().manual_seed(10)
    h, w = 9, 17
    q = torch.randn(1, 128, h, w, generator=g)
    k = torch.randn(1, 128, h, w, generator=g)
    v = torch.randn(1, 1024, h, w, generator=g)
    rkw = torch.randn(225, 128, 1, 1, generator=g) * 0.1
    rkb = torch.randn(225, generator=g) * 0.1
    ref = O.local_attention(q.double(), k.double(), v.double(), rkw.double(), rkb.double(), None, 1)[:, 0]
    tok = lambda t: t[0].permute(1, 2, 0).reshape(h * w, -1).contiguous().to(d)
    out = torch.empty(h * w, 1024, device=d)
    ops.local_attention(tok(q), tok(k), tok(v), rkw.view(225, 128).contiguous().to(d), rkb.to(d), None, out,
                        h, w, 1, 128, 1024)
    assert (out.cpu().double() - ref).abs().max().item() < 2e-5


@pytest.mark.parametrize("h,w", [(9, 17), (31, 54), (8, 6), (13, 5)])
def test_local_gated_tile_deaot(h, w):
    """Tiled DeAOT short-term kernel (8x6 query tiles, channels in chunks of 32) vs the fp64 oracle (attention.py:789-861 in the
    unfold form) and vs the generic per-warp kernel, incl. ragged tiles and maps smaller than the window."""
    from aot_benchmark_b200 import ops
    from oracle import aot_oracle as O
    d = _dev()
    g = torch.Generator().manual_seed(100 + h)
    q = torch.randn(1, 128, h, w, generator=g)
    k = torch.randn(1, 128, h, w, generator=g)
    v = torch.randn(1, 1024, h, w, generator=g)
    rkw = torch.randn(225, 128, 1, 1, generator=g) * 0.1
    rkb = torch.randn(225, generator=g) * 0.1
    ref = O.local_attention(q.double(), k.double(), v.double(), rkw.double(), rkb.double(), None, 1)[:, 0]
    tok = lambda t: t[0].permute(1, 2, 0).reshape(h * w, -1).contiguous().to(d)
    out = torch.full((h * w, 1024), float("nan"), device=d)
    ops.local_gated_tile(tok(q), tok(k), tok(v), rkw.view(225, 128).contiguous().to(d), rkb.to(d), out, h, w)
    gen = torch.empty(h * w, 1024, device=d)
    ops.local_attention(tok(q), tok(k), tok(v), rkw.view(225, 128).contiguous().to(d), rkb.to(d), None, gen, h, w, 1, 128, 1024)
    assert torch.isfinite(out).all()
    assert (out.cpu().double() - ref).abs().max().item() < 2e-5
    assert (out - gen).abs().max().item() < 2e-5


@pytest.mark.parametrize("align,ln", [(True, False), (True, True), (False, False)])
def test_id_embed(align, ln):
    from aot_benchmark_b200 import ops
    d = _dev()
    g = torch.Generator().manual_seed(11)
    k, pad = (17, 8) if align else (16, 0)
    Hm, Wm = (161, 241) if align else (160, 240)
    mask = torch.randint(0, 11, (1, 1, Hm // 8, Wm // 8), generator=g).float()
    mask = F.interpolate(mask, size=(Hm, Wm), mode="nearest")
    w = torch.randn(256, 11, k, k, generator=g) * 0.05
    b = torch.randn(256, generator=g) * 0.1
    onehot = (mask == torch.arange(11).view(1, -1, 1, 1)).float()
    ref = F.conv2d(onehot, w, b, 16, pad)
    ga, be = torch.randn(256, generator=g), torch.randn(256, generator=g)
    if ln:
        ref = F.layer_norm(ref.permute(0, 2, 3, 1), (256,), ga, be).permute(0, 3, 1, 2)
    ho, wo = ref.shape[2:]
    out = torch.empty(ho * wo, 256, device=d)
    ops.id_embed(mask[0, 0].to(d), _pack_conv(w).to(d), b.to(d), out, 256, 11, k, 16, pad,
                 ln_gamma=ga.to(d) if ln else None, ln_beta=be.to(d) if ln else None)
    assert (out.cpu() - ref[0].permute(1, 2, 0).reshape(ho * wo, 256)).abs().max().item() < 2e-5


def test_logits_postproc_argmax_nearest():
    from aot_benchmark_b200 import ops
    d = _dev()
    g = torch.Generator().manual_seed(12)
    lg = torch.randn(1, 11, 41, 61, generator=g)
    obj = 6
    ref_lo = lg.clone()
    ref_lo[:, obj + 1:] = -1e10
    ref = F.interpolate(ref_lo, size=(150, 230), mode="bilinear", align_corners=True)
    lo = torch.empty(1, 11, 41, 61, device=d)
    out = torch.empty(1, 11, 150, 230, device=d)
    ops.logits_postproc(lg.permute(0, 2, 3, 1).contiguous().to(d), lo, out, obj, True)
    assert torch.equal(lo.cpu(), ref_lo)
    assert (out.cpu()[:, :obj + 1] - ref[:, :obj + 1]).abs().max().item() < 2e-6
    label = torch.empty(1, 150, 230, device=d)
    ops.logits_argmax(lo, label, True)
    mism = (label.cpu() != out.cpu().argmax(1).float()).float().mean().item()
    assert mism < 1e-4
    small = torch.empty(1, 1, 97, 161, device=d)
    ops.nearest_resize(label.view(1, 1, 150, 230), small)
    assert torch.equal(small.cpu(), F.interpolate(label.view(1, 1, 150, 230).cpu(), size=(97, 161), mode="nearest"))


def test_bank_append():
    from aot_benchmark_b200 import ops
    d = _dev()
    bank = torch.zeros(100, 256, device=d)
    src = torch.randn(30, 512, device=d)
    ops.bank_append(src[:, 256:], bank, 40)
    assert torch.equal(bank[40:70], src[:, 256:]) and bank[:40].abs().max().item() == 0
    off = torch.tensor([70], dtype=torch.int32, device=d)
    ops.bank_append(src[:, :256], bank, 0, offset_dev=off)
    assert torch.equal(bank[70:100], src[:, :256])


def test_errors_are_loud():
    from aot_benchmark_b200 import ops
    from aot_benchmark_b200._lib import AotbError
    with pytest.raises(AotbError):
        ops.linear(torch.zeros(4, 4), torch.zeros(4, 4), None, torch.zeros(4, 4))  # CPU tensors are refused
    d = _dev()
    with pytest.raises(AotbError):
        ops.attention(torch.zeros(4, 40, device=d), torch.zeros(4, 40, device=d), torch.zeros(4, 40, device=d),
                      torch.zeros(4, 40, device=d), 1, 40, 40)  # unsupported head shape


@pytest.mark.parametrize("h,w", [(13, 22), (31, 54), (8, 8), (5, 40)])
def test_local_attention_tile_aot(h, w):
    """Halo-in-shared-memory kernel vs the fp64 oracle and vs the generic per-warp kernel."""
    from aot_benchmark_b200 import ops
    from oracle import aot_oracle as O
    d = _dev()
    g = torch.Generator().manual_seed(13 + h)
    H, dd = 8, 32
    q = torch.randn(1, 256, h, w, generator=g) * 2
    k = torch.randn(1, 256, h, w, generator=g)
    v = torch.randn(1, 256, h, w, generator=g)
    rkw = torch.randn(1800, 32, 1, 1, generator=g) * 0.2
    rkb = torch.randn(1800, generator=g) * 0.1
    rv = torch.randn(8, 32, 225, generator=g) * 0.3
    ref = O.local_attention(q.double(), k.double(), v.double(), rkw.double(), rkb.double(), rv.double(), H)[:, 0]
    tok = lambda t: t[0].permute(1, 2, 0).reshape(h * w, -1).contiguous().to(d)
    out = torch.full((h * w, 512), float("nan"), device=d)
    ops.local_attention_tile(tok(q), tok(k), tok(v), rkw.view(1800, 32).contiguous().to(d), rkb.to(d),
                             rv.permute(0, 2, 1).contiguous().to(d), out[:, 256:], h, w, H)
    assert (out[:, 256:].cpu().double() - ref).abs().max().item() < 2e-5
    out2 = torch.empty(h * w, 256, device=d)
    ops.local_attention(tok(q), tok(k), tok(v), rkw.view(1800, 32).contiguous().to(d), rkb.to(d), rv.to(d), out2,
                        h, w, H, dd, dd)
    assert (out[:, 256:] - out2).abs().max().item() < 1e-5


@pytest.mark.parametrize("align,ln", [(True, False), (True, True), (False, False)])
def test_id_embed_runs(align, ln):
    """Run-length / prefix-sum gather == dense conv of the one-hot mask (random blocky mask incl. invalid ids)."""
    from aot_benchmark_b200 import ops
    d = _dev()
    g = torch.Generator().manual_seed(21)
    k, pad = (17, 8) if align else (16, 0)
    Hm, Wm = (161, 241) if align else (160, 240)
    mask = torch.randint(0, 13, (1, 1, Hm // 5, Wm // 3), generator=g).float()      # ids 11, 12 are out of range
    mask = F.interpolate(mask, size=(Hm, Wm), mode="nearest")
    mask[0, 0, 5:9, 7:30] = 2.5                                                     # non-integer labels match nothing
    w = torch.randn(256, 11, k, k, generator=g) * 0.05
    b = torch.randn(256, generator=g) * 0.1
    onehot = (mask == torch.arange(11).view(1, -1, 1, 1)).float()
    ref = F.conv2d(onehot.double(), w.double(), b.double(), 16, pad)
    ga, be = torch.randn(256, generator=g), torch.randn(256, generator=g)
    if ln:
        ref = F.layer_norm(ref.permute(0, 2, 3, 1), (256,), ga.double(), be.double()).permute(0, 3, 1, 2)
    ho, wo = ref.shape[2:]
    t = w.double().permute(2, 3, 1, 0)
    pre = torch.zeros(k, k + 1, 11, 256, dtype=torch.float64)
    pre[:, 1:] = torch.cumsum(t, dim=1)
    out = torch.empty(ho * wo, 256, device=d)
    ops.id_embed_runs(mask[0, 0].to(d), pre.float().to(d), b.to(d), out, 256, 11, k, 16, pad,
                      ln_gamma=ga.to(d) if ln else None, ln_beta=be.to(d) if ln else None)
    err = (out.cpu().double() - ref[0].permute(1, 2, 0).reshape(ho * wo, 256)).abs().max().item()
    assert err < (2e-4 if ln else 3e-6), err


def test_soft_logit_aggregation_and_separate_labels_kernels():
    """Row f.2 kernels vs the reference's tensor expressions (aot_engine.py:515-533, 565-582)."""
    from aot_benchmark_b200 import ops
    d = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    E, per, H, W = 3, 10, 37, 53
    logits = [(torch.randn(1, 1 + per, H, W, generator=g) * 4).to(d) for _ in range(E)]
    logits[2][:, 6:] = -1e10                                  # unused ids of the last engine (aot_engine.py:368-370)
    out = torch.empty(1, 1 + E * per, H, W, device=d)
    ops.soft_logit_aggregation(logits, out, per)
    fg, bg = [], []
    for l in logits:
        p = torch.softmax(l, dim=1)
        bg.append(p[:, 0:1])
        fg.append(p[:, 1:1 + per])
    ref = torch.logit(torch.cat([torch.prod(torch.cat(bg, dim=1), dim=1, keepdim=True)] + fg, dim=1).clamp(1e-5, 1 - 1e-5))
    assert (out - ref).abs().max().item() < 2e-5
    assert torch.equal(out.argmax(1), ref.argmax(1))
    mask = torch.randint(0, 26, (1, 1, H, W), generator=g).float().to(d)
    sep = torch.empty(3, 1, 1, H, W, device=d)
    ops.separate_labels(mask, sep, per)
    for e in range(3):
        s_id, e_id = e * per + 1, (e + 1) * per
        fgm = ((mask >= s_id) & (mask <= e_id)).float()
        assert torch.equal(sep[e], (fgm * mask - s_id + 1) * fgm)


def test_frame_preprocess_kernel_vs_reference_fixture(golden_dir):
    """Row f.3: uint8 frame -> resize (cv2 INTER_CUBIC taps) + normalise + CHW in one kernel, against the tensors the REAL
    reference's MultiRestrictSize + MultiToTensor produced (tests/golden/io_side.pt)."""
    import os
    from aot_benchmark_b200.io_side import FramePreprocessor
    fx = torch.load(os.path.join(golden_dir, "io_side.pt"))
    img = fx["img"].numpy()
    for name, c in fx["cases"].items():
        kw = c["kw"]
        fp = FramePreprocessor(kw["max_short_edge"], kw["max_long_edge"], kw["flip"], kw["multi_scale"], kw["align_corners"])
        outs = fp(img)
        assert len(outs) == len(c["ref"])
        for o, r in zip(outs, c["ref"]):
            assert tuple(o.shape[1:]) == tuple(r.shape), name
            assert (o[0].cpu() - r).abs().max().item() < 2e-5, name
    # no-resize path: a frame that already has the network size
    fp = FramePreprocessor(None, 800, False, [1.0], True)
    same = torch.from_numpy(img[:113, :145].copy())
    o = fp(same.numpy())[0]
    from oracle import io_side as IO
    assert (o[0].cpu() - IO.preprocess(same.numpy(), None, 800, 1.0, True)).abs().max().item() < 1e-6


def test_async_mask_writer_from_device(tmp_path):
    import numpy as np
    from PIL import Image
    from aot_benchmark_b200.io_side import AsyncMaskWriter
    g = torch.Generator().manual_seed(3)
    masks = [torch.randint(0, 11, (1, 1, 60, 85), generator=g).float().cuda() for _ in range(5)]
    wr = AsyncMaskWriter(workers=2, ring=2)
    for i, m in enumerate(masks):
        wr.save(m, str(tmp_path / f"{i}.png"))
    wr.close()
    for i, m in enumerate(masks):
        assert np.array_equal(np.array(Image.open(tmp_path / f"{i}.png")), m[0, 0].cpu().numpy().astype(np.uint8))
