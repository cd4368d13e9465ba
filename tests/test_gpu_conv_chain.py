"""GPU: the persistent conv-chain kernel (csrc/conv_chain.cu) against the per-layer tensor-core conv and, through the engine,
against the real reference's goldens."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _chain(h0, w0, g, d):
    from aot_benchmark_b200 import ops
    layers = []

    def conv(x, cin, cout, k=1, stride=1, pad=0, res=None, in_layer=-1, res_layer=-1, act=1):
        H, W = x.shape[1], x.shape[2]
        ho, wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        w = (torch.randn(k * k * cin, cout, generator=g) / (k * k * cin) ** 0.5).to(d)
        wh, wl = ops.split_fp16(w)
        ops.register_tc_weights(w, wh, wl)
        out = torch.full((1, ho, wo, cout), float("nan"), device=d)
        layers.append(dict(x=x, w=w, bias=torch.randn(cout, generator=g).to(d), out=out, res=res, KH=k, stride=stride, pad=pad,
                           act=act, in_layer=in_layer, res_layer=res_layer))
        return out, len(layers) - 1

    x0 = torch.randn(1, h0, w0, 64, generator=g).to(d)
    cur, cur_i = x0, -1
    for (mid, cout, stride, nblk) in ((64, 256, 1, 2), (128, 512, 2, 2), (256, 1024, 2, 3)):
        for bi in range(nblk):
            s = stride if bi == 0 else 1
            cin = cur.shape[3]
            t1, i1 = conv(cur, cin, mid, in_layer=cur_i)
            t2, i2 = conv(t1, mid, mid, 3, s, 1, in_layer=i1)
            if bi == 0:
                res, ri = conv(cur, cin, cout, 1, s, 0, in_layer=cur_i, act=0)
            else:
                res, ri = cur, cur_i
            cur, cur_i = conv(t2, mid, cout, res=res, in_layer=i2, res_layer=ri)
    conv(cur, cur.shape[3], 256, in_layer=cur_i, act=0)
    return layers


@pytest.mark.parametrize("h0,w0", [(31, 45), (121, 213)])
def test_conv_chain_matches_per_layer_kernels(h0, w0):
    """Layer by layer (each layer of the reference pass reads the CHAIN's output of its producer, so errors do not compound):
    bit-identical where the layer is not split along K (same chunk order as the per-layer kernel without split-K), within fp32
    summation-order distance where it is; and the chain is bit-reproducible run to run (fixed split order)."""
    from aot_benchmark_b200 import ops
    from aot_benchmark_b200._lib import lib
    d = torch.device("cuda:0")
    g = torch.Generator().manual_seed(h0)
    layers = _chain(h0, w0, g, d)
    _, table = ops.conv_chain_dump(layers)
    chain = ops.ConvChain(layers, d)
    for rep in range(3):                                   # repeated launches: counters are reset by run()
        for l in layers:
            l["out"].fill_(float("nan"))
        chain.run()
        torch.cuda.synchronize()
    got = [l["out"].clone() for l in layers]
    assert all(torch.isfinite(a).all() for a in got)
    lib().aotb_set_conv_tiling(1 << 8)                     # per-layer kernel without split-K
    try:
        for i, l in enumerate(layers):
            ref = torch.empty_like(l["out"])
            ops.conv2d(l["x"], l["w"], l["bias"], ref, res=l.get("res"), KH=l["KH"], KW=l["KH"], stride=l["stride"],
                       pad=l["pad"], act=l["act"])
            torch.cuda.synchronize()
            if table[i][7] == 1:
                assert torch.equal(got[i], ref), f"layer {i}: max |d| = {(got[i] - ref).abs().max().item():.3e}"
            else:
                assert (got[i] - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item()), f"layer {i} (split-K)"
    finally:
        lib().aotb_set_conv_tiling(0)
    assert any(t[7] > 1 for t in table) or h0 > 100
    for l in layers:
        l["out"].fill_(float("nan"))
    chain.run()
    torch.cuda.synchronize()
    for i, l in enumerate(layers):
        assert torch.equal(got[i], l["out"]), f"layer {i}: the chain is not deterministic run to run"


@pytest.mark.parametrize("name", ["video_r50_aotl_small", "full_r50_aotl_480p"])
def test_engine_with_conv_chain_vs_reference_golden(name, golden_dir, monkeypatch):
    from aot_benchmark_b200 import ops
    from oracle import aot_oracle as O
    from oracle import weights as OW
    from oracle.fixtures import load_full_labels
    from test_gpu_engine import _build_cuda_engine
    monkeypatch.setattr(ops, "CONV_CHAIN", True)
    g = torch.load(os.path.join(golden_dir, f"{name}.pt"))
    sd = OW.build_state_dict(g["model"], seed=g["seed"], flavour=g["flavour"])
    frames, mask = O.synthetic_video(g["frames"], g["H"], g["W"], g["objs"], seed=1234 + g["seed"])
    full = name.startswith("full")
    ref_labels = load_full_labels(g) if full else [l.float() for l in g["ref_labels"]]
    eng = _build_cuda_engine(g["model"], sd, g["gap"])
    with torch.no_grad():
        lo, _ = O.run_video(eng, [f.cuda() for f in frames], mask.cuda(), g["objs"], tuple(g["out_size"]),
                            forced_masks=ref_labels)
    assert getattr(eng.aot_engines[0]._enc, "_chain", None) is not None, "the conv chain did not run"
    n = g["objs"] + 1
    if full:
        dmax = max((lo[t - 1].cpu()[:, :n] - g["ref_logits_lo"][t][:, :n]).abs().max().item() for t in g["logit_frames"])
    else:
        dmax = max((a.cpu()[:, :n] - b[:, :n]).abs().max().item() for a, b in zip(lo, g["ref_logits_lo"]))
    assert dmax < 1e-3, dmax
