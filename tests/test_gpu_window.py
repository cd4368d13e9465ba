"""GPU: the Swin-B encoder path (BASELINE configs[3]) through the C ABI -- window attention and
patch-merge kernels against the CPU oracle, the whole encoder against oracle.swin_forward (bit-exact to the reference's
SwinTransformer, see oracle/gen_golden.py), and the SwinB-AOTL / SwinB-DeAOTL engines against the committed goldens of
the real reference."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _oracle_window_core(qkv_w, qkv_b, table, x_ln, H, W, heads, shift):
    """softmax(q k^T * d^-0.5 + bias + mask) v before `proj`, in the reference's padded/rolled/partitioned form
    (oracle.swin_block up to the projection)."""
    from oracle import aot_oracle as O
    ws, C = 7, x_ln.shape[1]
    d = C // heads
    y = x_ln.view(H, W, C)
    pb, pr = (ws - H % ws) % ws, (ws - W % ws) % ws
    y = F.pad(y, (0, 0, 0, pr, 0, pb))
    Hp, Wp = H + pb, W + pr
    if shift:
        y = torch.roll(y, (-shift, -shift), (0, 1))
    nwy, nwx = Hp // ws, Wp // ws
    win = y.view(nwy, ws, nwx, ws, C).permute(0, 2, 1, 3, 4).reshape(-1, ws * ws, C)
    qkv = F.linear(win, qkv_w, qkv_b).view(-1, ws * ws, 3, heads, d).permute(2, 0, 3, 1, 4)
    att = (qkv[0] * d ** -0.5) @ qkv[1].transpose(-2, -1)
    att = att + table[O.swin_rel_index(ws).reshape(-1)].view(ws * ws, ws * ws, heads).permute(2, 0, 1).unsqueeze(0)
    if shift:
        att = att + O.swin_shift_mask(Hp, Wp, ws, shift, att.dtype).unsqueeze(1)
    o = (torch.softmax(att, -1) @ qkv[2]).transpose(1, 2).reshape(-1, ws * ws, C)
    o = o.view(nwy, nwx, ws, ws, C).permute(0, 2, 1, 3, 4).reshape(Hp, Wp, C)
    if shift:
        o = torch.roll(o, (shift, shift), (0, 1))
    return o[:H, :W].reshape(H * W, C)


@pytest.mark.parametrize("H,W,heads,shift", [(14, 21, 4, 0), (14, 21, 4, 3), (9, 13, 16, 3), (37, 65, 8, 0), (5, 3, 4, 3),
                                             (36, 52, 4, 3)])
def test_window_attention_kernel(H, W, heads, shift):
    from aot_benchmark_b200 import ops
    from oracle import aot_oracle as O
    g = torch.Generator().manual_seed(H * 100 + W + shift)
    C = heads * 32
    x = torch.randn(H * W, C, generator=g)
    qkv_w = torch.randn(3 * C, C, generator=g) / C ** 0.5
    qkv_w[:C] *= 3.0                                     # sharp attention: the mask / bias terms must be right
    qkv_b = torch.randn(3 * C, generator=g) * 0.5
    table = torch.randn(169, heads, generator=g)
    want = _oracle_window_core(qkv_w, qkv_b, table, x, H, W, heads, shift)
    qkv = F.linear(x, qkv_w, qkv_b).cuda()               # the kernel's input: qkv of the UN-padded tokens
    relb = table[O.swin_rel_index(7).reshape(-1)].view(49, 49, heads).permute(2, 0, 1).contiguous().cuda()
    out = torch.full((H * W, C), float("nan"), device="cuda")
    ops.window_attention(qkv, qkv_b.cuda(), relb, out, H, W, heads, shift)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()                      # every un-padded token is written exactly once
    assert (out.cpu() - want).abs().max().item() < 1e-4 * max(1.0, want.abs().max().item())


def test_window_attention_rejects_other_geometries():
    from aot_benchmark_b200 import ops
    from aot_benchmark_b200._lib import AotbError
    q = torch.zeros(49, 3 * 64, device="cuda")
    with pytest.raises(AotbError):
        ops.window_attention(q, torch.zeros(192, device="cuda"), torch.zeros(1, 49, 49, device="cuda"),
                             torch.zeros(49, 64, device="cuda"), 7, 7, 1, 0)       # head dim 64


@pytest.mark.parametrize("H,W,C", [(8, 12, 128), (9, 13, 256), (1, 5, 128), (37, 65, 256)])
def test_patch_merge_kernel(H, W, C):
    from aot_benchmark_b200 import ops
    g = torch.Generator().manual_seed(H + W)
    x = torch.randn(H * W, C, generator=g)
    y = F.pad(x.view(H, W, C), (0, 0, 0, W % 2, 0, H % 2))
    want = torch.cat([y[0::2, 0::2], y[1::2, 0::2], y[0::2, 1::2], y[1::2, 1::2]], -1).reshape(-1, 4 * C)   # :352-357
    out = torch.full(want.shape, float("nan"), device="cuda")
    ops.patch_merge(x.cuda(), out, H, W)
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), want)                   # pure data movement: bit-exact


@pytest.mark.parametrize("H,W", [(64, 96), (75, 118), (144, 208)])
def test_swin_encoder_vs_oracle(H, W):
    """Whole Swin-B encoder + projector on the GPU vs the oracle (itself bit-exact to the reference's SwinTransformer).
    (75, 118) exercises patch-embed padding and odd patch merges."""
    from aot_benchmark_b200 import EngineConfig, build_vos_model, engine, plan
    from oracle import aot_oracle as O
    from oracle import weights as OW
    sd = OW.build_state_dict("swinb_aotl", seed=0)
    cfg = EngineConfig("t", "swinb_aotl")
    model = build_vos_model(cfg.MODEL_VOS, cfg)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    img = torch.randn(1, 3, H, W, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        want = O.encode_image(sd, O.OracleConfig("swinb_aotl"), img)
        enc = engine._Encoder(plan.get_plan(model), H, W)
        st = torch.cuda.current_stream().cuda_stream
        for rep in range(3):            # eager, captured, replayed: all three must agree with the oracle
            got = enc(img.cuda(), st)
            torch.cuda.synchronize()
            for a, b in zip(got, want):
                assert tuple(a.shape) == tuple(b.shape)
                assert (a.cpu() - b).abs().max().item() < 5e-4 * max(1.0, b.abs().max().item()), rep


@pytest.mark.parametrize("name", ["swinb_aotl_small", "swinb_deaotl_small"])
def test_swin_engine_vs_reference_golden(name, golden_dir):
    from oracle import aot_oracle as O
    from oracle import weights as OW
    from test_gpu_engine import _build_cuda_engine, _tie_band_ok
    g = torch.load(os.path.join(golden_dir, f"video_{name}.pt"))
    sd = OW.build_state_dict(g["model"], seed=g["seed"], flavour=g["flavour"])
    assert OW.checksum(sd) == g["weights_checksum"], "seeded weights are not reproducible on this machine"
    frames, mask = O.synthetic_video(g["frames"], g["H"], g["W"], g["objs"], seed=1234 + g["seed"])
    eng = _build_cuda_engine(g["model"], sd, g["gap"])
    forced = [l.float() for l in g["ref_labels"]]
    with torch.no_grad():
        lo, labels = O.run_video(eng, [f.cuda() for f in frames], mask.cuda(), g["objs"], tuple(g["out_size"]),
                                 forced_masks=forced)
    n = g["objs"] + 1
    dmax = max((a.cpu()[:, :n] - b[:, :n]).abs().max().item() for a, b in zip(lo, g["ref_logits_lo"]))
    assert dmax < 1e-3, f"max |dlogit| vs reference = {dmax}"   # north-star tolerance (fp32 logits)
    assert _tie_band_ok(lo, g["ref_logits_lo"], labels, g["ref_labels"], tuple(g["out_size"]), n, align=False) == 0
    total = sum(b.numel() for b in g["ref_labels"])
    mism = sum((a.cpu().to(torch.uint8) != b).sum().item() for a, b in zip(labels, g["ref_labels"]))
    assert mism <= 2e-4 * total


def test_full_size_cfg4_geometry_tensor_core_vs_fp32_cuda_core_paths():
    """BASELINE configs[3] geometry (SwinB-AOTL, 592x1040 -> 148x260 / 74x130 / 37x65 maps, 10 objects, gap 5) is too big for
    the CPU oracle in a test: at full size the tensor-core path (tcgen05 GEMMs + attention, CUDA graphs) and the fp32
    CUDA-core path (no graphs) must agree on logits and masks, the run must be deterministic, and the bank must grow as
    the reference schedule says."""
    from aot_benchmark_b200 import engine as engine_mod
    from aot_benchmark_b200 import ops
    from oracle import aot_oracle as O
    from oracle import weights as OW
    from test_gpu_engine import _build_cuda_engine
    sd = OW.build_state_dict("swinb_aotl", seed=11)
    T = 8
    frames, mask = O.synthetic_video(T, 592, 1040, 10, seed=4)
    frames = [f.cuda() for f in frames]
    mask = mask.cuda()
    runs = {}
    for name, (lt, conv, graphs) in {"tc": ("tc_exact", "tc", True), "tc2": ("tc_exact", "tc", True),
                                     "simt": ("simt", "simt", False)}.items():
        old = (engine_mod.LT_IMPL, ops.CONV_IMPL, engine_mod.USE_GRAPHS)
        engine_mod.LT_IMPL, ops.CONV_IMPL, engine_mod.USE_GRAPHS = lt, conv, graphs
        try:
            eng = _build_cuda_engine("swinb_aotl", sd, 5)
            with torch.no_grad():
                lo, labels = O.run_video(eng, frames, mask, 10, (480, 854),
                                         forced_masks=runs["tc"][1] if name != "tc" else None)
            runs[name] = (lo, labels, eng.aot_engines[0].bank_len, eng.aot_engines[0].enc_hw)
        finally:
            engine_mod.LT_IMPL, ops.CONV_IMPL, engine_mod.USE_GRAPHS = old
    lo, labels, bank_len, N = runs["tc"]
    assert N == 37 * 65 == 2405
    assert bank_len == N * (1 + (T - 1) // 5)
    for a, b in zip(lo, runs["tc2"][0]):
        assert torch.equal(a, b)                                    # run-to-run determinism
    dmax = max((a[:, :11] - b[:, :11]).abs().max().item() for a, b in zip(lo, runs["simt"][0]))
    assert dmax < 1e-3, dmax
    # label differences between the two arithmetic paths are only legitimate inside the tie band
    from test_gpu_engine import _tie_band_ok
    assert _tie_band_ok(lo, [t.cpu() for t in runs["simt"][0]], labels, runs["simt"][1], (480, 854), 11, align=False) == 0
    mism = sum((a != b).sum().item() for a, b in zip(labels, runs["simt"][1]))
    total = sum(a.numel() for a in labels)
    assert mism <= 1e-3 * total, (mism, total)
