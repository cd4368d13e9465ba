"""CPU: host-side logic of the Swin-B encoder path (BASELINE config 4) -- state_dict contract,
weight packing (plan.py) and kernel orchestration (engine._Encoder._swin) -- checked against the oracle with the CUDA
entry points replaced by contract emulations (tests/emu_ops.py).  The kernels themselves are checked on the GPU
(tests/test_gpu_window.py)."""
import pytest
import torch

from aot_benchmark_b200 import EngineConfig, build_vos_model
from oracle import aot_oracle as O


def _swin_sd(seed=0):
    from oracle import weights as OW
    return OW.build_state_dict("swinb_aotl", seed=seed)


def test_swin_state_dict_contract():
    cfg = EngineConfig("t", "swinb_aotl")
    sd = build_vos_model(cfg.MODEL_VOS, cfg).state_dict()
    # SURVEY Appendix F / swin_transformer.py:571-640 names, spot-checked with shapes
    assert sd["encoder.patch_embed.proj.weight"].shape == (128, 3, 4, 4)
    assert sd["encoder.layers.2.blocks.17.attn.qkv.weight"].shape == (1536, 512)
    assert sd["encoder.layers.0.blocks.1.attn.relative_position_bias_table"].shape == (169, 4)
    assert sd["encoder.layers.1.downsample.reduction.weight"].shape == (512, 1024)
    assert "encoder.layers.2.downsample.reduction.weight" not in sd
    assert sd["encoder.layers.0.blocks.0.attn.relative_position_index"].dtype == torch.int64
    assert sd["encoder.norm2.weight"].shape == (512,) and sd["encoder_projector.weight"].shape == (256, 512, 1, 1)
    assert sd["patch_wise_id_bank.weight"].shape == (256, 11, 16, 16)      # align_corners False: k16 s16 p0 (aot.py:58-63)
    assert torch.equal(sd["encoder.layers.0.blocks.0.attn.relative_position_index"], O.swin_rel_index(7))


@pytest.mark.parametrize("hw", [(64, 96), (75, 118)])
def test_swin_encoder_orchestration_matches_oracle(monkeypatch, hw):
    """plan._swin + engine._Encoder._swin driven through the kernel-contract emulations == oracle.swin_forward.
    (64, 96): 16x24 / 8x12 / 4x6 maps (window padding at every stage, shifted masks);
    (75, 118): patch-embed padding, odd maps in both patch merges."""
    from aot_benchmark_b200 import engine, ops, plan
    import emu_ops
    emu_ops.install(monkeypatch, ops)
    monkeypatch.setattr(plan.Plan, "_require_cuda", staticmethod(lambda dev: None))
    monkeypatch.setattr(engine, "_cur_stream", lambda: 0)
    monkeypatch.setattr(engine, "USE_GRAPHS", False)
    cfg = EngineConfig("t", "swinb_aotl")
    model = build_vos_model(cfg.MODEL_VOS, cfg).eval()
    sd = _swin_sd()
    model.load_state_dict(sd)
    P = plan.Plan(model)
    H, W = hw
    img = torch.randn(1, 3, H, W, generator=torch.Generator().manual_seed(3))
    enc = engine._Encoder(P, H, W)
    with torch.no_grad():
        got = enc(img, 0)
        want = O.encode_image(sd, O.OracleConfig("swinb_aotl"), img)
    assert len(got) == 4
    for g, w in zip(got, want):
        assert tuple(g.shape) == tuple(w.shape)
        assert (g - w).abs().max().item() < 2e-4 * max(1.0, w.abs().max().item())
