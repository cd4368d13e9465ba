"""CPU: host-side logic of the drop-in engines -- weight packing (plan.py: FrozenBN folding, fused projections, ID-bank
prefix sums), buffer wiring, call order, memory-bank bookkeeping, short-term ring, multi-engine facade -- checked against
the oracle (pinned to the real reference by tests/golden) with every C-ABI entry point replaced by a torch-CPU emulation of
its documented contract (tests/emu_ops.py).  The CUDA kernels themselves are checked on the GPU (tests/test_gpu_*.py);
this suite is what lets the engine be refactored without a GPU at hand."""
import os

import pytest
import torch

from oracle import aot_oracle as O
from oracle import weights as OW


def _engine(model_name, sd, gap, skip=None):
    from aot_benchmark_b200 import EngineConfig, build_engine, build_vos_model
    cfg = EngineConfig("t", model_name)
    model = build_vos_model(cfg.MODEL_VOS, cfg).eval()
    model.load_state_dict(sd, strict=True)
    eng = build_engine(cfg.MODEL_ENGINE, phase="eval", aot_model=model, gpu_id=0, long_term_mem_gap=gap,
                       short_term_mem_skip=cfg.TEST_SHORT_TERM_MEM_SKIP if skip is None else skip)
    eng.eval()
    return eng


_GOLDEN_CLIPS = ["aott_raw_257", "r50_aotl_small", "r50_deaotl_small", "deaott_small", "swinb_aotl_small", "swinb_deaotl_small",
                 "aott_skip2", "deaott_skip3"]
# the fp32 CUDA-core long-term attention (LT_IMPL=simt) is a separate orchestration only for the AOT models; Swin clips run once
_SIMT_CLIPS = ["aott_raw_257", "r50_aotl_small", "aott_skip2"]


@pytest.mark.parametrize("name,lt_impl", [(n, "tc_exact") for n in _GOLDEN_CLIPS] + [(n, "simt") for n in _SIMT_CLIPS])
def test_engine_orchestration_vs_reference_golden(monkeypatch, golden_dir, name, lt_impl):
    import emu_ops
    from aot_benchmark_b200 import engine
    emu_ops.install_engine(monkeypatch)
    monkeypatch.setattr(engine, "LT_IMPL", lt_impl)
    g = torch.load(os.path.join(golden_dir, f"video_{name}.pt"))
    T = min(g["frames"], 3 if g["model"].startswith("swinb") else 6)
    sd = OW.build_state_dict(g["model"], seed=g["seed"], flavour=g["flavour"])
    frames, mask = O.synthetic_video(g["frames"], g["H"], g["W"], g["objs"], seed=1234 + g["seed"])
    eng = _engine(g["model"], sd, g["gap"], g.get("skip"))
    forced = [l.float() for l in g["ref_labels"]]
    with torch.no_grad():
        lo, labels = O.run_video(eng, frames[:T], mask, g["objs"], tuple(g["out_size"]), forced_masks=forced)
    n = g["objs"] + 1
    dmax = max((a[:, :n] - b[:, :n]).abs().max().item() for a, b in zip(lo, g["ref_logits_lo"]))
    assert dmax < 2e-4, f"max |dlogit| vs reference = {dmax}"
    e0 = eng.aot_engines[0]
    assert e0.bank_len == e0.enc_hw * (1 + (T - 1) // g["gap"])          # reference frame + every gap-th frame


@pytest.mark.parametrize("case", ["aott_multi14_events", "deaott_multi14_events"])
def test_multi_engine_and_new_objects_orchestration(monkeypatch, golden_dir, case):
    import emu_ops
    emu_ops.install_engine(monkeypatch)
    g = torch.load(os.path.join(golden_dir, f"events_{case}.pt"))
    sd = OW.build_state_dict(g["model"], seed=g["seed"])
    frames, full = O.synthetic_video(g["frames"], g["H"], g["W"], 14, seed=g["video_seed"])
    first = torch.where(full <= g["first_objs"], full, torch.zeros_like(full))
    eng = _engine(g["model"], sd, g["gap"])
    with torch.no_grad():
        lo, _ = O.run_video_events(eng, frames, first, g["first_objs"], tuple(g["out_size"]),
                                   new_objects={g["event_frame"]: g["new_label"].float()},
                                   forced_masks=[l.float() for l in g["ref_labels"]])
    assert len(eng.aot_engines) == 2
    for a, b, n in zip(lo, g["ref_logits"], g["live_channels"]):
        assert (a[:, :n] - b).abs().max().item() < 2e-4


@pytest.mark.parametrize("name", ["r50_deaotl_small", "deaott_skip3"])
def test_deaot_gemm_long_term_attention_orchestration(monkeypatch, golden_dir, name):
    """AOTB_DEAOT_LT=gemm: Q K^T GEMM -> row softmax over the live keys -> P V GEMM over split-fp16 copies of the bank
    (kept up to date at append time, including a capacity growth) gives the reference's logits."""
    import emu_ops
    from aot_benchmark_b200 import engine
    emu_ops.install_engine(monkeypatch)
    monkeypatch.setattr(engine, "DEAOT_LT", "gemm")
    monkeypatch.setattr(engine, "BANK_INIT_FRAMES", 1)           # force bank re-allocations mid-clip
    g = torch.load(os.path.join(golden_dir, f"video_{name}.pt"))
    sd = OW.build_state_dict(g["model"], seed=g["seed"], flavour=g["flavour"])
    frames, mask = O.synthetic_video(g["frames"], g["H"], g["W"], g["objs"], seed=1234 + g["seed"])
    eng = _engine(g["model"], sd, g["gap"], g.get("skip"))
    with torch.no_grad():
        lo, _ = O.run_video(eng, frames, mask, g["objs"], tuple(g["out_size"]),
                            forced_masks=[l.float() for l in g["ref_labels"]])
    e0 = eng.aot_engines[0]
    assert e0._gemm_lt and e0.bank_len > e0.enc_hw                 # grew past the initial capacity
    n = g["objs"] + 1
    dmax = max((a[:, :n] - b[:, :n]).abs().max().item() for a, b in zip(lo, g["ref_logits_lo"]))
    assert dmax < 2e-4, f"max |dlogit| vs reference = {dmax}"


@pytest.mark.parametrize("model_name,H,W,objs", [("aott", 100, 150, 0), ("aott", 83, 61, 1), ("deaott", 70, 95, 10),
                                                 ("aotb", 97, 113, 4), ("deaotl", 81, 129, 3)])
def test_engine_edge_cases_vs_oracle(monkeypatch, model_name, H, W, objs):
    """No objects at all, a single object on a tiny odd-sized frame, the maximum object count of one engine, and the
    multi-layer MobileNetV2 configurations (AOT-B: 3 LSTT layers; DeAOT-L: 3 GPM layers): the product engine (emulated
    entry points) against the oracle on the same seeded inputs, free-running (no teacher forcing)."""
    import emu_ops
    emu_ops.install_engine(monkeypatch)
    sd = OW.build_state_dict(model_name, seed=5)
    frames, mask = O.synthetic_video(4, H, W, objs, seed=17)
    oe = O.OracleEngine(sd, O.OracleConfig(model_name), long_term_mem_gap=2)
    eng = _engine(model_name, sd, 2)
    with torch.no_grad():
        o_lo, o_labels = O.run_video(oe, frames, mask, objs, (H, W))
        c_lo, c_labels = O.run_video(eng, frames, mask, objs, (H, W), forced_masks=o_labels)
    n = objs + 1
    for a, b in zip(c_lo, o_lo):
        assert (a[:, :n] - b[:, :n]).abs().max().item() < 2e-4
    if objs == 0:
        assert all(int(l.max()) == 0 for l in c_labels)          # nothing but background can be predicted


@pytest.mark.parametrize("name", ["r50_deaotl_small", "deaott_skip3"])
def test_deaot_fused_tc_long_term_attention_orchestration(monkeypatch, golden_dir, name):
    """AOTB_DEAOT_LT=tc: the packed-operand bank copies ([C/32][cap][64] split-fp16 rows), the Q packing with the 1/T
    division, the KV-split policy + merge and the capacity growth around the fused kernel give the reference's logits."""
    import emu_ops
    from aot_benchmark_b200 import engine
    emu_ops.install_engine(monkeypatch)
    monkeypatch.setattr(engine, "DEAOT_LT", "tc")
    monkeypatch.setattr(engine, "BANK_INIT_FRAMES", 1)
    g = torch.load(os.path.join(golden_dir, f"video_{name}.pt"))
    sd = OW.build_state_dict(g["model"], seed=g["seed"], flavour=g["flavour"])
    frames, mask = O.synthetic_video(g["frames"], g["H"], g["W"], g["objs"], seed=1234 + g["seed"])
    eng = _engine(g["model"], sd, g["gap"], g.get("skip"))
    with torch.no_grad():
        lo, _ = O.run_video(eng, frames, mask, g["objs"], tuple(g["out_size"]),
                            forced_masks=[l.float() for l in g["ref_labels"]])
    e0 = eng.aot_engines[0]
    assert e0._gp_tc and e0.bank_len > e0.enc_hw
    n = g["objs"] + 1
    dmax = max((a[:, :n] - b[:, :n]).abs().max().item() for a, b in zip(lo, g["ref_logits_lo"]))
    assert dmax < 2e-4, f"max |dlogit| vs reference = {dmax}"
