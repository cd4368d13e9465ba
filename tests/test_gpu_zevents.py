"""GPU: the multi-engine facade (> 10 objects, soft_logit_aggregation) and objects that first appear mid-video
(a second reference frame with live memory) through the drop-in AOTInferEngine, against the committed golden of the
real reference's AOTInferEngine (oracle/gen_golden.py::events_case; aot_engine.py:485-635, evaluator.py:362-402)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["aott_multi14_events", "deaott_multi14_events"])
def test_multi_engine_and_new_objects_vs_reference_golden(golden_dir, case):
    from oracle import aot_oracle as O
    from oracle import weights as OW
    from test_gpu_engine import _build_cuda_engine
    g = torch.load(os.path.join(golden_dir, f"events_{case}.pt"))
    sd = OW.build_state_dict(g["model"], seed=g["seed"])
    assert OW.checksum(sd) == g["weights_checksum"]
    frames, full = O.synthetic_video(g["frames"], g["H"], g["W"], 14, seed=g["video_seed"])
    first = torch.where(full <= g["first_objs"], full, torch.zeros_like(full))
    new = {g["event_frame"]: g["new_label"].float().cuda()}
    eng = _build_cuda_engine(g["model"], sd, g["gap"])
    with torch.no_grad():
        lo, _ = O.run_video_events(eng, [f.cuda() for f in frames], first.cuda(), g["first_objs"],
                                   tuple(g["out_size"]), new_objects=new,
                                   forced_masks=[l.float() for l in g["ref_labels"]])
    torch.cuda.synchronize()
    assert len(eng.aot_engines) == 2
    for t, (a, b, n) in enumerate(zip(lo, g["ref_logits"], g["live_channels"])):
        assert a.shape[1] >= n
        d = (a.cpu()[:, :n] - b).abs().max().item()
        assert d < 1e-3, f"frame {t + 1}: max |dlogit| vs reference = {d}"      # north-star tolerance


@pytest.mark.parametrize("name", ["aott_skip2", "deaott_skip3"])
def test_short_term_skip_vs_reference_golden(name, golden_dir):
    """TEST_SHORT_TERM_MEM_SKIP > 1 (ring of previous frames, aot_engine.py:329-332) against the real reference."""
    from oracle import aot_oracle as O
    from oracle import weights as OW
    from aot_benchmark_b200 import EngineConfig, build_engine, build_vos_model
    g = torch.load(os.path.join(golden_dir, f"video_{name}.pt"))
    sd = OW.build_state_dict(g["model"], seed=g["seed"], flavour=g["flavour"])
    assert OW.checksum(sd) == g["weights_checksum"]
    frames, mask = O.synthetic_video(g["frames"], g["H"], g["W"], g["objs"], seed=1234 + g["seed"])
    cfg = EngineConfig("t", g["model"])
    model = build_vos_model(cfg.MODEL_VOS, cfg)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    eng = build_engine(cfg.MODEL_ENGINE, phase="eval", aot_model=model, gpu_id=0, long_term_mem_gap=g["gap"],
                       short_term_mem_skip=g["skip"])
    eng.eval()
    with torch.no_grad():
        lo, _ = O.run_video(eng, [f.cuda() for f in frames], mask.cuda(), g["objs"], tuple(g["out_size"]),
                            forced_masks=[l.float() for l in g["ref_labels"]])
    n = g["objs"] + 1
    dmax = max((a.cpu()[:, :n] - b[:, :n]).abs().max().item() for a, b in zip(lo, g["ref_logits_lo"]))
    assert dmax < 1e-3, f"max |dlogit| vs reference = {dmax}"
