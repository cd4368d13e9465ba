"""GPU: the drop-in engines at the BASELINE geometries (configs[1] R50-AOTL and configs[2] R50-DeAOTL at 481x849 -> 480x854,
10 objects, long-term gap 5; configs[3] SwinB-AOTL at 592x1040) against goldens produced by the REAL reference at those sizes
(oracle/gen_golden.py --only full; the reference ran them on CPU in the build container).  Teacher-forced with the
reference's own label maps; logits within the north-star 1e-3, label mismatches only inside the reference's tie band."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = ["r50_aotl_480p", "r50_deaotl_480p", "swinb_aotl_592"]


def _engine(model_name, sd, gap):
    from aot_benchmark_b200 import EngineConfig, build_engine, build_vos_model
    cfg = EngineConfig("t", model_name)
    model = build_vos_model(cfg.MODEL_VOS, cfg)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    eng = build_engine(cfg.MODEL_ENGINE, phase="eval", aot_model=model, gpu_id=0, long_term_mem_gap=gap,
                       short_term_mem_skip=cfg.TEST_SHORT_TERM_MEM_SKIP)
    return eng.eval(), cfg


@pytest.mark.parametrize("name", CASES)
def test_full_geometry_vs_reference_golden(name, golden_dir):
    from oracle import aot_oracle as O
    from oracle import weights as OW
    from oracle.fixtures import load_full_labels
    g = torch.load(os.path.join(golden_dir, f"full_{name}.pt"))
    sd = OW.build_state_dict(g["model"], seed=g["seed"], flavour=g["flavour"])
    assert OW.checksum(sd) == g["weights_checksum"], "seeded weights are not reproducible on this machine"
    frames, mask = O.synthetic_video(g["frames"], g["H"], g["W"], g["objs"], seed=1234 + g["seed"])
    eng, cfg = _engine(g["model"], sd, g["gap"])
    ref_labels = load_full_labels(g)
    with torch.no_grad():
        lo, labels = O.run_video(eng, [f.cuda() for f in frames], mask.cuda(), g["objs"], tuple(g["out_size"]),
                                 forced_masks=ref_labels)
    e0 = eng.aot_engines[0]
    N = e0.enc_hw
    assert e0.bank_len == N * (1 + (g["frames"] - 1) // g["gap"])          # the bank grew (aot_engine.py:334-338)
    n = g["objs"] + 1
    dmax, bad = 0.0, 0
    for t in g["logit_frames"]:
        a, b = lo[t - 1].cpu()[:, :n], g["ref_logits_lo"][t][:, :n]
        d = (a - b).abs().max().item()
        dmax = max(dmax, d)
        mm = labels[t - 1].cpu().to(torch.uint8) != ref_labels[t - 1].to(torch.uint8)
        if mm.any():                                                        # tie band
            up = F.interpolate(b, size=tuple(g["out_size"]), mode="bilinear", align_corners=cfg.MODEL_ALIGN_CORNERS)
            top2 = up.topk(2, dim=1).values
            margin = (top2[:, 0] - top2[:, 1]).unsqueeze(1)
            bad += int((mm & (margin > 4 * d + 1e-5)).sum().item())
    print(f"{name}: max |dlogit| vs the real reference = {dmax:.3e}")
    assert dmax < 1e-3, dmax
    assert bad == 0
    total = sum(b.numel() for b in ref_labels)
    mism = sum((a.cpu().to(torch.uint8) != b.to(torch.uint8)).sum().item() for a, b in zip(labels, ref_labels))
    assert mism <= 2e-4 * total, (mism, total)
