#!/usr/bin/env python
"""Uninitialised-read hunt: fill the caching allocator's free blocks with a poison pattern, then run the same short clip
through the engine eagerly and through captured CUDA graphs and compare bit for bit (tests/test_gpu_engine.py::
test_cuda_graph_replay_matches_eager is this comparison without the poison).  A kernel that reads a padding row, a workspace
tail or a stale partial shows up as NaN in, or a difference between, the runs.

    python scripts/poison_check.py [nan|big|zero] [model]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aot_benchmark_b200 import EngineConfig, build_engine, build_vos_model  # noqa: E402
from aot_benchmark_b200 import engine as engine_mod  # noqa: E402
from oracle import aot_oracle as O  # noqa: E402  (test infrastructure: synthetic clip + the evaluator's loop)
from oracle import weights as OW  # noqa: E402


def poison(kind, gib=24):
    """Allocate `gib` GiB in blocks of several sizes, fill them, free them: later torch.empty() calls get the pattern."""
    val = {"nan": float("nan"), "big": 3.0e4, "zero": 0.0}[kind]
    blocks = []
    for mb in (2048, 512, 64, 8, 1):
        n = max(1, int(gib * 1024 / 5 / mb))
        for _ in range(min(n, 400)):
            blocks.append(torch.full((mb * 262144,), val, device="cuda"))
    for kb in (512, 64, 4):
        for _ in range(400):
            blocks.append(torch.full((kb * 256,), val, device="cuda"))
    torch.cuda.synchronize()
    del blocks


def build(model_name, sd, gap):
    cfg = EngineConfig("t", model_name)
    model = build_vos_model(cfg.MODEL_VOS, cfg)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    eng = build_engine(cfg.MODEL_ENGINE, phase="eval", aot_model=model, gpu_id=0, long_term_mem_gap=gap,
                       short_term_mem_skip=cfg.TEST_SHORT_TERM_MEM_SKIP)
    eng.eval()
    return eng


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "nan"
    model_name = sys.argv[2] if len(sys.argv) > 2 else "r50_aotl"
    sd = OW.build_state_dict(model_name, seed=7)
    frames, mask = O.synthetic_video(10, 161, 241, 6, seed=21)
    frames = [f.cuda() for f in frames]
    mask = mask.cuda()
    res = {}
    for use in (False, True):
        engine_mod.USE_GRAPHS = use
        poison(kind)
        eng = build(model_name, sd, 2)
        outs = []
        for rep in range(2):
            poison(kind, gib=8)
            with torch.no_grad():
                lo, labels = O.run_video(eng, frames, mask, 6, (160, 240))
            outs.append(([t.clone() for t in lo], labels))
        res[use] = outs
    bad = 0
    for name, A, B in (("eager vs graph, video 1", res[False][0], res[True][0]), ("eager vs graph, video 2", res[False][1], res[True][1]),
                       ("eager video 1 vs 2", res[False][0], res[False][1]), ("graph video 1 vs 2", res[True][0], res[True][1])):
        for f, (a, b) in enumerate(zip(A[0], B[0])):
            nan = int(torch.isnan(a).sum() + torch.isnan(b).sum())
            d = (a - b).abs().max().item() if not nan else float("nan")
            if nan or d != 0.0:
                bad += 1
                print(f"{kind} {model_name}: {name}: frame {f + 1}: max|dlogit| = {d}, NaNs = {nan}")
    print(f"{kind} {model_name}: {'CLEAN' if bad == 0 else str(bad) + ' differing frames'}")


if __name__ == "__main__":
    main()
