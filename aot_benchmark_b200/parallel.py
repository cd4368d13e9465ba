"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

Two ways the path shards:

1. video-level data parallelism -- whole clips are independent units (the reference's scheme:
   tools/eval.py:100-106, networks/managers/evaluator.py:216-235).  No collective on the data path;
   `partition_videos` is the static equivalent of the reference's work queue, `reduce_max_ms` is the
   only exchange (timing).
2. split-KV long-term attention (BASELINE config 4) -- the memory bank is sharded by memory frame
   round-robin over ranks (`frame_owner`); every rank computes un-normalised partials (row max m, row sum l,
   O) over its shard with aotb_attention_f32 / aotb_lt_attn_tc_f16x2, `gather_partials` all-gathers them and
   aotb_attn_merge_f32 performs the exact log-sum-exp merge locally (one collective per layer).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def partition_videos(num_videos: int, rank: int, world: int) -> List[int]:
    """Static round-robin assignment of clip indices to ranks."""
    return list(range(rank, num_videos, world))


def frame_owner(mem_frame_idx: int, world: int) -> int:
    """Rank that keeps memory frame `mem_frame_idx` of a sharded long-term bank (round-robin)."""
    return mem_frame_idx % world


def local_frames(num_mem_frames: int, rank: int, world: int) -> List[int]:
    return [f for f in range(num_mem_frames) if frame_owner(f, world) == rank]


def gather_partials(dist, Opart: torch.Tensor, Mpart: torch.Tensor, Lpart: torch.Tensor
                    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """All-gather one rank's split-KV partial (O [N,C], M [H,N], L [H,N]) -> stacked [R,...] tensors in rank
    order, ready for ops.attn_merge.  Payload per layer at N=2405: 2.46 MB + 2*77 KB per rank."""
    world = dist.get_world_size()
    outs = []
    for t in (Opart, Mpart, Lpart):
        buf = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(buf, t.contiguous())
        outs.append(torch.stack(buf, dim=0))
    return outs[0], outs[1], outs[2]


def reduce_max_ms(dist, ms: float, device) -> float:
    """Device time of a multi-rank step = max over ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
