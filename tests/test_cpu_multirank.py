"""CPU, world_size 2, gloo: the host-side logic of the N > 1 paths (video partition, sharded-bank
ownership, partial gather + exact LSE merge)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, REPO)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from aot_benchmark_b200 import parallel as P
    torch.manual_seed(0)                      # replicated inputs on every rank
    H, d, N, frames, per = 8, 32, 50, 5, 40
    Q = torch.randn(N, H * d) * 3
    K = torch.randn(frames * per, H * d)
    V = torch.randn(frames * per, H * d)
    mine = P.local_frames(frames, rank, world)
    rows = torch.cat([torch.arange(f * per, (f + 1) * per) for f in mine])
    q = (Q / d ** 0.5).view(N, H, d).permute(1, 0, 2)
    k = K[rows].view(-1, H, d).permute(1, 2, 0)
    v = V[rows].view(-1, H, d).permute(1, 0, 2)
    s = q @ k
    m = s.max(-1).values                       # [H, N]
    p = torch.exp(s - m.unsqueeze(-1))
    l = p.sum(-1)
    o = (p @ v).permute(1, 0, 2).reshape(N, H * d)
    Og, Mg, Lg = P.gather_partials(dist, o, m, l)
    mm = Mg.max(0).values
    w = torch.exp(Mg - mm)                     # [R, H, N]
    num = (w.permute(0, 2, 1).unsqueeze(-1) * Og.view(world, N, H, d)).sum(0)
    den = (w * Lg).sum(0).permute(1, 0).unsqueeze(-1)
    merged = (num / den).reshape(N, H * d)
    full = torch.softmax(q @ K.view(-1, H, d).permute(1, 2, 0), -1) @ V.view(-1, H, d).permute(1, 0, 2)
    full = full.permute(1, 0, 2).reshape(N, H * d)
    err = (merged - full).abs().max().item()
    t = P.reduce_max_ms(dist, 10.0 + rank, torch.device("cpu"))
    vids = P.partition_videos(7, rank, world)
    ret[rank] = (err, t, vids, mine)
    dist.destroy_process_group()


def test_two_rank_sharded_bank_merge_and_partition():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29517 + os.getpid() % 200, ret), nprocs=world, join=True)
    assert len(ret) == 2
    for rank in range(world):
        err, t, vids, mine = ret[rank]
        assert err < 1e-5
        assert abs(t - 11.0) < 1e-9
    assert sorted(ret[0][2] + ret[1][2]) == list(range(7))
    assert sorted(ret[0][3] + ret[1][3]) == list(range(5))
