"""CPU, world_size 2, gloo: the host-side logic of the N > 1 paths (video partition, sharded-bank
ownership, partial gather + exact LSE merge)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class P:
    """Reference formulation of the two ways the path shards, used only to check the arithmetic of the exact merge
    under gloo: round-robin ownership of clips / memory frames, all-gather of (O, m, l), max over ranks of the device time."""

    @staticmethod
    def partition_videos(num_videos, rank, world):
        return list(range(rank, num_videos, world))

    @staticmethod
    def local_frames(num_mem_frames, rank, world):
        return [f for f in range(num_mem_frames) if f % world == rank]

    @staticmethod
    def gather_partials(dist, Opart, Mpart, Lpart):
        world = dist.get_world_size()
        outs = []
        for t in (Opart, Mpart, Lpart):
            buf = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(buf, t.contiguous())
            outs.append(torch.stack(buf, dim=0))
        return outs

    @staticmethod
    def reduce_max_ms(dist, ms, device):
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())


def _worker(rank, world, port, ret):
    sys.path.insert(0, REPO)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
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


class _MP:
    """pytest's monkeypatch is not available inside spawned workers: plain setattr is enough there."""
    def setattr(self, obj, name, val):
        setattr(obj, name, val)


def _engine_worker(rank, world, port, ret):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import emu_ops
    from aot_benchmark_b200 import EngineConfig, build_engine, build_vos_model
    from oracle import aot_oracle as O
    from oracle import weights as OW
    import test_cpu_graph_static as gs
    gs._install(_MP())            # emulated entry points + a tracer in place of GraphCache: "replays" must repeat the captured launches
    name, H, W, objs, T = "aott", 65, 81, 2, 6
    sd = OW.build_state_dict(name, seed=1)
    cfg = EngineConfig("t", name)
    model = build_vos_model(cfg.MODEL_VOS, cfg).eval()
    model.load_state_dict(sd)
    frames, mask = O.synthetic_video(T, H, W, objs, seed=3)
    outs = {}
    for mode in ("plain", "sharded"):
        eng = build_engine(cfg.MODEL_ENGINE, phase="eval", aot_model=model, gpu_id=0, long_term_mem_gap=1)
        if mode == "sharded":
            eng.enable_kv_sharding(rank, world)
        for video in range(2 if mode == "sharded" else 1):       # second video: kept workspace, kept "graphs", bank restarted
            with torch.no_grad():
                lo, labels = O.run_video(eng, frames, mask, objs, (H, W),
                                         forced_masks=outs["plain"][1] if mode == "sharded" else None)
            if video == 1:
                assert all(torch.equal(a, b) for a, b in zip(lo, outs[mode][0]))
            outs[mode] = (lo, labels)
        if mode == "sharded":
            e0 = eng.aot_engines[0]
            local_rows, mem_frames, n = e0.bank_len, e0._mem_frames, e0.enc_hw
    d = max((a[:, :objs + 1] - b[:, :objs + 1]).abs().max().item() for a, b in zip(outs["plain"][0], outs["sharded"][0]))
    ret[rank] = (d, local_rows, mem_frames, n, gs.TracingGraphCache.replays)
    dist.destroy_process_group()


def test_two_rank_sharded_engine_matches_unsharded():
    """BASELINE configs[3] mechanism end to end on CPU: two ranks run the drop-in engine (C-ABI entry points emulated) with
    the long-term bank sharded by memory frame; the all-gather + exact merge must reproduce the unsharded logits, and
    the memory frames must be split between the ranks."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_engine_worker, args=(world, 29717 + os.getpid() % 200, ret), nprocs=world, join=True)
    assert len(ret) == 2
    rows = 0
    for rank in range(world):
        d, local_rows, mem_frames, n, replays = ret[rank]
        assert replays > 10                          # the sharded LSTT / update / decode bodies were "replayed" and matched
        assert d < 1e-4, f"rank {rank}: sharded vs unsharded max |dlogit| = {d}"
        assert mem_frames == 6                       # reference frame + 5 propagated frames at gap 1
        rows += local_rows
    assert rows == 6 * ret[0][3]                      # every memory frame lives on exactly one rank
    assert ret[0][1] == ret[1][1] == 3 * ret[0][3]    # round-robin: three frames each


def test_sharded_engine_peer_memory_exchange(monkeypatch):
    """AOTB_SHARD_XCHG=p2p: the partials live in per-rank buffers that every rank's merge reads in place behind a
    barrier (torch symmetric memory + aotb_attn_merge_peers_f32 on GPUs).  Here: two ranks as two threads of one process, the
    allocator replaced by an in-process stand-in (shared tensors + threading.Barrier), entry points emulated -- the
    sharded logits must equal the unsharded ones, for a 1-layer (second barrier) and a 3-layer model."""
    import threading
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import emu_ops
    from aot_benchmark_b200 import EngineConfig, build_engine, build_vos_model, engine
    from oracle import aot_oracle as O
    from oracle import weights as OW
    emu_ops.install_engine(monkeypatch)
    monkeypatch.setattr(engine, "SHARD_XCHG", "p2p")
    world = 2

    class Registry:
        def __init__(self):
            self.lock, self.bufs, self.count = threading.Lock(), {}, {}
            self.bar = threading.Barrier(world)

    class Group:                       # what enable_kv_sharding receives as `group`
        def __init__(self, rank, reg):
            self.rank, self.reg = rank, reg

    class Handle:
        def __init__(self, reg):
            self.reg = reg

        def barrier(self, channel=0):
            self.reg.bar.wait(timeout=120)

    def fake_alloc(numel, device, group):
        reg = group.reg
        with reg.lock:
            k = reg.count.get(group.rank, 0)          # the k-th allocation of this rank pairs with the k-th of its peers
            reg.count[group.rank] = k + 1
            for r in range(world):
                reg.bufs.setdefault((k, r), torch.zeros(numel))
        reg.bar.wait(timeout=120)                     # rendezvous
        return Handle(reg), (lambda r, sizes, off: reg.bufs[(k, r)][off:off + int(torch.Size(sizes).numel())].view(sizes))

    monkeypatch.setattr(engine, "_symm_alloc", fake_alloc)
    for name in ("aott", "aotb"):
        H, W, objs, T = 65, 81, 2, 6
        sd = OW.build_state_dict(name, seed=1)
        cfg = EngineConfig("t", name)
        model = build_vos_model(cfg.MODEL_VOS, cfg).eval()
        model.load_state_dict(sd)
        frames, mask = O.synthetic_video(T, H, W, objs, seed=3)
        plain = build_engine(cfg.MODEL_ENGINE, phase="eval", aot_model=model, gpu_id=0, long_term_mem_gap=1)
        with torch.no_grad():
            p_lo, p_labels = O.run_video(plain, frames, mask, objs, (H, W))
        reg, res, errs = Registry(), {}, []

        def worker(rank):
            try:
                eng = build_engine(cfg.MODEL_ENGINE, phase="eval", aot_model=model, gpu_id=0, long_term_mem_gap=1)
                eng.enable_kv_sharding(rank, world, Group(rank, reg))
                with torch.no_grad():
                    lo, _ = O.run_video(eng, frames, mask, objs, (H, W), forced_masks=p_labels)
                res[rank] = (lo, eng.aot_engines[0].bank_len, eng.aot_engines[0].enc_hw)
            except Exception as e:                     # pragma: no cover
                errs.append(repr(e))
                reg.bar.abort()

        ths = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
        for t in ths:
            t.start()
        for t in ths:
            t.join(timeout=600)
        assert not errs, errs
        for rank in range(world):
            lo, rows, n = res[rank]
            d = max((a[:, :objs + 1] - b[:, :objs + 1]).abs().max().item() for a, b in zip(p_lo, lo))
            assert d < 1e-4, f"{name} rank {rank}: peer-memory sharded vs unsharded max |dlogit| = {d}"
            assert rows == 3 * n
