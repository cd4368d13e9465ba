"""CPU: the tile program of the persistent conv-chain kernel (csrc/conv_chain.cu) -- host-side planner only (no launch):
dependency ranges against a brute-force receptive field, dependencies point backwards in program order, and a discrete
simulation of 148 CTAs walking their round-robin tile lists in order (every role blocks on its own inputs) terminates."""
import numpy as np
import torch

from aot_benchmark_b200 import ops


def _resnet_like_chain(h0=31, w0=45):
    """Three bottleneck stages at reduced size on CPU tensors (pointers only; nothing is launched)."""
    f = lambda *s: torch.zeros(s)
    layers = []
    reg = []

    def conv(x, cin, cout, k=1, stride=1, pad=0, res=None, in_layer=-1, res_layer=-1):
        H, W = x.shape[1], x.shape[2]
        ho, wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        w = torch.zeros(k * k * cin, cout)
        wh, wl = ops.split_fp16(w)
        ops.register_tc_weights(w, wh, wl)
        reg.append(w)
        out = f(1, ho, wo, cout)
        layers.append(dict(x=x, w=w, bias=f(cout), out=out, res=res, KH=k, stride=stride, pad=pad, act=1, in_layer=in_layer,
                           res_layer=res_layer))
        return out, len(layers) - 1

    cur, cur_i = f(1, h0, w0, 64), -1
    for (mid, cout, stride, nblk) in ((64, 256, 1, 2), (128, 512, 2, 2), (256, 1024, 2, 2)):
        for bi in range(nblk):
            s = stride if bi == 0 else 1
            cin = cur.shape[3]
            t1, i1 = conv(cur, cin, mid, in_layer=cur_i)
            t2, i2 = conv(t1, mid, mid, 3, s, 1, in_layer=i1)
            if bi == 0:
                res, ri = conv(cur, cin, cout, 1, s, 0, in_layer=cur_i)
            else:
                res, ri = cur, cur_i
            cur, cur_i = conv(t2, mid, cout, res=res, in_layer=i2, res_layer=ri)
    return layers, reg


def _patch_chk(monkeypatch):
    monkeypatch.setattr(ops, "_chk", lambda *a: None)            # CPU tensors: only their shapes / pointers are used


def test_tile_program_dependencies_cover_the_receptive_field(monkeypatch):
    _patch_chk(monkeypatch)
    layers, _ = _resnet_like_chain()
    tiles, table = ops.conv_chain_dump(layers)
    assert len(table) == len(layers)
    first = {}
    for i, t in enumerate(tiles):
        first.setdefault((t[0], t[1]), i)
    for i, (li, mt, nt, lo, hi, k0, k1, sp) in enumerate(tiles):
        L = layers[li]
        x, out = L["x"], L["out"]
        H, W, Ho, Wo = x.shape[1], x.shape[2], out.shape[1], out.shape[2]
        k, s, p = L.get("KH", 1), L.get("stride", 1), L.get("pad", 0)
        M = Ho * Wo
        need = set()
        for m in range(mt * 128, min(mt * 128 + 128, M)):
            oy, ox = divmod(m, Wo)
            for ky in range(k):
                for kx in range(k):
                    iy, ix = oy * s - p + ky, ox * s - p + kx
                    if 0 <= iy < H and 0 <= ix < W:
                        need.add((iy * W + ix) // 128)
        assert need and min(need) >= lo and max(need) <= hi, (i, li, mt, lo, hi, sorted(need))
        assert hi - lo <= (max(need) - min(need)) + 2 * ((W + 127) // 128 + 1)        # conservative, but not the whole layer
        if L["in_layer"] >= 0:                                                           # dependencies point backwards
            for pm in range(lo, hi + 1):
                assert first[(L["in_layer"], pm)] < i
        if L.get("res") is not None and L["res_layer"] >= 0:
            assert first[(L["res_layer"], mt)] < i
    # split-K work items: consecutive in program order, chunk ranges partition [0, chunks), the last split finishes the tile
    by_tile = {}
    for i, (li, mt, nt, lo, hi, k0, k1, sp) in enumerate(tiles):
        by_tile.setdefault((li, mt, nt), []).append((i, k0, k1, sp))
    assert any(table[li][7] > 1 for li in range(len(layers))), "the 256-channel 3x3 layers at this size should be split"
    for (li, mt, nt), items in by_tile.items():
        S, chunks = table[li][7], table[li][9]
        assert [it[3] for it in items] == list(range(S)) and [it[0] for it in items] == list(range(items[0][0], items[0][0] + S))
        assert items[0][1] == 0 and items[-1][2] == chunks and all(a[2] == b[1] for a, b in zip(items, items[1:]))
        assert all(it[2] > it[1] for it in items)
    # counters: one per (layer, m-tile); targets = n-tiles of the producer
    for li, (M, BN, off, in_off, res_off, in_need, res_need, S, part_off, chunks) in enumerate(table):
        assert BN in (64, 128) and layers[li]["out"].shape[3] % BN == 0
        if layers[li]["in_layer"] >= 0:
            pl = layers[li]["in_layer"]
            assert in_off == table[pl][2] and in_need == layers[pl]["out"].shape[3] // table[pl][1]


def test_round_robin_in_order_execution_terminates(monkeypatch):
    """148 CTAs, tile i on CTA i % 148, each CTA strictly in order; a tile can run when its producer m-tiles (and its residual
    m-tile) have all their n-tiles finished.  Random finishing order among runnable tiles: must always complete."""
    _patch_chk(monkeypatch)
    layers, _ = _resnet_like_chain(45, 61)
    tiles, table = ops.conv_chain_dump(layers)
    n_cta = 148
    rng = np.random.default_rng(0)
    for trial in range(5):
        nxt = list(range(min(n_cta, len(tiles))))                       # next tile index of every CTA
        done = np.zeros(4096, dtype=np.int64)
        finished = 0
        while finished < len(tiles):
            runnable = []
            for c, i in enumerate(nxt):
                if i >= len(tiles):
                    continue
                li, mt, nt, lo, hi, k0, k1, sp = tiles[i]
                M, BN, off, in_off, res_off, in_need, res_need, S, part_off, chunks = table[li]
                ntn = layers[li]["out"].shape[3] // BN
                ok = in_off < 0 or all(done[in_off + m] >= in_need for m in range(lo, hi + 1))
                if sp == S - 1:                  # the finishing item also needs the residual and the other splits' partials
                    ok = ok and (res_off < 0 or done[res_off + mt] >= res_need)
                    ok = ok and (S == 1 or done[part_off + mt * ntn + nt] >= S - 1)
                if ok:
                    runnable.append(c)
            assert runnable, f"deadlock after {finished} tiles"
            c = runnable[rng.integers(len(runnable))]
            li, mt, nt, _, _, _, _, sp = tiles[nxt[c]]
            S, ntn = table[li][7], layers[li]["out"].shape[3] // table[li][1]
            if sp == S - 1:
                done[table[li][2] + mt] += 1
            else:
                done[table[li][8] + mt * ntn + nt] += 1
            nxt[c] += n_cta
            finished += 1
