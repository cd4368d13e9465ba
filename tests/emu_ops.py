"""TEST INFRASTRUCTURE ONLY: torch-CPU emulations of a few C-ABI entry points' *contracts* (include/aotb200.h).

There is no GPU in the build container, so the host-side orchestration of a new path (weight packing in plan.py,
buffer wiring and call order in engine.py) is checked here by swapping these emulations in for `aot_benchmark_b200.ops`
and comparing the result with the oracle.  Each emulation follows the index arithmetic of the CUDA kernel it stands
for (not the oracle's formulation), so the test also cross-checks that arithmetic against the reference semantics.
Nothing under aot_benchmark_b200/ imports this module; the product has no CPU path.
"""
import math

import torch
import torch.nn.functional as F


def _act(x, act):
    if act == 1:
        return F.relu(x)
    if act == 2:
        return F.gelu(x)
    if act == 3:
        return F.silu(x)
    if act == 4:
        return F.relu6(x)
    return x


def image_to_nhwc4(img, out, stream=None):
    out.zero_()
    out[..., :3] = img.permute(0, 2, 3, 1)
    return out


def conv2d(x, w, bias, out, res=None, KH=1, KW=1, stride=1, pad=0, dil=1, act=0, stream=None):
    Cin, Cout = x.shape[3], w.shape[1]
    wt = w.view(KH, KW, Cin, Cout).permute(3, 2, 0, 1)
    y = F.conv2d(x.permute(0, 3, 1, 2), wt, bias, stride, pad, dil).permute(0, 2, 3, 1)
    if res is not None:
        y = y + res
    out.copy_(_act(y, act))
    return out


def linear(x, wt, bias, out, res=None, act=0, stream=None):
    y = x @ wt
    if bias is not None:
        y = y + bias
    if res is not None:
        y = y + res
    out.copy_(_act(y, act))
    return out


def layernorm(x, gamma, beta, out, add=None, out2=None, stream=None):
    y = F.layer_norm(x, (x.shape[1],), gamma, beta, 1e-5)
    out.copy_(y)
    if out2 is not None:
        out2.copy_(y + add)
    return out


def window_attention(qkv, qkv_bias, rel_bias, out, H, W, heads, shift, window=7, stream=None):
    """Literal restatement of window_attn_kernel's addressing (csrc/window_attn.cu)."""
    WS, D = window, 32
    C = out.shape[1]
    assert C == heads * D and qkv.shape == (H * W, 3 * C)
    Hp, Wp = -(-H // WS) * WS, -(-W // WS) * WS
    T = WS * WS
    ty, tx = torch.arange(T) // WS, torch.arange(T) % WS
    scale = 0.17677669529663687
    for wy in range(Hp // WS):
        for wx in range(Wp // WS):
            ys, xs = wy * WS + ty, wx * WS + tx
            y, x = (ys + shift) % Hp, (xs + shift) % Wp
            valid = (y < H) & (x < W)
            src = torch.where(valid, y * W + x, torch.zeros_like(y))
            rows = torch.where(valid[:, None], qkv[src], qkv_bias[None, :].expand(T, -1))
            if shift > 0:
                ry = torch.where(ys < Hp - WS, 0, torch.where(ys < Hp - shift, 1, 2))
                rx = torch.where(xs < Wp - WS, 0, torch.where(xs < Wp - shift, 1, 2))
            else:
                ry = rx = torch.zeros(T, dtype=torch.long)
            reg = ry * 3 + rx
            mask = torch.where(reg[:, None] != reg[None, :], -100.0, 0.0)
            for h in range(heads):
                q = rows[:, h * D:(h + 1) * D] * scale
                k = rows[:, C + h * D:C + (h + 1) * D]
                v = rows[:, 2 * C + h * D:2 * C + (h + 1) * D]
                s = q @ k.t() + rel_bias[h] + mask
                o = torch.softmax(s, dim=-1) @ v
                out[src[valid], h * D:(h + 1) * D] = o[valid]
    return out


def patch_merge(x, out, H, W, stream=None):
    C = x.shape[1]
    H2, W2 = (H + 1) // 2, (W + 1) // 2
    out.zero_()
    o = out.view(H2, W2, 4, C)
    xm = x.view(H, W, C)
    for q in range(4):
        dy, dx = q & 1, q >> 1
        sub = xm[dy::2, dx::2]
        o[:sub.shape[0], :sub.shape[1], q] = sub
    return out


def eltwise(op, a, b, out, scalar=0.0, stream=None):
    # op: 0 copy, 1 a+b, 2 a*b, 3 silu(a), 4 silu(a)*b, 5 fill(scalar)
    if op == 0:
        out.copy_(a)
    elif op == 1:
        out.copy_(a + b)
    elif op == 2:
        out.copy_(a * b)
    elif op == 3:
        out.copy_(F.silu(a))
    elif op == 4:
        out.copy_(F.silu(a) * b)
    elif op == 5:
        out.fill_(scalar)
    else:
        raise ValueError(op)
    return out


def nchw_to_nhwc(x, out, stream=None):
    out.copy_(x.permute(0, 2, 3, 1))
    return out


def nhwc_to_nchw(x, out, stream=None):
    out.copy_(x.permute(0, 3, 1, 2))
    return out


def maxpool3x3s2(x, out, stream=None):
    out.copy_(F.max_pool2d(x.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1))
    return out


def dwconv(x, w, bias, out, K=5, stride=1, pad=2, dil=1, act=0, stream=None):
    C = x.shape[3]
    wt = w.view(K, K, C).permute(2, 0, 1).unsqueeze(1)               # [K*K, C] -> [C, 1, K, K]
    y = F.conv2d(x.permute(0, 3, 1, 2), wt, bias, stride, pad, dil, C).permute(0, 2, 3, 1)
    out.copy_(_act(y, act))
    return out


def bilinear(x, out, align_corners, stream=None):
    y = F.interpolate(x.permute(0, 3, 1, 2), size=(out.shape[1], out.shape[2]), mode="bilinear",
                      align_corners=bool(align_corners))
    out.copy_(y.permute(0, 2, 3, 1))
    return out


def groupnorm_workspace(B, G, device):
    return torch.empty(1, dtype=torch.float64, device=device)


def groupnorm(x, gamma, beta, out, G, act, workspace, stream=None):
    y = F.group_norm(x.permute(0, 2, 1), G, gamma, beta, 1e-5).permute(0, 2, 1)      # x [B, P, C]
    out.copy_(_act(y, act))
    return out


def _heads_attention(Q, K, V, H, d_qk, d_v):
    """softmax((Q / sqrt(d_qk)) K^T) V per head -> (O un-normalised? no: normalised O, row max m, row sum l)."""
    N, Tk = Q.shape[0], K.shape[0]
    q = (Q / math.sqrt(d_qk)).view(N, H, d_qk).permute(1, 0, 2)
    k = K.view(Tk, H, d_qk).permute(1, 2, 0)
    v = V.view(Tk, H, d_v).permute(1, 0, 2)
    s = q @ k
    m = s.max(dim=-1).values                                              # [H, N]
    p = torch.exp(s - m.unsqueeze(-1))
    l = p.sum(-1)
    o = p @ v                                                             # [H, N, d_v] un-normalised
    return o, m, l


def attention(Q, K, V, O, H, d_qk, d_v, Tk=None, Tk_dev=None, Mout=None, Lout=None, stream=None):
    tk = int(Tk_dev.item()) if Tk_dev is not None else (K.shape[0] if Tk is None else int(Tk))
    o, m, l = _heads_attention(Q, K[:tk], V[:tk], H, d_qk, d_v)
    if Mout is not None:
        Mout.copy_(m)
        Lout.copy_(l)
    else:
        o = o / l.unsqueeze(-1)
    O.copy_(o.permute(1, 0, 2).reshape(Q.shape[0], H * d_v))
    return O


def attn_merge(Opart, Mpart, Lpart, O, H, d_v, stream=None):
    """Exact log-sum-exp merge of R partials: Opart [R,N,H*d_v] un-normalised, Mpart/Lpart [R,H,N]."""
    R, N = Opart.shape[0], Opart.shape[1]
    m = Mpart.max(dim=0).values                                            # [H, N]
    w = torch.exp(Mpart - m.unsqueeze(0))                                  # [R, H, N]; exp(-inf - m) = 0
    w = torch.where(torch.isfinite(Mpart), w, torch.zeros_like(w))
    l = (w * Lpart).sum(0)
    o = (Opart.view(R, N, H, d_v) * w.permute(0, 2, 1).unsqueeze(-1)).sum(0) / l.t().unsqueeze(-1)
    O.copy_(o.reshape(N, H * d_v))
    return O


def attn_merge_peers(Oparts, Mparts, Lparts, O, splits, H, d_v, stream=None):
    Og = torch.cat([t[:splits] for t in Oparts], dim=0)
    Mg = torch.cat([t[:splits] for t in Mparts], dim=0)
    Lg = torch.cat([t[:splits] for t in Lparts], dim=0)
    return attn_merge(Og, Mg, Lg, O, H, d_v)


def tc_pack_rows(src, dst, row_off=0, div=1.0, row_off_dev=None, stream=None):
    """fp32 [rows, H*32] -> fp16 [H, cap, 64] rows [off, off+rows) as [hi(32) | lo(32)], values / div first."""
    Hh = dst.shape[0]
    off = int(row_off_dev.item()) if row_off_dev is not None else int(row_off)
    rows = src.shape[0]
    x = (src / div if div != 1.0 else src).view(rows, Hh, 32).permute(1, 0, 2)
    hi = x.half()
    lo = (x - hi.float()).half()
    dst[:, off:off + rows, :32] = hi
    dst[:, off:off + rows, 32:] = lo
    return dst


def lt_attention_tc(Qp, Kp, Vp, N, Tk, O=None, Tk_dev=None, splits=1, exact=True, part=None, dbg=None, stream=None,
                    merge=True, variant=None):
    """Packed-operand attention: operands are hi + lo (exact) or hi only (fast); Q was divided by T when packed."""
    tk = int(Tk_dev.item()) if Tk_dev is not None else int(Tk)
    Hh = Qp.shape[0]

    def unpack(P, rows):
        hi, lo = P[:, :rows, :32].float(), P[:, :rows, 32:].float()
        return hi + lo if exact else hi
    q, k, v = unpack(Qp, N), unpack(Kp, tk), (Vp[:, :tk, :32].float() + Vp[:, :tk, 32:].float())
    tiles = (tk + 127) // 128
    per = (tiles + splits - 1) // splits
    parts = []
    for z in range(splits):
        k0, k1 = min(z * per * 128, tk), min((z + 1) * per * 128, tk)
        if k1 > k0:
            s = q @ k[:, k0:k1].transpose(1, 2)
            m = s.max(-1).values
            p = torch.exp(s - m.unsqueeze(-1))
            parts.append((p @ v[:, k0:k1], m, p.sum(-1)))
        else:
            parts.append((torch.zeros(Hh, N, 32), torch.full((Hh, N), float("-inf")), torch.zeros(Hh, N)))
    if splits == 1:
        o, m, l = parts[0]
        O.copy_((o / l.unsqueeze(-1)).permute(1, 0, 2).reshape(N, Hh * 32))
        return O
    Op, Mp, Lp = part
    for z, (o, m, l) in enumerate(parts):
        Op[z].copy_(o.permute(1, 0, 2).reshape(N, Hh * 32))
        Mp[z].copy_(m)
        Lp[z].copy_(l)
    if merge:
        attn_merge(Op, Mp, Lp, O, Hh, 32)
    return O


def _to2d(x, h, w):
    return x.view(h, w, -1).permute(2, 0, 1).unsqueeze(0)


def local_attention(q, k, v, relk_w, relk_b, relv, out, h, w, H, d_att, d_v, stream=None):
    """q,k [hw, H*d_att], v [hw, H*d_v]; relk_w [H*225, d_att], relk_b [H*225], relv [H, d_v, 225] | None."""
    from oracle import aot_oracle as O
    core = O.local_attention(_to2d(q, h, w), _to2d(k, h, w), _to2d(v, h, w), relk_w.view(H * 225, d_att, 1, 1), relk_b,
                             relv, H)
    out.copy_(core.reshape(h * w, H * d_v))
    return out


def local_attention_tile(q, k, v, relk_w, relk_b, relv_t, out, h, w, H, stream=None):
    return local_attention(q, k, v, relk_w, relk_b, relv_t.permute(0, 2, 1).contiguous(), out, h, w, H, 32, 32)


def id_embed(mask, wt, bias, out, C, nid, ksize, stride, pad, ln_gamma=None, ln_beta=None, stream=None):
    """mask [Hm, Wm] float ids; wt [(ky*K+kx)*nid + id, C] -> out [ho*wo, C] (+ LayerNorm for DeAOT)."""
    onehot = (mask.view(1, 1, *mask.shape) == torch.arange(nid, dtype=mask.dtype).view(1, -1, 1, 1)).to(mask.dtype)
    w = wt.view(ksize, ksize, nid, C).permute(3, 2, 0, 1)
    e = F.conv2d(onehot, w, bias, stride, pad)[0].permute(1, 2, 0).reshape(-1, C)
    if ln_gamma is not None:
        e = F.layer_norm(e, (C,), ln_gamma, ln_beta, 1e-5)
    out.copy_(e)
    return out


def id_embed_runs(mask, wp, bias, out, C, nid, ksize, stride, pad, ln_gamma=None, ln_beta=None, stream=None):
    """wp [K, K+1, nid, C] = exclusive prefix sums along kx of the table."""
    wt = (wp[:, 1:] - wp[:, :-1]).reshape(ksize * ksize * nid, C)
    return id_embed(mask, wt, bias, out, C, nid, ksize, stride, pad, ln_gamma, ln_beta)


def logits_postproc(logits_nhwc, lowres_nchw, out_nchw, obj_num, align_corners, stream=None):
    lo = logits_nhwc.reshape(logits_nhwc.shape[-3], logits_nhwc.shape[-2], -1).permute(2, 0, 1).unsqueeze(0).clone()
    lo[:, obj_num + 1:] = -1e10
    lowres_nchw.copy_(lo)
    if out_nchw is not None:
        out_nchw.copy_(F.interpolate(lo, size=tuple(out_nchw.shape[-2:]), mode="bilinear",
                                     align_corners=bool(align_corners)))
    return out_nchw


def logits_argmax(lowres_nchw, label, align_corners, stream=None):
    up = F.interpolate(lowres_nchw.view(1, *lowres_nchw.shape[-3:]), size=tuple(label.shape[-2:]), mode="bilinear",
                       align_corners=bool(align_corners))
    label.copy_(up.argmax(1).to(label.dtype).view(label.shape))
    return label


def nearest_resize(x, out, stream=None):
    out.copy_(F.interpolate(x.view(1, 1, *x.shape[-2:]), size=tuple(out.shape[-2:]), mode="nearest").view(out.shape))
    return out


def bank_append(src, bank, offset, offset_dev=None, stream=None):
    off = int(offset_dev.item()) if offset_dev is not None else int(offset)
    bank[off:off + src.shape[0], :src.shape[1]] = src
    return bank


def counter_add(counter, delta, stream=None):
    counter += int(delta)
    return counter


def linear_tc(x, wh, wl, bias, out, res=None, act=0, stream=None):
    y = x @ (wh.float() + wl.float()).t()
    if bias is not None:
        y = y + bias
    if res is not None:
        y = y + res
    out.copy_(_act(y, act))
    return out


def split_rows(src, hi, lo, row_off=0, row_off_dev=None, stream=None):
    off = int(row_off_dev.item()) if row_off_dev is not None else int(row_off)
    h = src.half()
    hi[off:off + src.shape[0], :src.shape[1]] = h
    lo[off:off + src.shape[0], :src.shape[1]] = (src - h.float()).half()


def split_cols(src, hiT, loT, col_off=0, col_off_dev=None, stream=None):
    off = int(col_off_dev.item()) if col_off_dev is not None else int(col_off)
    h = src.half()
    hiT[:src.shape[1], off:off + src.shape[0]] = h.t()
    loT[:src.shape[1], off:off + src.shape[0]] = (src - h.float()).half().t()


def row_softmax(S, cols, Tk, scale, Tk_dev=None, stream=None):
    live = min(int(Tk_dev.item()) if Tk_dev is not None else int(Tk), cols)
    p = torch.softmax(S[:, :live] * scale, dim=1)
    S[:, :cols] = 0
    S[:, :live] = p
    return S


def gp_attention_tc(Qp, Kp, Vp, N, Tk, O=None, Tk_dev=None, splits=1, exact=True, part=None, stream=None, merge=True):
    """Packed operands, one 'head' per 32 channels; 64-key tiles split over `splits`."""
    tk = int(Tk_dev.item()) if Tk_dev is not None else int(Tk)

    def unpack(P, rows, lo=True):              # [C/32, cap, 64] -> [rows, C]
        x = P[:, :rows, :32].float() + (P[:, :rows, 32:].float() if lo else 0)
        return x.permute(1, 0, 2).reshape(rows, -1)
    q, k, v = unpack(Qp, N, exact), unpack(Kp, tk, exact), unpack(Vp, tk)
    dv = v.shape[1]
    tiles = (tk + 63) // 64
    per = (tiles + splits - 1) // splits
    parts = []
    for z in range(splits):
        k0, k1 = min(z * per * 64, tk), min((z + 1) * per * 64, tk)
        if k1 > k0:
            s = q @ k[k0:k1].t()
            m = s.max(-1).values
            p = torch.exp(s - m.unsqueeze(-1))
            parts.append((p @ v[k0:k1], m, p.sum(-1)))
        else:
            parts.append((torch.zeros(N, dv), torch.full((N,), float("-inf")), torch.zeros(N)))
    if splits == 1:
        o, m, l = parts[0]
        O.copy_(o / l.unsqueeze(-1))
        return O
    Op, Mp, Lp = part
    for z, (o, m, l) in enumerate(parts):
        Op[z].copy_(o)
        Mp[z, 0].copy_(m)
        Lp[z, 0].copy_(l)
    if merge:
        attn_merge(Op, Mp, Lp, O, 1, dv)
    return O


def separate_labels(mask, out, max_obj, stream=None):
    """Contract of aotb_separate_labels_f32: engine e keeps ids (e*max_obj, (e+1)*max_obj], renumbered from 1."""
    for e in range(out.shape[0]):
        lo, hi = e * max_obj + 1, (e + 1) * max_obj
        keep = (mask >= lo) & (mask <= hi)
        out[e].copy_(torch.where(keep, mask - lo + 1, torch.zeros_like(mask)))
    return out


def soft_logit_aggregation(logits, out, max_obj, stream=None):
    """Contract of aotb_soft_logit_aggregation_f32 (aot_engine.py:565-582)."""
    probs = [torch.softmax(l, dim=1) for l in logits]
    bg = torch.ones_like(probs[0][:, 0:1])
    for p in probs:
        bg = bg * p[:, 0:1]
    merged = torch.cat([bg] + [p[:, 1:1 + max_obj] for p in probs], dim=1).clamp(1e-5, 1 - 1e-5)
    out.copy_(torch.log(merged / (1 - merged)))
    return out


def local_gated_tile(q, k, v, relk_w, relk_b, out, h, w, stream=None):
    """Contract of aotb_local_gated_tile_f32 == aotb_local_attention_f32 for the DeAOT head shape."""
    return local_attention(q, k, v, relk_w, relk_b, None, out, h, w, 1, 128, 1024, stream=stream)


EMULATED = ("local_gated_tile", "separate_labels", "soft_logit_aggregation", "image_to_nhwc4", "conv2d", "linear", "layernorm", "window_attention", "patch_merge", "eltwise",
            "nchw_to_nhwc", "nhwc_to_nchw", "maxpool3x3s2", "dwconv", "bilinear", "groupnorm_workspace", "groupnorm",
            "attention", "attn_merge", "attn_merge_peers", "tc_pack_rows", "lt_attention_tc", "local_attention", "local_attention_tile",
            "id_embed", "id_embed_runs", "logits_postproc", "logits_argmax", "nearest_resize", "bank_append",
            "counter_add", "linear_tc", "split_rows", "split_cols", "row_softmax", "gp_attention_tc")


class _FakeStream:
    cuda_stream = 0


def install(monkeypatch, ops_module):
    g = globals()
    for name in EMULATED:
        monkeypatch.setattr(ops_module, name, g[name])


def install_engine(monkeypatch):
    """Everything the engines need to run on CPU tensors through the emulated entry points (tests only)."""
    from aot_benchmark_b200 import engine, ops, plan
    install(monkeypatch, ops)
    monkeypatch.setattr(plan.Plan, "_require_cuda", staticmethod(lambda dev: None))
    monkeypatch.setattr(engine, "_cur_stream", lambda: 0)
    monkeypatch.setattr(engine, "USE_GRAPHS", False)
    monkeypatch.setattr(engine, "SUB_ENGINE_STREAMS", False)
    monkeypatch.setattr(engine.AOTEngine, "_check_img", lambda self, img: None)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _FakeStream())
