"""Drop-in eval engines: the reference's protocol, executed by the sm_100a kernels.

Mirrors networks/engines/aot_engine.py (AOTEngine :13-482, AOTInferEngine :485-635) and
networks/engines/deaot_engine.py (DeAOTEngine :9-56, DeAOTInferEngine :59-94): same class and
method names, arguments, state attributes and error behaviour, so networks/managers/evaluator.py
and tools/demo.py drive it unedited.  Differences are internal:

* activations live in NHWC / token-major fp32 buffers allocated once per video;
* the long-term memory is a pre-allocated append buffer (bank) instead of torch.cat'ed tensors
  (new frames are appended, not prepended -- attention is permutation invariant over keys);
* every FLOP of the per-frame path runs in libaotb200.so; there is no eager fallback and the
  engines refuse CPU tensors.

Training (AOTEngine.forward, aot_engine.py:33-108) is a "next" row of SURVEY 8(f) and raises.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .plan import get_plan

A_NONE, A_RELU, A_GELU, A_SILU, A_RELU6 = ops.ACT_NONE, ops.ACT_RELU, ops.ACT_GELU, ops.ACT_SILU, ops.ACT_RELU6

# bench.py hook: when set to a list, every long-term attention launch appends
# (start_event, end_event, algorithmic_flops) so the roofline is measured live, per launch.
LT_PROBE = None
# long-term attention implementation for the AOT head shape (8 x 32):
#   "tc_exact" tcgen05 kernel, split-fp16 operands, fp32-faithful   (lt_attn_tc.cu)
#   "tc_fast"  tcgen05 kernel, single fp16 pass for Q K^T and P
#   "simt"     fp32 CUDA-core flash kernel                          (attention_simt.cu)
import os as _os
LT_IMPL = _os.environ.get("AOTB_LT_IMPL", "tc_exact")
LOCAL_IMPL = _os.environ.get("AOTB_LOCAL_IMPL", "tile")      # "tile" (halo in smem) | "warp" (generic kernel)
# DeAOT long-term attention (1 head, d_qk 128, d_v 1024): "tc" = fused tcgen05 flash kernel (gp_attn_tc.cu, default since
# round 2: 475 us per launch on the cfg3 clip against 661 for "gemm" and 4189 for "simt", profiles/r02_trip1_summary.md) |
# "gemm" = tensor-core GEMM (Q K^T) -> row softmax -> tensor-core GEMM (P V) over split-fp16 operand copies of the bank
# (deaot_lt.cu) | "simt" = fp32 CUDA-core flash kernel
DEAOT_LT = _os.environ.get("AOTB_DEAOT_LT", "tc")
# exchange step of the sharded long-term bank (BASELINE configs[3]): "nccl" = three all-gathers of the (m, l, O) partials +
# local merge; "p2p" = the partials live in a torch symmetric-memory allocation and every rank's merge kernel reads its
# peers' partials in place over NVLink (aotb_attn_merge_peers_f32) behind one device-side barrier -- no NCCL on the data
# path.  "p2p" was written without multi-GPU access (logic checked on CPU with an in-process stand-in for the allocator).
SHARD_XCHG = _os.environ.get("AOTB_SHARD_XCHG", "nccl")
SHARD_SMAX = 16          # split capacity of the symmetric partial buffers
SUB_ENGINE_STREAMS = _os.environ.get("AOTB_SUB_ENGINE_STREAMS", "1") == "1"   # > 10 objects: sub-engines on concurrent streams
SHARD_GRAPHS = _os.environ.get("AOTB_SHARD_GRAPHS", "1") == "1"   # capture the LSTT call (incl. the exchange) in sharded mode


def _symm_alloc(numel, device, group):
    """-> (handle, rank -> flat fp32 view of that rank's buffer) for a symmetric-memory allocation of `numel` floats."""
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(numel, dtype=torch.float32, device=device)
    hdl = symm_mem.rendezvous(t, group if group is not None else dist.group.WORLD)
    return hdl, (lambda r, sizes, off: hdl.get_buffer(r, sizes, torch.float32, off))


GEMM_GROW_FRAMES = int(_os.environ.get("AOTB_GEMM_GROW_FRAMES", "8"))   # bank growth step of the GEMM path (memory frames)
_LT_NAMES = {"simt": "attn_f32_kernel<32,32> (fp32 SIMT flash attention)",
             "tc_exact": "lt_attn_tc_kernel (tcgen05 fp16x2 exact: 6+16 MMAs/tile)",
             "tc_fast": "lt_attn_tc_kernel (tcgen05 fp16 fast: 2+8 MMAs/tile)"}
_DEAOT_LT_NAMES = {"simt": "attn_f32_kernel<128,256> (fp32 SIMT flash attention, DeAOT 1 x 128 / 1024 head)",
                   "gemm": "conv_tc_kernel (Q K^T) -> row_softmax_kernel -> conv_tc_kernel (P V): tcgen05 GEMMs over split-fp16 "
                           "copies of the bank (deaot_lt.cu)",
                   "tc": "gp_attn_tc_kernel (fused tcgen05 flash attention, 128 queries x 128 value channels per CTA)"}


def deaot_lt_kernel_name():
    """The DeAOT long-term attention implementation actually in use (bench.py's roofline entry names it)."""
    return _DEAOT_LT_NAMES.get(DEAOT_LT, DEAOT_LT)


LT_KERNEL_NAME = _LT_NAMES.get(LT_IMPL, LT_IMPL) + (f", softmax layout '{ops.LT_VARIANT}'" if LT_IMPL.startswith("tc") else "")


# Whole-call CUDA graphs (encoder / LSTT / decoder / memory update are each captured once per video geometry and
# replayed; the live key count and the bank append offset are read from a device counter inside the kernels).
USE_GRAPHS = _os.environ.get("AOTB_GRAPHS", "1") == "1"
# programmatic dependent launch: kernel N+1's prologue (barrier init, TMEM allocation, descriptor prefetch) overlaps
# kernel N's tail; every kernel waits (griddepcontrol.wait) before reading its inputs
USE_PDL = _os.environ.get("AOTB_PDL", "1") == "1"     # programmatic dependent launch: +2 % (profiles/r01_trip14)
CONV_TILING = _os.environ.get("AOTB_CONV_TILING", "model")   # "model" (fitted cost model) | "narrow" (old heuristic)
_CONV_TILING_MASK = {"model": 0, "narrow": 1}


def _apply_pdl():
    """Push the launch-policy knobs into the library (cheap; called at the start of every clip)."""
    from ._lib import lib, check
    lib().aotb_set_pdl(1 if USE_PDL else 0)
    check(lib().aotb_set_conv_tiling(_CONV_TILING_MASK[CONV_TILING]), "aotb_set_conv_tiling")
BANK_INIT_FRAMES = int(_os.environ.get("AOTB_BANK_FRAMES", "24"))   # initial long-term bank capacity (memory frames)


REPLAYED_KERNELS = [0]      # kernels launched through graph replays (bench.py adds this to the library's eager count)


class GraphCache:
    """key -> [eager_runs, CUDAGraph | None, result, kernels_in_graph]; first call runs eagerly (warm-up: lazy module
    loads, cudaFuncSetAttribute, buffer allocation), second call is captured, later calls replay."""

    def __init__(self):
        self.slots = {}

    def clear(self):
        self.slots.clear()

    def run(self, key, fn, enabled=True):
        if not (enabled and USE_GRAPHS and LT_PROBE is None):
            return fn()
        slot = self.slots.get(key)
        if slot is None:
            slot = self.slots[key] = [0, None, None]
        if slot[1] is not None:
            slot[1].replay()
            REPLAYED_KERNELS[0] += slot[3]
            return slot[2]
        if slot[0] < 1:
            slot[0] += 1
            return fn()
        from ._lib import lib
        n0 = lib().aotb_launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = fn()
        n_kernels = int(lib().aotb_launch_count() - n0)     # kernels recorded into this graph
        g.replay()                     # capture records only; replay produces this call's results
        slot[1], slot[2] = g, out
        slot.append(n_kernels)
        return out


def _cur_stream():
    return torch.cuda.current_stream().cuda_stream


def lt_splits(n_queries, heads, tk, sms=148, variant=None):
    """KV-split count for the tensor-core kernel: fill whole waves of CTA slots while keeping >= 2 x 128 keys per split.
    One CTA = 256 queries x 1 head x 1 split at one CTA per SM ("tile" / "groups" / "ahead" layouts) or 128 queries x 1
    head x 1 split at two CTAs per SM ("pair" layout)."""
    pair = (ops.LT_VARIANT if variant is None else variant) == "pair"
    qrows, slots = (128, 2 * sms) if pair else (256, sms)
    base = ((n_queries + qrows - 1) // qrows) * heads
    tiles = (tk + 127) // 128
    effs = []
    for s in range(1, 17):
        if s > 1 and tiles < 2 * s:
            break
        ctas = base * s
        effs.append((s, ctas / (((ctas + slots - 1) // slots) * slots)))
    top = max(e for _, e in effs)
    return next(s for s, e in effs if e >= top - 0.05)     # fewest splits within 5 % of the best wave fill


def _pos_emb_sine(h, w, npf=128):
    """networks/layers/position.py:49-74 (normalize=True, scale=2*pi, T=1e4), as a host-computed
    constant table [h*w, 2*npf] (computed once per video, aot_engine.py:225-228)."""
    y = torch.arange(h, dtype=torch.float32).view(h, 1).expand(h, w)
    x = torch.arange(w, dtype=torch.float32).view(1, w).expand(h, w)
    eps = 1e-6
    y = y / (y[-1:, :] + eps) * (2 * math.pi)
    x = x / (x[:, -1:] + eps) * (2 * math.pi)
    dim_t = torch.arange(npf, dtype=torch.float32)
    dim_t = 10000 ** (2 * (dim_t // 2) / npf)
    px = x[:, :, None] / dim_t
    py = y[:, :, None] / dim_t
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=3).flatten(2)
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=3).flatten(2)
    return torch.cat((py, px), dim=2).reshape(h * w, 2 * npf).contiguous()


class EncEmbs(list):
    """``curr_enc_embs``: list of NCHW feature views [4x, 8x, 16x, 16x-projected] like the reference
    (aot.py:81-84) plus the NHWC tensors the kernels use (``.nhwc``)."""
    nhwc = None


# =====================================================================================
# image encoder (shared by all sub-engines of an infer engine)
# =====================================================================================
class _Encoder:
    def __init__(self, plan, H, W):
        self.plan = plan
        self.H, self.W = H, W
        dev = plan.device
        self.bufs = {}
        self.dev = dev

    def _buf(self, key, shape):
        b = self.bufs.get(key)
        if b is None or tuple(b.shape) != tuple(shape):
            b = torch.empty(shape, dtype=torch.float32, device=self.dev)
            self.bufs[key] = b
        return b

    @staticmethod
    def _osz(n, k, s, p, d=1):
        return (n + 2 * p - d * (k - 1) - 1) // s + 1

    def __call__(self, img, st):
        P = self.plan
        if img.dim() != 4 or img.shape[0] != 1 or img.shape[1] != 3:
            raise ValueError("expected an image tensor [1,3,H,W]")
        if P.encoder_name == "swin_base" and (img.shape[2] % 4 or img.shape[3] % 4):
            # PatchEmbed zero-pads right/bottom to a multiple of the 4x4 patch (swin_transformer.py:476-481)
            img = torch.nn.functional.pad(img, (0, -img.shape[3] % 4, 0, -img.shape[2] % 4))
        H, W = img.shape[2], img.shape[3]
        if getattr(self, "_gkey", None) != (id(P), H, W):
            self.graphs = GraphCache()
            self._gkey = (id(P), H, W)
        x = self._buf("in", (1, H, W, 4))
        ops.image_to_nhwc4(img.float(), x, stream=st)        # caller's tensor -> static NHWC4 input (eager)
        nhwc = self.graphs.run("enc", lambda: self._body(x))
        out = EncEmbs(t.permute(0, 3, 1, 2) for t in nhwc)
        out.nhwc = nhwc
        return out

    def _body(self, x):
        P = self.plan
        st = _cur_stream()
        if P.encoder_name == "resnet50" and ops.CONV_CHAIN and ops.CONV_IMPL == "tc":
            return self._resnet_chain(x, st)
        if P.encoder_name == "resnet50":
            feats = self._resnet(x, st)
        elif P.encoder_name == "swin_base":
            feats = self._swin(x, st)
        else:
            feats = self._mobilenet(x, st)
        f16 = feats[-1]
        proj = self._buf("proj", (1, f16.shape[1], f16.shape[2], P.C))
        ops.conv2d(f16, P.proj.w, P.proj.b, proj, stream=st)
        return [feats[0], feats[1], feats[2], proj]

    def _resnet(self, x, st):
        e = self.plan.enc
        H, W = x.shape[1], x.shape[2]
        h1, w1 = self._osz(H, 7, 2, 3), self._osz(W, 7, 2, 3)
        c1 = self._buf("stem", (1, h1, w1, 64))
        ops.conv2d(x, e.stem.w, e.stem.b, c1, KH=7, KW=7, stride=2, pad=3, act=A_RELU, stream=st)
        h, w = self._osz(h1, 3, 2, 1), self._osz(w1, 3, 2, 1)
        cur = self._buf("pool", (1, h, w, 64))
        ops.maxpool3x3s2(c1, cur, stream=st)
        feats = []
        for si, blocks in enumerate(e.stages):
            for bi, b in enumerate(blocks):
                ho, wo = self._osz(h, 3, b.stride, 1), self._osz(w, 3, b.stride, 1)
                t1 = self._buf(f"s{si}b{bi}t1", (1, h, w, b.c1.cout))
                ops.conv2d(cur, b.c1.w, b.c1.b, t1, act=A_RELU, stream=st)
                t2 = self._buf(f"s{si}b{bi}t2", (1, ho, wo, b.c2.cout))
                ops.conv2d(t1, b.c2.w, b.c2.b, t2, KH=3, KW=3, stride=b.stride, pad=1, act=A_RELU, stream=st)
                if b.down is not None:
                    res = self._buf(f"s{si}ds", (1, ho, wo, b.down.cout))
                    ops.conv2d(cur, b.down.w, b.down.b, res, stride=b.stride, stream=st)
                else:
                    res = cur
                out = self._buf(f"s{si}o{bi % 2}", (1, ho, wo, b.c3.cout))
                ops.conv2d(t2, b.c3.w, b.c3.b, out, res=res, act=A_RELU, stream=st)
                cur, h, w = out, ho, wo
            feats.append(cur)
        return feats

    def _resnet_chain(self, x, st):
        """ResNet-50 stages + projector as ONE persistent dataflow kernel (ops.ConvChain, csrc/conv_chain.cu): the 7x7 stem and
        the max-pool stay separate launches, the 52 bottleneck convs and the 1x1 projector become a tile program.  Every layer
        output gets its own buffer (an address must not change twice inside one launch: the L1 caches are not coherent)."""
        P = self.plan
        e = P.enc
        H, W = x.shape[1], x.shape[2]
        h1, w1 = self._osz(H, 7, 2, 3), self._osz(W, 7, 2, 3)
        c1 = self._buf("stem", (1, h1, w1, 64))
        ops.conv2d(x, e.stem.w, e.stem.b, c1, KH=7, KW=7, stride=2, pad=3, act=A_RELU, stream=st)
        h, w = self._osz(h1, 3, 2, 1), self._osz(w1, 3, 2, 1)
        pool = self._buf("pool", (1, h, w, 64))
        ops.maxpool3x3s2(c1, pool, stream=st)
        key = (id(P), H, W)
        ch = getattr(self, "_chain", None)
        if ch is None or ch[0] != key:
            layers, feats = [], []
            cur, cur_i = pool, -1                      # tensor + index of the chain layer that produces it (-1: the pool)
            for si, blocks in enumerate(e.stages):
                for bi, b in enumerate(blocks):
                    ho, wo = self._osz(h, 3, b.stride, 1), self._osz(w, 3, b.stride, 1)
                    t1 = self._buf(f"c{si}b{bi}t1", (1, h, w, b.c1.cout))
                    layers.append(dict(x=cur, w=b.c1.w, bias=b.c1.b, out=t1, act=A_RELU, in_layer=cur_i))
                    i1 = len(layers) - 1
                    t2 = self._buf(f"c{si}b{bi}t2", (1, ho, wo, b.c2.cout))
                    layers.append(dict(x=t1, w=b.c2.w, bias=b.c2.b, out=t2, KH=3, stride=b.stride, pad=1, act=A_RELU, in_layer=i1))
                    i2 = len(layers) - 1
                    if b.down is not None:
                        res = self._buf(f"c{si}ds", (1, ho, wo, b.down.cout))
                        layers.append(dict(x=cur, w=b.down.w, bias=b.down.b, out=res, stride=b.stride, in_layer=cur_i))
                        res_i = len(layers) - 1
                    else:
                        res, res_i = cur, cur_i
                    out = self._buf(f"c{si}o{bi}", (1, ho, wo, b.c3.cout))
                    layers.append(dict(x=t2, w=b.c3.w, bias=b.c3.b, out=out, res=res, act=A_RELU, in_layer=i2, res_layer=res_i))
                    cur, cur_i, h, w = out, len(layers) - 1, ho, wo
                feats.append(cur)
            proj = self._buf("proj", (1, h, w, P.C))
            layers.append(dict(x=cur, w=P.proj.w, bias=P.proj.b, out=proj, in_layer=cur_i))
            ch = self._chain = (key, ops.ConvChain(layers, self.dev, stream=st), feats, proj)
        ch[1].run(st)
        return [ch[2][0], ch[2][1], ch[2][2], ch[3]]

    def _swin(self, x4, st):
        """SwinTransformer.forward (swin_transformer.py:684-716) for 'swin_base': tokens stay one [H*W, C] matrix per
        stage (= the NHWC map), every Linear is a tensor-core GEMM with the residual / GELU fused in its finish, the
        window partition / shift / padding / mask live inside window_attn_kernel, and the per-stage output norms
        write the NHWC feature maps the decoder reads."""
        e = self.plan.enc
        H, W, C = x4.shape[1] // 4, x4.shape[2] // 4, e.embed
        pe = self._buf("pe", (1, H, W, C))
        ops.conv2d(x4, e.patch.w, e.patch.b, pe, KH=4, KW=4, stride=4, pad=0, stream=st)          # PatchEmbed :473-489
        x = self._buf("s0x", (H * W, C))
        ops.layernorm(pe.view(H * W, C), e.patch_norm[0], e.patch_norm[1], x, stream=st)
        feats = []
        for si, stg in enumerate(e.stages):
            N = H * W
            ln = self._buf(f"s{si}ln", (N, C))
            qkv = self._buf(f"s{si}qkv", (N, 3 * C))
            att = self._buf(f"s{si}att", (N, C))
            hid = self._buf(f"s{si}hid", (N, 4 * C))
            for b in stg.blocks:                                                            # SwinTransformerBlock :257-323
                ops.layernorm(x, b.norm1[0], b.norm1[1], ln, stream=st)
                ops.linear(ln, b.qkv_w, b.qkv_b, qkv, stream=st)
                ops.window_attention(qkv, b.qkv_b, b.relb, att, H, W, stg.heads, b.shift, window=e.window, stream=st)
                ops.linear(att, b.proj_w, b.proj_b, x, res=x, stream=st)                    # x = shortcut + proj(attn)
                ops.layernorm(x, b.norm2[0], b.norm2[1], ln, stream=st)
                ops.linear(ln, b.fc1_w, b.fc1_b, hid, act=A_GELU, stream=st)
                ops.linear(hid, b.fc2_w, b.fc2_b, x, res=x, stream=st)                      # x = x + mlp(norm2(x))
            f = self._buf(f"s{si}f", (1, H, W, C))
            ops.layernorm(x, stg.norm[0], stg.norm[1], f.view(N, C), stream=st)             # norm{i} on the stage output
            feats.append(f)
            if stg.down is not None:                                                        # PatchMerging :339-365
                H2, W2 = (H + 1) // 2, (W + 1) // 2
                mg = self._buf(f"s{si}mg", (H2 * W2, 4 * C))
                ops.patch_merge(x, mg, H, W, stream=st)
                mln = self._buf(f"s{si}mln", (H2 * W2, 4 * C))
                ops.layernorm(mg, stg.down.norm[0], stg.down.norm[1], mln, stream=st)
                x = self._buf(f"s{si + 1}x", (H2 * W2, 2 * C))
                ops.linear(mln, stg.down.w, stg.down.b, x, stream=st)
                H, W, C = H2, W2, 2 * C
        return feats

    def _mobilenet(self, x, st):
        e = self.plan.enc
        H, W = x.shape[1], x.shape[2]
        h, w = self._osz(H, 3, 2, 1), self._osz(W, 3, 2, 1)
        cur = self._buf("stem", (1, h, w, 32))
        ops.conv2d(x, e.stem.w, e.stem.b, cur, KH=3, KW=3, stride=2, pad=1, act=A_RELU6, stream=st)
        feats = []
        for i, b in enumerate(e.blocks):
            y = cur
            if b.expand is not None:
                t = self._buf(f"b{i}e", (1, h, w, b.expand.cout))
                ops.conv2d(y, b.expand.w, b.expand.b, t, act=A_RELU6, stream=st)
                y = t
            pad = b.dil  # (3-1)//2*dil, mobilenetv2.py:41-42
            ho, wo = self._osz(h, 3, b.stride, pad, b.dil), self._osz(w, 3, b.stride, pad, b.dil)
            t = self._buf(f"b{i}d", (1, ho, wo, b.dw.cout))
            ops.dwconv(y, b.dw.w, b.dw.b, t, K=3, stride=b.stride, pad=pad, dil=b.dil, act=A_RELU6, stream=st)
            o = self._buf(f"b{i}o", (1, ho, wo, b.pw.cout))
            ops.conv2d(t, b.pw.w, b.pw.b, o, res=cur if b.res else None, stream=st)
            cur, h, w = o, ho, wo
            if b.tap:
                feats.append(cur)
        last = self._buf("last", (1, h, w, e.last.cout))
        ops.conv2d(cur, e.last.w, e.last.b, last, act=A_RELU6, stream=st)
        feats.append(last)
        return feats


# =====================================================================================
# single engine (<= max_obj_num objects)
# =====================================================================================
class AOTEngine(nn.Module):
    def __init__(self, aot_model, gpu_id=0, long_term_mem_gap=9999, short_term_mem_skip=1):
        super().__init__()
        self.cfg = aot_model.cfg
        self.align_corners = aot_model.cfg.MODEL_ALIGN_CORNERS
        self.AOT = aot_model
        self.max_obj_num = aot_model.max_obj_num
        self.gpu_id = gpu_id
        self.long_term_mem_gap = long_term_mem_gap
        self.short_term_mem_skip = short_term_mem_skip
        self.losses = None
        self._enc = None
        self._ws = None
        self._ws_key = None
        self._P = None
        self.graphs = GraphCache()
        self.tk_dev = None
        self.kv_shard = None          # (rank, world, process_group) when the long-term bank is sharded over GPUs
        self.restart_engine()

    # ------------------------------------------------------------------ protocol
    def forward(self, *args, **kwargs):
        raise NotImplementedError("training step (BASELINE config 5) is a 'next' row (SURVEY 8f); the B200 engine "
                                  "implements the eval protocol")

    def restart_engine(self, batch_size=1, enable_id_shuffle=False):
        if batch_size != 1 or enable_id_shuffle:
            raise NotImplementedError("batch_size > 1 / id shuffle are training-only (SURVEY 8f)")
        self.batch_size = 1
        self.frame_step = 0
        self.last_mem_step = -1
        self.enable_id_shuffle = False
        self.freeze_id = False
        self.obj_nums = None
        self.pos_emb = None
        self.enc_size_2d = None
        self.enc_hw = None
        self.input_size_2d = None
        self.bank_len = 0
        self.enable_offline_enc = False
        self.curr_enc_embs = None
        self.curr_id_embs = None
        self.pred_id_logits = None
        self._have_lstt = False
        self._mem_frames = 0          # memory frames stored so far (global count, all shards)
        if self.tk_dev is not None:
            self.tk_dev.zero_()

    def enable_kv_sharding(self, rank, world, group=None):
        """BASELINE config 4: shard the long-term bank by memory frame round-robin over `world` ranks.  Every rank
        runs the rest of the network redundantly on identical inputs; rank `f % world` keeps memory frame f; each
        rank's long-term attention produces un-normalised partials (m, l, O) over its shard which are all-gathered
        (one NCCL collective per layer) and merged exactly (log-sum-exp) on every rank."""
        if not (LT_IMPL.startswith("tc") and not self._plan_is_deaot()):
            raise NotImplementedError("sharded long-term attention is implemented for the AOT tensor-core kernel")
        self.kv_shard = (int(rank), int(world), group)

    def _plan_is_deaot(self):
        return self.AOT.cfg.MODEL_VOS == "deaot"

    def update_size(self, input_size, enc_size):
        self.input_size_2d = tuple(int(s) for s in input_size)
        self.enc_size_2d = tuple(int(s) for s in enc_size)
        self.enc_hw = self.enc_size_2d[0] * self.enc_size_2d[1]

    # ------------------------------------------------------------------ buffers
    def _plan(self):
        # resolved once per reference frame (get_plan walks every parameter to detect reloads)
        if self._P is None:
            self._P = get_plan(self.AOT)
        return self._P

    def _alloc(self):
        P = self._plan()
        dev = P.device
        N = self.enc_hw
        C = P.C
        L = P.L
        key = (id(P), N, tuple(self.enc_size_2d), tuple(self.input_size_2d), LT_IMPL, DEAOT_LT)
        if self._ws is not None and self._ws_key == key:
            # same geometry and weights as the previous video: keep buffers and captured graphs
            self.bank_len = 0
            self.tk_dev.zero_()
            self._st_ring = []
            return
        self._ws_key = key
        self.graphs.clear()
        self.tk_dev = torch.zeros(1, dtype=torch.int32, device=dev)    # live rows of the long-term bank
        f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        ws = type("WS", (), {})()
        ws.N = N
        ws.gn_ws = ops.groupnorm_workspace(1, 32, dev)
        ws.id_emb = f(N, C)
        if not P.deaot:
            ws.x = f(N, C)
            ws.ln = f(N, C)
            ws.ln_pos = f(N, C)
            ws.qk = f(N, 2 * C)
            ws.v = f(N, C)
            ws.core = f(N, 2 * C)
            ws.tmp = f(N, C)
            ws.ff = f(N, 4 * C)
            ws.ff2 = f(N, 4 * C)
            ws.cat = f(N, (L + 1) * C)
            self.curr_Q = [f(N, C) for _ in range(L)]
            self.curr_V = [f(N, C) for _ in range(L)]
            self.st_K = [f(N, C) for _ in range(L)]
            self.st_V = [f(N, C) for _ in range(L)]
            self._kdim, self._vdim = C, C
        else:
            d = C // 2
            ws.xz = f(N, 2 * C)
            ws.ln = f(N, C)
            ws.qv = f(N, d + 2 * C)
            ws.catU = f(N, 4 * C)
            ws.idin = f(N, 2 * C)
            ws.core = f(N, 4 * C)
            ws.gated = f(N, 4 * C)
            ws.dw = f(N, 8 * C)
            ws.c = f(N, 2 * C)
            ws.sa_qk = f(N, d)
            ws.sa_v = f(N, 4 * C)
            ws.sa_u = f(N, 4 * C)
            ws.cat = f(N, 2 * C)
            self.curr_Q = [f(N, d) for _ in range(L)]
            self.curr_V = [f(N, 2 * C) for _ in range(L)]
            self.curr_IDV = [None] + [f(N, C) for _ in range(L - 1)]
            self.st_K = [f(N, d) for _ in range(L)]
            self.st_V = [f(N, 4 * C) for _ in range(L)]   # cat[V | ID_V] (transformer.py:625-626)
            self._kdim, self._vdim = d, 4 * C
        ws.mask = f(*self.input_size_2d)                            # static copy of the caller's label map
        self._dec_out = {}
        cap = (min(BANK_INIT_FRAMES, GEMM_GROW_FRAMES) if (P.deaot and DEAOT_LT == "gemm") else BANK_INIT_FRAMES) * N
        self.bank_cap = cap
        self.bank_K = [f(cap, self._kdim) for _ in range(L)]
        self.bank_V = [f(cap, self._vdim) for _ in range(L)]
        self.bank_len = 0
        self._gemm_lt = bool(P.deaot and DEAOT_LT == "gemm")
        if self._gemm_lt:
            self._alloc_gemm_lt(cap, ws=ws)
        self._gp_tc = bool(P.deaot and DEAOT_LT == "tc")
        if self._gp_tc:
            hz = lambda *s: torch.zeros(s, dtype=torch.float16, device=dev)
            ws.gpQp = hz(self._kdim // 32, ((N + 127) // 128) * 128, 64)
            ncap = ((N + 63) // 64) * 64 + 64                      # K / V of the CURRENT frame (self-attention, reference frame)
            ws.gpSaK = hz(self._kdim // 32, ncap, 64)
            ws.gpSaV = hz(self._vdim // 32, ncap, 64)
            self.bank_gpK = [hz(self._kdim // 32, cap, 64) for _ in range(L)]     # split-fp16 rows, one "head" / 32 channels
            self.bank_gpV = [hz(self._vdim // 32, cap, 64) for _ in range(L)]
            ws.gp_part = {}
        self._tc = LT_IMPL.startswith("tc") and (not P.deaot) and (C // P.H == 32)
        if self._tc:
            hz = lambda *s: torch.zeros(s, dtype=torch.float16, device=dev)
            ws.Qp = hz(P.H, ((N + 255) // 256) * 256, 64)
            ws.saKp = hz(P.H, ((N + 127) // 128) * 128, 64)          # self-attention K / V of the current frame
            ws.saVp = hz(P.H, ((N + 127) // 128) * 128, 64)
            self.bank_Kp = [hz(P.H, cap, 64) for _ in range(L)]     # split-fp16 copies read by TMA
            self.bank_Vp = [hz(P.H, cap, 64) for _ in range(L)]
            ws.part = {}
        self._st_ring = []
        self._ws = ws
        self._dec_bufs = {}

    def _alloc_gemm_lt(self, cap, old_len=0, old=None, ws=None):
        """Split-fp16 operand copies of the DeAOT bank for the GEMM formulation: keys [cap64][d] (weights of Q K^T),
        values transposed [4C][cap64] (weights of P V), zero beyond the live rows, plus the score matrix [N][cap64]."""
        P = self._plan()
        dev = P.device
        capw = ((cap + 63) // 64) * 64
        hz = lambda *s: torch.zeros(s, dtype=torch.float16, device=dev)
        Kh, Kl = [hz(capw, self._kdim) for _ in range(P.L)], [hz(capw, self._kdim) for _ in range(P.L)]
        Vh, Vl = [hz(self._vdim, capw) for _ in range(P.L)], [hz(self._vdim, capw) for _ in range(P.L)]
        if old is not None:
            for new_l, old_l in zip((Kh, Kl), old[:2]):
                for a, b in zip(new_l, old_l):
                    a[:old_len].copy_(b[:old_len])
            for new_l, old_l in zip((Vh, Vl), old[2:]):
                for a, b in zip(new_l, old_l):
                    a[:, :old_len].copy_(b[:, :old_len])
        self.bank_Kh, self.bank_Kl, self.bank_VhT, self.bank_VlT = Kh, Kl, Vh, Vl
        self._capw = capw
        (self._ws if ws is None else ws).S = torch.empty((self.enc_hw, capw), dtype=torch.float32, device=dev)

    def _bank_reserve(self, rows):
        if self.bank_len + rows <= self.bank_cap:
            return
        new_cap = max(2 * self.bank_cap, self.bank_len + rows)
        if getattr(self, "_gemm_lt", False):
            # the GEMM formulation of DeAOT's long-term attention costs O(capacity), not O(live keys): grow in steps of
            # GEMM_GROW_FRAMES memory frames (a few graph re-captures per clip) instead of doubling
            new_cap = max(self.bank_cap + GEMM_GROW_FRAMES * self.enc_hw, self.bank_len + rows)
        for lst in (self.bank_K, self.bank_V):
            for i, old in enumerate(lst):
                nb = torch.empty((new_cap, old.shape[1]), dtype=torch.float32, device=old.device)
                nb[: self.bank_len].copy_(old[: self.bank_len])
                lst[i] = nb
        if self._tc:
            for lst in (self.bank_Kp, self.bank_Vp):
                for i, old in enumerate(lst):
                    nb = torch.zeros((old.shape[0], new_cap, 64), dtype=torch.float16, device=old.device)
                    nb[:, : self.bank_len].copy_(old[:, : self.bank_len])
                    lst[i] = nb
        if getattr(self, "_gemm_lt", False):
            self._alloc_gemm_lt(new_cap, self.bank_len, (self.bank_Kh, self.bank_Kl, self.bank_VhT, self.bank_VlT))
        if getattr(self, "_gp_tc", False):
            for lst in (self.bank_gpK, self.bank_gpV):
                for i, old in enumerate(lst):
                    nb = torch.zeros((old.shape[0], new_cap, 64), dtype=torch.float16, device=old.device)
                    nb[:, : self.bank_len].copy_(old[:, : self.bank_len])
                    lst[i] = nb
        self.bank_cap = new_cap
        self.graphs.clear()            # captured launches point at the old bank

    # ------------------------------------------------------------------ reference-shaped views
    @property
    def long_term_memories(self):
        if self.bank_len == 0:
            return None
        P = self._plan()
        n = self.bank_len
        out = []
        for li in range(P.L):
            K = self.bank_K[li][:n].unsqueeze(1)
            V = self.bank_V[li][:n]
            if P.deaot:
                c2 = 2 * P.C
                out.append([K, V[:, :c2].unsqueeze(1), None, V[:, c2:].unsqueeze(1)])
            else:
                out.append([K, V.unsqueeze(1)])
        return out

    @property
    def short_term_memories(self):
        if not self._have_lstt:
            return None
        P = self._plan()
        h, w = self.enc_size_2d
        to2d = lambda t: t.view(h, w, 1, -1).permute(2, 3, 0, 1)
        out = []
        for li in range(P.L):
            if P.deaot:
                c2 = 2 * P.C
                out.append([to2d(self.st_K[li]), to2d(self.st_V[li][:, :c2]), None, to2d(self.st_V[li][:, c2:])])
            else:
                out.append([to2d(self.st_K[li]), to2d(self.st_V[li])])
        return out

    # ------------------------------------------------------------------ steps
    def _encode(self, img, st):
        if self._enc is None:
            self._enc = _Encoder(self._plan(), img.shape[2], img.shape[3])
        self._enc.plan = self._plan()
        return self._enc(img, st)

    def _check_img(self, img):
        if not img.is_cuda:
            raise RuntimeError("aot_benchmark_b200 engines run on CUDA tensors only (there is no CPU path)")

    def assign_identity_from_mask(self, mask, st):
        """one_hot_mask + get_id_emb (aot_engine.py:168-179) fused as a gather (K4)."""
        P = self._plan()
        ws = self._ws
        if mask.dim() == 4 and mask.shape[1] != 1:
            # probability / one-hot input [1, 11, H, W]: dense conv through the same weight table
            x = torch.empty((1, mask.shape[2], mask.shape[3], mask.shape[1]), dtype=torch.float32, device=mask.device)
            ops.nchw_to_nhwc(mask.float().contiguous(), x, stream=st)
            ops.conv2d(x, P.id_wt, P.id_b, ws.id_emb.view(1, *self.enc_size_2d, P.C), KH=P.id_k, KW=P.id_k,
                       stride=P.id_stride, pad=P.id_pad, stream=st)
            if P.deaot:
                ops.layernorm(ws.id_emb, P.id_norm[0], P.id_norm[1], ws.id_emb, stream=st)
            return ws.id_emb
        m2 = mask.reshape(mask.shape[-2], mask.shape[-1]).float().contiguous()
        if tuple(m2.shape) == tuple(ws.mask.shape):
            ops.eltwise(ops.EW_COPY, m2, None, ws.mask, stream=st)      # static buffer: graph replays read it
            m2 = ws.mask
        self._id_from_static_mask(m2, st)
        return ws.id_emb

    def _id_from_static_mask(self, m2, st):
        P = self._plan()
        if P.C == 256:
            ops.id_embed_runs(m2, P.id_wp, P.id_b, self._ws.id_emb, P.C, P.nid, P.id_k, P.id_stride, P.id_pad,
                              ln_gamma=P.id_norm[0] if P.deaot else None, ln_beta=P.id_norm[1] if P.deaot else None,
                              stream=st)
        else:
            ops.id_embed(m2, P.id_wt, P.id_b, self._ws.id_emb, P.C, P.nid, P.id_k, P.id_stride, P.id_pad,
                         ln_gamma=P.id_norm[0] if P.deaot else None, ln_beta=P.id_norm[1] if P.deaot else None, stream=st)

    def add_reference_frame(self, img=None, mask=None, frame_step=-1, obj_nums=None, img_embs=None):
        if self.obj_nums is None and obj_nums is None:
            print('No objects for reference frame!')
            exit()
        elif obj_nums is not None:
            self.obj_nums = obj_nums
        if frame_step == -1:
            frame_step = self.frame_step
        if img_embs is None and img is None:
            print('No image for reference frame!')
            exit()
        if mask is None:
            print('No mask for reference frame!')
            exit()
        st = torch.cuda.current_stream().cuda_stream
        _apply_pdl()
        self._P = get_plan(self.AOT)   # picks up load_state_dict / .to() done since the last video
        if img_embs is None:
            self._check_img(img)
            img_embs = self._encode(img, st)
        if self.input_size_2d is None:
            f16 = img_embs.nhwc[-1]
            in_size = img.shape[2:] if img is not None else (f16.shape[1] * 16, f16.shape[2] * 16)
            self.update_size(in_size, (f16.shape[1], f16.shape[2]))
            self._alloc()
        self.curr_enc_embs = img_embs
        if self.pos_emb is None:
            # the table lives in the workspace, i.e. exactly as long as the captured graphs that read it: re-creating it per
            # video left the graphs of the previous video replaying a freed address (same block again only by allocator luck)
            if getattr(self._ws, "pos_emb", None) is None:
                self._ws.pos_emb = _pos_emb_sine(*self.enc_size_2d, npf=self._plan().C // 2).to(self._plan().device)
            self.pos_emb = self._ws.pos_emb
        id_emb = self.assign_identity_from_mask(mask, st)
        self.curr_id_embs = id_emb
        self._lstt_forward(img_embs, id_emb, st)
        # lstt_long_memories of a reference frame = its own fused K/V (transformer.py:337-341)
        if self._claim_memory_frame():
            self._bank_reserve(self.enc_hw)
            self._append_short_to_bank(st)
            self.bank_len += self.enc_hw
        self.last_mem_step = self.frame_step
        self._have_lstt = True

    def match_propogate_one_frame(self, img=None, img_embs=None):
        self.frame_step += 1
        st = torch.cuda.current_stream().cuda_stream
        if img_embs is None:
            self._check_img(img)
            img_embs = self._encode(img, st)
        self.curr_enc_embs = img_embs
        splits = lt_splits(self.enc_hw, self._plan().H, max(self.bank_len, 1)) if getattr(self, "_tc", False) else 0
        if getattr(self, "_gp_tc", False):
            splits = self._gp_splits(self.bank_len)       # the fused DeAOT kernel's split count is part of the captured body
        if self.kv_shard is not None:
            # sharded bank: the captured body also depends on the shard split count (a function of the GLOBAL memory-frame
            # count, identical on every rank) and on whether this rank holds any memory frame yet
            splits = ("shard", self._shard_splits(), self.bank_len > 0, SHARD_XCHG)
        self.graphs.run(("lstt", splits, img_embs.nhwc[-1].data_ptr()),
                        lambda: self._lstt_forward(img_embs, None, _cur_stream()),
                        enabled=self.short_term_mem_skip <= 1 and (self.kv_shard is None or SHARD_GRAPHS))

    def update_short_term_memory(self, curr_mask, curr_id_emb=None, skip_long_term_update=False):
        st = torch.cuda.current_stream().cuda_stream
        append = False
        if self.frame_step - self.last_mem_step >= self.long_term_mem_gap:
            append = not skip_long_term_update
            self.last_mem_step = self.frame_step
        append = append and self._claim_memory_frame()       # sharded bank: only the owner rank stores this memory frame
        if append:
            self._bank_reserve(self.enc_hw)          # may re-allocate (and drop graphs) before anything is captured
        ws = self._ws
        label_map = curr_id_emb is None and not (curr_mask.dim() == 4 and curr_mask.shape[1] != 1) and \
            tuple(curr_mask.shape[-2:]) == tuple(ws.mask.shape)
        if label_map and self.short_term_mem_skip <= 1:
            m2 = curr_mask.reshape(ws.mask.shape).float().contiguous()
            ops.eltwise(ops.EW_COPY, m2, None, ws.mask, stream=st)

            def body():
                s2 = _cur_stream()
                self._id_from_static_mask(ws.mask, s2)
                self._fuse_memories(ws.id_emb, s2)
                if append:
                    self._append_short_to_bank(s2)
            self.graphs.run(("upd", append), body)
        else:
            id_emb = self.assign_identity_from_mask(curr_mask, st) if curr_id_emb is None else curr_id_emb
            self._fuse_memories(id_emb, st)
            if append:
                self._append_short_to_bank(st)
        if append:
            self.bank_len += self.enc_hw

    def _claim_memory_frame(self):
        """Count one more memory frame (global count, all shards) and say whether THIS rank stores it: always without
        sharding, rank `f % world` for memory frame f with the bank sharded."""
        f = self._mem_frames
        self._mem_frames += 1
        return self.kv_shard is None or f % self.kv_shard[1] == self.kv_shard[0]

    def _append_short_to_bank(self, st):
        """Append the newest fused K/V of every layer at the device-side row counter, then advance it (kernels only: the
        host-side row count and capacity are the caller's business, so the body can be captured in a graph)."""
        N = self.enc_hw
        K_src, V_src = self._latest_kv()
        off = self.tk_dev
        for li in range(self._plan().L):
            ops.bank_append(K_src[li], self.bank_K[li], 0, offset_dev=off, stream=st)
            ops.bank_append(V_src[li], self.bank_V[li], 0, offset_dev=off, stream=st)
            if self._tc:
                ops.tc_pack_rows(K_src[li], self.bank_Kp[li], 0, row_off_dev=off, stream=st)
                ops.tc_pack_rows(V_src[li], self.bank_Vp[li], 0, row_off_dev=off, stream=st)
            if self._gemm_lt:
                ops.split_rows(K_src[li], self.bank_Kh[li], self.bank_Kl[li], row_off_dev=off, stream=st)
                ops.split_cols(V_src[li], self.bank_VhT[li], self.bank_VlT[li], col_off_dev=off, stream=st)
            if getattr(self, "_gp_tc", False):
                ops.tc_pack_rows(K_src[li], self.bank_gpK[li], 0, row_off_dev=off, stream=st)
                ops.tc_pack_rows(V_src[li], self.bank_gpV[li], 0, row_off_dev=off, stream=st)
        ops.counter_add(off, N, stream=st)

    def _latest_kv(self):
        return self._new_K, self._new_V

    # ------------------------------------------------------------------ LSTT (AOT)
    def _lstt_forward(self, embs, id_emb, st):
        P = self._plan()
        ws = self._ws
        N, C, H = self.enc_hw, P.C, P.H
        h, w = self.enc_size_2d
        d = C // H
        proj = embs.nhwc[-1].view(N, C)
        ops.eltwise(ops.EW_COPY, proj, None, ws.x, stream=st)
        ops.eltwise(ops.EW_COPY, proj, None, ws.cat[:, :C], stream=st)
        x = ws.x
        is_ref = id_emb is not None
        if is_ref:
            stK, stV = self._next_short_slot()
        else:
            stK, stV = self.st_K, self.st_V
        for li in range(P.L):
            Lw = P.layers[li]
            # 1) self-attention (transformer.py:321-326)
            ops.layernorm(x, Lw.norm1[0], Lw.norm1[1], ws.ln, add=self.pos_emb, out2=ws.ln_pos, stream=st)
            ops.linear(ws.ln_pos, Lw.sa_qk_w, Lw.sa_qk_b, ws.qk, stream=st)
            ops.linear(ws.ln, Lw.sa_v_w, Lw.sa_v_b, ws.v, stream=st)
            if self._tc:
                self._tc_attention(ws.qk[:, :C], ws.qk[:, C:], ws.v, None, None, N, ws.core[:, :C], st)
            else:
                ops.attention(ws.qk[:, :C], ws.qk[:, C:], ws.v, ws.core[:, :C], H, d, d, Tk=N, stream=st)
            ops.linear(ws.core[:, :C], Lw.sa_proj_w, Lw.sa_proj_b, x, res=x, stream=st)
            # 2) long + short term (transformer.py:329-352)
            cQ, cV = self.curr_Q[li], self.curr_V[li]
            ops.layernorm(x, Lw.norm2[0], Lw.norm2[1], cV, stream=st)
            ops.linear(cV, Lw.linQ_w, Lw.linQ_b, cQ, stream=st)
            if is_ref:
                ops.eltwise(ops.EW_ADD, cV, id_emb, ws.tmp, stream=st)
                ops.linear(ws.tmp, Lw.linV_w, Lw.linV_b, stV[li], stream=st)      # fuse_key_value_id :364-367
                ops.eltwise(ops.EW_COPY, cQ, None, stK[li], stream=st)
                gK, gV, Tk = stK[li], stV[li], N
            else:
                gK, gV, Tk = self.bank_K[li], self.bank_V[li], self.bank_len
            self._long_term_attention(li, cQ, gK, gV, Tk, ws.core[:, :C], st)
            if d == 32 and LOCAL_IMPL == "tile":
                ops.local_attention_tile(cQ, stK[li], stV[li], Lw.relk_w, Lw.relk_b, Lw.relv_t, ws.core[:, C:], h, w, H,
                                         stream=st)
            else:
                ops.local_attention(cQ, stK[li], stV[li], Lw.relk_w, Lw.relk_b, Lw.relv, ws.core[:, C:], h, w, H, d, d,
                                    stream=st)
            ops.linear(ws.core, Lw.lst_proj_w, Lw.lst_proj_b, x, res=x, stream=st)
            # 3) feed-forward (transformer.py:354-359, basic.py:27-35)
            ops.layernorm(x, Lw.norm3[0], Lw.norm3[1], ws.ln, stream=st)
            ops.linear(ws.ln, Lw.lin1_w, Lw.lin1_b, ws.ff, stream=st)
            ops.groupnorm(ws.ff.view(1, N, 4 * C), Lw.gn[0], Lw.gn[1], ws.ff.view(1, N, 4 * C), 32, A_GELU, ws.gn_ws,
                          stream=st)
            ops.dwconv(ws.ff.view(1, h, w, 4 * C), Lw.dw_w, None, ws.ff2.view(1, h, w, 4 * C), K=5, pad=2, stream=st)
            ops.linear(ws.ff2, Lw.lin2_w, Lw.lin2_b, x, res=x, stream=st)
            # decoder norms (transformer.py:124-135) written straight into the decoder input
            ops.layernorm(x, Lw.dec_norm[0], Lw.dec_norm[1], ws.cat[:, (li + 1) * C:(li + 2) * C], stream=st)
        if is_ref:
            self._commit_short_slot(stK, stV)
        self._have_lstt = True

    def _long_term_attention(self, li, Q, K, V, Tk, out, st):
        P = self._plan()
        d = P.C // P.H
        probe = LT_PROBE
        use_tc = self._tc and (K is self.bank_K[li])
        if use_tc:
            ops.tc_pack_rows(Q, self._ws.Qp, 0, div=math.sqrt(d), stream=st)      # Q / T (attention.py:82)
        if probe is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        if use_tc and self.kv_shard is not None:
            self._sharded_attention(li, out, st)
        elif self.kv_shard is not None and K is self.bank_K[li]:
            raise ops.AotbError("sharded long-term bank needs the tensor-core attention kernel (8 heads x 32); this model's "
                                "head shape runs on the fp32 kernel, which has no partial (m, l, O) outputs")
        elif use_tc:
            self._tc_attention(None, None, None, self.bank_Kp[li], self.bank_Vp[li], Tk, out, st, Tk_dev=self.tk_dev)
        elif self._tc:
            self._tc_attention(Q, K, V, None, None, Tk, out, st)     # reference frame: Tk = N, own K/V
        else:
            ops.attention(Q, K, V, out, P.H, d, d, Tk=Tk, Tk_dev=self.tk_dev if K is self.bank_K[li] else None,
                          stream=st)
        if probe is not None:
            e1.record()
            probe.append((e0, e1, 4.0 * Q.shape[0] * Tk * P.C))

    def _tc_attention(self, Q, K, V, Kp, Vp, Tk, out, st, Tk_dev=None):
        """softmax(Q K^T / T) V on the tcgen05 kernel.  Q/K/V fp32 [rows, C] are packed into the split-fp16
        operand buffers first unless already-packed banks (Kp, Vp) are given (then Q was packed by the caller)."""
        P = self._plan()
        ws = self._ws
        d = P.C // P.H
        N = self.enc_hw
        if Q is not None:
            ops.tc_pack_rows(Q, ws.Qp, 0, div=math.sqrt(d), stream=st)
        if Kp is None:
            ops.tc_pack_rows(K, ws.saKp, 0, stream=st)
            ops.tc_pack_rows(V, ws.saVp, 0, stream=st)
            Kp, Vp = ws.saKp, ws.saVp
        splits = lt_splits(N, P.H, Tk)
        part = None
        if splits > 1:
            part = ws.part.get(splits)
            if part is None:
                fz = lambda *s: torch.empty(s, dtype=torch.float32, device=out.device)
                part = (fz(splits, N, P.C), fz(splits, P.H, N), fz(splits, P.H, N))
                ws.part[splits] = part
        ops.lt_attention_tc(ws.Qp, Kp, Vp, N, Tk, O=out, Tk_dev=Tk_dev, splits=splits, exact=(LT_IMPL == "tc_exact"),
                            part=part, stream=st)

    def _shard_splits(self):
        """KV-split count of the per-rank partial attention in sharded mode: a function of the GLOBAL memory-frame count, so it
        is identical on every rank (the exchanged buffers must have the same shape everywhere)."""
        world = self.kv_shard[1]
        frames_per_rank = (self._mem_frames + world - 1) // world
        s = max(2, lt_splits(self.enc_hw, self._plan().H, max(frames_per_rank, 1) * self.enc_hw))
        if s > SHARD_SMAX:
            raise ops.AotbError(f"sharded long-term attention: {s} KV splits exceed the exchange buffer capacity {SHARD_SMAX}")
        return s

    def _sharded_attention(self, li, out, st):
        """Split-KV over ranks: local partials -> ONE all-gather of the packed [O | m | l] block -> exact LSE
        merge straight out of the gathered buffer (per-rank slices, no re-layout).  Q was packed by the caller."""
        import torch.distributed as dist
        rank, world, group = self.kv_shard
        P = self._plan()
        ws = self._ws
        N, C, H = self.enc_hw, P.C, P.H
        splits = self._shard_splits()                                            # identical on every rank
        if st != torch.cuda.current_stream().cuda_stream:
            raise ops.AotbError("sharded long-term attention must run on torch's current stream (the collective / the "
                                "symmetric-memory barrier are issued there)")
        if SHARD_XCHG == "p2p":
            return self._sharded_attention_p2p(li, out, st, splits)
        key = ("shard", splits)
        bufs = ws.part.get(key)
        if bufs is None:
            nO, nM = splits * N * C, splits * H * N
            mine = torch.empty(nO + 2 * nM, dtype=torch.float32, device=out.device)
            allr = torch.empty(world * (nO + 2 * nM), dtype=torch.float32, device=out.device)
            sl = lambda t: (t[:nO].view(splits, N, C), t[nO:nO + nM].view(splits, H, N), t[nO + nM:].view(splits, H, N))
            per_rank = [sl(allr[r * (nO + 2 * nM):(r + 1) * (nO + 2 * nM)]) for r in range(world)]
            bufs = ws.part[key] = (mine, allr, sl(mine), per_rank)
        mine, allr, (Op, Mp, Lp), per_rank = bufs
        if self.bank_len > 0:
            ops.lt_attention_tc(ws.Qp, self.bank_Kp[li], self.bank_Vp[li], N, self.bank_len, O=None, Tk_dev=self.tk_dev,
                                splits=splits, exact=(LT_IMPL == "tc_exact"), part=(Op, Mp, Lp), stream=st, merge=False)
        else:                                  # this rank holds no memory frame yet: neutral partial
            ops.eltwise(ops.EW_FILL, None, None, mine[:Op.numel()].view(1, -1), scalar=0.0, stream=st)
            ops.eltwise(ops.EW_FILL, None, None, Mp.view(1, -1), scalar=float("-inf"), stream=st)
            ops.eltwise(ops.EW_FILL, None, None, Lp.view(1, -1), scalar=0.0, stream=st)
        dist.all_gather_into_tensor(allr, mine, group=group)
        ops.attn_merge_peers([v[0] for v in per_rank], [v[1] for v in per_rank], [v[2] for v in per_rank], out, splits, H,
                             C // H, stream=st)

    def _sharded_attention_p2p(self, li, out, st, splits):
        """Exchange over peer memory: the local partials are written into this rank's slice of a symmetric allocation, one
        device-side barrier makes every rank's partials of this layer visible, and the merge kernel reads all ranks'
        partials in place (NVLink P2P loads).  One allocation per layer, so the next write of a buffer (this layer, next
        frame) is separated from its readers by the barriers of the other layers; single-layer models add a second barrier."""
        rank, world, group = self.kv_shard
        P = self._plan()
        ws = self._ws
        N, C, H = self.enc_hw, P.C, P.H
        key = ("p2p", li)
        buf = ws.part.get(key)
        if buf is None:
            nO, nM = SHARD_SMAX * N * C, SHARD_SMAX * H * N
            hdl, view = _symm_alloc(nO + 2 * nM, out.device, group)
            views = [(view(r, (SHARD_SMAX, N, C), 0), view(r, (SHARD_SMAX, H, N), nO), view(r, (SHARD_SMAX, H, N), nO + nM))
                     for r in range(world)]
            buf = ws.part[key] = (hdl, views)
        hdl, views = buf
        Op, Mp, Lp = (t[:splits] for t in views[rank])
        if self.bank_len > 0:
            ops.lt_attention_tc(ws.Qp, self.bank_Kp[li], self.bank_Vp[li], N, self.bank_len, O=None, Tk_dev=self.tk_dev,
                                splits=splits, exact=(LT_IMPL == "tc_exact"), part=(Op, Mp, Lp), stream=st, merge=False)
        else:                                  # this rank holds no memory frame yet: neutral partial
            ops.eltwise(ops.EW_FILL, None, None, Op.reshape(1, -1), scalar=0.0, stream=st)
            ops.eltwise(ops.EW_FILL, None, None, Mp.reshape(1, -1), scalar=float("-inf"), stream=st)
            ops.eltwise(ops.EW_FILL, None, None, Lp.reshape(1, -1), scalar=0.0, stream=st)
        hdl.barrier()
        ops.attn_merge_peers([v[0] for v in views], [v[1] for v in views], [v[2] for v in views], out, splits, H, C // H,
                             stream=st)
        if P.L < 2:
            hdl.barrier()

    # short-term memory slots (TEST_SHORT_TERM_MEM_SKIP ring, aot_engine.py:329-332)
    def _next_short_slot(self):
        if self.short_term_mem_skip <= 1:
            return self.st_K, self.st_V
        K = [torch.empty_like(t) for t in self.st_K]
        V = [torch.empty_like(t) for t in self.st_V]
        return K, V

    def _commit_short_slot(self, K, V, reset=True):
        self._new_K, self._new_V = K, V
        if self.short_term_mem_skip <= 1:
            return
        if reset:
            self._st_ring = [(K, V)]
        else:
            self._st_ring.append((K, V))
            self._st_ring = self._st_ring[-self.short_term_mem_skip:]
        self.st_K, self.st_V = self._st_ring[0]

    def _fuse_memories(self, id_emb, st):
        """update_short_term_memory core (aot_engine.py:315-332): K = curr_K, V = linear_V(curr_V + id)."""
        P = self._plan()
        ws = self._ws
        K, V = self._next_short_slot()
        for li in range(P.L):
            Lw = P.layers[li]
            ops.eltwise(ops.EW_ADD, self.curr_V[li], id_emb, ws.tmp, stream=st)
            ops.linear(ws.tmp, Lw.linV_w, Lw.linV_b, V[li], stream=st)
            ops.eltwise(ops.EW_COPY, self.curr_Q[li], None, K[li], stream=st)
        self._commit_short_slot(K, V, reset=False)

    # ------------------------------------------------------------------ decoder (fpn.py:34-58)
    def _dbuf(self, key, shape):
        b = self._dec_bufs.get(key)
        if b is None or tuple(b.shape) != tuple(shape):
            b = torch.empty(shape, dtype=torch.float32, device=self._plan().device)
            self._dec_bufs[key] = b
        return b

    def _decode(self, st):
        P = self._plan()
        D = P.dec
        ws = self._ws
        ac = P.align_corners
        x4, x8, x16, _ = self.curr_enc_embs.nhwc
        h, w = self.enc_size_2d
        C = P.C
        gws = ws.gn_ws

        def conv_gn(x, blk, key, k, pad, res=None):
            o = self._dbuf(key, (1, x.shape[1], x.shape[2], blk.cout))
            ops.conv2d(x, blk.w, blk.b, o, KH=k, KW=k, pad=pad, stream=st)
            ov = o.view(1, -1, blk.cout)
            ops.groupnorm(ov, blk.gn[0], blk.gn[1], ov, 8, A_RELU, gws, stream=st)
            return o

        cat = ws.cat.view(1, h, w, -1)
        x = conv_gn(cat, D.conv_in, "in", 1, 0)
        a = self._dbuf("a16", (1, h, w, D.adapter_16x.cout))
        ops.conv2d(x16, D.adapter_16x.w, D.adapter_16x.b, a, res=x, stream=st)
        x = conv_gn(a, D.conv_16x, "c16", 3, 1)
        up = self._dbuf("up8", (1, x8.shape[1], x8.shape[2], x.shape[3]))
        ops.bilinear(x, up, ac, stream=st)
        a = self._dbuf("a8", (1, x8.shape[1], x8.shape[2], D.adapter_8x.cout))
        ops.conv2d(x8, D.adapter_8x.w, D.adapter_8x.b, a, res=up, stream=st)
        x = conv_gn(a, D.conv_8x, "c8", 3, 1)
        up = self._dbuf("up4", (1, x4.shape[1], x4.shape[2], x.shape[3]))
        ops.bilinear(x, up, ac, stream=st)
        a = self._dbuf("a4", (1, x4.shape[1], x4.shape[2], D.adapter_4x.cout))
        ops.conv2d(x4, D.adapter_4x.w, D.adapter_4x.b, a, res=up, stream=st)
        x = conv_gn(a, D.conv_4x, "c4", 3, 1)
        lg = self._dbuf("logit", (1, x.shape[1], x.shape[2], D.conv_out.cout))
        ops.conv2d(x, D.conv_out.w, D.conv_out.b, lg, stream=st)
        return lg

    def decode_current_logits(self, output_size=None):
        """aot_engine.py:356-380.  The returned tensor (and ``pred_id_logits``) are per-engine static buffers that
        the next call overwrites -- clone them to keep a frame's logits."""
        P = self._plan()
        size = None if output_size is None else (int(output_size[0]), int(output_size[1]))
        obj = int(self.obj_nums[0])

        def body():
            st = _cur_stream()
            lg = self._decode(st)
            h4, w4, NC = lg.shape[1], lg.shape[2], lg.shape[3]
            bufs = self._dec_out.get(size)
            if bufs is None:
                lo = torch.empty((1, NC, h4, w4), dtype=torch.float32, device=lg.device)
                out = None if size is None else torch.empty((1, NC) + size, dtype=torch.float32, device=lg.device)
                bufs = self._dec_out[size] = (lo, out)
            ops.logits_postproc(lg, bufs[0], bufs[1], obj, P.align_corners, stream=st)
            return bufs

        lo, out = self.graphs.run(("dec", size, obj, self.curr_enc_embs.nhwc[0].data_ptr()), body)
        self.pred_id_logits = lo
        return lo if out is None else out

    def predict_current_mask(self, output_size=None, return_prob=False):
        """aot_engine.py:382-396 (argmax of the upsampled logits; fused K9 kernel)."""
        if output_size is None:
            output_size = self.input_size_2d
        st = torch.cuda.current_stream().cuda_stream
        oh, ow = int(output_size[0]), int(output_size[1])
        label = torch.empty((1, oh, ow), dtype=torch.float32, device=self.pred_id_logits.device)
        ops.logits_argmax(self.pred_id_logits, label, self._plan().align_corners, stream=st)
        if return_prob:
            raise NotImplementedError("return_prob is used by the training path only (SURVEY 8f)")
        return label.long()


class DeAOTEngine(AOTEngine):
    """networks/engines/deaot_engine.py:9-56 -- GatedPropagationModule stack (transformer.py:501-665)."""

    def __init__(self, aot_model, gpu_id=0, long_term_mem_gap=9999, short_term_mem_skip=1,
                 layer_loss_scaling_ratio=2.):
        super().__init__(aot_model, gpu_id, long_term_mem_gap, short_term_mem_skip)
        self.layer_loss_scaling_ratio = layer_loss_scaling_ratio

    def _gated_tail(self, core, U, dw_w, out_slice, h, w, st):
        """(attn @ V) * U -> depthwise 5x5 (attention.py:707-709 / :855-857); projection is fused later."""
        ws = self._ws
        N = core.shape[0]
        C4 = core.shape[1]
        ops.eltwise(ops.EW_MUL, core, U, ws.gated, stream=st)
        ops.dwconv(ws.gated.view(1, h, w, C4), dw_w, None, out_slice.unflatten(0, (1, h, w)), K=5, pad=2, stream=st)

    def _lstt_forward(self, embs, id_emb, st):
        P = self._plan()
        ws = self._ws
        N, C = self.enc_hw, P.C
        h, w = self.enc_size_2d
        d = C // 2
        C2, C4 = 2 * C, 4 * C
        proj = embs.nhwc[-1].view(N, C)
        x, z = ws.xz[:, :C], ws.xz[:, C:]
        ops.eltwise(ops.EW_COPY, proj, None, x, stream=st)
        ops.eltwise(ops.EW_FILL, None, None, z, scalar=0.0, stream=st)        # tgt_id = 0 (transformer.py:603)
        is_ref = id_emb is not None
        if is_ref:
            stK, stV = self._next_short_slot()
        else:
            stK, stV = self.st_K, self.st_V
        for li in range(P.L):
            Lw = P.layers[li]
            cQ, cV = self.curr_Q[li], self.curr_V[li]
            ops.layernorm(x, Lw.norm1[0], Lw.norm1[1], ws.ln, stream=st)
            ops.linear(ws.ln, Lw.qv_w, Lw.qv_b, ws.qv, stream=st)
            ops.eltwise(ops.EW_COPY, ws.qv[:, :d], None, cQ, stream=st)
            ops.eltwise(ops.EW_SILU, ws.qv[:, d:], None, cV, stream=st)          # curr_V = silu(.) :599
            ops.linear(ws.ln, Lw.u_w, Lw.u_b, ws.catU[:, :C2], act=A_SILU, stream=st)
            if li == 0:
                ops.eltwise(ops.EW_FILL, None, None, ws.catU[:, C2:], scalar=1.0, stream=st)   # :604-605
                cIDV = None
            else:
                cIDV = self.curr_IDV[li]
                ops.layernorm(z, Lw.id_norm1[0], Lw.id_norm1[1], cIDV, stream=st)
                ops.linear(cIDV, Lw.idu_w, Lw.idu_b, ws.catU[:, C2:], act=A_SILU, stream=st)    # :610-611
            if is_ref:
                ops.eltwise(ops.EW_COPY, cQ, None, stK[li], stream=st)
                ops.eltwise(ops.EW_COPY, cV, None, stV[li][:, :C2], stream=st)
                self._fuse_id(li, cIDV, id_emb, stV[li][:, C2:], st)
                gK, gV, Tk = stK[li], stV[li], N
            else:
                gK, gV, Tk = self.bank_K[li], self.bank_V[li], self.bank_len
            probe = LT_PROBE if not is_ref else None
            if probe is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            if self._gp_tc and not is_ref:
                # fused tcgen05 kernel (gp_attn_tc.cu): 128 queries x 128 value channels per CTA, KV splits to fill the GPU
                self._gp_attention(cQ, None, None, self.bank_gpK[li], self.bank_gpV[li], Tk, self.tk_dev, ws.core, st)
            elif self._gp_tc:
                self._gp_attention(cQ, gK, gV, None, None, N, None, ws.core, st)     # reference frame: its own K / V
            elif self._gemm_lt and not is_ref:
                # S = Q K^T -> softmax(S / T) over the live keys -> P V, all on the tensor-core GEMM (deaot_lt.cu)
                ops.linear_tc(cQ, self.bank_Kh[li], self.bank_Kl[li], None, ws.S, stream=st)
                ops.row_softmax(ws.S, self._capw, Tk, 1.0 / math.sqrt(d), Tk_dev=self.tk_dev, stream=st)
                ops.linear_tc(ws.S, self.bank_VhT[li], self.bank_VlT[li], None, ws.core, stream=st)
            else:
                ops.attention(cQ, gK, gV, ws.core, 1, d, C4, Tk=Tk, Tk_dev=None if is_ref else self.tk_dev, stream=st)
            if probe is not None:
                e1.record()
                probe.append((e0, e1, 2.0 * N * Tk * (d + C4)))      # FLOPs = 2*N*Tk*(d_qk + d_v), SURVEY 8d
            self._gated_tail(ws.core, ws.catU, Lw.lt_dw, ws.dw[:, :C4], h, w, st)
            if LOCAL_IMPL == "tile" and d == 128 and C4 == 1024:
                ops.local_gated_tile(cQ, stK[li], stV[li], Lw.relk_w, Lw.relk_b, ws.core, h, w, stream=st)
            else:
                ops.local_attention(cQ, stK[li], stV[li], Lw.relk_w, Lw.relk_b, None, ws.core, h, w, 1, d, C4, stream=st)
            self._gated_tail(ws.core, ws.catU, Lw.st_dw, ws.dw[:, C4:], h, w, st)
            # [tgt | tgt_id] += proj_lt(.) + proj_st(.)   (transformer.py:633-641) as one K = 8C GEMM
            ops.linear(ws.dw, Lw.lst_proj_w, Lw.lst_proj_b, ws.xz, res=ws.xz, stream=st)
            # gated self-attention on both streams (transformer.py:644-653, attention.py:648-669)
            ops.layernorm(x, Lw.norm2[0], Lw.norm2[1], ws.c[:, :C], stream=st)
            ops.layernorm(z, Lw.id_norm2[0], Lw.id_norm2[1], ws.c[:, C:], stream=st)
            ops.linear(ws.c, Lw.sa_qk_w, Lw.sa_qk_b, ws.sa_qk, stream=st)
            ops.linear(ws.c[:, :C], Lw.sa_v1[0], Lw.sa_v1[1], ws.sa_v[:, :C2], act=A_SILU, stream=st)
            ops.linear(ws.c[:, C:], Lw.sa_v2[0], Lw.sa_v2[1], ws.sa_v[:, C2:], act=A_SILU, stream=st)
            ops.linear(ws.c[:, :C], Lw.sa_u1[0], Lw.sa_u1[1], ws.sa_u[:, :C2], act=A_SILU, stream=st)
            ops.linear(ws.c[:, C:], Lw.sa_u2[0], Lw.sa_u2[1], ws.sa_u[:, C2:], act=A_SILU, stream=st)
            if self._gp_tc:
                self._gp_attention(ws.sa_qk, ws.sa_qk, ws.sa_v, None, None, N, None, ws.core, st)
            else:
                ops.attention(ws.sa_qk, ws.sa_qk, ws.sa_v, ws.core, 1, d, C4, Tk=N, stream=st)
            self._gated_tail(ws.core, ws.sa_u, Lw.sa_dw, ws.dw[:, :C4], h, w, st)
            ops.linear(ws.dw[:, :C4], Lw.sa_proj_w, Lw.sa_proj_b, ws.xz, res=ws.xz, stream=st)
        # final GroupNorm1D(2C, groups=2) (transformer.py:197-200,241) -> decoder input
        ops.groupnorm(ws.xz.view(1, N, C2), P.final_gn[0], P.final_gn[1], ws.cat.view(1, N, C2), 2, A_NONE, ws.gn_ws,
                      stream=st)
        if is_ref:
            self._commit_short_slot(stK, stV)
        self._have_lstt = True

    def _gp_attention(self, Q, K, V, Kp, Vp, Tk, Tk_dev, out, st):
        """softmax(Q K^T / T) V for the DeAOT head shape (1 x 128 / 1024) on the fused tcgen05 kernel.  Q fp32 [N, 128] is
        packed (with the 1/T of attention.py:672) here; K / V fp32 of the current frame are packed unless already-packed
        bank copies (Kp, Vp) are given."""
        ws = self._ws
        N = self.enc_hw
        ops.tc_pack_rows(Q, ws.gpQp, 0, div=math.sqrt(self._kdim), stream=st)
        if Kp is None:
            ops.tc_pack_rows(K, ws.gpSaK, 0, stream=st)
            ops.tc_pack_rows(V, ws.gpSaV, 0, stream=st)
            Kp, Vp = ws.gpSaK, ws.gpSaV
        splits = self._gp_splits(Tk)
        part = None
        if splits > 1:
            part = ws.gp_part.get(splits)
            if part is None:
                fz = lambda *s: torch.empty(s, dtype=torch.float32, device=out.device)
                part = ws.gp_part[splits] = (fz(splits, N, out.shape[1]), fz(splits, 1, N), fz(splits, 1, N))
        ops.gp_attention_tc(ws.gpQp, Kp, Vp, N, Tk, O=out, Tk_dev=Tk_dev, splits=splits, exact=True, part=part, stream=st)

    def _gp_splits(self, Tk):
        """KV-split count of the fused DeAOT long-term attention kernel (one CTA = 128 queries x 128 value channels x one
        split): about two waves of CTAs, at least four 64-key tiles per split."""
        base = ((self.enc_hw + 127) // 128) * (4 * self._plan().C // 128)
        tiles = (max(Tk, 1) + 63) // 64
        return max(1, min(tiles // 4 if tiles >= 8 else 1, max(1, (2 * 148) // base)))

    def _fuse_id(self, li, cIDV, id_emb, out, st):
        """GatedPropagationModule.fuse_key_value_id (transformer.py:659-665)."""
        Lw = self._plan().layers[li]
        ws = self._ws
        C = self._plan().C
        if cIDV is None:
            ops.linear(id_emb, Lw.idv_w, Lw.idv_b, out, act=A_SILU, stream=st)
        else:
            ops.eltwise(ops.EW_COPY, cIDV, None, ws.idin[:, :C], stream=st)
            ops.eltwise(ops.EW_COPY, id_emb, None, ws.idin[:, C:], stream=st)
            ops.linear(ws.idin, Lw.idv_w, Lw.idv_b, out, act=A_SILU, stream=st)

    def _fuse_memories(self, id_emb, st):
        """deaot_engine.py:20-45: K, V unchanged; ID_V = fuse_key_value_id(None, curr_ID_V, id_emb)."""
        P = self._plan()
        C2 = 2 * P.C
        K, V = self._next_short_slot()
        for li in range(P.L):
            ops.eltwise(ops.EW_COPY, self.curr_Q[li], None, K[li], stream=st)
            ops.eltwise(ops.EW_COPY, self.curr_V[li], None, V[li][:, :C2], stream=st)
            self._fuse_id(li, self.curr_IDV[li], id_emb, V[li][:, C2:], st)
        self._commit_short_slot(K, V, reset=False)


# =====================================================================================
# multi-object facade (aot_engine.py:485-635)
# =====================================================================================
class AOTInferEngine(nn.Module):
    _engine_cls = AOTEngine

    def __init__(self, aot_model, gpu_id=0, long_term_mem_gap=9999, short_term_mem_skip=1, max_aot_obj_num=None):
        super().__init__()
        self.cfg = aot_model.cfg
        self.AOT = aot_model
        if max_aot_obj_num is None or max_aot_obj_num > aot_model.max_obj_num:
            self.max_aot_obj_num = aot_model.max_obj_num
        else:
            self.max_aot_obj_num = max_aot_obj_num
        self.gpu_id = gpu_id
        self.long_term_mem_gap = long_term_mem_gap
        self.short_term_mem_skip = short_term_mem_skip
        self.aot_engines = []
        self._kv_shard = None
        self.restart_engine()

    def enable_kv_sharding(self, rank, world, group=None):
        """Shard the long-term memory bank over `world` ranks (see AOTEngine.enable_kv_sharding)."""
        self._kv_shard = (rank, world, group)
        for e in self.aot_engines:
            e.enable_kv_sharding(rank, world, group)

    def restart_engine(self):
        # keep the engines (and their device buffers) across videos; just reset their state
        self._pool = getattr(self, "_pool", []) + list(self.aot_engines)
        self.aot_engines = []
        self.obj_nums = None

    # ------------------------------------------------------------------ > max_aot_obj_num objects
    # ceil(objects / 10) sub-engines share one encoder pass.  Their LSTT / decoder / memory-update work is independent, so
    # every sub-engine after the first runs on its own side stream (forked from and joined to the caller's stream inside
    # each protocol call) and the small per-engine kernels overlap on the GPU instead of queueing behind one another as in
    # the reference's Python loop (aot_engine.py:584-630); mask separation and logit aggregation are one kernel each.
    def _engine_streams(self):
        n = len(self.aot_engines)
        pool = getattr(self, "_side_streams", [])
        while len(pool) < n - 1:
            pool.append(torch.cuda.Stream())
        self._side_streams = pool
        return pool[: n - 1]

    def _run_engines(self, fn):
        """fn(index, engine) for every sub-engine; engine 0 on the current stream, the others concurrently on side streams."""
        engines = self.aot_engines
        if len(engines) == 1 or not SUB_ENGINE_STREAMS:
            return [fn(i, e) for i, e in enumerate(engines)]
        cur = torch.cuda.current_stream()
        side = self._engine_streams()
        for s in side:
            s.wait_stream(cur)                      # fork: everything queued so far (shared encoding, masks) is visible
        outs = [None] * len(engines)
        for i in range(1, len(engines)):
            with torch.cuda.stream(side[i - 1]):
                outs[i] = fn(i, engines[i])
        outs[0] = fn(0, engines[0])
        for s in side:
            cur.wait_stream(s)                      # join
        return outs

    def separate_mask(self, mask, obj_nums):
        """aot_engine.py:515-545 -> (per-engine masks, per-engine object counts).  Label maps go through one kernel
        (ids [10e+1, 10e+10] -> 1..10 for engine e); the probability form ([K, ...] foreground stack) slices channels."""
        n = len(self.aot_engines)
        if mask is None:
            return [None] * n
        if n == 1:
            return [mask], [obj_nums]
        per = self.max_aot_obj_num
        counts = [per] * n
        if obj_nums % per > 0:
            counts[-1] = obj_nums % per
        if mask.dim() == 3 or mask.shape[0] == 1:
            m = mask.float().contiguous()
            out = torch.empty((n,) + tuple(m.shape), dtype=torch.float32, device=m.device)
            ops.separate_labels(m, out, per)
            return [out[e] for e in range(n)], counts
        probs = []
        for e in range(n):
            fg = mask[e * per + 1:(e + 1) * per + 1]
            probs.append(torch.cat([1. - fg.sum(dim=1, keepdim=True), fg], dim=1))
        return probs, counts

    def soft_logit_aggregation(self, all_logits):
        """aot_engine.py:565-582: identity for one engine, otherwise the fused aggregation kernel."""
        if len(all_logits) == 1:
            return all_logits[0]
        per = self.max_aot_obj_num
        first = all_logits[0]
        key = (len(all_logits), tuple(first.shape[-2:]))
        buf = getattr(self, "_agg_out", None)
        if buf is None or buf[0] != key:
            buf = self._agg_out = (key, torch.empty((1, 1 + len(all_logits) * per) + tuple(first.shape[-2:]),
                                                    dtype=torch.float32, device=first.device))
        return ops.soft_logit_aggregation([t if t.is_contiguous() else t.contiguous() for t in all_logits], buf[1], per)

    def add_reference_frame(self, img, mask, obj_nums, frame_step=-1):
        if isinstance(obj_nums, list):
            obj_nums = obj_nums[0]
        self.obj_nums = obj_nums
        want = max(-(-int(obj_nums) // self.max_aot_obj_num), 1)          # ceil(objects / max_aot_obj_num), at least one
        while len(self.aot_engines) < want:
            eng = self._pool.pop(0) if self._pool else self._engine_cls(self.AOT, self.gpu_id, self.long_term_mem_gap,
                                                                       self.short_term_mem_skip)
            eng.restart_engine()
            eng.eval()
            if self._kv_shard is not None:
                eng.enable_kv_sharding(*self._kv_shard)
            self.aot_engines.append(eng)
        masks, counts = self.separate_mask(mask, obj_nums)
        # engine 0 encodes the frame; the others reuse its feature maps (aot_engine.py:596-607)
        first = self.aot_engines[0]
        first.add_reference_frame(img, masks[0], obj_nums=[counts[0]], frame_step=frame_step)
        embs = first.curr_enc_embs
        if len(self.aot_engines) > 1:
            self._run_engines(lambda i, e: None if i == 0 else e.add_reference_frame(
                img, masks[i], obj_nums=[counts[i]], frame_step=frame_step, img_embs=embs))
        self.update_size()

    def match_propogate_one_frame(self, img=None):
        first = self.aot_engines[0]
        if len(self.aot_engines) == 1:
            return first.match_propogate_one_frame(img)
        first._check_img(img)
        embs = first._encode(img, torch.cuda.current_stream().cuda_stream)       # shared by all sub-engines
        self._run_engines(lambda i, e: e.match_propogate_one_frame(img, img_embs=embs))

    def decode_current_logits(self, output_size=None):
        all_logits = self._run_engines(lambda i, e: e.decode_current_logits(output_size))
        return self.soft_logit_aggregation(all_logits)

    def update_memory(self, curr_mask, skip_long_term_update=False):
        masks, _ = self.separate_mask(curr_mask, self.obj_nums)
        self._run_engines(lambda i, e: e.update_short_term_memory(masks[i], skip_long_term_update=skip_long_term_update))

    # BASELINE.json's north_star names this method; the reference's real name is update_memory
    update_short_long_term_memory = update_memory

    def update_size(self):
        self.input_size_2d = self.aot_engines[0].input_size_2d
        self.enc_size_2d = self.aot_engines[0].enc_size_2d
        self.enc_hw = self.aot_engines[0].enc_hw

    @property
    def pred_id_logits(self):
        return self.aot_engines[0].pred_id_logits if self.aot_engines else None


class DeAOTInferEngine(AOTInferEngine):
    _engine_cls = DeAOTEngine


_ENGINES = {("aotengine", "train"): AOTEngine, ("aotengine", "eval"): AOTInferEngine,
            ("deaotengine", "train"): DeAOTEngine, ("deaotengine", "eval"): DeAOTInferEngine}


def build_engine(name, phase='train', **kwargs):
    """networks/engines/__init__.py:5-21: same names, same keyword arguments, NotImplementedError for anything else."""
    cls = _ENGINES.get((name, phase))
    if cls is None:
        raise NotImplementedError
    return cls(**kwargs)
