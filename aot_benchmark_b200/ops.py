"""Thin tensor-level wrappers over the C ABI (include/aotb200.h).

PyTorch is used here only for device memory (tensors) and streams.  Every function launches
hand-written sm_100a kernels from libaotb200.so on ``stream`` (a raw cudaStream_t int, default:
torch's current stream) and raises ``AotbError`` on failure -- there is no eager fallback.

Layout conventions: activations are NHWC ``[B,H,W,C]`` or token matrices ``[rows, C]`` whose last
stride is 1; channel-sliced views are fine (the row stride is passed as ``ld``).
"""
from __future__ import annotations

import os

import torch

from ._lib import AotbError, check, lib

ACT_NONE, ACT_RELU, ACT_GELU, ACT_SILU, ACT_RELU6 = 0, 1, 2, 3, 4
EW_COPY, EW_ADD, EW_MUL, EW_SILU, EW_SILU_MUL, EW_FILL = 0, 1, 2, 3, 4, 5


def _st(stream):
    return torch.cuda.current_stream().cuda_stream if stream is None else stream


def _chk(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda or t.dtype != torch.float32:
            raise AotbError(f"aot_benchmark_b200 kernels need float32 CUDA tensors, got {t.dtype} on {t.device}")
        if t.dim() >= 1 and t.stride(-1) != 1:
            raise AotbError("innermost stride must be 1")


def _p(t):
    return None if t is None else t.data_ptr()


def _nhwc_ld(x):
    B, H, W, C = x.shape
    ld = x.stride(2)
    if x.stride(1) != W * ld or (B > 1 and x.stride(0) != H * W * ld):
        raise AotbError("NHWC tensor must be dense over pixels")
    return ld


# fp32 weight (by data_ptr) -> (wh, wl) split-fp16 [Cout, K] copies for the tensor-core conv (registered by plan.py)
_TC_WEIGHTS = {}
CONV_IMPL = os.environ.get("AOTB_CONV_IMPL", "tc")     # "tc" (tcgen05, fp16x2 split) | "simt" (fp32 CUDA cores)


_TC_WS = {}
# bench.py hook: when set to a list, every tensor-core conv / linear launch appends (start_event, end_event, flops) with
# flops = 2 * M * N * K (algorithmic: one multiply-add per weight per output) so the conv family's roofline is measured live
CONV_PROBE = None


class _ConvProbe:
    def __init__(self, flops, stream):
        self.on = CONV_PROBE is not None and stream in (None, torch.cuda.current_stream().cuda_stream)
        self.flops = flops

    def __enter__(self):
        if self.on:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if self.on:
            self.e1.record()
            CONV_PROBE.append((self.e0, self.e1, self.flops))


def _tc_workspace(device):
    """Per-device scratch handed to the tensor-core conv (reserved by the C-ABI; currently unused by the kernel)."""
    ws = _TC_WS.get(device)
    if ws is None:
        ws = _TC_WS[device] = torch.zeros(1 << 20, dtype=torch.uint8, device=device)
    return ws


def register_tc_weights(w, wh, wl):
    _TC_WEIGHTS[w.data_ptr()] = (wh, wl, w)     # keep w alive so the pointer key stays unique


def split_fp16(w_kn):
    """fp32 [K, N] -> (hi, lo) fp16 [N, Kpad] (K zero-padded to a multiple of 64) with hi + lo ~= w to 2^-22."""
    K = w_kn.shape[0]
    wt = w_kn.t().contiguous()
    if K % 64:
        wt = torch.nn.functional.pad(wt, (0, 64 - K % 64))
    hi = wt.half()
    lo = (wt - hi.float()).half()
    return hi.contiguous(), lo.contiguous()


# ---- a chain of tensor-core convs as one persistent dataflow kernel (csrc/conv_chain.cu)
CONV_CHAIN = os.environ.get("AOTB_CONV_CHAIN", "0") == "1"


def _chain_struct():
    import ctypes

    class ChainLayer(ctypes.Structure):
        _fields_ = [(n, ctypes.c_void_p) for n in ("inp", "wh", "wl", "bias", "res", "out")] + \
                   [(n, ctypes.c_int) for n in ("H", "W", "Cin", "ldin", "Cout", "ldout", "ldres", "KH", "KW", "stride", "pad",
                                                "act", "in_layer", "res_layer")]
    return ChainLayer


def conv_chain_layers(layers):
    """layers: list of dicts (x [1,H,W,Cin] NHWC view, w fp32 [K, Cout] registered for the tensor-core path, bias, out, res,
    KH, stride, pad, act, in_layer, res_layer) -> ctypes array of aotb_chain_layer (keeps the tensors alive via the dicts)."""
    CL = _chain_struct()
    arr = (CL * len(layers))()
    for i, l in enumerate(layers):
        x, out, res = l["x"], l["out"], l.get("res")
        t = _TC_WEIGHTS.get(l["w"].data_ptr())
        if t is None:
            raise AotbError("conv_chain: layer weights are not registered for the tensor-core path")
        _chk(x, l["bias"], out, res)
        B, H, W, Cin = x.shape
        if B != 1:
            raise AotbError("conv_chain: batch 1 only")
        a = arr[i]
        a.inp, a.wh, a.wl = x.data_ptr(), t[0].data_ptr(), t[1].data_ptr()
        a.bias, a.res, a.out = _p(l["bias"]), _p(res), out.data_ptr()
        a.H, a.W, a.Cin, a.ldin = H, W, Cin, _nhwc_ld(x)
        a.Cout, a.ldout, a.ldres = out.shape[3], _nhwc_ld(out), (_nhwc_ld(res) if res is not None else 0)
        a.KH = a.KW = l.get("KH", 1)
        a.stride, a.pad, a.act = l.get("stride", 1), l.get("pad", 0), l.get("act", ACT_NONE)
        a.in_layer, a.res_layer = l.get("in_layer", -1), l.get("res_layer", -1)
    return arr


def conv_chain_dump(layers):
    """Host-only: -> (work items [n, 8] = layer, m-tile, n-tile, dep_lo, dep_hi, k0, k1, split; layer table [L, 10] = M, BN,
    counters offset, producer / residual counters offset (-1: none), in_need, res_need, splits, partial counters offset, chunks)."""
    import ctypes
    arr = layers if not isinstance(layers, list) else conv_chain_layers(layers)
    n = len(arr)
    nbytes, ntiles, ncnt = ctypes.c_size_t(), ctypes.c_int(), ctypes.c_int()
    check(lib().aotb_conv_chain_plan(ctypes.addressof(arr), n, ctypes.addressof(nbytes), ctypes.addressof(ntiles),
                                     ctypes.addressof(ncnt)), "aotb_conv_chain_plan")
    tiles = (ctypes.c_int * (8 * ntiles.value))()
    lay = (ctypes.c_int * (10 * n))()
    check(lib().aotb_conv_chain_dump(ctypes.addressof(arr), n, ctypes.addressof(tiles), ntiles.value, ctypes.addressof(lay)),
          "aotb_conv_chain_dump")
    return [list(tiles[8 * i:8 * i + 8]) for i in range(ntiles.value)], [list(lay[10 * i:10 * i + 10]) for i in range(n)]


class ConvChain:
    """A built chain program: `run(stream)` clears the dependency counters and launches the persistent kernel."""

    def __init__(self, layers, device, stream=None):
        import ctypes
        self.layers = layers                                   # keeps every tensor of the program alive
        arr = conv_chain_layers(layers)
        self.n = len(layers)
        nbytes, ntiles, ncnt = ctypes.c_size_t(), ctypes.c_int(), ctypes.c_int()
        check(lib().aotb_conv_chain_plan(ctypes.addressof(arr), self.n, ctypes.addressof(nbytes), ctypes.addressof(ntiles),
                                         ctypes.addressof(ncnt)), "aotb_conv_chain_plan")
        self.ntiles, self.ncounters = ntiles.value, ncnt.value
        self.program = torch.zeros(nbytes.value + 256, dtype=torch.uint8, device=device)
        self._base = (self.program.data_ptr() + 255) // 256 * 256
        check(lib().aotb_conv_chain_build(ctypes.addressof(arr), self.n, self._base, nbytes.value, _st(stream)),
              "aotb_conv_chain_build")

    def run(self, stream=None):
        check(lib().aotb_conv_chain_run(self._base, self.n, self.ntiles, self.ncounters, _st(stream)), "aotb_conv_chain_run")

    def profile(self):
        """AOTB_CHAIN_PROF=1: per work item [picked up, inputs complete, accumulator complete, published] in ns (globaltimer)."""
        off = lib().aotb_conv_chain_prof_offset(self.n, self.ntiles, self.ncounters) + (self._base - self.program.data_ptr())
        torch.cuda.synchronize()
        return self.program[off:off + self.ntiles * 32].view(torch.int64).view(self.ntiles, 4).cpu()


def conv2d_tc(x, wh, wl, bias, out, res=None, KH=1, KW=1, stride=1, pad=0, act=ACT_NONE, stream=None):
    """Tensor-core conv: x [B,H,W,Cin] fp32, wh/wl [Cout, KH*KW*Cin] fp16."""
    _chk(x, bias, out, res)
    B, H, W, Cin = x.shape
    Cout = wh.shape[0]
    ws = _tc_workspace(x.device)
    with _ConvProbe(2.0 * out.shape[0] * out.shape[1] * out.shape[2] * Cout * KH * KW * Cin, stream):
        check(lib().aotb_conv2d_nhwc_tc(_p(x), wh.data_ptr(), wl.data_ptr(), _p(bias), _p(res), _p(out), B, H, W, Cin,
                                        _nhwc_ld(x), Cout, _nhwc_ld(out), _nhwc_ld(res) if res is not None else 0, KH, KW,
                                        stride, pad, act, ws.data_ptr(), ws.numel(), _st(stream)), "aotb_conv2d_nhwc_tc")
    return out


def conv2d(x, w, bias, out, res=None, KH=1, KW=1, stride=1, pad=0, dil=1, act=ACT_NONE, stream=None):
    """x [B,H,W,Cin], w [KH*KW*Cin, Cout], out [B,Ho,Wo,Cout] (+res like out)."""
    if CONV_IMPL == "tc" and dil == 1:
        t = _TC_WEIGHTS.get(w.data_ptr())
        if t is not None:
            return conv2d_tc(x, t[0], t[1], bias, out, res=res, KH=KH, KW=KW, stride=stride, pad=pad, act=act,
                             stream=stream)
    _chk(x, w, bias, out, res)
    B, H, W, Cin = x.shape
    Cout = w.shape[1]
    check(lib().aotb_conv2d_nhwc_f32(_p(x), _p(w), _p(bias), _p(res), _p(out), B, H, W, Cin, _nhwc_ld(x), Cout,
                                     _nhwc_ld(out), _nhwc_ld(res) if res is not None else 0, KH, KW, stride, pad,
                                     dil, act, _st(stream)), "aotb_conv2d_nhwc_f32")
    return out


def linear(x, wt, bias, out, res=None, act=ACT_NONE, stream=None):
    """x [M,K], wt [K,N], out [M,N] (+res [M,N]); res may alias out (in-place accumulate)."""
    if CONV_IMPL == "tc":
        t = _TC_WEIGHTS.get(wt.data_ptr())
        if t is not None:
            _chk(x, bias, out, res)
            M, K = x.shape
            N = wt.shape[1]
            ws = _tc_workspace(x.device)
            with _ConvProbe(2.0 * M * N * K, stream):
                check(lib().aotb_conv2d_nhwc_tc(_p(x), t[0].data_ptr(), t[1].data_ptr(), _p(bias), _p(res), _p(out), 1, M, 1,
                                                K, x.stride(0), N, out.stride(0), res.stride(0) if res is not None else 0,
                                                1, 1, 1, 0, act, ws.data_ptr(), ws.numel(), _st(stream)),
                      "aotb_conv2d_nhwc_tc")
            return out
    _chk(x, wt, bias, out, res)
    M, K = x.shape
    N = wt.shape[1]
    check(lib().aotb_linear_f32(_p(x), _p(wt), _p(bias), _p(res), _p(out), M, K, x.stride(0), N, out.stride(0),
                                res.stride(0) if res is not None else 0, act, _st(stream)), "aotb_linear_f32")
    return out


def linear_tc(x, wh, wl, bias, out, res=None, act=ACT_NONE, stream=None):
    """out[M][N] = act(x[M][K] @ W^T + bias + res) on the tensor-core GEMM with explicitly given split-fp16 weights
    wh / wl [N][K] (K % 64 == 0, N % 64 == 0) -- e.g. operand copies of the memory bank."""
    _chk(x, bias, out, res)
    if wh.dtype != torch.float16 or wl.dtype != torch.float16 or wh.shape != wl.shape or not wh.is_contiguous() \
            or not wl.is_contiguous():
        raise AotbError("linear_tc: wh / wl must be contiguous fp16 tensors of the same shape [N, K]")
    M, K = x.shape
    N = wh.shape[0]
    if wh.shape[1] != K or K % 64 or N % 64 or out.shape[0] != M or out.shape[1] != N:
        raise AotbError(f"linear_tc: shapes x {tuple(x.shape)}, w {tuple(wh.shape)}, out {tuple(out.shape)}")
    ws = _tc_workspace(x.device)
    check(lib().aotb_conv2d_nhwc_tc(_p(x), wh.data_ptr(), wl.data_ptr(), _p(bias), _p(res), _p(out), 1, M, 1, K,
                                    x.stride(0), N, out.stride(0), res.stride(0) if res is not None else 0, 1, 1, 1, 0,
                                    act, ws.data_ptr(), ws.numel(), _st(stream)), "aotb_conv2d_nhwc_tc")
    return out


def split_rows(src, hi, lo, row_off=0, row_off_dev=None, stream=None):
    """src fp32 [rows, C] -> hi / lo fp16 [cap, ldw] rows [row_off, row_off + rows)."""
    _chk(src)
    rows, C = src.shape
    check(lib().aotb_split_rows_f16x2(_p(src), src.stride(0), hi.data_ptr(), lo.data_ptr(), hi.stride(0), rows, C,
                                      int(row_off), row_off_dev.data_ptr() if row_off_dev is not None else None,
                                      _st(stream)), "aotb_split_rows_f16x2")


def split_cols(src, hiT, loT, col_off=0, col_off_dev=None, stream=None):
    """src fp32 [rows, C] -> columns [col_off, col_off + rows) of hiT / loT fp16 [C, cap]."""
    _chk(src)
    rows, C = src.shape
    check(lib().aotb_split_cols_f16x2(_p(src), src.stride(0), hiT.data_ptr(), loT.data_ptr(), hiT.stride(0), rows, C,
                                      int(col_off), col_off_dev.data_ptr() if col_off_dev is not None else None,
                                      _st(stream)), "aotb_split_cols_f16x2")


def row_softmax(S, cols, Tk, scale, Tk_dev=None, stream=None):
    """In place on S [N, >= cols]: softmax(scale * S[r, :live]) in columns [0, live), zeros in [live, cols)."""
    _chk(S)
    check(lib().aotb_row_softmax_f32(_p(S), S.stride(0), S.shape[0], int(cols), int(Tk),
                                     Tk_dev.data_ptr() if Tk_dev is not None else None, float(scale), _st(stream)),
          "aotb_row_softmax_f32")
    return S


def nchw_to_nhwc(x, out, stream=None):
    _chk(x, out)
    B, C, H, W = x.shape
    check(lib().aotb_nchw_to_nhwc_f32(_p(x.contiguous()), _p(out), B, C, H * W, _st(stream)), "aotb_nchw_to_nhwc_f32")
    return out


def image_to_nhwc4(img, out, stream=None):
    """img [1,3,H,W] -> out [1,H,W,4] (4th channel zero)."""
    _chk(img, out)
    check(lib().aotb_image_to_nhwc4_f32(_p(img.contiguous()), _p(out), img.shape[2] * img.shape[3], _st(stream)),
          "aotb_image_to_nhwc4_f32")
    return out


def nhwc_to_nchw(x, out, stream=None):
    _chk(x, out)
    B, H, W, C = x.shape
    check(lib().aotb_nhwc_to_nchw_f32(_p(x), _p(out), B, C, H * W, _st(stream)), "aotb_nhwc_to_nchw_f32")
    return out


def maxpool3x3s2(x, out, stream=None):
    _chk(x, out)
    B, H, W, C = x.shape
    check(lib().aotb_maxpool3x3s2_nhwc_f32(_p(x), _p(out), B, H, W, C, _st(stream)), "aotb_maxpool3x3s2_nhwc_f32")
    return out


def dwconv(x, w, bias, out, K=5, stride=1, pad=2, dil=1, act=ACT_NONE, stream=None):
    """x [B,H,W,C], w [K*K, C]."""
    _chk(x, w, bias, out)
    B, H, W, C = x.shape
    check(lib().aotb_dwconv_nhwc_f32(_p(x), _p(w), _p(bias), _p(out), B, H, W, C, _nhwc_ld(x), _nhwc_ld(out), K, K,
                                     stride, pad, dil, act, _st(stream)), "aotb_dwconv_nhwc_f32")
    return out


def bilinear(x, out, align_corners, stream=None):
    _chk(x, out)
    B, H, W, C = x.shape
    check(lib().aotb_bilinear_nhwc_f32(_p(x), _p(out), B, H, W, C, out.shape[1], out.shape[2],
                                       1 if align_corners else 0, _st(stream)), "aotb_bilinear_nhwc_f32")
    return out


def eltwise(op, a, b, out, scalar=0.0, stream=None):
    """2-D strided element-wise op on [rows, cols] views."""
    _chk(a, b, out)
    rows, cols = out.shape
    check(lib().aotb_eltwise_f32(op, _p(a), a.stride(0) if a is not None else 0, _p(b),
                                 b.stride(0) if b is not None else 0, _p(out), out.stride(0), rows, cols,
                                 float(scalar), _st(stream)), "aotb_eltwise_f32")
    return out


def layernorm(x, gamma, beta, out, add=None, out2=None, stream=None):
    _chk(x, gamma, beta, out, add, out2)
    rows, C = x.shape
    check(lib().aotb_layernorm_f32(_p(x), x.stride(0), _p(gamma), _p(beta), _p(add),
                                   add.stride(0) if add is not None else 0, _p(out), out.stride(0), _p(out2),
                                   out2.stride(0) if out2 is not None else 0, rows, C, _st(stream)),
          "aotb_layernorm_f32")
    return out


def window_attention(qkv, qkv_bias, rel_bias, out, H, W, heads, shift, window=7, stream=None):
    """Swin (S)W-MSA core: qkv [H*W, 3C], rel_bias [heads, 49, 49], out [H*W, C]."""
    _chk(qkv, qkv_bias, rel_bias, out)
    C = out.shape[1]
    if qkv.shape[0] != H * W or qkv.shape[1] != 3 * C or out.shape[0] != H * W or not rel_bias.is_contiguous():
        raise AotbError("window_attention: qkv must be [H*W, 3C], out [H*W, C], rel_bias contiguous")
    check(lib().aotb_window_attention_f32(_p(qkv), qkv.stride(0), _p(qkv_bias), _p(rel_bias), _p(out), out.stride(0),
                                          H, W, C, heads, window, shift, _st(stream)), "aotb_window_attention_f32")
    return out


def patch_merge(x, out, H, W, stream=None):
    """x [H*W, C] -> out [ceil(H/2)*ceil(W/2), 4C] (PatchMerging gather)."""
    _chk(x, out)
    C = x.shape[1]
    if x.shape[0] != H * W or out.shape[0] != ((H + 1) // 2) * ((W + 1) // 2) or out.shape[1] != 4 * C:
        raise AotbError("patch_merge: shape mismatch")
    check(lib().aotb_patch_merge_f32(_p(x), x.stride(0), _p(out), out.stride(0), H, W, C, _st(stream)),
          "aotb_patch_merge_f32")
    return out


def groupnorm_workspace(B, G, device):
    n = lib().aotb_groupnorm_workspace_bytes(B, G)
    return torch.zeros((n + 7) // 8, dtype=torch.float64, device=device)      # zero: the launch counter lives in it


def groupnorm(x, gamma, beta, out, G, act, workspace, stream=None):
    """x/out [B, P, C] (any NHWC flattened over pixels)."""
    _chk(x, gamma, beta, out)
    B, P, C = x.shape
    check(lib().aotb_groupnorm_nhwc_f32(_p(x), x.stride(1), _p(gamma), _p(beta), _p(out), out.stride(1), B, P, C, G,
                                        act, workspace.data_ptr(), _st(stream)), "aotb_groupnorm_nhwc_f32")
    return out


def attention(Q, K, V, O, H, d_qk, d_v, Tk=None, Tk_dev=None, Mout=None, Lout=None, stream=None):
    """Q [N, H*d_qk], K [>=Tk, H*d_qk], V [>=Tk, H*d_v], O [N, H*d_v]."""
    _chk(Q, K, V, O, Mout, Lout)
    N = Q.shape[0]
    tk = K.shape[0] if Tk is None else int(Tk)
    check(lib().aotb_attention_f32(_p(Q), Q.stride(0), _p(K), K.stride(0), _p(V), V.stride(0), _p(O), O.stride(0),
                                   N, tk, Tk_dev.data_ptr() if Tk_dev is not None else None, H, d_qk, d_v,
                                   _p(Mout), _p(Lout), _st(stream)), "aotb_attention_f32")
    return O


def attn_merge(Opart, Mpart, Lpart, O, H, d_v, stream=None):
    _chk(Opart, Mpart, Lpart, O)
    R, N = Opart.shape[0], Opart.shape[1]
    check(lib().aotb_attn_merge_f32(_p(Opart), _p(Mpart), _p(Lpart), _p(O), R, N, H, d_v, O.stride(0), _st(stream)),
          "aotb_attn_merge_f32")
    return O


def attn_merge_peers(Oparts, Mparts, Lparts, O, splits, H, d_v, stream=None):
    """Merge split-KV partials that live in `len(Oparts)` ranks' buffers (local tensor + peer views of a symmetric-memory
    allocation): Oparts[r] [>=splits, N, H*d_v], Mparts[r] / Lparts[r] [>=splits, H, N]."""
    import ctypes
    _chk(O, *Oparts, *Mparts, *Lparts)
    R, N = len(Oparts), O.shape[0]
    arr = lambda ts: (ctypes.c_void_p * R)(*[t.data_ptr() for t in ts])
    check(lib().aotb_attn_merge_peers_f32(arr(Oparts), arr(Mparts), arr(Lparts), R, int(splits), _p(O), N, H, d_v,
                                          O.stride(0), _st(stream)), "aotb_attn_merge_peers_f32")
    return O


def local_attention(q, k, v, relk_w, relk_b, relv, out, h, w, H, d_att, d_v, stream=None):
    """q,k [hw, H*d_att], v [hw, H*d_v], out [hw, H*d_v]."""
    _chk(q, k, v, relk_w, relk_b, relv, out)
    check(lib().aotb_local_attention_f32(_p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(relk_w),
                                         _p(relk_b), _p(relv), _p(out), out.stride(0), h, w, H, d_att, d_v,
                                         _st(stream)), "aotb_local_attention_f32")
    return out


def local_gated_tile(q, k, v, relk_w, relk_b, out, h, w, stream=None):
    """DeAOT head shape (1 x 128 / 1024, no relative_emb_v): halo-in-shared-memory kernel; q, k [hw, 128], v / out [hw, 1024]."""
    _chk(q, k, v, relk_w, relk_b, out)
    if q.shape[1] != 128 or k.shape[1] != 128 or v.shape[1] != 1024 or out.shape[1] != 1024 or not relk_w.is_contiguous():
        raise AotbError("local_gated_tile: q / k [hw, 128], v / out [hw, 1024], contiguous relative_emb_k weights [225, 128]")
    check(lib().aotb_local_gated_tile_f32(_p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(relk_w), _p(relk_b),
                                          _p(out), out.stride(0), h, w, _st(stream)), "aotb_local_gated_tile_f32")
    return out


def local_attention_tile(q, k, v, relk_w, relk_b, relv_t, out, h, w, H, stream=None):
    """AOT head shape (d = 32): halo-in-shared-memory kernel; relv_t [H, 225, 32]."""
    _chk(q, k, v, relk_w, relk_b, relv_t, out)
    check(lib().aotb_local_attention_tile_f32(_p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(relk_w),
                                              _p(relk_b), _p(relv_t), _p(out), out.stride(0), h, w, H, _st(stream)),
          "aotb_local_attention_tile_f32")
    return out


def id_embed(mask, wt, bias, out, C, nid, ksize, stride, pad, ln_gamma=None, ln_beta=None, stream=None):
    """mask [Hm, Wm] float ids -> out [ho*wo, C]."""
    _chk(mask, wt, bias, out, ln_gamma, ln_beta)
    Hm, Wm = mask.shape
    check(lib().aotb_id_embed_f32(_p(mask), Hm, Wm, _p(wt), _p(bias), _p(ln_gamma), _p(ln_beta), _p(out),
                                  out.stride(0), C, nid, ksize, stride, pad, _st(stream)), "aotb_id_embed_f32")
    return out


def id_embed_runs(mask, wp, bias, out, C, nid, ksize, stride, pad, ln_gamma=None, ln_beta=None, stream=None):
    """Run-length form of id_embed; wp [ksize, ksize+1, nid, C] exclusive prefix sums along kx."""
    _chk(mask, wp, bias, out, ln_gamma, ln_beta)
    Hm, Wm = mask.shape
    check(lib().aotb_id_embed_runs_f32(_p(mask), Hm, Wm, _p(wp), _p(bias), _p(ln_gamma), _p(ln_beta), _p(out),
                                       out.stride(0), C, nid, ksize, stride, pad, _st(stream)), "aotb_id_embed_runs_f32")
    return out


def logits_postproc(logits_nhwc, lowres_nchw, out_nchw, obj_num, align_corners, stream=None):
    _chk(logits_nhwc, lowres_nchw, out_nchw)
    h, w, NC = logits_nhwc.shape[-3:]
    Ho, Wo = (out_nchw.shape[-2], out_nchw.shape[-1]) if out_nchw is not None else (0, 0)
    check(lib().aotb_logits_postproc_f32(_p(logits_nhwc), _p(lowres_nchw), _p(out_nchw), h, w, NC, obj_num, Ho, Wo,
                                         1 if align_corners else 0, _st(stream)), "aotb_logits_postproc_f32")
    return out_nchw


def logits_argmax(lowres_nchw, label, align_corners, stream=None):
    _chk(lowres_nchw, label)
    NC, h, w = lowres_nchw.shape[-3:]
    Ho, Wo = label.shape[-2:]
    check(lib().aotb_logits_argmax_f32(_p(lowres_nchw), _p(label), h, w, NC, Ho, Wo, 1 if align_corners else 0,
                                       _st(stream)), "aotb_logits_argmax_f32")
    return label


def soft_logit_aggregation(logits, out, max_obj, stream=None):
    """logits: list of E contiguous NCHW maps [1, 1 + max_obj, H, W]; out [1, 1 + E * max_obj, H, W]."""
    import ctypes
    _chk(out, *logits)
    E = len(logits)
    HW = out.shape[-2] * out.shape[-1]
    for t in logits:
        if not t.is_contiguous() or t.shape[1] != 1 + max_obj or t.shape[-2] * t.shape[-1] != HW:
            raise AotbError("soft_logit_aggregation: logit maps must be contiguous [1, 1 + max_obj, H, W] of one size")
    if not out.is_contiguous() or out.shape[1] != 1 + E * max_obj:
        raise AotbError("soft_logit_aggregation: out must be contiguous [1, 1 + E * max_obj, H, W]")
    arr = (ctypes.c_void_p * E)(*[t.data_ptr() for t in logits])
    check(lib().aotb_soft_logit_aggregation_f32(arr, E, int(max_obj), _p(out), HW, _st(stream)),
          "aotb_soft_logit_aggregation_f32")
    return out


def separate_labels(mask, out, max_obj, stream=None):
    """mask: contiguous label map with HW elements; out [E, ...HW...] receives the per-engine renumbered label maps."""
    _chk(mask, out)
    E, HW = out.shape[0], mask.numel()
    if not mask.is_contiguous() or not out.is_contiguous() or out.numel() != E * HW:
        raise AotbError("separate_labels: mask [HW] and out [E, HW] must be contiguous")
    check(lib().aotb_separate_labels_f32(_p(mask), E, int(max_obj), _p(out), HW, _st(stream)), "aotb_separate_labels_f32")
    return out


def preprocess_bgr_u8(img_u8, out, taps=None, flip=False, stream=None):
    """img_u8 uint8 [H, W, 3] (device); out fp32 [1, 3, Ho, Wo]; taps = (ix, cx, iy, cy) device tables or None (same size)."""
    if img_u8.dtype != torch.uint8 or not img_u8.is_cuda or not img_u8.is_contiguous() or img_u8.dim() != 3 or img_u8.shape[2] != 3:
        raise AotbError("preprocess_bgr_u8: image must be a contiguous uint8 CUDA tensor [H, W, 3]")
    _chk(out)
    if not out.is_contiguous():
        raise AotbError("preprocess_bgr_u8: out must be contiguous [1, 3, Ho, Wo]")
    H, W = int(img_u8.shape[0]), int(img_u8.shape[1])
    Ho, Wo = int(out.shape[-2]), int(out.shape[-1])
    if taps is None:
        ptrs = (None, None, None, None)
    else:
        ix, cx, iy, cy = taps
        if ix.dtype != torch.int32 or iy.dtype != torch.int32 or cx.dtype != torch.float32 or cy.dtype != torch.float32 \
                or tuple(ix.shape) != (Wo, 4) or tuple(cx.shape) != (Wo, 4) or tuple(iy.shape) != (Ho, 4) or tuple(cy.shape) != (Ho, 4):
            raise AotbError("preprocess_bgr_u8: taps must be (int32 [Wo,4], fp32 [Wo,4], int32 [Ho,4], fp32 [Ho,4])")
        ptrs = tuple(t.data_ptr() for t in (ix, cx, iy, cy))
    check(lib().aotb_preprocess_bgr_u8(img_u8.data_ptr(), H, W, ptrs[0], ptrs[1], ptrs[2], ptrs[3], _p(out), Ho, Wo,
                                       1 if flip else 0, _st(stream)), "aotb_preprocess_bgr_u8")
    return out


def label_to_u8(label, out_u8, stream=None):
    _chk(label)
    if out_u8.dtype != torch.uint8 or not out_u8.is_cuda or not out_u8.is_contiguous() or not label.is_contiguous() \
            or out_u8.numel() != label.numel():
        raise AotbError("label_to_u8: contiguous float32 label and uint8 output of the same size on the device")
    check(lib().aotb_label_to_u8(_p(label), out_u8.data_ptr(), label.numel(), _st(stream)), "aotb_label_to_u8")
    return out_u8


def nearest_resize(x, out, stream=None):
    _chk(x, out)
    H, W = x.shape[-2:]
    Ho, Wo = out.shape[-2:]
    check(lib().aotb_nearest_resize_f32(_p(x), _p(out), H, W, Ho, Wo, _st(stream)), "aotb_nearest_resize_f32")
    return out


def bank_append(src, bank, offset, offset_dev=None, stream=None):
    _chk(src, bank)
    rows, cols = src.shape
    check(lib().aotb_bank_append_f32(_p(src), src.stride(0), _p(bank), bank.stride(0), rows, cols, int(offset),
                                     offset_dev.data_ptr() if offset_dev is not None else None, _st(stream)),
          "aotb_bank_append_f32")
    return bank


def counter_add(counter, delta, stream=None):
    """counter: int32 CUDA tensor [1]; *counter += delta on the stream (bank row counter)."""
    if not counter.is_cuda or counter.dtype != torch.int32:
        raise AotbError("counter must be an int32 CUDA tensor")
    check(lib().aotb_counter_add(counter.data_ptr(), int(delta), _st(stream)), "aotb_counter_add")
    return counter


# ------------------------------------------------------------------ tensor-core long-term attention
def tc_pack_rows(src, dst, row_off=0, div=1.0, row_off_dev=None, stream=None):
    """src fp32 [rows, H*32] -> dst fp16 [H, cap, 64] rows [row_off, row_off+rows) as [hi(32) | lo(32)]."""
    _chk(src)
    if dst.dtype != torch.float16 or not dst.is_cuda or not dst.is_contiguous():
        raise AotbError("packed operand buffer must be a contiguous fp16 CUDA tensor [H, cap, 64]")
    H, cap, _ = dst.shape
    rows = src.shape[0]
    check(lib().aotb_tc_pack_rows_f16x2(_p(src), src.stride(0), dst.data_ptr(), cap, rows, H, int(row_off),
                                        row_off_dev.data_ptr() if row_off_dev is not None else None, float(div),
                                        _st(stream)), "aotb_tc_pack_rows_f16x2")
    return dst


# softmax layout of the tensor-core attention kernel:
#   "tile"   all 16 softmax warps on one 128x128 score tile at a time (4 threads per row)
#   "groups" two groups of 8 warps, one per query tile, running out of phase (2 threads per row, 64 scores in registers)
#   "ahead"  three score buffers in TMEM: S runs one tile ahead of the softmax and the TMEM read of the next tile is
#            issued under the ex2 pass of the current one (bit-identical to "tile", 16 % slower: profiles/r02_summary.md)
#   "pair"   two co-resident CTAs per SM (128 queries, 64-key tiles, three score buffers each) whose softmax phases
#            interleave on the MUFU pipe; packed-fp32 softmax arithmetic (lt_attn_tc2.cu)
LT_VARIANT = os.environ.get("AOTB_LT_VARIANT", "tile")
# 1: the mbarrier waits on the softmax -> MMA -> softmax chain poll instead of sleeping with a suspend-time hint
LT_SPIN = os.environ.get("AOTB_LT_SPIN", "0") == "1"


def lt_attention_tc(Qp, Kp, Vp, N, Tk, O=None, Tk_dev=None, splits=1, exact=True, part=None, dbg=None, stream=None,
                    merge=True, variant=None):
    """Qp [H, Nq_cap, 64], Kp/Vp [H, kv_cap, 64] packed fp16x2; O [N, H*32] fp32.
    With splits > 1, `part` = (Opart [S,N,H*32], Mpart [S,H,N], Lpart [S,H,N]) and O receives the merge.
    `variant` (default: AOTB_LT_VARIANT) selects the softmax layout: "tile", "groups" or "ahead"."""
    v = LT_VARIANT if variant is None else variant
    if v not in ("tile", "groups", "ahead", "pair"):
        raise AotbError(f"unknown long-term attention variant '{v}' (tile | groups | ahead | pair)")
    mode = (1 if exact else 0) | (2 if v == "groups" else 0) | (4 if LT_SPIN else 0) | (8 if v == "ahead" else 0) | \
        (16 if v == "pair" else 0)
    H, nq_cap, _ = Qp.shape
    kv_cap = Kp.shape[1]
    if splits > 1:
        Op, Mp, Lp = part
    else:
        Op = Mp = Lp = None
    check(lib().aotb_lt_attn_tc_f16x2(Qp.data_ptr(), nq_cap, Kp.data_ptr(), Vp.data_ptr(), kv_cap, N, int(Tk),
                                      Tk_dev.data_ptr() if Tk_dev is not None else None, H,
                                      _p(O) if splits == 1 else None, O.stride(0) if O is not None else 0,
                                      _p(Op), _p(Mp), _p(Lp), splits, mode, _p(dbg), _st(stream)),
          "aotb_lt_attn_tc_f16x2")
    if splits > 1 and merge:
        attn_merge(Op, Mp, Lp, O, H, 32, stream=stream)
    return O


def gp_attention_tc(Qp, Kp, Vp, N, Tk, O=None, Tk_dev=None, splits=1, exact=True, part=None, stream=None, merge=True):
    """Fused DeAOT long-term attention (EXPERIMENTAL): Qp [4, Nq_cap, 64], Kp [4, kv_cap, 64], Vp [dv/32, kv_cap, 64] packed
    fp16x2 (one 'head' per 32 channels); O [N, dv] fp32.  With splits > 1, `part` = (Opart [S,N,dv], Mpart [S,1,N],
    Lpart [S,1,N]) and O receives the merge."""
    nq_cap, kv_cap, dv = Qp.shape[1], Kp.shape[1], Vp.shape[0] * 32
    if splits > 1:
        Op, Mp, Lp = part
    else:
        Op = Mp = Lp = None
    mode = (1 if exact else 0) | (4 if LT_SPIN else 0)
    check(lib().aotb_gp_attn_tc_f16x2(Qp.data_ptr(), nq_cap, Kp.data_ptr(), Vp.data_ptr(), kv_cap, N, int(Tk),
                                      Tk_dev.data_ptr() if Tk_dev is not None else None, dv,
                                      _p(O) if splits == 1 else None, O.stride(0) if O is not None else 0,
                                      _p(Op), _p(Mp), _p(Lp), splits, mode, _st(stream)), "aotb_gp_attn_tc_f16x2")
    if splits > 1 and merge:
        attn_merge(Op, Mp, Lp, O, 1, dv, stream=stream)
    return O
