// Identity-bank embedding (K4), logit post-processing (K9) and memory-bank append (K10).
//
// K4  one_hot_mask (utils/image.py:69-74) -> patch_wise_id_bank Conv2d(11->C, k17 s16 p8 | k16 s16 p0)
//     (networks/models/aot.py:50-63,76-79; aot_engine.py:168-179) [+ LayerNorm over C for DeAOT,
//     deaot.py:51-55].  A convolution of a one-hot image is a gather-sum of weight columns:
//        id[y,x,c] = b[c] + sum_{ky,kx in frame} W[c, mask[s*y+ky-p, s*x+kx-p], ky, kx]
//     so the [1,11,H,W] one-hot tensor (18 MB/frame at 480p) and 96 % of the conv FLOPs vanish.
// K9  aot_engine.py:367-378: ids > obj_num are set to -1e10 at stride-4 resolution, then
//     F.interpolate(bilinear, align_corners=cfg) to the output size (NCHW result for the caller).
// K10 aot_engine.py:291-305 re-copies the whole bank with torch.cat on every update; here the new
//     frame's rows are written in place into a pre-allocated bank (row order is append, not the
//     reference's prepend -- softmax attention is permutation invariant over keys).
#include "common.cuh"

namespace aotb {

// weights re-laid out as wt[(ky*KW + kx) * NID + id][C]
template <bool LN>
__global__ void __launch_bounds__(256) id_embed_kernel(const float* __restrict__ mask, int Hm, int Wm,
                                                       const float* __restrict__ wt, const float* __restrict__ bias,
                                                       const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                                                       float* __restrict__ out, int ldo, int ho, int wo, int C,
                                                       int NID, int KS, int stride, int pad) {
    pdl_sync();
    __shared__ int ids[17 * 17];
    __shared__ float red[2][8];
    const int pix = blockIdx.x;
    const int oy = pix / wo, ox = pix - oy * wo;
    for (int t = threadIdx.x; t < KS * KS; t += blockDim.x) {
        const int ky = t / KS, kx = t - ky * KS;
        const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
        int id = -1;
        if (iy >= 0 && iy < Hm && ix >= 0 && ix < Wm) {
            const float v = __ldg(mask + (size_t)iy * Wm + ix);
            const int iv = (int)v;
            if ((float)iv == v && iv >= 0 && iv < NID) id = iv;   // (mask == arange).float()
        }
        ids[t] = id;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float acc = 0.f;
        for (int t = 0; t < KS * KS; ++t) {
            const int id = ids[t];
            if (id >= 0) acc += __ldg(wt + ((size_t)t * NID + id) * C + c);
        }
        acc += __ldg(bias + c);
        if constexpr (!LN) out[(size_t)pix * ldo + c] = acc;
        else {
            // C == blockDim.x == 256 in the LN variant (checked on the host)
            float s = warp_sum(acc);
            const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
            if (lane == 0) red[0][wid] = s;
            __syncthreads();
            float tot = 0.f;
            for (int w2 = 0; w2 < 8; ++w2) tot += red[0][w2];
            const float mean = tot / (float)C;
            const float d = acc - mean;
            float q = warp_sum(d * d);
            if (lane == 0) red[1][wid] = q;
            __syncthreads();
            float tq = 0.f;
            for (int w2 = 0; w2 < 8; ++w2) tq += red[1][w2];
            const float rstd = rsqrtf(tq / (float)C + 1e-5f);
            out[(size_t)pix * ldo + c] = d * rstd * __ldg(ln_g + c) + __ldg(ln_b + c);
        }
    }
}

// Run-length variant: a label map is piecewise constant, so along each of the KS window rows the ids form a few runs.
// With wp[ky][j][id][c] = sum_{kx < j} W[c, id, ky, kx] (exclusive prefix along kx, built in float64 by the host) a run
// [a, b) of one id contributes wp[ky][b][id] - wp[ky][a][id]: ~2 table rows per run instead of one per tap
// (typically ~40 instead of 289 one-KB rows per output pixel).
template <bool LN>
__global__ void __launch_bounds__(256) id_embed_runs_kernel(const float* __restrict__ mask, int Hm, int Wm,
                                                            const float* __restrict__ wp, const float* __restrict__ bias,
                                                            const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                                                            float* __restrict__ out, int ldo, int ho, int wo, int C,
                                                            int NID, int KS, int stride, int pad) {
    pdl_sync();
    __shared__ int ids[17 * 17];
    __shared__ int run_pos[17 * 18];     // per row: up to KS runs, stored as (start | end<<8 | id<<16)
    __shared__ int run_cnt[17];
    __shared__ float red[2][8];
    const int pix = blockIdx.x;
    const int oy = pix / wo, ox = pix - oy * wo;
    for (int t = threadIdx.x; t < KS * KS; t += blockDim.x) {
        const int ky = t / KS, kx = t - ky * KS;
        const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
        int id = -1;
        if (iy >= 0 && iy < Hm && ix >= 0 && ix < Wm) {
            const float v = __ldg(mask + (size_t)iy * Wm + ix);
            const int iv = (int)v;
            if ((float)iv == v && iv >= 0 && iv < NID) id = iv;
        }
        ids[t] = id;
    }
    __syncthreads();
    if (threadIdx.x < KS) {
        const int ky = threadIdx.x;
        int n = 0, start = 0, cur = ids[ky * KS];
        for (int kx = 1; kx <= KS; ++kx) {
            const int v = (kx < KS) ? ids[ky * KS + kx] : -2;
            if (v != cur) {
                if (cur >= 0) run_pos[ky * 18 + n++] = start | (kx << 8) | (cur << 16);
                start = kx;
                cur = v;
            }
        }
        run_cnt[ky] = n;
    }
    __syncthreads();
    const int c = threadIdx.x;      // C == blockDim.x == 256 (checked on the host)
    float acc = 0.f;
    for (int ky = 0; ky < KS; ++ky) {
        const int n = run_cnt[ky];
        const float* row = wp + (size_t)ky * (KS + 1) * NID * C + c;
        for (int r = 0; r < n; ++r) {
            const int e = run_pos[ky * 18 + r];
            const int a = e & 0xff, b = (e >> 8) & 0xff, id = e >> 16;
            acc += __ldg(row + ((size_t)b * NID + id) * C) - __ldg(row + ((size_t)a * NID + id) * C);
        }
    }
    acc += __ldg(bias + c);
    if constexpr (!LN) out[(size_t)pix * ldo + c] = acc;
    else {
        float s = warp_sum(acc);
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) red[0][wid] = s;
        __syncthreads();
        float tot = 0.f;
        for (int w2 = 0; w2 < 8; ++w2) tot += red[0][w2];
        const float mean = tot / (float)C;
        const float d = acc - mean;
        float q = warp_sum(d * d);
        if (lane == 0) red[1][wid] = q;
        __syncthreads();
        float tq = 0.f;
        for (int w2 = 0; w2 < 8; ++w2) tq += red[1][w2];
        const float rstd = rsqrtf(tq / (float)C + 1e-5f);
        out[(size_t)pix * ldo + c] = d * rstd * __ldg(ln_g + c) + __ldg(ln_b + c);
    }
}

__device__ __forceinline__ void bl_src(int dst, int in_sz, int out_sz, int align, int& i0, int& i1, float& l1) {
    float src;
    if (align) {
        const float scale = out_sz > 1 ? (float)(in_sz - 1) / (float)(out_sz - 1) : 0.f;
        src = scale * dst;
    } else {
        const float scale = (float)in_sz / (float)out_sz;
        src = scale * (dst + 0.5f) - 0.5f;
        if (src < 0.f) src = 0.f;
    }
    i0 = (int)src;
    if (i0 > in_sz - 1) i0 = in_sz - 1;
    i1 = i0 + (i0 < in_sz - 1 ? 1 : 0);
    l1 = src - (float)i0;
}

// logits_nhwc [h][w][NC] -> lowres NCHW [NC][h][w] with ids > obj_num masked to -1e10
__global__ void logits_mask_kernel(const float* __restrict__ in, float* __restrict__ lo, int h, int w, int NC,
                                   int obj_num) {
    pdl_sync();
    const int total = NC * h * w;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i / (h * w), r = i - c * h * w;
        lo[i] = (c > obj_num) ? -1e10f : __ldg(in + (size_t)r * NC + c);
    }
}

// lowres NCHW [NC][h][w] -> out NCHW [NC][Ho][Wo], bilinear
__global__ void logits_upsample_kernel(const float* __restrict__ lo, float* __restrict__ out, int h, int w, int NC,
                                       int Ho, int Wo, int align) {
    pdl_sync();
    const size_t total = (size_t)NC * Ho * Wo;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int ox = i % Wo;
        const size_t t = i / Wo;
        const int oy = t % Ho, c = t / Ho;
        int y0, y1, x0, x1;
        float ly, lx;
        bl_src(oy, h, Ho, align, y0, y1, ly);
        bl_src(ox, w, Wo, align, x0, x1, lx);
        const float* b = lo + (size_t)c * h * w;
        const float hy = 1.f - ly, hx = 1.f - lx;
        out[i] = hy * (hx * b[y0 * w + x0] + lx * b[y0 * w + x1]) + ly * (hx * b[y1 * w + x0] + lx * b[y1 * w + x1]);
    }
}

// fused K9 fast path: bilinear upsample of the masked low-res logits + argmax over ids -> label map
// (evaluator.py:339-361 collapses to argmax(logits) for one engine without TTA).  First maximum
// wins on ties, like torch.argmax.
__global__ void logits_argmax_kernel(const float* __restrict__ lo, float* __restrict__ label, int h, int w, int NC,
                                     int Ho, int Wo, int align) {
    pdl_sync();
    const int total = Ho * Wo;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int oy = i / Wo, ox = i - oy * Wo;
        int y0, y1, x0, x1;
        float ly, lx;
        bl_src(oy, h, Ho, align, y0, y1, ly);
        bl_src(ox, w, Wo, align, x0, x1, lx);
        const float hy = 1.f - ly, hx = 1.f - lx;
        float best = -INFINITY;
        int bi = 0;
        for (int c = 0; c < NC; ++c) {
            const float* b = lo + (size_t)c * h * w;
            const float v =
                hy * (hx * b[y0 * w + x0] + lx * b[y0 * w + x1]) + ly * (hx * b[y1 * w + x0] + lx * b[y1 * w + x1]);
            if (v > best) { best = v; bi = c; }
        }
        label[i] = (float)bi;
    }
}

// nearest-neighbour resize of a label map (F.interpolate(mode='nearest'), evaluator.py:418-421):
// src = floor(dst * in / out)
__global__ void nearest_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, int Ho, int Wo) {
    pdl_sync();
    const int total = Ho * Wo;
    const float sy = (float)H / (float)Ho, sx = (float)W / (float)Wo;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int oy = i / Wo, ox = i - oy * Wo;
        int iy = (int)floorf(oy * sy), ix = (int)floorf(ox * sx);
        iy = iy < H - 1 ? iy : H - 1;
        ix = ix < W - 1 ? ix : W - 1;
        out[i] = in[(size_t)iy * W + ix];
    }
}


// soft_logit_aggregation (aot_engine.py:565-582) for E sub-engines of max_obj objects each, fused: per output pixel
//   prob_e = softmax over the 1 + max_obj channels of engine e;  bg = prod_e prob_e[0];
//   merged = clamp([bg, prob_0[1:], prob_1[1:], ...], 1e-5, 1 - 1e-5);  out = log(merged / (1 - merged))   (torch.logit)
// One thread per pixel: the E x (1 + max_obj) logits of a pixel are read once (channel planes are HW apart, so a warp reads
// 128 contiguous bytes per channel) and the 1 + E * max_obj merged logits written once.  The reference materialises E softmax
// tensors, two concatenations, a product, a clamp and a logit (7 passes over E x 18 MB at 480p).
struct AggArgs { const float* logits[8]; };

template <int NC>
__global__ void soft_logit_aggregation_kernel(const AggArgs a, int E, float* __restrict__ out, int HW) {
    pdl_sync();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        float bg = 1.f;
        for (int e = 0; e < E; ++e) {
            const float* l = a.logits[e] + i;
            float v[NC];
            float m = -INFINITY;
#pragma unroll
            for (int c = 0; c < NC; ++c) { v[c] = l[(size_t)c * HW]; m = fmaxf(m, v[c]); }
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < NC; ++c) { v[c] = expf(v[c] - m); sum += v[c]; }
            bg *= v[0] / sum;
            float* o = out + (size_t)(1 + e * (NC - 1)) * HW + i;
#pragma unroll
            for (int c = 1; c < NC; ++c) {
                const float p = fminf(fmaxf(v[c] / sum, 1e-5f), 1.f - 1e-5f);
                o[(size_t)(c - 1) * HW] = logf(p / (1.f - p));
            }
        }
        const float p = fminf(fmaxf(bg, 1e-5f), 1.f - 1e-5f);
        out[i] = logf(p / (1.f - p));
    }
}


// separate_mask for label maps (aot_engine.py:515-533): engine e keeps ids [e*max_obj + 1, (e+1)*max_obj], renumbered from 1,
// everything else becomes background.  One pass writes all E maps (the reference builds E boolean masks and 3 E temporaries).
__global__ void separate_labels_kernel(const float* __restrict__ mask, int E, int max_obj, float* __restrict__ out, int HW) {
    pdl_sync();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        const float m = mask[i];
        for (int e = 0; e < E; ++e) {
            const float lo = (float)(e * max_obj + 1), hi = (float)((e + 1) * max_obj);
            out[(size_t)e * HW + i] = (m >= lo && m <= hi) ? m - lo + 1.f : 0.f;
        }
    }
}

// rows x cols copy into bank at row offset (host value or device counter)
__global__ void bank_append_kernel(const float* __restrict__ src, int lds, float* __restrict__ bank, int ldb,
                                   int rows, int cols4, int offset, const int* __restrict__ offset_dev) {
    pdl_sync();
    const int off = offset_dev ? *offset_dev : offset;
    const size_t total = (size_t)rows * cols4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = i / cols4, c = (i - (size_t)r * cols4) * 4;
        *reinterpret_cast<float4*>(bank + (size_t)(off + r) * ldb + c) =
            *reinterpret_cast<const float4*>(src + (size_t)r * lds + c);
    }
}

__global__ void counter_add_kernel(int* ctr, int delta) {
    pdl_sync(); *ctr += delta; }

}  // namespace aotb

using namespace aotb;

extern "C" int aotb_id_embed_f32(const float* mask, int Hm, int Wm, const float* wt, const float* bias,
                                 const float* ln_gamma, const float* ln_beta, float* out, int ldo, int C, int nid,
                                 int ksize, int stride, int pad, void* stream) {
    AOTB_REQUIRE(mask && wt && bias && out && Hm > 0 && Wm > 0, "aotb_id_embed_f32: bad args");
    AOTB_REQUIRE(ksize <= 17 && ksize > 0 && stride > 0, "aotb_id_embed_f32: kernel size > 17");
    const int ho = (Hm + 2 * pad - ksize) / stride + 1, wo = (Wm + 2 * pad - ksize) / stride + 1;
    AOTB_REQUIRE(ho > 0 && wo > 0, "aotb_id_embed_f32: empty output");
    cudaStream_t st = (cudaStream_t)stream;
    if (ln_gamma) {
        AOTB_REQUIRE(C == 256 && ln_beta, "aotb_id_embed_f32: fused LayerNorm needs C == 256");
        launch(id_embed_kernel<true>, dim3(ho * wo), dim3(256), 0, st, mask, Hm, Wm, wt, bias, ln_gamma, ln_beta, out, ldo, ho, wo, C,
                                                       nid, ksize, stride, pad);
    } else {
        launch(id_embed_kernel<false>, dim3(ho * wo), dim3(256), 0, st, mask, Hm, Wm, wt, bias, nullptr, nullptr, out, ldo, ho, wo, C,
                                                        nid, ksize, stride, pad);
    }
    return check_launch("aotb_id_embed_f32");
}

// wp: exclusive prefix sums of the ID-bank weights along kx: [KS][KS+1][nid][C] (C must be 256).
extern "C" int aotb_id_embed_runs_f32(const float* mask, int Hm, int Wm, const float* wp, const float* bias,
                                      const float* ln_gamma, const float* ln_beta, float* out, int ldo, int C, int nid,
                                      int ksize, int stride, int pad, void* stream) {
    AOTB_REQUIRE(mask && wp && bias && out && Hm > 0 && Wm > 0, "aotb_id_embed_runs_f32: bad args");
    AOTB_REQUIRE(ksize <= 17 && ksize > 0 && stride > 0 && C == 256 && nid < 128, "aotb_id_embed_runs_f32: unsupported shape");
    const int ho = (Hm + 2 * pad - ksize) / stride + 1, wo = (Wm + 2 * pad - ksize) / stride + 1;
    AOTB_REQUIRE(ho > 0 && wo > 0, "aotb_id_embed_runs_f32: empty output");
    cudaStream_t st = (cudaStream_t)stream;
    if (ln_gamma) {
        AOTB_REQUIRE(ln_beta, "aotb_id_embed_runs_f32: ln_beta");
        launch(id_embed_runs_kernel<true>, dim3(ho * wo), dim3(256), 0, st, mask, Hm, Wm, wp, bias, ln_gamma, ln_beta, out, ldo, ho, wo,
                                                            C, nid, ksize, stride, pad);
    } else {
        launch(id_embed_runs_kernel<false>, dim3(ho * wo), dim3(256), 0, st, mask, Hm, Wm, wp, bias, nullptr, nullptr, out, ldo, ho, wo,
                                                             C, nid, ksize, stride, pad);
    }
    return check_launch("aotb_id_embed_runs_f32");
}

extern "C" int aotb_logits_postproc_f32(const float* logits_nhwc, float* lowres_nchw, float* out_nchw, int h, int w,
                                        int NC, int obj_num, int Ho, int Wo, int align_corners, void* stream) {
    AOTB_REQUIRE(logits_nhwc && lowres_nchw && h > 0 && w > 0 && NC > 0, "aotb_logits_postproc_f32: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    launch(logits_mask_kernel, dim3(cdiv(NC * h * w, 256)), dim3(256), 0, st, logits_nhwc, lowres_nchw, h, w, NC, obj_num);
    if (out_nchw) {
        AOTB_REQUIRE(Ho > 0 && Wo > 0, "aotb_logits_postproc_f32: bad output size");
        const size_t total = (size_t)NC * Ho * Wo;
        int g = (int)((total + 255) / 256);
        if (g > 148 * 16) g = 148 * 16;
        launch(logits_upsample_kernel, dim3(g), dim3(256), 0, st, lowres_nchw, out_nchw, h, w, NC, Ho, Wo, align_corners);
    }
    return check_launch("aotb_logits_postproc_f32", out_nchw ? 2 : 1);
}

extern "C" int aotb_logits_argmax_f32(const float* lowres_nchw, float* label, int h, int w, int NC, int Ho, int Wo,
                                      int align_corners, void* stream) {
    AOTB_REQUIRE(lowres_nchw && label && h > 0 && w > 0 && NC > 0 && Ho > 0 && Wo > 0,
                 "aotb_logits_argmax_f32: bad args");
    launch(logits_argmax_kernel, dim3(cdiv(Ho * Wo, 256)), dim3(256), 0, (cudaStream_t)stream, lowres_nchw, label, h, w, NC, Ho, Wo,
                                                                               align_corners);
    return check_launch("aotb_logits_argmax_f32");
}

extern "C" int aotb_nearest_resize_f32(const float* in, float* out, int H, int W, int Ho, int Wo, void* stream) {
    AOTB_REQUIRE(in && out && H > 0 && W > 0 && Ho > 0 && Wo > 0, "aotb_nearest_resize_f32: bad args");
    launch(nearest_kernel, dim3(cdiv(Ho * Wo, 256)), dim3(256), 0, (cudaStream_t)stream, in, out, H, W, Ho, Wo);
    return check_launch("aotb_nearest_resize_f32");
}


// logits: E device pointers to NCHW fp32 maps [1 + max_obj][HW] (the sub-engines' upsampled logits); out [1 + E*max_obj][HW].
extern "C" int aotb_soft_logit_aggregation_f32(const float* const* logits, int n_engines, int max_obj, float* out, int HW,
                                               void* stream) {
    AOTB_REQUIRE(logits && out && n_engines >= 1 && n_engines <= 8 && HW > 0, "aotb_soft_logit_aggregation_f32: bad args");
    AOTB_REQUIRE(max_obj == 10, "aotb_soft_logit_aggregation_f32: built for MODEL_MAX_OBJ_NUM = 10 (got %d)", max_obj);
    AggArgs a;
    for (int e = 0; e < 8; ++e) a.logits[e] = e < n_engines ? logits[e] : nullptr;
    for (int e = 0; e < n_engines; ++e) AOTB_REQUIRE(a.logits[e], "aotb_soft_logit_aggregation_f32: null logit map");
    int g = cdiv(HW, 256);
    if (g > 148 * 8) g = 148 * 8;
    launch(soft_logit_aggregation_kernel<11>, dim3(g), dim3(256), 0, (cudaStream_t)stream, a, n_engines, out, HW);
    return check_launch("aotb_soft_logit_aggregation_f32");
}


extern "C" int aotb_separate_labels_f32(const float* mask, int n_engines, int max_obj, float* out, int HW, void* stream) {
    AOTB_REQUIRE(mask && out && n_engines >= 1 && max_obj >= 1 && HW > 0, "aotb_separate_labels_f32: bad args");
    int g = cdiv(HW, 256);
    if (g > 148 * 8) g = 148 * 8;
    launch(separate_labels_kernel, dim3(g), dim3(256), 0, (cudaStream_t)stream, mask, n_engines, max_obj, out, HW);
    return check_launch("aotb_separate_labels_f32");
}

extern "C" int aotb_bank_append_f32(const float* src, int lds, float* bank, int ldb, int rows, int cols, int offset,
                                    const int* offset_dev, void* stream) {
    AOTB_REQUIRE(src && bank && rows > 0 && cols > 0 && cols % 4 == 0 && lds % 4 == 0 && ldb % 4 == 0,
                 "aotb_bank_append_f32: bad args");
    const size_t total = (size_t)rows * (cols / 4);
    int g = (int)((total + 255) / 256);
    if (g > 148 * 8) g = 148 * 8;
    launch(bank_append_kernel, dim3(g), dim3(256), 0, (cudaStream_t)stream, src, lds, bank, ldb, rows, cols / 4, offset, offset_dev);
    return check_launch("aotb_bank_append_f32");
}

extern "C" int aotb_counter_add(int* counter, int delta, void* stream) {
    AOTB_REQUIRE(counter, "aotb_counter_add: null");
    launch(counter_add_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, counter, delta);
    return check_launch("aotb_counter_add");
}
