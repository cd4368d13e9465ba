// Short-term local-window attention (K2 / K2'), fp32, one warp per (query pixel, head).
//
// Restates MultiheadLocalAttentionV2.forward (networks/layers/attention.py:308-376) and
// LocalGatedPropagation.forward (:789-861) in the form of SURVEY Appendix C -- i.e. the
// reference's own `unfold` definition (:343-348 / :830-835) of what the absent third-party
// spatial_correlation_sampler computes -- without the [1,H,225,N] score tensor, the boolean
// scatter of local2global (:378-417) or the dense N x N matmul (:366-368):
//
//   r[wi]  = relative_emb_k(q)[g*225+wi]            (grouped 1x1 conv + bias on UNSCALED q, :327)
//   s[wi]  = (q/T).k[y+dy,x+dx] + r[wi]             in frame;   r[wi] - 1e8 outside (:355-357)
//   p      = softmax_wi(s)                          225 taps, wi = (dy+7)*15 + (dx+7)
//   o[c]   = sum_wi p[wi] * ( v[y+dy,x+dx][c] + relative_emb_v[g][c][wi] )   (:363-371; no emb_v in DeAOT)
//
// Layout: q,k [HW][ldq/ldk] (head g at columns g*D), v [HW][ldv] (head g at g*DV), out [HW][ldo].
// Phase 1 puts window taps on lanes (each lane owns <= 8 taps and walks the D channels with
// 128-bit loads); phase 2 puts channels on lanes so every tap is one coalesced row read.
#include "common.cuh"

namespace aotb {

constexpr int LW = 15, LR = 7, LTAPS = 225;

struct LocalArgs {
    const float* q; int ldq;
    const float* k; int ldk;
    const float* v; int ldv;
    const float* relk_w;   // [H*225][D]
    const float* relk_b;   // [H*225]
    const float* relv;     // [H][DV][225] or null
    float* out; int ldo;
    int h, w, H;
    float T;
};

template <int D, int DV, bool HAS_RELV, bool STAGE_WK>
__global__ void __launch_bounds__(256) local_attn_kernel(const LocalArgs p) {
    constexpr int WARPS = 8;
    constexpr int WKS = D + 4;                 // padded row stride of the staged rel-k weights
    extern __shared__ __align__(16) float smem[];
    float* q_raw = smem;                       // [WARPS][D]
    float* q_scl = q_raw + WARPS * D;          // [WARPS][D]
    float* prob = q_scl + WARPS * D;           // [WARPS][232]
    float* wk_s = prob + WARPS * 232;          // [225][WKS]           (STAGE_WK)
    float* relv_s = wk_s + (STAGE_WK ? LTAPS * WKS : 0);  // [DV][225] (HAS_RELV, DV == 32)

    const int g = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = p.h * p.w;
    const int qi = blockIdx.x * WARPS + warp;

    if constexpr (STAGE_WK) {
        const float* src = p.relk_w + (size_t)g * LTAPS * D;
        for (int f = threadIdx.x; f < LTAPS * (D / 4); f += blockDim.x) {
            const int r = f / (D / 4), c = (f % (D / 4)) * 4;
            *reinterpret_cast<float4*>(wk_s + r * WKS + c) = __ldg(reinterpret_cast<const float4*>(src + r * D + c));
        }
    }
    if constexpr (HAS_RELV) {
        const float* src = p.relv + (size_t)g * DV * LTAPS;
        for (int f = threadIdx.x; f < DV * LTAPS; f += blockDim.x) relv_s[f] = __ldg(src + f);
    }
    if (qi < N) {
        for (int c = lane; c < D; c += 32) {
            const float x = __ldg(p.q + (size_t)qi * p.ldq + g * D + c);
            q_raw[warp * D + c] = x;
            q_scl[warp * D + c] = x / p.T;   // true division (attention.py:330)
        }
    }
    __syncthreads();
    if (qi >= N) return;

    const int y = qi / p.w, x = qi - y * p.w;
    const float* qr = q_raw + warp * D;
    const float* qs = q_scl + warp * D;

    // ---- phase 1: scores, taps on lanes
    float s[8];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int wi = lane + 32 * j;
        float sc = -INFINITY;
        if (wi < LTAPS) {
            const int dy = wi / LW - LR, dx = wi % LW - LR;
            const int yy = y + dy, xx = x + dx;
            const bool inside = (yy >= 0 && yy < p.h && xx >= 0 && xx < p.w);
            float rel = __ldg(p.relk_b + g * LTAPS + wi);
            float dot = 0.f;
            const float* wrow = STAGE_WK ? (wk_s + wi * WKS) : (p.relk_w + ((size_t)g * LTAPS + wi) * D);
            const float* krow = p.k + (size_t)(inside ? (yy * p.w + xx) : 0) * p.ldk + g * D;
#pragma unroll 4
            for (int c = 0; c < D; c += 4) {
                const float4 qa = *reinterpret_cast<const float4*>(qr + c);
                float4 wv;
                if constexpr (STAGE_WK) wv = *reinterpret_cast<const float4*>(wrow + c);
                else wv = __ldg(reinterpret_cast<const float4*>(wrow + c));
                rel = fmaf(wv.x, qa.x, rel); rel = fmaf(wv.y, qa.y, rel);
                rel = fmaf(wv.z, qa.z, rel); rel = fmaf(wv.w, qa.w, rel);
                if (inside) {
                    const float4 qb = *reinterpret_cast<const float4*>(qs + c);
                    const float4 kv = __ldg(reinterpret_cast<const float4*>(krow + c));
                    dot = fmaf(qb.x, kv.x, dot); dot = fmaf(qb.y, kv.y, dot);
                    dot = fmaf(qb.z, kv.z, dot); dot = fmaf(qb.w, kv.w, dot);
                }
            }
            sc = inside ? (dot + rel) : (rel - 1e8f);
        }
        s[j] = sc;
        mx = fmaxf(mx, sc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        s[j] = (lane + 32 * j < LTAPS) ? expf(s[j] - mx) : 0.f;
        sum += s[j];
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    float* pw = prob + warp * 232;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (lane + 32 * j < LTAPS) pw[lane + 32 * j] = s[j] * inv;
    __syncwarp();

    // ---- phase 2: aggregate, channels on lanes
    constexpr int NV = DV / 128 > 0 ? DV / 128 : 1;  // float4 groups per lane when DV >= 128
    if constexpr (DV == 32) {
        float acc = 0.f;
        for (int wi = 0; wi < LTAPS; ++wi) {
            const int dy = wi / LW - LR, dx = wi % LW - LR;
            const int yy = y + dy, xx = x + dx;
            const float pv = pw[wi];
            if (yy >= 0 && yy < p.h && xx >= 0 && xx < p.w)
                acc = fmaf(pv, __ldg(p.v + (size_t)(yy * p.w + xx) * p.ldv + g * DV + lane), acc);
            if constexpr (HAS_RELV) acc = fmaf(pv, relv_s[lane * LTAPS + wi], acc);
        }
        p.out[(size_t)qi * p.ldo + g * DV + lane] = acc;
    } else {
        static_assert(DV == 32 || DV % 128 == 0, "DV");
        static_assert(!HAS_RELV || DV == 32, "relative_emb_v only exists in the AOT head shape");
        float4 acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int wi = 0; wi < LTAPS; ++wi) {
            const int dy = wi / LW - LR, dx = wi % LW - LR;
            const int yy = y + dy, xx = x + dx;
            if (yy < 0 || yy >= p.h || xx < 0 || xx >= p.w) continue;
            const float pv = pw[wi];
            const float* vrow = p.v + (size_t)(yy * p.w + xx) * p.ldv + g * DV;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const float4 vv = __ldg(reinterpret_cast<const float4*>(vrow + i * 128 + lane * 4));
                acc[i].x = fmaf(pv, vv.x, acc[i].x); acc[i].y = fmaf(pv, vv.y, acc[i].y);
                acc[i].z = fmaf(pv, vv.z, acc[i].z); acc[i].w = fmaf(pv, vv.w, acc[i].w);
            }
        }
#pragma unroll
        for (int i = 0; i < NV; ++i)
            *reinterpret_cast<float4*>(p.out + (size_t)qi * p.ldo + g * DV + i * 128 + lane * 4) = acc[i];
    }
}

template <int D, int DV, bool HAS_RELV, bool STAGE_WK>
static int launch_local(const LocalArgs& a, cudaStream_t st) {
    const size_t smem = sizeof(float) * (size_t)(8 * D * 2 + 8 * 232 + (STAGE_WK ? LTAPS * (D + 4) : 0) +
                                                 (HAS_RELV ? DV * LTAPS : 0));
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(local_attn_kernel<D, DV, HAS_RELV, STAGE_WK>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("aotb_local_attention_f32: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return AOTB_ERR_CUDA;
        }
        configured = true;
    }
    dim3 grid(cdiv(a.h * a.w, 8), a.H);
    local_attn_kernel<D, DV, HAS_RELV, STAGE_WK><<<grid, 256, smem, st>>>(a);
    return check_launch("aotb_local_attention_f32");
}

}  // namespace aotb

using namespace aotb;

extern "C" int aotb_local_attention_f32(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                                        const float* relk_w, const float* relk_b, const float* relv, float* out,
                                        int ldo, int h, int w, int H, int d_att, int d_v, void* stream) {
    AOTB_REQUIRE(q && k && v && relk_w && relk_b && out && h > 0 && w > 0 && H > 0,
                 "aotb_local_attention_f32: bad args");
    AOTB_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0, "aotb_local_attention_f32: ld %% 4");
    LocalArgs a;
    a.q = q; a.ldq = ldq; a.k = k; a.ldk = ldk; a.v = v; a.ldv = ldv;
    a.relk_w = relk_w; a.relk_b = relk_b; a.relv = relv; a.out = out; a.ldo = ldo;
    a.h = h; a.w = w; a.H = H; a.T = sqrtf((float)d_att);
    cudaStream_t st = (cudaStream_t)stream;
    if (d_att == 32 && d_v == 32 && relv) return launch_local<32, 32, true, true>(a, st);
    if (d_att == 128 && d_v == 1024 && !relv) return launch_local<128, 1024, false, false>(a, st);
    set_error("aotb_local_attention_f32: unsupported head shape d_att=%d d_v=%d relv=%d", d_att, d_v, relv != nullptr);
    return AOTB_ERR_UNSUPPORTED;
}
