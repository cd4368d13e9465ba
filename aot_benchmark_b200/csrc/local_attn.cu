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
    pdl_sync();
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
    launch(local_attn_kernel<D, DV, HAS_RELV, STAGE_WK>, dim3(grid), dim3(256), smem, st, a);
    return check_launch("aotb_local_attention_f32");
}


// ---------------------------------------------------------------------------------------------------
// Tiled variant for the AOT head shape (d = dv = 32): one CTA = one head x an 8x8 tile of query pixels.
// The (8+14)^2 K halo is staged once in shared memory with coalesced 128-byte row loads (the per-warp kernel
// above issues one scattered 16-byte load per lane and tap), scores are computed with taps on lanes, the
// softmax probabilities of the tile's 64 queries are parked in shared memory, then the same halo buffer is
// refilled with V and the aggregate runs with channels on lanes.
// relv_t is relative_emb_v transposed to [H][225][32] so the per-tap row is one coalesced 128-byte read.
template <int TY, int TX>
__global__ void __launch_bounds__(512, 1) local_attn_tile_kernel(const LocalArgs p, const float* __restrict__ relv_t) {
    pdl_sync();
    constexpr int D = 32, HH = TY + 2 * LR, HWD = TX + 2 * LR, NPOS = HH * HWD, LD = 33;
    constexpr int NT = 512, QPW = TY * TX / (NT / 32);
    extern __shared__ __align__(16) float smem[];
    float* halo = smem;                 // [NPOS][LD]
    float* wk = halo + NPOS * LD;       // [225][LD]
    float* bk = wk + LTAPS * LD;        // [225]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_x = (p.w + TX - 1) / TX;
    const int ty0 = (blockIdx.x / tiles_x) * TY, tx0 = (blockIdx.x % tiles_x) * TX;
    const int g = blockIdx.y;

    for (int f = tid; f < LTAPS * 8; f += NT) {
        const int r = f >> 3, c4 = (f & 7) * 4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(p.relk_w + ((size_t)g * LTAPS + r) * D + c4));
        float* d = wk + r * LD + c4;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    for (int t = tid; t < LTAPS; t += NT) bk[t] = __ldg(p.relk_b + g * LTAPS + t);

    auto load_halo = [&](const float* src, int ld) {
        for (int f = tid; f < NPOS * 8; f += NT) {
            const int pos = f >> 3, c4 = (f & 7) * 4;
            const int hy = pos / HWD, hx = pos - hy * HWD;
            const int yy = ty0 - LR + hy, xx = tx0 - LR + hx;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (yy >= 0 && yy < p.h && xx >= 0 && xx < p.w)
                v = __ldg(reinterpret_cast<const float4*>(src + (size_t)(yy * p.w + xx) * ld + g * D + c4));
            float* d = halo + pos * LD + c4;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
    };
    load_halo(p.k, p.ldk);
    __syncthreads();

    // ---- phase 1: scores + softmax, taps on lanes; probabilities of all TY*TX queries parked in shared memory
    float* prob = bk + LTAPS + 7;       // [TY*TX][PLD]
    constexpr int PLD = 228;
    const float invT = 1.f / p.T;
#pragma unroll 1
    for (int qi = 0; qi < QPW; ++qi) {
        const int ql = warp * QPW + qi;
        const int ly = ql / TX, lx = ql - ly * TX;
        const int y = ty0 + ly, x = tx0 + lx;
        if (y >= p.h || x >= p.w) continue;      // warp-uniform
        float qv[D];
        const float4* qp = reinterpret_cast<const float4*>(p.q + (size_t)(y * p.w + x) * p.ldq + g * D);
#pragma unroll
        for (int c = 0; c < D / 4; ++c) {
            const float4 t = __ldg(qp + c);
            qv[4 * c] = t.x; qv[4 * c + 1] = t.y; qv[4 * c + 2] = t.z; qv[4 * c + 3] = t.w;
        }
        float sc[8];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int wi = lane + 32 * j;
            float s1 = -INFINITY;
            if (wi < LTAPS) {
                const int dy = wi / LW, dx = wi - dy * LW;       // 0..14
                const int yy = y + dy - LR, xx = x + dx - LR;
                const bool inside = (yy >= 0 && yy < p.h && xx >= 0 && xx < p.w);
                const float* kr = halo + ((ly + dy) * HWD + lx + dx) * LD;
                const float* wr = wk + wi * LD;
                float dot = 0.f, rel = bk[wi];
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    dot = fmaf(qv[c], kr[c], dot);
                    rel = fmaf(wr[c], qv[c], rel);
                }
                s1 = inside ? fmaf(dot, invT, rel) : (rel - 1e8f);
            }
            sc[j] = s1;
            mx = fmaxf(mx, s1);
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sc[j] = (lane + 32 * j < LTAPS) ? expf(sc[j] - mx) : 0.f;
            sum += sc[j];
        }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (lane + 32 * j < LTAPS) prob[ql * PLD + lane + 32 * j] = sc[j] * inv;
    }
    __syncthreads();          // every warp is done with the K halo
    load_halo(p.v, p.ldv);
    __syncthreads();

    // ---- phase 2: aggregate, channels on lanes
    const float* rv = relv_t + (size_t)g * LTAPS * D + lane;
#pragma unroll 1
    for (int qi = 0; qi < QPW; ++qi) {
        const int ql = warp * QPW + qi;
        const int ly = ql / TX, lx = ql - ly * TX;
        const int y = ty0 + ly, x = tx0 + lx;
        if (y >= p.h || x >= p.w) continue;     // warp-uniform
        const float* pq = prob + ql * PLD;
        float acc0 = 0.f, acc1 = 0.f;
        for (int dy = 0; dy < LW; ++dy) {
            const float* vrow = halo + ((ly + dy) * HWD + lx) * LD + lane;   // zero outside the frame
#pragma unroll
            for (int dx = 0; dx < LW; ++dx) {
                const int wi = dy * LW + dx;
                const float pv = pq[wi];
                const float vv = vrow[dx * LD] + __ldg(rv + wi * D);
                if (dx & 1) acc1 = fmaf(pv, vv, acc1); else acc0 = fmaf(pv, vv, acc0);
            }
        }
        p.out[(size_t)(y * p.w + x) * p.ldo + g * D + lane] = acc0 + acc1;
    }
}

static int launch_local_tile(const LocalArgs& a, const float* relv_t, cudaStream_t st) {
    constexpr int TY = 8, TX = 8;
    const size_t smem = sizeof(float) * (size_t)((TY + 14) * (TX + 14) * 33 + LTAPS * 33 + LTAPS + 8 + TY * TX * 228);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(local_attn_tile_kernel<TY, TX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) {
            set_error("aotb_local_attention_tile_f32: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return AOTB_ERR_CUDA;
        }
        configured = true;
    }
    dim3 grid(cdiv(a.h, TY) * cdiv(a.w, TX), a.H);
    launch(local_attn_tile_kernel<TY, TX>, dim3(grid), dim3(512), smem, st, a, relv_t);
    return check_launch("aotb_local_attention_tile_f32");
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

// Tiled kernel for the AOT head shape (d_att = d_v = 32).  relv_t = relative_emb_v transposed to [H][225][32].
extern "C" int aotb_local_attention_tile_f32(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                                             const float* relk_w, const float* relk_b, const float* relv_t, float* out,
                                             int ldo, int h, int w, int H, void* stream) {
    AOTB_REQUIRE(q && k && v && relk_w && relk_b && relv_t && out && h > 0 && w > 0 && H > 0,
                 "aotb_local_attention_tile_f32: bad args");
    AOTB_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0, "aotb_local_attention_tile_f32: ld %% 4");
    LocalArgs a;
    a.q = q; a.ldq = ldq; a.k = k; a.ldk = ldk; a.v = v; a.ldv = ldv;
    a.relk_w = relk_w; a.relk_b = relk_b; a.relv = nullptr; a.out = out; a.ldo = ldo;
    a.h = h; a.w = w; a.H = H; a.T = sqrtf(32.f);
    return launch_local_tile(a, relv_t, (cudaStream_t)stream);
}
