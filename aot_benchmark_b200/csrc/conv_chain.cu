// A whole chain of convolutions / linears as ONE persistent kernel with tile-level dataflow (round 2).
//
// conv_tc.cu launches one kernel per layer: at the sizes of a 480p frame a layer is one or two waves of 128-pixel tiles, so
// every launch pays prologue (barrier init, TMEM allocation) -> first operand chunk -> K loop -> finish in lockstep on all
// SMs and the next layer cannot start before the last CTA has drained (profiles/r02_trip2_conv_microbench.json: of the
// 10.5 us of a layer-3 1x1 conv, 2.3 us are MMAs).  The ResNet-50 stages (networks/encoders/resnet.py:34-54,140-157, FrozenBN
// folded) are 52 such launches per frame.  Here the whole chain is a PROGRAM of tiles executed by one CTA per SM:
//
//   program     tiles (layer, m-tile of 128 output pixels, n-tile of BN channels) in layer order, dealt round-robin to the
//               CTAs; every CTA walks its tiles in program order.  A tile names the m-tiles of its input (and residual) layer
//               it reads; `done[layer][m-tile]` counts the finished n-tiles of an m-tile (release / acquire at gpu scope).
//               A tile starts as soon as ITS inputs are complete -- no grid barrier, no launch boundary: layers overlap at
//               tile granularity (a 1x1 conv needs the same m-tile of its input, a 3x3 the neighbouring image rows).
//               Dependencies always point backwards in program order and every CTA executes in program order, so the
//               earliest unfinished tile can always run: no deadlock.
//   roles       448 threads.  warps 0-7: A producers (fp32 NHWC gather -> fp16 hi / lo split -> 128B-swizzled smem, exactly
//               conv_tc.cu's producer); warp 12: TMA producer of the pre-split weights (never waits for a dependency: the
//               weights of the next tile stream in while the activations are still being produced elsewhere); warp 13:
//               tcgen05.mma issuer; warps 8-11: epilogue.  The 3-stage operand ring and its mbarrier phases run on across
//               tiles; TWO accumulators in TMEM (2 x 128 columns) let the epilogue of tile k drain while tile k+1
//               accumulates.
//   epilogue    tcgen05.ld 32 columns -> warp-private padded smem staging (the TMEM layout has one row per lane; storing it
//               directly costs 32 half-filled sectors per instruction) -> rows re-read 4 at a time, bias + residual +
//               activation, coalesced 16-byte stores -> __threadfence -> one release-add on done[layer][m-tile].
//   arithmetic  identical to conv_tc.cu: per 64-deep K chunk 4 x (Ah Wh + Al Wh + Ah Wl), fp32 accumulation in TMEM, same chunk
//               order => bit-identical outputs to the per-layer kernel without split-K.
#include "common.cuh"
#include "tc_common.cuh"
#include <algorithm>
#include <cstring>
#include <vector>

namespace aotb {
namespace tc {

struct ChainLayer {                 // device copy of one layer (64-byte aligned array)
    const float* in;
    const float* bias;
    const float* res;
    float* out;
    int H, W, Cin, ldin;
    int Ho, Wo, Cout, ldout, ldres;
    int KH, KW, stride, pad;
    int M, nchunks, act, BN;
    int in_done, res_done;          // offset of the producing layer's counters in `done`, -1 = ready before the launch
    int in_need, res_need;          // n-tiles per m-tile of the producing layer
    int done_off;                   // this layer's counters
    int pad0;
};
struct ChainTile { int layer, mt, nt, dep_lo, dep_hi, pad0, pad1, pad2; };

struct ChainArgs {
    const ChainLayer* layers;
    const ChainTile* tiles;
    const CUtensorMap* tmaps;       // [2 * nlayers]: Wh, Wl
    int* done;
    int ntiles;
};

constexpr int CH_STAGES = 3, CH_THREADS = 448;
constexpr int CH_A_BYTES = 128 * 128;                      // one 128 x 64 half tile
constexpr int CH_B_BYTES = 128 * 128;                      // BN <= 128 rows of 128 B
constexpr int CH_STAGE_BYTES = 2 * CH_A_BYTES + 2 * CH_B_BYTES;
constexpr int CH_STG_LD = 36;                              // floats per staging row (32 + 4: conflict-free 16-byte accesses)
constexpr int CH_STG_BYTES = 4 * 32 * CH_STG_LD * 4;

struct ChRowInfo { int pix_base, iy0, ix0, valid; };

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// wait until done[lo..hi] >= need (bounded: a protocol bug traps instead of hanging the GPU)
__device__ __forceinline__ void wait_done(const int* done, int lo, int hi, int need) {
    for (int m = lo; m <= hi; ++m) {
        uint32_t spins = 0;
        while (ld_acquire_gpu(done + m) < need) {
            __nanosleep(40);
            if (++spins > (1u << 23)) __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d_g(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

__global__ void __launch_bounds__(CH_THREADS, 1) conv_chain_kernel(const ChainArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* staging = reinterpret_cast<float*>(smem + CH_STAGES * CH_STAGE_BYTES);
    ChRowInfo* rinfo = reinterpret_cast<ChRowInfo*>(smem + CH_STAGES * CH_STAGE_BYTES + CH_STG_BYTES);      // [2][128]
    uint64_t* a_full = reinterpret_cast<uint64_t*>(rinfo + 256);
    uint64_t* b_full = a_full + CH_STAGES;
    uint64_t* s_free = b_full + CH_STAGES;
    uint64_t* acc_full = s_free + CH_STAGES;      // [2]
    uint64_t* acc_free = acc_full + 2;            // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_trigger();
    if (tid == 0) {
        for (int s = 0; s < CH_STAGES; ++s) { mbar_init(&a_full[s], 256); mbar_init(&b_full[s], 1); mbar_init(&s_free[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_free[i], 128); }
        fence_mbar_init();
    }
    if (warp == 13) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();                     // the chain's first input was written by the kernel before this one

    if (warp < 8) {
        // ======================= A producers =======================
        const int q = tid & 15, rsub = tid >> 4;      // this thread serves rows i*16 + rsub, i = 0..7, 16-byte segment q
        const uint32_t soff = rsub * 128 + (((q >> 1) ^ (rsub & 7)) << 4) + ((q & 1) << 3);
        int g0 = 0, seq = 0;
        for (int ti = blockIdx.x; ti < a.ntiles; ti += gridDim.x, ++seq) {
            const ChainTile t = a.tiles[ti];
            const ChainLayer& L = a.layers[t.layer];
            const float* in = L.in;
            const int H = L.H, W = L.W, Cin = L.Cin, ldin = L.ldin, KW = L.KW, taps = L.KH * L.KW, nchunks = L.nchunks;
            ChRowInfo* ri = rinfo + (seq & 1) * 128;
            if (tid < 128) {
                const int m = t.mt * 128 + tid;
                ChRowInfo r;
                if (m < L.M) {
                    const int oy = m / L.Wo, ox = m - oy * L.Wo;           // batch 1
                    r.pix_base = 0; r.iy0 = oy * L.stride - L.pad; r.ix0 = ox * L.stride - L.pad; r.valid = 1;
                } else {
                    r.pix_base = 0; r.iy0 = 0; r.ix0 = 0; r.valid = 0;
                }
                ri[tid] = r;
            }
            if (tid == 0 && L.in_done >= 0) wait_done(a.done + L.in_done, t.dep_lo, t.dep_hi, L.in_need);
            asm volatile("bar.sync 2, 256;" ::: "memory");
            const int cpt = Cin >> 6;                                          // 64-wide chunks per filter tap
            int rowoff[8];                                                     // element offset of the row's pixel, -1 = zero padding
            int cur_tap = -1;
            auto load_chunk = [&](int kc, float4* v) {
                const int tap = kc / cpt, c0 = ((kc - tap * cpt) << 6) + q * 4;
                if (tap != cur_tap) {
                    cur_tap = tap;
                    const bool kvalid = tap < taps;
                    const int ky = tap / KW, kx = tap - ky * KW;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const ChRowInfo r = ri[i * 16 + rsub];
                        const int iy = r.iy0 + ky, ix = r.ix0 + kx;
                        rowoff[i] = (kvalid && r.valid && iy >= 0 && iy < H && ix >= 0 && ix < W) ? (iy * W + ix) * ldin : -1;
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)        // plain (coherent) loads: the rows were written earlier in THIS launch
                    v[i] = rowoff[i] >= 0 ? *reinterpret_cast<const float4*>(in + rowoff[i] + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            auto store_chunk = [&](int it, const float4* v) {
                const int g = g0 + it, s = g % CH_STAGES;
                if (g >= CH_STAGES) mbar_wait(&s_free[s], ((g / CH_STAGES) - 1) & 1);
                uint8_t* Ah = smem + s * CH_STAGE_BYTES + soff;
                uint8_t* Al = Ah + CH_A_BYTES;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const __half2 h0 = __floats2half2_rn(v[i].x, v[i].y), h1 = __floats2half2_rn(v[i].z, v[i].w);
                    const __half2 l0 = __floats2half2_rn(v[i].x - __low2float(h0), v[i].y - __high2float(h0));
                    const __half2 l1 = __floats2half2_rn(v[i].z - __low2float(h1), v[i].w - __high2float(h1));
                    uint2 ph, pl;
                    ph.x = *reinterpret_cast<const uint32_t*>(&h0); ph.y = *reinterpret_cast<const uint32_t*>(&h1);
                    pl.x = *reinterpret_cast<const uint32_t*>(&l0); pl.y = *reinterpret_cast<const uint32_t*>(&l1);
                    *reinterpret_cast<uint2*>(Ah + i * 2048) = ph;
                    *reinterpret_cast<uint2*>(Al + i * 2048) = pl;
                }
                fence_proxy_async();
                mbar_arrive(&a_full[s]);
            };
            float4 v0[8], v1[8], v2[8];
            if (nchunks > 0) load_chunk(0, v0);
            if (nchunks > 1) load_chunk(1, v1);
            for (int it = 0; it < nchunks; it += 3) {
                if (it + 2 < nchunks) load_chunk(it + 2, v2);
                store_chunk(it, v0);
                if (it + 1 < nchunks) {
                    if (it + 3 < nchunks) load_chunk(it + 3, v0);
                    store_chunk(it + 1, v1);
                }
                if (it + 2 < nchunks) {
                    if (it + 4 < nchunks) load_chunk(it + 4, v1);
                    store_chunk(it + 2, v2);
                }
            }
            g0 += nchunks;
        }
    } else if (warp < 12) {
        // ======================= epilogue (4 warps: TMEM lane quadrants) =======================
        const int ew = warp - 8, etid = tid - 256;
        float* stg = staging + ew * 32 * CH_STG_LD;
        const uint32_t trow = tmem + ((uint32_t)(ew * 32) << 16);
        const int r4 = lane >> 3, c4 = (lane & 7) * 4;          // coalesced phase: 4 rows per instruction, 8 lanes per row
        int seq = 0;
        for (int ti = blockIdx.x; ti < a.ntiles; ti += gridDim.x, ++seq) {
            const ChainTile t = a.tiles[ti];
            const ChainLayer& L = a.layers[t.layer];
            const int acc = seq & 1, BN = L.BN, M = L.M, act = L.act, ldout = L.ldout, ldres = L.ldres;
            const float* bias = L.bias;
            const float* res = L.res;
            float* out = L.out;
            const int m0 = t.mt * 128 + ew * 32, n0 = t.nt * BN;
            if (res && L.res_done >= 0) {            // uniform over the 128 epilogue threads
                if (etid == 0) wait_done(a.done + L.res_done, t.mt, t.mt, L.res_need);
                asm volatile("bar.sync 3, 128;" ::: "memory");
            }
            mbar_wait(&acc_full[acc], (seq >> 1) & 1);
            tc_fence_after();
            for (int c = 0; c < BN; c += 32) {
                uint32_t r[32];
                tmem_ld32(trow + acc * 128 + c, r);
                tmem_wait_ld();
                if (c + 32 >= BN) {                  // the accumulator has been read completely: hand it back to the MMA warp
                    tc_fence_before();
                    mbar_arrive(&acc_free[acc]);
                }
                float* srow = stg + lane * CH_STG_LD;
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(srow + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                       __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                __syncwarp();
                const int n = n0 + c + c4;
                const float4 b4 = bias ? *reinterpret_cast<const float4*>(bias + n) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int row = it * 4 + r4, m = m0 + row;
                    float4 o = *reinterpret_cast<const float4*>(stg + row * CH_STG_LD + c4);
                    if (m < M) {
                        o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
                        if (res) {
                            const float4 rs = *reinterpret_cast<const float4*>(res + (size_t)m * ldres + n);
                            o.x += rs.x; o.y += rs.y; o.z += rs.z; o.w += rs.w;
                        }
                        o.x = apply_act(o.x, act); o.y = apply_act(o.y, act);
                        o.z = apply_act(o.z, act); o.w = apply_act(o.w, act);
                        *reinterpret_cast<float4*>(out + (size_t)m * ldout + n) = o;
                    }
                }
                __syncwarp();
            }
            __threadfence();                         // this thread's stores are visible at gpu scope ...
            asm volatile("bar.sync 3, 128;" ::: "memory");
            if (etid == 0) red_release_gpu_add(a.done + L.done_off + t.mt, 1);        // ... before the tile is published
        }
    } else if (warp == 12) {
        // ======================= weight TMA producer =======================
        if (elect_one()) {
            int g0 = 0;
            for (int ti = blockIdx.x; ti < a.ntiles; ti += gridDim.x) {
                const ChainTile t = a.tiles[ti];
                const ChainLayer& L = a.layers[t.layer];
                const CUtensorMap* th = a.tmaps + 2 * t.layer;
                const int nchunks = L.nchunks, BN = L.BN, n0 = t.nt * BN;
                for (int kc = 0; kc < nchunks; ++kc) {
                    const int g = g0 + kc, s = g % CH_STAGES;
                    if (g >= CH_STAGES) mbar_wait(&s_free[s], ((g / CH_STAGES) - 1) & 1);
                    uint8_t* Bh = smem + s * CH_STAGE_BYTES + 2 * CH_A_BYTES;
                    mbar_arrive_expect_tx(&b_full[s], 2 * BN * 128);
                    tma_load_2d_g(Bh, th, &b_full[s], kc * 64, n0);
                    tma_load_2d_g(Bh + CH_B_BYTES, th + 1, &b_full[s], kc * 64, n0);
                }
                g0 += nchunks;
            }
        }
    } else {
        // ======================= MMA issuer =======================
        if (elect_one()) {
            constexpr uint32_t IDESC128 = idesc_f16(128, 128, 0, 0), IDESC64 = idesc_f16(128, 64, 0, 0);
            const uint64_t dAh0 = smem_desc_sw128(smem_u32(smem));
            const uint64_t dAl0 = dAh0 + (CH_A_BYTES >> 4);
            const uint64_t dBh0 = dAl0 + (CH_A_BYTES >> 4);
            const uint64_t dBl0 = dBh0 + (CH_B_BYTES >> 4);
            int g0 = 0, seq = 0;
            for (int ti = blockIdx.x; ti < a.ntiles; ti += gridDim.x, ++seq) {
                const ChainTile t = a.tiles[ti];
                const ChainLayer& L = a.layers[t.layer];
                const int nchunks = L.nchunks, acc = seq & 1;
                const uint32_t idesc = L.BN == 128 ? IDESC128 : IDESC64;
                const uint32_t d = tmem + acc * 128;
                if (seq >= 2) { mbar_wait(&acc_free[acc], ((seq >> 1) - 1) & 1); tc_fence_after(); }
                for (int kc = 0; kc < nchunks; ++kc) {
                    const int g = g0 + kc, s = g % CH_STAGES;
                    const uint32_t ph = (g / CH_STAGES) & 1;
                    mbar_wait(&a_full[s], ph);
                    mbar_wait(&b_full[s], ph);
                    tc_fence_after();
                    const uint64_t so = (uint64_t)(s * (CH_STAGE_BYTES >> 4));
                    const uint64_t ah = dAh0 + so, al = dAl0 + so, bh = dBh0 + so, bl = dBl0 + so;
                    mma_ss(d, ah, bh, idesc, kc ? 1u : 0u);
                    mma_ss(d, al, bh, idesc, 1u);
                    mma_ss(d, ah, bl, idesc, 1u);
#pragma unroll
                    for (int ks = 1; ks < 4; ++ks) {
                        mma_ss(d, ah + 2 * ks, bh + 2 * ks, idesc, 1u);
                        mma_ss(d, al + 2 * ks, bh + 2 * ks, idesc, 1u);
                        mma_ss(d, ah + 2 * ks, bl + 2 * ks, idesc, 1u);
                    }
                    mma_commit(&s_free[s]);
                }
                mma_commit(&acc_full[acc]);
                g0 += nchunks;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 13) tmem_dealloc<256>(tmem);
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace tc
}  // namespace aotb

using namespace aotb;

// Host-side description of one layer of a chain (caller-owned device pointers, same conventions as aotb_conv2d_nhwc_tc).
// in_layer / res_layer: index (in this array) of the layer that PRODUCES `in` / `res` inside the chain, or -1 when the
// tensor is complete before the chain starts.
struct aotb_chain_layer {
    const float* in;
    const void* wh;
    const void* wl;
    const float* bias;
    const float* res;
    float* out;
    int H, W, Cin, ldin, Cout, ldout, ldres, KH, KW, stride, pad, act, in_layer, res_layer;
};

namespace {
struct ChainPlan {
    std::vector<tc::ChainLayer> layers;
    std::vector<tc::ChainTile> tiles;
    int ndone = 0;
    size_t off_tiles = 0, off_tmaps = 0, off_done = 0, bytes = 0;
};

int plan_chain(const aotb_chain_layer* ls, int n, ChainPlan& P) {
    std::vector<int> mtiles(n), ntn(n);
    for (int i = 0; i < n; ++i) {
        const aotb_chain_layer& l = ls[i];
        AOTB_REQUIRE(l.in && l.wh && l.wl && l.out, "aotb_conv_chain: layer %d: null pointer", i);
        AOTB_REQUIRE(l.Cin % 64 == 0 && l.Cout % 64 == 0, "aotb_conv_chain: layer %d: Cin and Cout must be multiples of 64", i);
        AOTB_REQUIRE(l.ldin % 4 == 0 && l.ldout % 4 == 0 && (!l.res || l.ldres % 4 == 0), "aotb_conv_chain: layer %d: strides", i);
        AOTB_REQUIRE(l.in_layer < i && l.res_layer < i, "aotb_conv_chain: layer %d: producers must come earlier in the chain", i);
        tc::ChainLayer d{};
        d.in = l.in; d.bias = l.bias; d.res = l.res; d.out = l.out;
        d.H = l.H; d.W = l.W; d.Cin = l.Cin; d.ldin = l.ldin;
        d.Ho = (l.H + 2 * l.pad - l.KH) / l.stride + 1;
        d.Wo = (l.W + 2 * l.pad - l.KW) / l.stride + 1;
        AOTB_REQUIRE(d.Ho > 0 && d.Wo > 0, "aotb_conv_chain: layer %d: empty output", i);
        d.Cout = l.Cout; d.ldout = l.ldout; d.ldres = l.ldres;
        d.KH = l.KH; d.KW = l.KW; d.stride = l.stride; d.pad = l.pad;
        d.M = d.Ho * d.Wo;
        d.nchunks = l.KH * l.KW * l.Cin / 64;
        d.act = l.act;
        d.BN = (l.Cout % 128 == 0) ? 128 : 64;
        mtiles[i] = cdiv(d.M, 128);
        ntn[i] = l.Cout / d.BN;
        d.done_off = P.ndone;
        P.ndone += mtiles[i];
        d.in_done = d.res_done = -1;
        d.in_need = d.res_need = 0;
        if (l.in_layer >= 0) {
            AOTB_REQUIRE(P.layers[l.in_layer].M == l.H * l.W, "aotb_conv_chain: layer %d: input geometry does not match its producer", i);
            d.in_done = P.layers[l.in_layer].done_off; d.in_need = ntn[l.in_layer];
        }
        if (l.res && l.res_layer >= 0) {
            AOTB_REQUIRE(P.layers[l.res_layer].M == d.M, "aotb_conv_chain: layer %d: residual geometry does not match its producer", i);
            d.res_done = P.layers[l.res_layer].done_off; d.res_need = ntn[l.res_layer];
        }
        P.layers.push_back(d);
    }
    for (int i = 0; i < n; ++i) {
        const tc::ChainLayer& d = P.layers[i];
        for (int mt = 0; mt < mtiles[i]; ++mt) {
            const int m0 = mt * 128, m1 = std::min(m0 + 127, d.M - 1);
            int lo, hi;
            if (d.KH == 1 && d.KW == 1 && d.stride == 1 && d.pad == 0) {
                lo = mt; hi = m1 / 128;                                   // the same pixels of the input
            } else {                                                      // every input image row the tile's output rows touch
                const int oy0 = m0 / d.Wo, oy1 = m1 / d.Wo;
                const int iy0 = std::max(0, oy0 * d.stride - d.pad), iy1 = std::min(d.H - 1, oy1 * d.stride - d.pad + d.KH - 1);
                lo = (iy0 * d.W) / 128; hi = (iy1 * d.W + d.W - 1) / 128;
            }
            for (int nt = 0; nt < ntn[i]; ++nt) P.tiles.push_back(tc::ChainTile{i, mt, nt, lo, hi, 0, 0, 0});
        }
    }
    P.off_tiles = tc::align_up(P.layers.size() * sizeof(tc::ChainLayer), 256);
    P.off_tmaps = tc::align_up(P.off_tiles + P.tiles.size() * sizeof(tc::ChainTile), 256);
    P.off_done = tc::align_up(P.off_tmaps + 2 * (size_t)n * sizeof(CUtensorMap), 256);
    P.bytes = P.off_done + tc::align_up((size_t)P.ndone * sizeof(int), 256);
    return AOTB_OK;
}

int make_tmap_w(CUtensorMap* out, const void* base, int K, int Cout, int BN) {
    static tc::PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
            set_error("cuTensorMapEncodeTiled entry point unavailable");
            return AOTB_ERR_CUDA;
        }
        fn = (tc::PFN_encodeTiled)p;
    }
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(chain weights) failed (%d)", (int)r);
        return AOTB_ERR_CUDA;
    }
    return AOTB_OK;
}
}  // namespace

// Size of the device-resident program of a chain (layer table, tile list, tensor maps, dependency counters).
extern "C" int aotb_conv_chain_plan(const void* layers, int nlayers, size_t* program_bytes, int* ntiles, int* ncounters) {
    AOTB_REQUIRE(layers && nlayers > 0 && program_bytes && ntiles && ncounters, "aotb_conv_chain_plan: bad args");
    ChainPlan P;
    int rc = plan_chain((const aotb_chain_layer*)layers, nlayers, P);
    if (rc != AOTB_OK) return rc;
    *program_bytes = P.bytes; *ntiles = (int)P.tiles.size(); *ncounters = P.ndone;
    return AOTB_OK;
}

// Host-only view of the tile program (tests / diagnostics): 5 ints per tile (layer, m-tile, n-tile, dep_lo, dep_hi) and per layer
// (M, BN, counters offset, producer counters offset or -1, residual counters offset or -1, in_need, res_need).
extern "C" int aotb_conv_chain_dump(const void* layers, int nlayers, int* tiles5, int max_tiles, int* layers7) {
    AOTB_REQUIRE(layers && nlayers > 0 && tiles5 && layers7, "aotb_conv_chain_dump: bad args");
    ChainPlan P;
    int rc = plan_chain((const aotb_chain_layer*)layers, nlayers, P);
    if (rc != AOTB_OK) return rc;
    AOTB_REQUIRE((int)P.tiles.size() <= max_tiles, "aotb_conv_chain_dump: %zu tiles exceed the buffer", P.tiles.size());
    for (size_t i = 0; i < P.tiles.size(); ++i) {
        const tc::ChainTile& t = P.tiles[i];
        int* o = tiles5 + 5 * i;
        o[0] = t.layer; o[1] = t.mt; o[2] = t.nt; o[3] = t.dep_lo; o[4] = t.dep_hi;
    }
    for (int i = 0; i < nlayers; ++i) {
        const tc::ChainLayer& d = P.layers[i];
        int* o = layers7 + 7 * i;
        o[0] = d.M; o[1] = d.BN; o[2] = d.done_off; o[3] = d.in_done; o[4] = d.res_done; o[5] = d.in_need; o[6] = d.res_need;
    }
    return AOTB_OK;
}

// Write the program into `program` (device memory, >= program_bytes, 256-byte aligned).  Call once per geometry, outside any
// stream capture (it copies from host memory and synchronises the stream).
extern "C" int aotb_conv_chain_build(const void* layers, int nlayers, void* program, size_t program_bytes, void* stream) {
    AOTB_REQUIRE(layers && nlayers > 0 && program && ((uintptr_t)program % 256) == 0, "aotb_conv_chain_build: bad args");
    const aotb_chain_layer* ls = (const aotb_chain_layer*)layers;
    ChainPlan P;
    int rc = plan_chain(ls, nlayers, P);
    if (rc != AOTB_OK) return rc;
    AOTB_REQUIRE(program_bytes >= P.bytes, "aotb_conv_chain_build: program buffer too small (%zu < %zu)", program_bytes, P.bytes);
    std::vector<uint8_t> host(P.bytes, 0);
    memcpy(host.data(), P.layers.data(), P.layers.size() * sizeof(tc::ChainLayer));
    memcpy(host.data() + P.off_tiles, P.tiles.data(), P.tiles.size() * sizeof(tc::ChainTile));
    for (int i = 0; i < nlayers; ++i) {
        const int K = ls[i].KH * ls[i].KW * ls[i].Cin;
        CUtensorMap th, tl;
        if ((rc = make_tmap_w(&th, ls[i].wh, K, ls[i].Cout, P.layers[i].BN)) != AOTB_OK) return rc;
        if ((rc = make_tmap_w(&tl, ls[i].wl, K, ls[i].Cout, P.layers[i].BN)) != AOTB_OK) return rc;
        memcpy(host.data() + P.off_tmaps + (2 * (size_t)i) * sizeof(CUtensorMap), &th, sizeof(CUtensorMap));
        memcpy(host.data() + P.off_tmaps + (2 * (size_t)i + 1) * sizeof(CUtensorMap), &tl, sizeof(CUtensorMap));
    }
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemcpyAsync(program, host.data(), P.bytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        set_error("aotb_conv_chain_build: %s", cudaGetErrorString(e));
        return AOTB_ERR_CUDA;
    }
    return AOTB_OK;
}

// Run a built program: clears the dependency counters and launches the persistent kernel (graph-capturable).
extern "C" int aotb_conv_chain_run(void* program, int nlayers, int ntiles, int ncounters, void* stream) {
    AOTB_REQUIRE(program && nlayers > 0 && ntiles > 0 && ncounters > 0, "aotb_conv_chain_run: bad args");
    const size_t off_tiles = tc::align_up((size_t)nlayers * sizeof(tc::ChainLayer), 256);
    const size_t off_tmaps = tc::align_up(off_tiles + (size_t)ntiles * sizeof(tc::ChainTile), 256);
    const size_t off_done = tc::align_up(off_tmaps + 2 * (size_t)nlayers * sizeof(CUtensorMap), 256);
    uint8_t* base = (uint8_t*)program;
    tc::ChainArgs a;
    a.layers = (const tc::ChainLayer*)base;
    a.tiles = (const tc::ChainTile*)(base + off_tiles);
    a.tmaps = (const CUtensorMap*)(base + off_tmaps);
    a.done = (int*)(base + off_done);
    a.ntiles = ntiles;
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int smem = tc::CH_STAGES * tc::CH_STAGE_BYTES + tc::CH_STG_BYTES + 256 * (int)sizeof(tc::ChRowInfo) + 16 * 8 + 16 + 1024;
    static bool configured = false;
    static int sms = 0;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tc::conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        int dev = 0;
        if (e == cudaSuccess) e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) {
            set_error("aotb_conv_chain_run: %s", cudaGetErrorString(e));
            return AOTB_ERR_CUDA;
        }
        configured = true;
    }
    cudaError_t e = cudaMemsetAsync(a.done, 0, (size_t)ncounters * sizeof(int), st);
    if (e != cudaSuccess) {
        set_error("aotb_conv_chain_run: memset: %s", cudaGetErrorString(e));
        return AOTB_ERR_CUDA;
    }
    const int grid = ntiles < sms ? ntiles : sms;            // one CTA per SM: all CTAs are co-resident (the dataflow needs it)
    launch(tc::conv_chain_kernel, dim3(grid), dim3(tc::CH_THREADS), smem, st, a);
    return check_launch("aotb_conv_chain_run");
}
