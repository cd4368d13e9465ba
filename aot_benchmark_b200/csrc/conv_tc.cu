// Implicit-GEMM convolution / linear layer on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), NHWC fp32
// in / fp32 out, fp32-faithful through the split-fp16 ("fp16x2") scheme of lt_attn_tc.cu.
//
//   out[m][n] = act( sum_k A(m,k) * W[k][n] + bias[n] + res[m][n] )      m -> (b, oy, ox), k -> (ky, kx, ci)
//
// Same reference sites as conv_igemm.cu (resnet.py:34-54,140-157 with FrozenBatchNorm2d folded,
// aot.py:19-21,83, fpn.py:34-58, every nn.Linear of transformer.py:321-367,582-665).  Eligibility:
// Cin % 4 == 0, Cout % 64 == 0, dilation 1 (everything on the R50-AOTL path except the 11-channel conv_out;
// the 7x7 stem runs on a 4-channel zero-padded copy of the image).
//
// One CTA = 128 output pixels x BN output channels, 320 threads:
//   warps 0-7  A producers (two chunks of loads in flight per thread): gather the fp32 activation rows of a 64-wide K chunk straight from NHWC global
//              memory (im2col is never materialised; padding / stride handled per row), split every value into
//              hi = fp16(x), lo = fp16(x - hi) and store both 128x64 tiles in the 128B-swizzled K-major layout
//              the UMMA descriptor expects; afterwards the same warps run the epilogue
//              (tcgen05.ld -> bias / residual / activation -> fp32 NHWC stores).
//   warp 8     TMA producer for the pre-split weights (Wh, Wl as [Cout][K] fp16, K-major), 3-stage ring.
//   warp 9     tcgen05.mma issuer: per 64-wide chunk 4 k-steps x (Ah*Wh + Al*Wh + Ah*Wl) into a 128 x BN fp32
//              accumulator in TMEM.
#include "common.cuh"
#include "tc_common.cuh"

namespace aotb {
namespace tc {

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

struct ConvTcArgs {
    const float* in;
    const float* bias;
    const float* res;
    float* out;
    int B, H, W, Cin, ldin;
    int Ho, Wo, Cout, ldout, ldres;
    int KH, KW, stride, pad;
    int M, nchunks;
    int act;
    int splits;          // split-K: the `splits` CTAs (blockIdx.z) of one output tile form a thread-block cluster; CTA z
                         // handles chunks [z*per, (z+1)*per) and the partial tiles are summed over distributed smem
    int spin;            // 1: mbarrier waits without the suspend hint
    long long* prof;     // diagnostic: 12 clock64 stamps per CTA (see aotb_set_conv_tiling), or null
};

static int g_conv_tiling = 0;      // aotb_set_conv_tiling

struct RowInfo { int pix_base, iy0, ix0, valid; };

// Finish of one output tile by the 256 epilogue threads.  The staging tile is [128][BN + 4] fp32; with split-K the S
// CTAs of the cluster each own 128 / S rows and sum that slice of every peer's staging buffer in rank order.  A thread
// keeps one 4-channel column (its bias is loaded once) and walks rows; NB rows are processed per batch with every load
// of the batch (S remote tiles + residual) issued before the first use -- a load-use-load chain costs one DSMEM / L2
// round trip per row and was 6 us per tile.
template <int BN, int S, int NB>
__device__ __forceinline__ void conv_finish_tile(const ConvTcArgs& a, const uint8_t* smem, int tid, int m0, int n0, int zrank) {
    constexpr int LD = BN + 4, C4 = BN / 4, RSTEP = 256 / C4, ROWS = 128 / S, ITERS = ROWS / RSTEP;
    constexpr int B = NB < ITERS ? NB : ITERS;
    static_assert(ITERS >= 1 && ITERS % B == 0, "finish tiling");
    const int c = (tid % C4) * 4, n = n0 + c;
    const int r0 = zrank * ROWS + tid / C4;
    const float4 b4 = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float* lbase = reinterpret_cast<const float*>(smem) + r0 * LD + c;
    const uint32_t sbase = smem_u32(lbase);
#pragma unroll 1
    for (int i0 = 0; i0 < ITERS; i0 += B) {
        float4 v[B][S], rs[B];
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const int ro = (i0 + b) * RSTEP;
            if (S == 1) {
                v[b][0] = *reinterpret_cast<const float4*>(lbase + ro * LD);
            } else {
#pragma unroll
                for (int z = 0; z < S; ++z) v[b][z] = dsmem_ld_f4(dsmem_addr(sbase + (uint32_t)(ro * LD) * 4u, (uint32_t)z));
            }
            const int m = m0 + r0 + ro;
            rs[b] = (a.res && m < a.M) ? *reinterpret_cast<const float4*>(a.res + (size_t)m * a.ldres + n)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int b = 0; b < B; ++b) {
            float4 o = v[b][0];
#pragma unroll
            for (int z = 1; z < S; ++z) { o.x += v[b][z].x; o.y += v[b][z].y; o.z += v[b][z].z; o.w += v[b][z].w; }
            o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
            o.x += rs[b].x; o.y += rs[b].y; o.z += rs[b].z; o.w += rs[b].w;
            o.x = apply_act(o.x, a.act); o.y = apply_act(o.y, a.act);
            o.z = apply_act(o.z, a.act); o.w = apply_act(o.w, a.act);
            const int m = m0 + r0 + (i0 + b) * RSTEP;
            if (m < a.M) *reinterpret_cast<float4*>(a.out + (size_t)m * a.ldout + n) = o;
        }
    }
}

// Finish without split-K with the global loads taken off its critical path: the bias and the residual rows of the FIRST batch are
// already in registers (`pre`, loaded by finish_prefetch before the thread waited for the accumulator, i.e. under the tail of the
// K loop), and inside the loop the residual rows of batch k+1 are requested before batch k is computed and stored.  The staging
// tile is read with ld.shared so that the compiler does not have to order those reads behind the global stores.
template <int BN>
struct FinishPre { float4 b4; float4 rs[8]; };

template <int BN>
__device__ __forceinline__ void finish_prefetch(const ConvTcArgs& a, int tid, int m0, int n0, FinishPre<BN>& pre) {
    constexpr int C4 = BN / 4, RSTEP = 256 / C4;
    const int n = n0 + (tid % C4) * 4, r0 = tid / C4;
    pre.b4 = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const int m = m0 + r0 + b * RSTEP;
        pre.rs[b] = (a.res && m < a.M) ? *reinterpret_cast<const float4*>(a.res + (size_t)m * a.ldres + n)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

template <int BN>
__device__ __forceinline__ void conv_finish_tile_pre(const ConvTcArgs& a, const uint8_t* smem, int tid, int m0, int n0,
                                                     const FinishPre<BN>& pre) {
    constexpr int LD = BN + 4, C4 = BN / 4, RSTEP = 256 / C4, ITERS = 128 / RSTEP, B = 8;
    static_assert(ITERS % B == 0, "finish tiling");
    const int c = (tid % C4) * 4, n = n0 + c, r0 = tid / C4;
    const uint32_t sbase = smem_u32(reinterpret_cast<const float*>(smem) + r0 * LD + c);
    float4 rs[B];
#pragma unroll
    for (int b = 0; b < B; ++b) rs[b] = pre.rs[b];
#pragma unroll 1
    for (int i0 = 0; i0 < ITERS; i0 += B) {
        float4 v[B], nx[B];
#pragma unroll
        for (int b = 0; b < B; ++b)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[b].x), "=f"(v[b].y), "=f"(v[b].z), "=f"(v[b].w)
                         : "r"(sbase + (uint32_t)((i0 + b) * RSTEP * LD) * 4u));
        if (i0 + B < ITERS) {                      // residual rows of the next batch, in flight while this one is stored
#pragma unroll
            for (int b = 0; b < B; ++b) {
                const int m = m0 + r0 + (i0 + B + b) * RSTEP;
                nx[b] = (a.res && m < a.M) ? *reinterpret_cast<const float4*>(a.res + (size_t)m * a.ldres + n)
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int b = 0; b < B; ++b) {
            float4 o = v[b];
            o.x += pre.b4.x; o.y += pre.b4.y; o.z += pre.b4.z; o.w += pre.b4.w;
            o.x += rs[b].x; o.y += rs[b].y; o.z += rs[b].z; o.w += rs[b].w;
            o.x = apply_act(o.x, a.act); o.y = apply_act(o.y, a.act);
            o.z = apply_act(o.z, a.act); o.w = apply_act(o.w, a.act);
            const int m = m0 + r0 + (i0 + b) * RSTEP;
            if (m < a.M) *reinterpret_cast<float4*>(a.out + (size_t)m * a.ldout + n) = o;
        }
        if (i0 + B < ITERS) {
#pragma unroll
            for (int b = 0; b < B; ++b) rs[b] = nx[b];
        }
    }
}

template <int BN, int STAGES>
struct ConvSmem {
    static constexpr int A_BYTES = 128 * 128;          // one 128 x 64 half tile
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int STG_LD = BN + 4;              // split-K staging tile [128][BN + 4] fp32, aliases the stages
    static_assert(128 * STG_LD * 4 <= STAGES * STAGE_BYTES, "staging tile must fit in the operand stages");
    static constexpr int TOTAL = STAGES * STAGE_BYTES + 128 * (int)sizeof(RowInfo) + (3 * STAGES + 1) * 8 + 16 + 1024;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(320, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWl, const ConvTcArgs a) {
    
    using SM = ConvSmem<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    RowInfo* rinfo = reinterpret_cast<RowInfo*>(smem + STAGES * SM::STAGE_BYTES);
    uint64_t* a_full = reinterpret_cast<uint64_t*>(rinfo + 128);
    uint64_t* b_full = a_full + STAGES;
    uint64_t* s_free = b_full + STAGES;
    uint64_t* acc_full = s_free + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * 128, n0 = blockIdx.y * BN;
    auto W = [&](uint64_t* bar, uint32_t parity) {
        if (a.spin) mbar_wait_spin(bar, parity); else mbar_wait(bar, parity);
    };
    long long* prof = a.prof ? a.prof + 12 * ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) : nullptr;
    auto stamp = [&](int slot) { if (prof) prof[slot] = clock64(); };
    if (tid == 0) stamp(0);
    pdl_trigger();      // the next kernel may start its prologue; it waits for this grid before reading our output

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&a_full[s], 8); mbar_init(&b_full[s], 1); mbar_init(&s_free[s], 1); }
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    if (tid < 128) {
        const int m = m0 + tid;
        RowInfo ri;
        if (m < a.M) {
            const int HoWo = a.Ho * a.Wo;
            const int b = m / HoWo, r = m - b * HoWo;
            const int oy = r / a.Wo, ox = r - oy * a.Wo;
            ri.pix_base = b * a.H * a.W;
            ri.iy0 = oy * a.stride - a.pad;
            ri.ix0 = ox * a.stride - a.pad;
            ri.valid = 1;
        } else {
            ri.pix_base = 0; ri.iy0 = 0; ri.ix0 = 0; ri.valid = 0;
        }
        rinfo[tid] = ri;
    }
    if (warp == 9) tmem_alloc<BN>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) stamp(1);
    pdl_wait();         // activations / residual below were written by earlier kernels
    const int cpt = (a.Cin % 64 == 0) ? (a.Cin >> 6) : 0;  // 64-wide chunks per filter tap (0: general Cin % 4 path)
    const int per = (a.nchunks + a.splits - 1) / a.splits;
    const int kbeg = blockIdx.z * per;
    const int nloc = max(0, min(per, a.nchunks - kbeg));     // chunks of this CTA (local index it <-> chunk kbeg + it)

    FinishPre<BN> pre;
    if (warp < 8) {
        // ======================= A producers (8 warps, register double-buffered) =======================
        const int q = tid & 15, rsub = tid >> 4;      // rsub 0..15; this thread serves rows i*16 + rsub, i = 0..7
        // Per-row source pointers for the current filter tap (nullptr = zero padding / out of range); recomputed only when
        // the tap changes (never for 1x1 convs and linears, every Cin/64 chunks for 3x3), so the steady-state work per
        // 16-byte segment is one LDG, the hi/lo split and two 8-byte swizzled STS.
        const float* rowptr[8];
        int cur_tap = -1;
        auto load_chunk = [&](int kc, float4* v) {
            int tap, c0;
            if (cpt > 0) { tap = kc / cpt; c0 = ((kc - tap * cpt) << 6) + q * 4; }       // Cin % 64 == 0
            else { const int k = kc * 64 + q * 4; tap = k / a.Cin; c0 = k - tap * a.Cin; }  // Cin % 4 == 0 (stem)
            if (tap != cur_tap) {
                cur_tap = tap;
                const bool kvalid = tap < a.KH * a.KW;                                     // zero padding of K
                const int ky = tap / a.KW, kx = tap - ky * a.KW;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const RowInfo ri = rinfo[i * 16 + rsub];
                    const int iy = ri.iy0 + ky, ix = ri.ix0 + kx;
                    rowptr[i] = (kvalid && ri.valid && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W)
                                    ? a.in + (size_t)(ri.pix_base + iy * a.W + ix) * a.ldin
                                    : nullptr;
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
                v[i] = rowptr[i] ? __ldg(reinterpret_cast<const float4*>(rowptr[i] + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        // byte offset of this thread's 8-byte slot inside a 128x64 half tile (128B swizzle; row & 7 == rsub & 7)
        const uint32_t soff = rsub * 128 + (((q >> 1) ^ (rsub & 7)) << 4) + ((q & 1) << 3);
        auto store_chunk = [&](int it, const float4* v) {
            const int s = it % STAGES;
            if (it >= STAGES) W(&s_free[s], ((it / STAGES) - 1) & 1);
            uint8_t* Ah = smem + s * SM::STAGE_BYTES + soff;
            uint8_t* Al = Ah + SM::A_BYTES;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const __half2 h0 = __floats2half2_rn(v[i].x, v[i].y), h1 = __floats2half2_rn(v[i].z, v[i].w);
                const __half2 l0 = __floats2half2_rn(v[i].x - __low2float(h0), v[i].y - __high2float(h0));
                const __half2 l1 = __floats2half2_rn(v[i].z - __low2float(h1), v[i].w - __high2float(h1));
                uint2 ph, pl;
                ph.x = *reinterpret_cast<const uint32_t*>(&h0); ph.y = *reinterpret_cast<const uint32_t*>(&h1);
                pl.x = *reinterpret_cast<const uint32_t*>(&l0); pl.y = *reinterpret_cast<const uint32_t*>(&l1);
                *reinterpret_cast<uint2*>(Ah + i * 2048) = ph;      // row i*16 + rsub
                *reinterpret_cast<uint2*>(Al + i * 2048) = pl;
            }
            fence_proxy_async();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
            mbar_arrive_warp(&a_full[s]);
            if (tid == 0 && it == 0) stamp(2);
        };
        // three chunks of global loads in flight per thread (register ring), so the ~L2 latency of a chunk is
        // covered by the convert+store work of the two chunks before it
        float4 v0[8], v1[8], v2[8];
        if (nloc > 0) load_chunk(kbeg, v0);
        if (nloc > 1) load_chunk(kbeg + 1, v1);
        for (int it = 0; it < nloc; it += 3) {
            if (it + 2 < nloc) load_chunk(kbeg + it + 2, v2);
            store_chunk(it, v0);
            if (it + 1 < nloc) {
                if (it + 3 < nloc) load_chunk(kbeg + it + 3, v0);
                store_chunk(it + 1, v1);
            }
            if (it + 2 < nloc) {
                if (it + 4 < nloc) load_chunk(kbeg + it + 4, v1);
                store_chunk(it + 2, v2);
            }
        }
        // ======================= epilogue (warps 0-3: columns [0, BN/2), warps 4-7: [BN/2, BN)) =======================
        if (a.splits == 1) finish_prefetch<BN>(a, tid, m0, n0, pre);      // bias + first residual rows: under the last MMAs
        if (nloc > 0) {
            W(acc_full, 0);
            tc_fence_after();
        }
        if (tid == 0) stamp(5);
        const int wq = warp & 3;
        const uint32_t trow = tmem + ((uint32_t)(wq * 32) << 16);
        const int cbeg = (warp >> 2) * (BN / 2);
#pragma unroll 1
        for (int c = cbeg; c < cbeg + BN / 2; c += 32) {
            uint32_t r[32];
            if (nloc > 0) {
                tmem_ld32(trow + c, r);
                tmem_wait_ld();
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = 0u;
            }
            // accumulator tile -> this CTA's staging buffer (the operand stages are dead: every MMA has completed), so
            // the global stores below run along rows.  Row stride BN + 4 floats: an odd number of 16-byte units, so the
            // 32 rows of a warp do not collide.  (Storing straight from the TMEM layout -- one row per lane -- costs
            // 32 half-filled sectors per instruction: 9 us for a 128 x 256 tile against < 1 us this way.)
            float* srow = reinterpret_cast<float*>(smem) + (wq * 32 + lane) * SM::STG_LD + c;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(srow + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                   __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
    } else if (warp == 8) {
        // ======================= weight TMA producer =======================
        if (elect_one()) {
            tma_prefetch_desc(&tmWh); tma_prefetch_desc(&tmWl);
            for (int it = 0; it < nloc; ++it) {
                const int s = it % STAGES;
                if (it >= STAGES) W(&s_free[s], ((it / STAGES) - 1) & 1);
                uint8_t* Bh = smem + s * SM::STAGE_BYTES + 2 * SM::A_BYTES;
                mbar_arrive_expect_tx(&b_full[s], 2 * SM::B_BYTES);
                tma_load_2d(Bh, &tmWh, &b_full[s], (kbeg + it) * 64, n0);
                tma_load_2d(Bh + SM::B_BYTES, &tmWl, &b_full[s], (kbeg + it) * 64, n0);
            }
        }
    } else {
        // ======================= MMA issuer =======================
        if (elect_one()) {
            constexpr uint32_t IDESC = idesc_f16(128, BN, 0, 0);
            // descriptors of stage 0, built once; +STAGE_BYTES>>4 per stage, +2 per 32-byte k-slice
            const uint64_t dAh0 = smem_desc_sw128(smem_u32(smem));
            const uint64_t dAl0 = dAh0 + (SM::A_BYTES >> 4);
            const uint64_t dBh0 = dAl0 + (SM::A_BYTES >> 4);
            const uint64_t dBl0 = dBh0 + (SM::B_BYTES >> 4);
            for (int kc = 0; kc < nloc; ++kc) {
                const int s = kc % STAGES;
                const uint32_t ph = (kc / STAGES) & 1;
                W(&a_full[s], ph);
                W(&b_full[s], ph);
                tc_fence_after();
                if (kc == 0) stamp(3);
                const uint64_t so = (uint64_t)(s * (SM::STAGE_BYTES >> 4));
                const uint64_t ah = dAh0 + so, al = dAl0 + so, bh = dBh0 + so, bl = dBl0 + so;
                mma_ss(tmem, ah, bh, IDESC, kc ? 1u : 0u);
                mma_ss(tmem, al, bh, IDESC, 1u);
                mma_ss(tmem, ah, bl, IDESC, 1u);
#pragma unroll
                for (int ks = 1; ks < 4; ++ks) {
                    mma_ss(tmem, ah + 2 * ks, bh + 2 * ks, IDESC, 1u);
                    mma_ss(tmem, al + 2 * ks, bh + 2 * ks, IDESC, 1u);
                    mma_ss(tmem, ah + 2 * ks, bl + 2 * ks, IDESC, 1u);
                }
                mma_commit(&s_free[s]);
            }
            if (nloc > 0) mma_commit(acc_full);
            stamp(4);
        }
    }
    if (tid == 0) stamp(6);
    // Finish: CTA z of a split-K cluster owns rows [z*128/S, (z+1)*128/S) of the tile and reads that slice of every
    // peer's staging buffer through distributed shared memory, summing in rank order (deterministic); without split-K
    // (S = 1) it is the CTA's own buffer.  Then bias / residual / activation and coalesced row stores.  No global
    // partials, no second kernel.
    __syncwarp();
    if (a.splits > 1) cluster_sync_all(); else __syncthreads();
    if (tid == 0) stamp(8);
    if (warp < 8) {
        const int zr = a.splits > 1 ? (int)blockIdx.z : 0;
        if (a.splits == 1) conv_finish_tile_pre<BN>(a, smem, tid, m0, n0, pre);
        else if (a.splits == 2) conv_finish_tile<BN, 2, 4>(a, smem, tid, m0, n0, zr);
        else if (a.splits == 4) conv_finish_tile<BN, 4, 2>(a, smem, tid, m0, n0, zr);
        else conv_finish_tile<BN, 8, 2>(a, smem, tid, m0, n0, zr);
    }
    if (tid == 0) stamp(9);
    if (a.splits > 1) {
        __syncwarp();
        cluster_sync_all();      // nobody leaves (and frees its shared memory) while a peer may still read it
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) stamp(7);
    if (warp == 9) tmem_dealloc<BN>(tmem);
}

static int make_tmap_weights(CUtensorMap* out, const void* base, int Kpad, int Cout, int BN) {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
            set_error("cuTensorMapEncodeTiled entry point unavailable");
            return AOTB_ERR_CUDA;
        }
        fn = (PFN_encodeTiled)p;
    }
    cuuint64_t dims[2] = {(cuuint64_t)Kpad, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)Kpad * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(weights) failed (%d)", (int)r);
        return AOTB_ERR_CUDA;
    }
    return AOTB_OK;
}

template <int BN, int STAGES>
static int launch_conv_tc(const CUtensorMap& th, const CUtensorMap& tl, const ConvTcArgs& a, cudaStream_t st) {
    constexpr int smem = ConvSmem<BN, STAGES>::TOTAL;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
            set_error("aotb_conv2d_nhwc_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return AOTB_ERR_CUDA;
        }
        configured = true;
    }
    dim3 grid(cdiv(a.M, 128), a.Cout / BN, a.splits);
    launch_cluster(conv_tc_kernel<BN, STAGES>, dim3(grid), dim3(320), smem, st, a.splits, th, tl, a);
    return check_launch("aotb_conv2d_nhwc_tc");
}

}  // namespace tc
}  // namespace aotb

using namespace aotb;

extern "C" int aotb_set_conv_tiling(int mode) {
    AOTB_REQUIRE(mode >= 0 && mode < (1 << 12), "aotb_set_conv_tiling: mode is a 12-bit mask");
    tc::g_conv_tiling = mode;
    return AOTB_OK;
}

// wh / wl: pre-split weights [Cout][Kpad] fp16 (K = KH*KW*Cin ordered (ky,kx,ci), zero-padded to a multiple of 64).
extern "C" int aotb_conv2d_nhwc_tc(const float* in, const void* wh, const void* wl, const float* bias, const float* res,
                                   float* out, int B, int H, int W, int Cin, int ldin, int Cout, int ldout, int ldres,
                                   int KH, int KW, int stride, int pad, int act, void* workspace, size_t workspace_bytes,
                                   void* stream) {
    AOTB_REQUIRE(in && wh && wl && out, "aotb_conv2d_nhwc_tc: null pointer");
    AOTB_REQUIRE(Cin % 4 == 0 && Cout % 64 == 0, "aotb_conv2d_nhwc_tc: Cin must be a multiple of 4, Cout of 64");
    AOTB_REQUIRE(ldin % 4 == 0 && ldout % 4 == 0 && (!res || ldres % 4 == 0) && ((uintptr_t)in % 16 == 0) &&
                     ((uintptr_t)out % 16 == 0) && (!res || (uintptr_t)res % 16 == 0) && (!bias || (uintptr_t)bias % 16 == 0),
                 "aotb_conv2d_nhwc_tc: 16-byte alignment required");
    tc::ConvTcArgs a;
    a.in = in; a.bias = bias; a.res = res; a.out = out;
    a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.ldin = ldin;
    a.Ho = (H + 2 * pad - KH) / stride + 1;
    a.Wo = (W + 2 * pad - KW) / stride + 1;
    AOTB_REQUIRE(a.Ho > 0 && a.Wo > 0, "aotb_conv2d_nhwc_tc: empty output");
    a.Cout = Cout; a.ldout = ldout; a.ldres = ldres; a.KH = KH; a.KW = KW; a.stride = stride; a.pad = pad;
    a.M = B * a.Ho * a.Wo;
    const int K = ((KH * KW * Cin + 63) / 64) * 64;   // weights are zero-padded to a multiple of 64 along K
    a.nchunks = K / 64;
    a.act = act;
    // Tile policy: pick the N tile (64 / 128 / 256 dividing Cout) and the split-K cluster size (1 / 2 / 4 / 8) that
    // minimise a cost model fitted to profiles/r01_trip14_conv_sweep.json (B200, graph-replayed launches):
    //     T = waves * (ramp + chunks_per_cta * t_chunk[BN] + t_finish[BN]) + (S > 1 ? t_cluster : 0)
    //   t_chunk : one 64-deep K chunk = 12 MMAs.  A 128 x BN x 16 MMA reads (4 + BN/32) KB of shared memory per BN/2
    //             math cycles -- 192 / 128 / 96 B/clk for BN = 64 / 128 / 256 against 128 B/clk of smem bandwidth --
    //             so narrow tiles are smem-bound (0.63 us for a quarter of the work a BN = 256 chunk does in 0.80 us);
    //   t_finish: staging + (DSMEM reduction) + row stores of a 128 x BN fp32 tile; proportional to the tile bytes at
    //             ~20-26 GB/s per SM whether it is written straight out or summed over the cluster (DSMEM ~21 B/clk);
    //   ramp    : prologue + first activation chunk in shared memory, paid once per wave of CTAs.
    // bit 0 of the tuning mask ("narrow") restores the pre-model heuristic for A/B runs.
    const int mt = cdiv(a.M, 128);
    int BN = 64, best_s = 1;
    if ((tc::g_conv_tiling & 1) == 0) {
        float best = 1e30f;
        const int bns[3] = {256, 128, 64};
        const float t_chunk[3] = {0.80f, 0.66f, 0.63f}, t_fin[3] = {6.5f, 3.6f, 2.1f};
        for (int bi = 0; bi < 3; ++bi) {
            if (Cout % bns[bi]) continue;
            for (int sp = 1; sp <= 8; sp <<= 1) {
                if (sp > a.nchunks) break;
                const int ctas = mt * (Cout / bns[bi]) * sp;
                const int slots = sp == 8 ? 128 : (sp == 4 ? 132 : 148);        // co-resident CTAs with clusters of sp
                if (sp > 1 && ctas > slots) continue;                         // split-K only to fill a partial wave
                const int waves = cdiv(ctas, slots);
                const float t = waves * (2.0f + cdiv(a.nchunks, sp) * t_chunk[bi] + t_fin[bi]) + (sp > 1 ? 0.3f : 0.f);
                if (t < best) { best = t; BN = bns[bi]; best_s = sp; }
            }
        }
    } else {          // "narrow": the widest BN that still fills ~a wave on its own, else 64; split-K to ~one wave
        if (Cout % 256 == 0 && mt * (Cout / 256) >= 120) BN = 256;
        else if (Cout % 128 == 0 && mt * (Cout / 128) >= 120) BN = 128;
        const int ctas = mt * (Cout / BN);
        if (ctas < 100 && a.nchunks >= 4) {
            best_s = 8;
            while (best_s > 1 && (ctas * best_s > 160 || a.nchunks / best_s < 2)) best_s >>= 1;
        }
    }
    const int force_bn = (tc::g_conv_tiling >> 4) & 15, force_s = (tc::g_conv_tiling >> 8) & 15;
    if (force_bn) {
        BN = 32 << force_bn;
        AOTB_REQUIRE((BN == 64 || BN == 128 || BN == 256) && Cout % BN == 0, "aotb_conv2d_nhwc_tc: forced tile %d invalid", BN);
    }
    a.splits = force_bn ? 1 : best_s;
    const int ctas = mt * (Cout / BN);
    if (force_s) {
        AOTB_REQUIRE((force_s == 1 || force_s == 2 || force_s == 4 || force_s == 8) && force_s <= a.nchunks,
                     "aotb_conv2d_nhwc_tc: forced split %d invalid", force_s);
        a.splits = force_s;
    }
    a.spin = (tc::g_conv_tiling & 2) ? 1 : 0;
    a.prof = nullptr;
    if (tc::g_conv_tiling & 4) {      // diagnostic stamps go to the caller's workspace
        const size_t need = (size_t)ctas * a.splits * 12 * sizeof(long long);
        AOTB_REQUIRE(workspace && workspace_bytes >= need, "aotb_conv2d_nhwc_tc: profile mode needs %zu workspace bytes", need);
        a.prof = (long long*)workspace;
    }
    CUtensorMap th, tl;
    int rc;
    if ((rc = tc::make_tmap_weights(&th, wh, K, Cout, BN)) != AOTB_OK) return rc;
    if ((rc = tc::make_tmap_weights(&tl, wl, K, Cout, BN)) != AOTB_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (BN == 256) return tc::launch_conv_tc<256, 2>(th, tl, a, st);
    if (BN == 128) return tc::launch_conv_tc<128, 3>(th, tl, a, st);
    return tc::launch_conv_tc<64, 3>(th, tl, a, st);
}
