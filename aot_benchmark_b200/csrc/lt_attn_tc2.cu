// Long-term attention (K1), "pair" layout: TWO co-resident CTAs per SM instead of one wide CTA.
//
// Same arithmetic contract as lt_attn_tc.cu (networks/layers/attention.py:82-117; fp16x2 split operands, fp32 accumulate).
// Round-2 measurements (profiles/r02_summary.md) showed that every single-CTA organisation of the softmax ("tile",
// "groups", "ahead") lands at ~2400-2800 cycles per 128x128 score tile although no pipe is saturated: 16 softmax warps
// that move in lockstep through tcgen05.ld -> max -> exchange -> ex2 -> tcgen05.st leave the MUFU pipe (the real bound of a
// d = 32 head: 128 tensor FLOPs per exponential) idle for more than half of the tile.  Here the SM is shared by two
// independent CTAs whose phases drift apart freely -- while one reads TMEM or waits at its row-max barrier the other one
// feeds the MUFU pipe:
//   CTA        128 queries x 1 head x 1 KV split, 320 threads: warps 0-7 softmax, warp 8 TMA producer, warp 9 MMA issuer;
//              __launch_bounds__(320, 2); 80 KB of shared memory and 256 TMEM columns per CTA.
//   key tile   64 keys.  TMEM: S_0 | S_1 | S_2 (64 fp32 columns each) | O' (64) = 256 columns.  Score tile n lives in buffer
//              n % 3 and the MMA warp runs up to three tiles ahead (S(0..2) up front, then wait P(n) -> PV(n) -> S(n + 3)), so
//              a CTA's softmax never waits for the tensor pipe in steady state.
//   softmax    warp w owns TMEM lanes 32 (w % 4) .. +31 and key columns 32 (w / 4) .. +31 of the tile: two threads share a
//              query row, each reads its 32 scores once, the half-row maxima are exchanged through shared memory (one
//              256-thread named barrier), then ex2, row sums and the fp16 hi / lo split.  P_hi and P_lo of a thread's 32
//              keys overwrite its own 32 score columns (hi: [c, c + 16), lo: [c + 16, c + 32)).
//   issue diet the scale-and-shift, the row sums and the residuals run as packed fp32 pairs (fma.rn.f32x2 / add.rn.f32x2,
//              one issue slot per two scores); P_hi is the fp32 value truncated to 11 significant bits (one LOP3), so the
//              residual p - hi is exact and needs no fp16 -> fp32 unpack: 9.5 instructions per score pair instead of 14.
//   rescale    O' is rescaled (rarely: only when a row maximum grows) by the two threads of the row, 32 columns each, after
//              o_done says PV(n - 1) has completed; PV(n) is not issued before all 256 threads arrive on p_full(n).
#include "common.cuh"
#include "tc_common.cuh"

namespace aotb {
namespace tc {

constexpr int P_BM = 128, P_BN = 64, P_STAGES = 4, P_THREADS = 320, P_TMA_WARP = 8, P_MMA_WARP = 9;
constexpr int P_QBYTES = P_BM * 128;    // 128 rows x 128 B
constexpr int P_KVBYTES = P_BN * 128;   // 64 rows x 128 B (one K or V tile)
constexpr float P_LOG2E = 1.4426950408889634f;

struct LtArgs2 {
    int N, Tk;
    const int* Tk_dev;
    int H;
    float* O;
    int ldo;
    float* Opart;
    float* Mpart;
    float* Lpart;
    int splits;
    int spin;
};

struct __align__(8) BarriersP {
    uint64_t q_full;
    uint64_t kv_full[P_STAGES];
    uint64_t kv_free[P_STAGES];
    uint64_t s_full[3];     // S(n) complete in buffer n % 3; use k = n / 3 of a buffer completes phase k
    uint64_t p_full[3];     // 8 arrivals (one per softmax warp): P(n) written over S(n)
    uint64_t o_done;        // committed after every PV: PV(n) completes phase n
    uint64_t o_final;       // the last PV
    uint32_t tmem_base;
    float xmax[4][P_BM];    // [(n & 1) * 2 + key half][row]
    float xsum[2][P_BM];
};

static int make_tmap_rows64_box(CUtensorMap* out, const void* base, int rows, int heads, int box_rows) {
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
    cuuint64_t dims[3] = {64, (cuuint64_t)rows, (cuuint64_t)heads};
    cuuint64_t strides[2] = {128, (cuuint64_t)rows * 128};  // bytes, dims 1..2
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
        return AOTB_ERR_CUDA;
    }
    return AOTB_OK;
}

template <bool EXACT>
__global__ void __launch_bounds__(P_THREADS, 2)
lt_attn_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const LtArgs2 a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;                                   // 1 tile of 128 rows
    uint8_t* sK = sQ + P_QBYTES;                          // P_STAGES tiles of 64 rows
    uint8_t* sV = sK + P_STAGES * P_KVBYTES;
    BarriersP* B = reinterpret_cast<BarriersP*>(sV + P_STAGES * P_KVBYTES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * P_BM, h = blockIdx.y, z = blockIdx.z;
    pdl_trigger();

    if (tid == 0) {
        mbar_init(&B->q_full, 1);
        for (int s = 0; s < P_STAGES; ++s) { mbar_init(&B->kv_full[s], 1); mbar_init(&B->kv_free[s], 1); }
        for (int b = 0; b < 3; ++b) { mbar_init(&B->s_full[b], 1); mbar_init(&B->p_full[b], 2 * P_BM / 32); }
        mbar_init(&B->o_done, 1);
        mbar_init(&B->o_final, 1);
        fence_mbar_init();
    }
    if (warp == P_MMA_WARP) tmem_alloc<256>(&B->tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = B->tmem_base;
    pdl_wait();
    const int Tk = a.Tk_dev ? *a.Tk_dev : a.Tk;
    const int tiles_total = (Tk + P_BN - 1) / P_BN;
    const int per = (tiles_total + a.splits - 1) / a.splits;
    const int tb = z * per;
    int T = tiles_total - tb;
    T = T < 0 ? 0 : (T > per ? per : T);

    if (warp == P_TMA_WARP) {
        // ======================= TMA producer =======================
        if (elect_one() && T > 0) {
            tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
            mbar_arrive_expect_tx(&B->q_full, P_QBYTES);
            tma_load_3d(sQ, &tmQ, &B->q_full, 0, q0, h);
            for (int j = 0; j < T; ++j) {
                const int s = j % P_STAGES;
                if (j >= P_STAGES) mbar_wait(&B->kv_free[s], ((j / P_STAGES) - 1) & 1);
                mbar_arrive_expect_tx(&B->kv_full[s], 2 * P_KVBYTES);
                tma_load_3d(sK + s * P_KVBYTES, &tmK, &B->kv_full[s], 0, (tb + j) * P_BN, h);
                tma_load_3d(sV + s * P_KVBYTES, &tmV, &B->kv_full[s], 0, (tb + j) * P_BN, h);
            }
        }
    } else if (warp == P_MMA_WARP) {
        // ======================= MMA issuer =======================
        if (elect_one() && T > 0) {
            constexpr uint32_t IDESC_S = idesc_f16(128, 64, 0, 0);
            constexpr uint32_t IDESC_O = idesc_f16(128, 64, 0, 1);
            const uint64_t dQ = smem_desc_sw128(smem_u32(sQ));
            const uint64_t dK = smem_desc_sw128(smem_u32(sK)), dV = smem_desc_sw128(smem_u32(sV));
            auto issue_S = [&](int n) {
                const int s = n % P_STAGES;
                mbar_wait_cp(&B->kv_full[s], (n / P_STAGES) & 1, a.spin);
                tc_fence_after();
                const uint64_t k = dK + (uint64_t)(s * (P_KVBYTES >> 4));
                const uint32_t d = tmem + (n % 3) * 64;
                // k-slices of 16 halfs = 32 B inside the 128 B row: 0,1 = hi ; 2,3 = lo
                mma_ss(d, dQ, k, IDESC_S, 0);
                mma_ss(d, dQ + 2, k + 2, IDESC_S, 1);
                if (EXACT) {
                    mma_ss(d, dQ + 4, k, IDESC_S, 1);          // Ql Kh
                    mma_ss(d, dQ + 6, k + 2, IDESC_S, 1);
                    mma_ss(d, dQ, k + 4, IDESC_S, 1);          // Qh Kl
                    mma_ss(d, dQ + 2, k + 6, IDESC_S, 1);
                }
                mma_commit(&B->s_full[n % 3]);
            };
            auto issue_PV = [&](int n) {
                const int s = n % P_STAGES;
                const uint64_t v = dV + (uint64_t)(s * (P_KVBYTES >> 4));
                const uint32_t d = tmem + 192;
                const uint32_t p = tmem + (n % 3) * 64;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)      // P_hi of keys [16 kk, +16) at columns 32 (kk / 2) + 8 (kk % 2)
                    mma_ts(d, p + 32 * (kk >> 1) + 8 * (kk & 1), v + 128 * kk, IDESC_O, (kk > 0 || n > 0) ? 1u : 0u);
                if (EXACT) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)  // P_lo 16 columns further up in the same thread's score columns
                        mma_ts(d, p + 32 * (kk >> 1) + 16 + 8 * (kk & 1), v + 128 * kk, IDESC_O, 1);
                }
                mma_commit(&B->o_done);
                if (n + 1 == T) mma_commit(&B->o_final);
                mma_commit(&B->kv_free[s]);
            };
            mbar_wait(&B->q_full, 0);
            tc_fence_after();
            for (int n = 0; n < 3 && n < T; ++n) issue_S(n);
            for (int n = 0; n < T; ++n) {
                mbar_wait_cp(&B->p_full[n % 3], (n / 3) & 1, a.spin);
                tc_fence_after();
                issue_PV(n);
                if (n + 3 < T) issue_S(n + 3);
            }
        }
    } else {
        // ======================= softmax (8 warps, two threads per query row) =======================
        const int hf = warp >> 2, wq = warp & 3;           // key half of the tile, TMEM lane quadrant
        const int row = wq * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
        const uint32_t tO = tmem + lane_addr + 192 + hf * 32;      // this thread's 32 of the 64 O' columns (rescale)
        float m_used = -INFINITY;
        uint64_t lsum = pk2(0.f, 0.f);                              // two partial row sums, packed
        const uint64_t l2e2 = pk2(P_LOG2E, P_LOG2E);
        int b = 0;
        uint32_t par = 0;                                           // bit b: phase parity of the next use of buffer b
        for (int n = 0; n < T; ++n) {
            const uint32_t tS = tmem + lane_addr + b * 64 + hf * 32;
            mbar_wait_cp(&B->s_full[b], (par >> b) & 1u, a.spin);
            tc_fence_after();
            uint32_t sr[32];
            tmem_ld32(tS, sr);
            tmem_wait_ld();
            const int key0 = (tb + n) * P_BN + hf * 32;
            if (key0 + 32 > Tk) {                    // warp-uniform: only the last key tile of the bank is ragged
#pragma unroll
                for (int k = 0; k < 32; ++k)
                    if (key0 + k >= Tk) sr[k] = __float_as_uint(-INFINITY);
            }
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
                mx0 = fmaxf(mx0, __uint_as_float(sr[k]));
                mx1 = fmaxf(mx1, __uint_as_float(sr[k + 1]));
            }
            const int xb = (n & 1) * 2;
            B->xmax[xb + hf][row] = fmaxf(mx0, mx1);
            asm volatile("bar.sync 1, 256;" ::: "memory");                    // the 8 softmax warps
            const float mt = fmaxf(B->xmax[xb][row], B->xmax[xb + 1][row]);
            const float m_new = fmaxf(m_used, mt);
            const bool grow = (m_new > m_used) && (n > 0);
            if (__any_sync(0xffffffffu, grow)) {
                // O' must be quiescent: PV(n - 1) complete (o_done), PV(n) not issued before all threads arrive on p_full
                mbar_wait(&B->o_done, (uint32_t)((n - 1) & 1));
                tc_fence_after();
                const float f = grow ? ex2((m_used - m_new) * P_LOG2E) : 1.f;
                uint32_t orr[16];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    tmem_ld16(tO + 16 * c, orr);
                    tmem_wait_ld();
#pragma unroll
                    for (int k = 0; k < 16; ++k) orr[k] = __float_as_uint(__uint_as_float(orr[k]) * f);
                    tmem_st16(tO + 16 * c, orr);
                }
                float s0, s1;
                upk2(lsum, s0, s1);
                lsum = pk2(s0 * f, s1 * f);
            }
            m_used = m_new;
            const float negs = -m_new * P_LOG2E;
            const uint64_t neg2 = pk2(negs, negs);
            // p = 2^(s*log2e - m*log2e) two scores per FFMA2; hi = p truncated to 11 significant bits (exactly representable
            // in fp16 for p >= 2^-14), lo = p - hi exact in fp32; both packed to fp16 and written back 16 keys at a time
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t ph[8], pl[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int k = 16 * c + 2 * t;
                    float t0, t1;
                    upk2(fma2(pk2(__uint_as_float(sr[k]), __uint_as_float(sr[k + 1])), l2e2, neg2), t0, t1);
                    const float p0 = ex2(t0), p1 = ex2(t1);
                    const uint64_t p2 = pk2(p0, p1);
                    lsum = add2(lsum, p2);
                    if (EXACT) {
                        const float h0 = __uint_as_float(__float_as_uint(p0) & 0xFFFFE000u);
                        const float h1 = __uint_as_float(__float_as_uint(p1) & 0xFFFFE000u);
                        ph[t] = cvt_h2(h0, h1);
                        float r0, r1;
                        upk2(add2(p2, pk2(-h0, -h1)), r0, r1);
                        pl[t] = cvt_h2(r0, r1);
                    } else {
                        ph[t] = cvt_h2(p0, p1);
                    }
                }
                tmem_st8(tS + 8 * c, ph);                // keys [32 hf + 16 c, +16) -> columns [32 hf + 8 c, +8)
                if (EXACT) tmem_st8(tS + 16 + 8 * c, pl);
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive_warp(&B->p_full[b]);
            par ^= 1u << b;
            b = b == 2 ? 0 : b + 1;
        }

        // ---- epilogue: this thread finishes output channels [16 hf, 16 hf + 16) of its row
        float s0, s1;
        upk2(lsum, s0, s1);
        B->xsum[hf][row] = s0 + s1;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float l = B->xsum[0][row] + B->xsum[1][row];          // same order in both threads of the row
        const int q = q0 + row;
        float o[16];
        if (T > 0) {
            mbar_wait(&B->o_final, 0);
            tc_fence_after();
            uint32_t o0[16], o1[16];
            tmem_ld16(tmem + lane_addr + 192 + hf * 16, o0);
            tmem_ld16(tmem + lane_addr + 192 + 32 + hf * 16, o1);
            tmem_wait_ld();
#pragma unroll
            for (int k = 0; k < 16; ++k) o[k] = __uint_as_float(o0[k]) + __uint_as_float(o1[k]);
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) o[k] = 0.f;
        }
        if (q < a.N) {
            if (a.splits == 1) {
                const float inv = 1.f / l;
                float* dst = a.O + (size_t)q * a.ldo + h * 32 + hf * 16;
#pragma unroll
                for (int k = 0; k < 16; k += 4)
                    *reinterpret_cast<float4*>(dst + k) = make_float4(o[k] * inv, o[k + 1] * inv, o[k + 2] * inv, o[k + 3] * inv);
            } else {
                float* dst = a.Opart + ((size_t)z * a.N + q) * (a.H * 32) + h * 32 + hf * 16;
#pragma unroll
                for (int k = 0; k < 16; k += 4)
                    *reinterpret_cast<float4*>(dst + k) = make_float4(o[k], o[k + 1], o[k + 2], o[k + 3]);
                if (hf == 0) {
                    a.Mpart[((size_t)z * a.H + h) * a.N + q] = m_used;
                    a.Lpart[((size_t)z * a.H + h) * a.N + q] = l;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == P_MMA_WARP) tmem_dealloc<256>(tmem);
}

static size_t pair_smem_bytes() { return (size_t)P_QBYTES + 2 * P_STAGES * P_KVBYTES + sizeof(BarriersP) + 1024; }

// Host side of the "pair" layout (called from aotb_lt_attn_tc_f16x2 when mode bit 4 is set).
int launch_lt_attn_pair(const void* Qp, int Nq_cap, const void* Kp, const void* Vp, int kv_cap, int N, int Tk,
                        const int* Tk_dev, int H, float* O, int ldo, float* Opart, float* Mpart, float* Lpart, int splits,
                        int exact, int spin, cudaStream_t st) {
    CUtensorMap tq, tk, tv;
    int rc;
    if ((rc = make_tmap_rows64_box(&tq, Qp, Nq_cap, H, P_BM)) != AOTB_OK) return rc;
    if ((rc = make_tmap_rows64_box(&tk, Kp, kv_cap, H, P_BN)) != AOTB_OK) return rc;
    if ((rc = make_tmap_rows64_box(&tv, Vp, kv_cap, H, P_BN)) != AOTB_OK) return rc;
    const size_t smem = pair_smem_bytes();
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(lt_attn_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(lt_attn_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("aotb_lt_attn_tc_f16x2 (pair): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return AOTB_ERR_CUDA;
        }
        configured = true;
    }
    LtArgs2 a;
    a.N = N; a.Tk = Tk; a.Tk_dev = Tk_dev; a.H = H; a.O = O; a.ldo = ldo;
    a.Opart = Opart; a.Mpart = Mpart; a.Lpart = Lpart; a.splits = splits; a.spin = spin;
    dim3 grid(cdiv(N, P_BM), H, splits);
    if (exact) launch(lt_attn_pair_kernel<true>, grid, dim3(P_THREADS), smem, st, tq, tk, tv, a);
    else launch(lt_attn_pair_kernel<false>, grid, dim3(P_THREADS), smem, st, tq, tk, tv, a);
    return check_launch("aotb_lt_attn_tc_f16x2 (pair)");
}

}  // namespace tc
}  // namespace aotb
