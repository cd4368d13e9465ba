// DeAOT long-term attention (GatedPropagation, K1') fused on the 5th-gen tensor cores: Q K^T -> softmax -> P V with the
// scores staying in TMEM.  EXPERIMENTAL: written at the end of round 1 after the GPU budget was spent -- it compiles,
// its mbarrier / buffer protocol is model-checked (scripts/gp_attn_protocol_sim.py), its host logic is verified on CPU,
// but it has NOT run on a GPU yet (AOTB_DEAOT_LT=tc; tests gated by AOTB_TEST_VARIANTS=gp_tc).
//
// Reference computation: GatedPropagation.forward, networks/layers/attention.py:672-704, as DeAOT instantiates it
// (1 head, d_qk = 128, d_v = 1024: networks/layers/transformer.py:541-548): softmax((Q / T) K^T) V over the memory bank.
//
// Why it differs from lt_attn_tc.cu (8 heads x 32): 2304 tensor FLOPs per exponential instead of 128, so this shape is
// tensor-bound, not MUFU-bound; and O = 128 x 1024 fp32 does not fit the 512 TMEM columns of a CTA.  Plan:
//   grid     (query tiles of 128, d_v slices of 128, KV splits); a CTA recomputes Q K^T for its value slice (768 of the
//            1792 tensor-pipe cycles of a 64-key tile)
//   operands the split-fp16 row format of lt_attn_tc.cu, one "head" per 32 channels: Q, K packed [4][rows][64], V packed
//            [32][rows][64] (hi(32) | lo(32) halfs = 128 B rows, TMA 128B swizzle); every smem tile and UMMA descriptor is
//            the layout the AOT kernel already uses, with 64-key tiles so that Q (64 KB) + 2 K stages + 2 V stages fit
//   TMEM     S_0 | S_1 | S_2 (64 fp32 columns = 64 keys each) | O' = 4 chunks x [Vh | Vl] (64 columns each) = 448 columns.
//            P_hi and P_lo of the 16 keys a thread owns overwrite its own 16 score columns (8 + 8).
//   MMA      S(0..2) up front, then per key tile n: wait P(n) -> PV(n) -> S(n+3): scores run ahead of the softmax, the
//            tcgen05.ld of tile n+1 is issued under the ex2 pass of tile n (the "ahead" protocol of lt_attn_tc3_kernel).
//            exact: S = sum over the 4 channel chunks of (Qh Kh + Ql Kh + Qh Kl) = 24 MMAs 128x64x16;
//                   O'_c += (Ph + Pl) [Vh_c | Vl_c] = 32 MMAs 128x64x16.   fast: 8 + 16.
//   warps    0-15 softmax (TMEM lane quadrant w%4, key quarter w/4: 4 threads per row, 16 scores each), 16 K producer,
//            17 V producer (separate rings: K(n+3) must land while V(n) is still live), 18 MMA issuer.
#include "common.cuh"
#include "tc_common.cuh"

namespace aotb {
namespace tc {

constexpr int GP_BM = 128, GP_BK = 64, GP_THREADS = 608;
constexpr int GP_KWARP = 16, GP_VWARP = 17, GP_MMAWARP = 18;
constexpr int GP_QTILE = GP_BM * 128;      // 128 rows x 128 B
constexpr int GP_KVTILE = GP_BK * 128;     // 64 rows x 128 B
constexpr int GP_STAGE = 4 * GP_KVTILE;    // 4 channel chunks per stage
constexpr float GP_LOG2E = 1.4426950408889634f;

struct GpArgs {
    int N, Tk;
    const int* Tk_dev;
    int dv;              // total value width (1024)
    float* O;
    int ldo;
    float* Opart;
    float* Mpart;
    float* Lpart;
    int splits;
    int spin;
};

struct __align__(8) GpBarriers {
    uint64_t q_full;
    uint64_t k_full[2], k_free[2];
    uint64_t v_full[2], v_free[2];
    uint64_t s_full[3];     // S(n) complete in buffer n % 3; use k = n / 3 of a buffer completes phase k
    uint64_t p_full[3];     // 16 arrivals (one per softmax warp): P(n) written over S(n)
    uint64_t o_done[2];     // PV(n) complete, on barrier n & 1 (phase n >> 1).  Two barriers because a parity wait is only
                            // sound when the waiter is at most one phase behind: at tile n the softmax knows PV(n - 3) is
                            // complete (S(n) was issued behind it), i.e. barrier (n - 1) & 1 is at most one phase short.
    uint64_t o_final;       // every PV of this CTA complete (single phase; the epilogue cannot use o_done: it may be 2 behind)
    uint32_t tmem_base;
    float xmax[2][4][GP_BM];
    float xsum[4][GP_BM];
};

__device__ __forceinline__ void tmem_wait_ld16(uint32_t* r) {      // see tmem_wait_ld32
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}

// [heads][rows][64 halfs] operands with a box of 64 halfs x 64 rows x 1 head (the K / V tiles of this kernel)
static int make_tmap_rows64_box64(CUtensorMap* out, const void* base, int rows, int heads) {
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
    cuuint64_t strides[2] = {128, (cuuint64_t)rows * 128};
    cuuint32_t box[3] = {64, (cuuint32_t)GP_BK, 1};
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
__global__ void __launch_bounds__(GP_THREADS, 1)
gp_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const GpArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;                          // 4 channel-chunk tiles of 128 queries
    uint8_t* sK = sQ + 4 * GP_QTILE;             // 2 stages x 4 chunk tiles of 64 keys
    uint8_t* sV = sK + 2 * GP_STAGE;             // 2 stages x 4 chunk tiles of 64 keys (this CTA's 128 value channels)
    GpBarriers* B = reinterpret_cast<GpBarriers*>(sV + 2 * GP_STAGE);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * GP_BM, vs = blockIdx.y, z = blockIdx.z;
    pdl_trigger();

    if (tid == 0) {
        mbar_init(&B->q_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&B->k_full[s], 1); mbar_init(&B->k_free[s], 1);
            mbar_init(&B->v_full[s], 1); mbar_init(&B->v_free[s], 1);
        }
        for (int b = 0; b < 3; ++b) { mbar_init(&B->s_full[b], 1); mbar_init(&B->p_full[b], 4 * GP_BM / 32); }
        mbar_init(&B->o_done[0], 1);
        mbar_init(&B->o_done[1], 1);
        mbar_init(&B->o_final, 1);
        fence_mbar_init();
    }
    if (warp == GP_MMAWARP) tmem_alloc<512>(&B->tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = B->tmem_base;
    pdl_wait();
    const int Tk = a.Tk_dev ? *a.Tk_dev : a.Tk;
    const int tiles_total = (Tk + GP_BK - 1) / GP_BK;
    const int per = (tiles_total + a.splits - 1) / a.splits;
    const int tb = z * per;
    int T = tiles_total - tb;
    T = T < 0 ? 0 : (T > per ? per : T);

    if (warp == GP_KWARP) {
        // ======================= Q once, then the K ring =======================
        if (elect_one() && T > 0) {
            tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK);
            mbar_arrive_expect_tx(&B->q_full, 4 * GP_QTILE);
            for (int c = 0; c < 4; ++c) tma_load_3d(sQ + c * GP_QTILE, &tmQ, &B->q_full, 0, q0, c);
            for (int j = 0; j < T; ++j) {
                const int s = j & 1;
                if (j >= 2) mbar_wait(&B->k_free[s], ((j >> 1) - 1) & 1);        // S(j - 2) has consumed this stage
                mbar_arrive_expect_tx(&B->k_full[s], GP_STAGE);
                for (int c = 0; c < 4; ++c)
                    tma_load_3d(sK + s * GP_STAGE + c * GP_KVTILE, &tmK, &B->k_full[s], 0, (tb + j) * GP_BK, c);
            }
        }
    } else if (warp == GP_VWARP) {
        // ======================= the V ring (this CTA's 4 value chunks) =======================
        if (elect_one() && T > 0) {
            tma_prefetch_desc(&tmV);
            for (int j = 0; j < T; ++j) {
                const int s = j & 1;
                if (j >= 2) mbar_wait(&B->v_free[s], ((j >> 1) - 1) & 1);        // PV(j - 2) has consumed this stage
                mbar_arrive_expect_tx(&B->v_full[s], GP_STAGE);
                for (int c = 0; c < 4; ++c)
                    tma_load_3d(sV + s * GP_STAGE + c * GP_KVTILE, &tmV, &B->v_full[s], 0, (tb + j) * GP_BK, vs * 4 + c);
            }
        }
    } else if (warp == GP_MMAWARP) {
        // ======================= MMA issuer =======================
        if (elect_one() && T > 0) {
            constexpr uint32_t IDESC_S = idesc_f16(128, 64, 0, 0);     // A, B K-major: S[128 x 64] += Q_c[128 x 16] K_c[64 x 16]^T
            constexpr uint32_t IDESC_O = idesc_f16(128, 64, 0, 1);     // B MN-major: O'_c[128 x 64] += P[128 x 16] V_c[16 x 64]
            const uint64_t dQ = smem_desc_sw128(smem_u32(sQ)), dK = smem_desc_sw128(smem_u32(sK)),
                           dV = smem_desc_sw128(smem_u32(sV));
            // descriptor start-address field = byte offset >> 4: +2 per 32 B k-slice, +1024 per 16 KB Q chunk tile,
            // +512 per 8 KB K / V chunk tile, +2048 per 32 KB stage, +128 per 16 key rows of V
            auto issue_S = [&](int n) {
                const int s = n & 1;
                mbar_wait_cp(&B->k_full[s], (uint32_t)((n >> 1) & 1), a.spin);
                tc_fence_after();
                const uint32_t d = tmem + (n % 3) * 64;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint64_t q = dQ + (uint64_t)(c * (GP_QTILE >> 4));
                    const uint64_t k = dK + (uint64_t)(s * (GP_STAGE >> 4) + c * (GP_KVTILE >> 4));
                    mma_ss(d, q, k, IDESC_S, c > 0 ? 1u : 0u);        // k-slices 0,1 = hi ; 2,3 = lo
                    mma_ss(d, q + 2, k + 2, IDESC_S, 1);
                    if (EXACT) {
                        mma_ss(d, q + 4, k, IDESC_S, 1);              // Ql Kh
                        mma_ss(d, q + 6, k + 2, IDESC_S, 1);
                        mma_ss(d, q, k + 4, IDESC_S, 1);              // Qh Kl
                        mma_ss(d, q + 2, k + 6, IDESC_S, 1);
                    }
                }
                mma_commit(&B->k_free[s]);
                mma_commit(&B->s_full[n % 3]);
            };
            auto issue_PV = [&](int n) {
                const int s = n & 1;
                mbar_wait_cp(&B->v_full[s], (uint32_t)((n >> 1) & 1), a.spin);
                tc_fence_after();
                const uint32_t p = tmem + (n % 3) * 64;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t d = tmem + 192 + c * 64;
                    const uint64_t v = dV + (uint64_t)(s * (GP_STAGE >> 4) + c * (GP_KVTILE >> 4));
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)       // P_hi of keys [16 kk, +16) at columns [16 kk, 16 kk + 8)
                        mma_ts(d, p + 16 * kk, v + 128 * kk, IDESC_O, (kk > 0 || n > 0) ? 1u : 0u);
                    if (EXACT) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)   // P_lo at columns [16 kk + 8, 16 kk + 16)
                            mma_ts(d, p + 16 * kk + 8, v + 128 * kk, IDESC_O, 1);
                    }
                }
                mma_commit(&B->v_free[s]);
                mma_commit(&B->o_done[n & 1]);
                if (n + 1 == T) mma_commit(&B->o_final);
            };
            mbar_wait(&B->q_full, 0);
            tc_fence_after();
            for (int n = 0; n < 3 && n < T; ++n) issue_S(n);
            for (int n = 0; n < T; ++n) {
                mbar_wait_cp(&B->p_full[n % 3], (uint32_t)((n / 3) & 1), a.spin);
                tc_fence_after();
                issue_PV(n);
                if (n + 3 < T) issue_S(n + 3);
            }
        }
    } else {
        // ======================= softmax (16 warps, 4 threads per query row, 16 scores each) =======================
        const int qt = warp >> 2, wq = warp & 3;
        const int row = wq * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
        const uint32_t tO = tmem + lane_addr + 192 + qt * 64;      // this thread's quarter of the O' columns = chunk qt
        float m_used = -INFINITY, l0 = 0.f, l1 = 0.f;
        uint32_t srA[16], srB[16];
        auto tile = [&](uint32_t (&sr)[16], uint32_t (&srn)[16], const int n, const int b, const int bn, const bool has_next,
                        const uint32_t next_par) {
            const uint32_t tS = tmem + lane_addr + b * 64 + qt * 16;
            const int key0 = (tb + n) * GP_BK + qt * 16;
            if (key0 + 16 > Tk) {                    // warp-uniform: only the last key tile of the bank is ragged
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    if (key0 + k >= Tk) sr[k] = __float_as_uint(-INFINITY);
            }
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                mx0 = fmaxf(mx0, __uint_as_float(sr[k]));
                mx1 = fmaxf(mx1, __uint_as_float(sr[k + 1]));
            }
            B->xmax[n & 1][qt][row] = fmaxf(mx0, mx1);
            asm volatile("bar.sync 1, 512;" ::: "memory");
            const float mt = fmaxf(fmaxf(B->xmax[n & 1][0][row], B->xmax[n & 1][1][row]),
                                   fmaxf(B->xmax[n & 1][2][row], B->xmax[n & 1][3][row]));
            const float m_new = fmaxf(m_used, mt);
            const bool grow = (m_new > m_used) && (n > 0);
            if (__any_sync(0xffffffffu, grow)) {
                // O' must be quiescent: PV(n - 1) complete (and with it every earlier PV: in-order pipe), PV(n) not issued
                // before all threads arrive on p_full
                mbar_wait(&B->o_done[(n - 1) & 1], (uint32_t)(((n - 1) >> 1) & 1));
                tc_fence_after();
                const float f = grow ? ex2((m_used - m_new) * GP_LOG2E) : 1.f;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    uint32_t orr[16];
                    tmem_ld16(tO + 16 * cc, orr);
                    tmem_wait_ld();
#pragma unroll
                    for (int k = 0; k < 16; ++k) orr[k] = __float_as_uint(__uint_as_float(orr[k]) * f);
                    tmem_st16(tO + 16 * cc, orr);
                }
                l0 *= f; l1 *= f;
            }
            m_used = m_new;
            const float neg = m_new * GP_LOG2E;
            if (has_next) {                          // S(n + 1) was issued behind PV(n - 2): complete in steady state
                mbar_wait_cp(&B->s_full[bn], next_par, a.spin);
                tc_fence_after();
                tmem_ld16(tmem + lane_addr + bn * 64 + qt * 16, srn);
            }
            uint32_t ph[8], pl[8];
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const float p0 = ex2(fmaf(__uint_as_float(sr[2 * t]), GP_LOG2E, -neg));
                const float p1 = ex2(fmaf(__uint_as_float(sr[2 * t + 1]), GP_LOG2E, -neg));
                s0 += p0; s1 += p1;
                const __half2 hi = __floats2half2_rn(p0, p1);
                ph[t] = *reinterpret_cast<const uint32_t*>(&hi);
                if (EXACT) {
                    const __half2 lo = __floats2half2_rn(p0 - __low2float(hi), p1 - __high2float(hi));
                    pl[t] = *reinterpret_cast<const uint32_t*>(&lo);
                }
            }
            l0 += s0; l1 += s1;
            tmem_st8(tS, ph);                        // keys [16 qt, +16) -> columns [16 qt, 16 qt + 8)
            if (EXACT) tmem_st8(tS + 8, pl);         //                   -> columns [16 qt + 8, 16 qt + 16)
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive_warp(&B->p_full[b]);
            if (has_next) tmem_wait_ld16(srn);
        };
        if (T > 0) {
            mbar_wait(&B->s_full[0], 0);
            tc_fence_after();
            tmem_ld16(tmem + lane_addr + qt * 16, srA);
            tmem_wait_ld16(srA);
        }
        // buffer b of tile n and the parity of its s_full phase rotate with n: bit b of `par` = uses of buffer b so far (mod 2)
        int b = 0;
        uint32_t par = 1u;                           // the prologue consumed phase 0 of s_full[0]
        for (int n = 0; n < T; n += 2) {
            int bn = b == 2 ? 0 : b + 1;
            tile(srA, srB, n, b, bn, n + 1 < T, (par >> bn) & 1u);
            par ^= 1u << bn;
            b = bn;
            if (n + 1 < T) {
                bn = b == 2 ? 0 : b + 1;
                tile(srB, srA, n + 1, b, bn, n + 2 < T, (par >> bn) & 1u);
                par ^= 1u << bn;
                b = bn;
            }
        }

        // ---- epilogue: this thread finishes value channels [32 qt, 32 qt + 32) of the slice for its row
        B->xsum[qt][row] = l0 + l1;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        const float l = (B->xsum[0][row] + B->xsum[1][row]) + (B->xsum[2][row] + B->xsum[3][row]);
        const int q = q0 + row;
        if (T > 0) {
            mbar_wait(&B->o_final, 0);                               // the last PV (and every earlier one) complete
            tc_fence_after();
        }
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
            float o[16];
            if (T > 0) {
                uint32_t o0[16], o1[16];
                tmem_ld16(tO + 16 * hb, o0);                 // [Vh part | Vl part] of chunk qt: columns [0,32) | [32,64)
                tmem_ld16(tO + 32 + 16 * hb, o1);
                tmem_wait_ld();
#pragma unroll
                for (int k = 0; k < 16; ++k) o[k] = __uint_as_float(o0[k]) + __uint_as_float(o1[k]);
            } else {
#pragma unroll
                for (int k = 0; k < 16; ++k) o[k] = 0.f;
            }
            if (q < a.N) {
                const int col = vs * 128 + qt * 32 + hb * 16;
                if (a.splits == 1) {
                    const float inv = 1.f / l;
                    float* dst = a.O + (size_t)q * a.ldo + col;
#pragma unroll
                    for (int k = 0; k < 16; k += 4)
                        *reinterpret_cast<float4*>(dst + k) = make_float4(o[k] * inv, o[k + 1] * inv, o[k + 2] * inv, o[k + 3] * inv);
                } else {
                    float* dst = a.Opart + ((size_t)z * a.N + q) * a.dv + col;
#pragma unroll
                    for (int k = 0; k < 16; k += 4)
                        *reinterpret_cast<float4*>(dst + k) = make_float4(o[k], o[k + 1], o[k + 2], o[k + 3]);
                }
            }
        }
        if (a.splits > 1 && q < a.N && qt == 0 && vs == 0) {      // one head: Mpart / Lpart [splits][1][N]
            a.Mpart[(size_t)z * a.N + q] = m_used;
            a.Lpart[(size_t)z * a.N + q] = l;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == GP_MMAWARP) tmem_dealloc<512>(tmem);
}

}  // namespace tc
}  // namespace aotb

using namespace aotb;

// Qp [4][Nq_cap][64], Kp [4][kv_cap][64], Vp [dv/32][kv_cap][64]: split-fp16 rows packed by aotb_tc_pack_rows_f16x2 with
// one "head" per 32 channels (Q divided by T = sqrt(128) when packed).  O [N][ldo] fp32.  splits > 1 writes un-normalised
// partials (Opart [splits][N][dv], Mpart / Lpart [splits][1][N]) for aotb_attn_merge_f32 (H = 1, d_v = dv).
extern "C" int aotb_gp_attn_tc_f16x2(const void* Qp, int Nq_cap, const void* Kp, const void* Vp, int kv_cap, int N, int Tk,
                                     const int* Tk_dev, int dv, float* O, int ldo, float* Opart, float* Mpart,
                                     float* Lpart, int splits, int exact, void* stream) {
    AOTB_REQUIRE(Qp && Kp && Vp && N > 0 && (Tk > 0 || Tk_dev) && splits >= 1 && dv > 0 && dv % 128 == 0,
                 "aotb_gp_attn_tc_f16x2: bad args");
    AOTB_REQUIRE(Nq_cap >= ((N + 127) / 128) * 128, "aotb_gp_attn_tc_f16x2: Q buffer must be padded to 128 rows");
    AOTB_REQUIRE(splits == 1 ? (O != nullptr && ldo % 4 == 0) : (Opart && Mpart && Lpart),
                 "aotb_gp_attn_tc_f16x2: output buffers");
    AOTB_REQUIRE(((uintptr_t)Qp | (uintptr_t)Kp | (uintptr_t)Vp) % 128 == 0, "aotb_gp_attn_tc_f16x2: alignment");
    CUtensorMap tq, tk, tv;
    int rc;
    if ((rc = tc::make_tmap_rows64(&tq, Qp, Nq_cap, 4)) != AOTB_OK) return rc;
    if ((rc = tc::make_tmap_rows64_box64(&tk, Kp, kv_cap, 4)) != AOTB_OK) return rc;
    if ((rc = tc::make_tmap_rows64_box64(&tv, Vp, kv_cap, dv / 32)) != AOTB_OK) return rc;
    const size_t smem = (size_t)4 * tc::GP_QTILE + 4 * tc::GP_STAGE + sizeof(tc::GpBarriers) + 1024;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tc::gp_attn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(tc::gp_attn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("aotb_gp_attn_tc_f16x2: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return AOTB_ERR_CUDA;
        }
        configured = true;
    }
    tc::GpArgs a;
    a.N = N; a.Tk = Tk; a.Tk_dev = Tk_dev; a.dv = dv; a.O = O; a.ldo = ldo;
    a.Opart = Opart; a.Mpart = Mpart; a.Lpart = Lpart; a.splits = splits; a.spin = (exact >> 2) & 1;
    dim3 grid(cdiv(N, tc::GP_BM), dv / 128, splits);
    if (exact & 1)
        launch(tc::gp_attn_tc_kernel<true>, dim3(grid), dim3(tc::GP_THREADS), smem, (cudaStream_t)stream, tq, tk, tv, a);
    else
        launch(tc::gp_attn_tc_kernel<false>, dim3(grid), dim3(tc::GP_THREADS), smem, (cudaStream_t)stream, tq, tk, tv, a);
    return check_launch("aotb_gp_attn_tc_f16x2");
}
