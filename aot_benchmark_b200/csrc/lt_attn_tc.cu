// Long-term attention (K1) on the 5th-gen tensor cores: fused Q K^T -> softmax -> P V, AOT head shape
// (H heads x d = 32), scores never leave the SM.
//
// Reference computation: MultiheadAttention.forward(use_linear=False), networks/layers/attention.py:82-117,
// called at networks/layers/transformer.py:346 (long-term) and :324 (self-attention, Tk = N).
//
// Design (one CTA = 256 queries x 1 head x one KV split, 1 CTA / SM, 576 threads):
//   warp 16     TMA producer: Q tiles once, then a 3-stage ring of K/V tiles (128 keys) via
//               cp.async.bulk.tensor + mbarrier complete_tx
//   warp 17     tcgen05.mma issuer (one elected thread): S_i = Q_i K_j^T into TMEM, later O_i += P_i V_j
//   warps 0-15  softmax: all 16 warps work on ONE 128 x 128 score tile at a time, alternating between S_0 (query rows
//               0..127) and S_1 (rows 128..255).  Warp w owns TMEM lanes 32(w%4)..+31 and key columns 32(w/4)..+31, i.e.
//               four threads share a row: tcgen05.ld of its 32 scores ONCE -> quarter-row max, exchanged through shared
//               memory -> ex2 / partial row sum -> P (fp16 hi / lo) back into TMEM as the A operand of the PV MMA.
//               TMEM reads run at 64 B/clk/SM, so a 128 x 128 fp32 tile costs 1024 cycles per pass: the earlier
//               layouts (1 or 2 threads per row) could not keep a row in registers and read every tile twice
//               (max pass + exp pass), which capped the kernel at ~2 x 2048 cycles per key tile.  While the warps
//               are on S_i the tensor pipe runs PV_(1-i) and the next S_(1-i) (issue order PV_0, S_0', PV_1, S_1').
//   TMEM        S_0 | S_1 (128 fp32 columns each) | O_0 | O_1 | Plo_0 | Plo_1 (64 each).  P_hi of the 32 keys a thread
//               owns overwrites the first 16 of its own 32 S columns: keys [32t, 32t+32) -> columns [32t, 32t+16).
//
// Precision ("fp16x2"): every fp32 operand x is carried as hi = fp16(x), lo = fp16(x - hi); rows of the
// packed operands are [hi(32) | lo(32)] halfs = 128 bytes (the same bytes as fp32, one TMA swizzle atom).
//   exact mode  S = Qh Kh + Ql Kh + Qh Kl   (6 MMAs of 128x128x16),  O' = (Ph + Pl) [Vh | Vl]  (16 MMAs 128x64x16)
//   fast mode   S = Qh Kh                   (2 MMAs),               O' =  Ph       [Vh | Vl]  ( 8 MMAs)
// and O = O'[:, :32] + O'[:, 32:].  With d = 32 the kernel is MUFU(ex2)-bound (128 tensor FLOPs per
// exponential), so the extra MMAs of the exact mode ride in otherwise idle tensor-pipe slots.
// fp32 accumulation everywhere; Q is pre-divided by T (true division, attention.py:82) when packed.
#include "common.cuh"
#include "tc_common.cuh"
#include <cstdlib>

namespace aotb {
namespace tc {

constexpr int BM = 128, BN = 128, STAGES = 3, NTHREADS = 576, TMA_WARP = 16, MMA_WARP = 17;
constexpr int TILE_BYTES = BN * 128;  // 128 rows x 128 B
constexpr float LOG2E = 1.4426950408889634f;

int make_tmap_rows64(CUtensorMap* out, const void* base, int rows, int heads) {
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
    cuuint32_t box[3] = {64, 128, 1};
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

struct LtArgs {
    int N, Tk;
    const int* Tk_dev;
    int H;
    float* O;
    int ldo;
    float* Opart;
    float* Mpart;
    float* Lpart;
    int splits;
    int exact;
    int spin;    // 1: critical-path mbarrier waits poll (no suspend-time hint) instead of sleeping
    float* dbg;  // optional: CTA (0,0,0) dumps S_0(tile 0) [128][128] then O'_0 [128][64]
};

struct __align__(8) Barriers {
    uint64_t q_full;
    uint64_t kv_full[STAGES];
    uint64_t kv_free[STAGES];
    uint64_t s_full[2];
    uint64_t p_full[2];
    uint64_t o_final[2];
    uint32_t tmem_base;
    // partial row maxima / sums exchanged between the threads that share a query row:
    //   one-tile layout  (4 threads per row): xmax[tile parity * 4 + column quarter], xsum[tile * 4 + column quarter]
    //   two-group layout (2 threads per row): xmax[((j & 1) * 2 + group) * 2 + half], xsum[group * 2 + half]
    float xmax[8][BM];
    float xsum[8][BM];
};

template <bool EXACT, bool GROUPS>
__global__ void __launch_bounds__(NTHREADS, 1)
lt_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const LtArgs a) {
    
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;                                  // 2 tiles
    uint8_t* sK = sQ + 2 * TILE_BYTES;                   // STAGES tiles
    uint8_t* sV = sK + STAGES * TILE_BYTES;              // STAGES tiles
    Barriers* B = reinterpret_cast<Barriers*>(sV + STAGES * TILE_BYTES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * (2 * BM), h = blockIdx.y, z = blockIdx.z;
    pdl_trigger();      // the next kernel may start its prologue; it waits for this grid before reading our output

    if (tid == 0) {
        mbar_init(&B->q_full, 1);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&B->kv_full[s], 1); mbar_init(&B->kv_free[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&B->s_full[i], 1); mbar_init(&B->p_full[i], GROUPS ? 2 * BM / 32 : 4 * BM / 32); mbar_init(&B->o_final[i], 1); }
        fence_mbar_init();
    }
    if (warp == MMA_WARP) tmem_alloc<512>(&B->tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = B->tmem_base;
    pdl_wait();         // everything below reads tensors (and the key counter) written by earlier kernels
    const int Tk = a.Tk_dev ? *a.Tk_dev : a.Tk;
    const int tiles_total = (Tk + BN - 1) / BN;
    const int per = (tiles_total + a.splits - 1) / a.splits;
    const int tb = z * per;
    int T = tiles_total - tb;
    T = T < 0 ? 0 : (T > per ? per : T);

    if (warp == TMA_WARP) {
        // ======================= TMA producer =======================
        if (elect_one() && T > 0) {
            tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
            mbar_arrive_expect_tx(&B->q_full, 2 * TILE_BYTES);
            tma_load_3d(sQ, &tmQ, &B->q_full, 0, q0, h);
            tma_load_3d(sQ + TILE_BYTES, &tmQ, &B->q_full, 0, q0 + BM, h);
            for (int j = 0; j < T; ++j) {
                const int s = j % STAGES;
                if (j >= STAGES) mbar_wait(&B->kv_free[s], ((j / STAGES) - 1) & 1);
                mbar_arrive_expect_tx(&B->kv_full[s], 2 * TILE_BYTES);
                tma_load_3d(sK + s * TILE_BYTES, &tmK, &B->kv_full[s], 0, (tb + j) * BN, h);
                tma_load_3d(sV + s * TILE_BYTES, &tmV, &B->kv_full[s], 0, (tb + j) * BN, h);
            }
        }
    } else if (warp == MMA_WARP) {
        // ======================= MMA issuer =======================
        if (elect_one() && T > 0) {
            constexpr uint32_t IDESC_S = idesc_f16(128, 128, 0, 0);
            constexpr uint32_t IDESC_O = idesc_f16(128, 64, 0, 1);
            // Descriptors are built once; per-MMA work is a 64-bit add (byte offsets >> 4 land in the 14-bit
            // start-address field: +2 per 32 B k-slice, +1024 per 16 KB stage, +128 per 16 key rows of V).
            const uint64_t dQ0 = smem_desc_sw128(smem_u32(sQ)), dQ1 = smem_desc_sw128(smem_u32(sQ) + TILE_BYTES);
            const uint64_t dK = smem_desc_sw128(smem_u32(sK)), dV = smem_desc_sw128(smem_u32(sV));
            auto issue_S = [&](int i, int s) {
                const uint64_t q = i ? dQ1 : dQ0;
                const uint64_t k = dK + (uint64_t)(s * (TILE_BYTES >> 4));
                const uint32_t d = tmem + i * 128;
                // k-slices of 16 halfs = 32 B inside the 128 B row: 0,1 = hi ; 2,3 = lo
                mma_ss(d, q, k, IDESC_S, 0);
                mma_ss(d, q + 2, k + 2, IDESC_S, 1);
                if (EXACT) {
                    mma_ss(d, q + 4, k, IDESC_S, 1);          // Ql Kh
                    mma_ss(d, q + 6, k + 2, IDESC_S, 1);
                    mma_ss(d, q, k + 4, IDESC_S, 1);          // Qh Kl
                    mma_ss(d, q + 2, k + 6, IDESC_S, 1);
                }
            };
            auto issue_PV = [&](int i, int s, uint32_t acc) {
                const uint64_t v = dV + (uint64_t)(s * (TILE_BYTES >> 4));
                const uint32_t d = tmem + 256 + i * 64;
                const uint32_t p = tmem + i * 128;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    // one-tile layout : P_hi of keys [32t, 32t+32) sits at columns [32t, 32t+16)
                    // two-group layout: P_hi of keys [0,64) at columns [0,32), keys [64,128) at [64,96)
                    const uint32_t pc = GROUPS ? 8 * kk + (kk >= 4 ? 32 : 0) : 32 * (kk >> 1) + 8 * (kk & 1);
                    mma_ts(d, p + pc, v + 128 * kk, IDESC_O, (kk > 0) ? 1u : acc);
                }
                if (EXACT) {
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) mma_ts(d, tmem + 384 + i * 64 + 8 * kk, v + 128 * kk, IDESC_O, 1);
                }
            };
            mbar_wait(&B->q_full, 0);
            mbar_wait(&B->kv_full[0], 0);
            tc_fence_after();
            issue_S(0, 0); mma_commit(&B->s_full[0]);
            issue_S(1, 0); mma_commit(&B->s_full[1]);
            for (int j = 0; j < T; ++j) {
                const int s = j % STAGES;
                for (int i = 0; i < 2; ++i) {
                    mbar_wait_cp(&B->p_full[i], j & 1, a.spin);
                    tc_fence_after();
                    issue_PV(i, s, j > 0 ? 1u : 0u);
                    if (i == 1) mma_commit(&B->kv_free[s]);
                    if (j + 1 < T) {
                        const int s2 = (j + 1) % STAGES;
                        if (i == 0) { mbar_wait_cp(&B->kv_full[s2], ((j + 1) / STAGES) & 1, a.spin); tc_fence_after(); }
                        issue_S(i, s2);
                        mma_commit(&B->s_full[i]);
                    } else {
                        mma_commit(&B->o_final[i]);
                    }
                }
            }
        }
    } else if constexpr (GROUPS) {
        // ======================= softmax, two-group layout =======================
        // Warps 0-7 own query tile 0 (S_0, O_0), warps 8-15 query tile 1; inside a group warp w owns TMEM lanes
        // 32(w%4)..+31 and the key half (w/4)%2, i.e. two threads share a row and each keeps its 64 scores of a tile in
        // registers: ONE tcgen05.ld pass per tile (the earlier two-group kernel re-read every tile for the exp pass).
        // The groups run out of phase -- S_1 is issued ~900 cycles after S_0 -- so while one group is on the TMEM read
        // port (64 B/clk: 1024 clk per 128x128 tile) the other is on the MUFU pipe (16 ex2/clk: 1024 clk per tile)
        // instead of all 16 warps queuing on the same resource in lockstep.
        const int wg = warp >> 3, half = (warp >> 2) & 1, wq = warp & 3;
        const int row = wq * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
        const uint32_t tS = tmem + lane_addr + wg * 128 + half * 64;      // this thread's 64 score columns
        const uint32_t tO = tmem + lane_addr + 256 + wg * 64;
        const uint32_t tPl = tmem + lane_addr + 384 + wg * 64 + half * 32;
        const int q = q0 + wg * BM + row;
        const bool dump = a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && wg == 0;
        float m_used = -INFINITY, l0 = 0.f, l1 = 0.f;

        for (int j = 0; j < T; ++j) {
            mbar_wait_cp(&B->s_full[wg], j & 1, a.spin);
            tc_fence_after();
            const int key0 = (tb + j) * BN + half * 64;
            uint32_t sr[64];
            tmem_ld32(tS + 0, sr);
            tmem_ld32(tS + 32, sr + 32);
            tmem_wait_ld();
            if (dump && j == 0) {
#pragma unroll
                for (int k = 0; k < 64; ++k) a.dbg[row * 128 + half * 64 + k] = __uint_as_float(sr[k]);
            }
            if (key0 + 64 > Tk) {                    // warp-uniform: only the last key tile of the bank is ragged
#pragma unroll
                for (int k = 0; k < 64; ++k)
                    if (key0 + k >= Tk) sr[k] = __float_as_uint(-INFINITY);
            }
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int k = 0; k < 64; k += 2) {
                mx0 = fmaxf(mx0, __uint_as_float(sr[k]));
                mx1 = fmaxf(mx1, __uint_as_float(sr[k + 1]));
            }
            float mt = fmaxf(mx0, mx1);
            const int xb = ((j & 1) * 2 + wg) * 2;
            B->xmax[xb + half][row] = mt;
            asm volatile("bar.sync %0, 256;" ::"r"(1 + wg) : "memory");       // the 8 warps of this group
            mt = fmaxf(mt, B->xmax[xb + (half ^ 1)][row]);
            const float m_new = fmaxf(m_used, mt);
            const bool grow = (m_new > m_used) && (j > 0);
            if (__any_sync(0xffffffffu, grow)) {
                // rescale the running output of this warp's rows -- this thread's 32 of the 64 O' columns -- and its sum
                // (O_i is quiescent here: every MMA issued before S_i(j) has completed, PV_i(j) is not issued until
                // all 256 threads of the group arrive on p_full)
                const float f = grow ? ex2((m_used - m_new) * LOG2E) : 1.f;
                uint32_t orr[16];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    tmem_ld16(tO + half * 32 + 16 * c, orr);
                    tmem_wait_ld();
#pragma unroll
                    for (int k = 0; k < 16; ++k) orr[k] = __float_as_uint(__uint_as_float(orr[k]) * f);
                    tmem_st16(tO + half * 32 + 16 * c, orr);
                }
                l0 *= f; l1 *= f;
            }
            m_used = m_new;
            const float neg = m_used * LOG2E;
            // p = 2^(s*log2e - m*log2e), partial row sum, fp16 hi / lo split, back into TMEM 16 keys at a time.  P_hi of
            // chunk c lands on columns [8c, 8c+8) of this thread's own S columns (all 64 scores are in registers).
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t ph[8], pl[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const float p0 = ex2(fmaf(__uint_as_float(sr[16 * c + 2 * t]), LOG2E, -neg));
                    const float p1 = ex2(fmaf(__uint_as_float(sr[16 * c + 2 * t + 1]), LOG2E, -neg));
                    l0 += p0; l1 += p1;
                    const __half2 hi = __floats2half2_rn(p0, p1);
                    ph[t] = *reinterpret_cast<const uint32_t*>(&hi);
                    if (EXACT) {
                        const __half2 lo = __floats2half2_rn(p0 - __low2float(hi), p1 - __high2float(hi));
                        pl[t] = *reinterpret_cast<const uint32_t*>(&lo);
                    }
                }
                tmem_st8(tS + 8 * c, ph);
                if (EXACT) tmem_st8(tPl + 8 * c, pl);
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive_warp(&B->p_full[wg]);
        }

        // ---- epilogue: this thread finishes output channels [16*half, 16*half+16) of its row
        float l = l0 + l1;
        B->xsum[wg * 2 + half][row] = l;
        asm volatile("bar.sync %0, 256;" ::"r"(1 + wg) : "memory");
        l = B->xsum[wg * 2 + 0][row] + B->xsum[wg * 2 + 1][row];          // same order in both threads of the row
        float o[16];
        if (T > 0) {
            mbar_wait(&B->o_final[wg], 0);
            tc_fence_after();
            uint32_t o0[16], o1[16];
            tmem_ld16(tO + half * 16, o0);
            tmem_ld16(tO + 32 + half * 16, o1);
            tmem_wait_ld();
            if (dump) {
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    a.dbg[128 * 128 + row * 64 + half * 16 + k] = __uint_as_float(o0[k]);
                    a.dbg[128 * 128 + row * 64 + 32 + half * 16 + k] = __uint_as_float(o1[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) o[k] = __uint_as_float(o0[k]) + __uint_as_float(o1[k]);
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) o[k] = 0.f;
        }
        if (q < a.N) {
            if (a.splits == 1) {
                const float inv = 1.f / l;
                float* dst = a.O + (size_t)q * a.ldo + h * 32 + half * 16;
#pragma unroll
                for (int k = 0; k < 16; k += 4)
                    *reinterpret_cast<float4*>(dst + k) = make_float4(o[k] * inv, o[k + 1] * inv, o[k + 2] * inv, o[k + 3] * inv);
            } else {
                float* dst = a.Opart + ((size_t)z * a.N + q) * (a.H * 32) + h * 32 + half * 16;
#pragma unroll
                for (int k = 0; k < 16; k += 4)
                    *reinterpret_cast<float4*>(dst + k) = make_float4(o[k], o[k + 1], o[k + 2], o[k + 3]);
                if (half == 0) {
                    a.Mpart[((size_t)z * a.H + h) * a.N + q] = m_used;
                    a.Lpart[((size_t)z * a.H + h) * a.N + q] = l;
                }
            }
        }
    } else {
        // ======================= softmax (16 warps, one score tile at a time) =======================
        const int qt = warp >> 2, wq = warp & 3;           // column quarter, TMEM lane quadrant
        const int row = wq * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
        float m_used[2] = {-INFINITY, -INFINITY}, l0[2] = {0.f, 0.f}, l1[2] = {0.f, 0.f};

        for (int j = 0; j < T; ++j) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const uint32_t tS = tmem + lane_addr + i * 128 + qt * 32;      // this thread's 32 score columns
                const uint32_t tO = tmem + lane_addr + 256 + i * 64 + qt * 16;  // its 16 of the 64 O' columns
                const uint32_t tPl = tmem + lane_addr + 384 + i * 64 + qt * 16;
                const bool dump = a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && i == 0;
                mbar_wait_cp(&B->s_full[i], j & 1, a.spin);
                tc_fence_after();
                const int key0 = (tb + j) * BN + qt * 32;
                uint32_t sr[32];
                tmem_ld32(tS, sr);
                tmem_wait_ld();
                if (dump && j == 0) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) a.dbg[row * 128 + qt * 32 + k] = __uint_as_float(sr[k]);
                }
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
                const int par = i;                        // tiles alternate 0,1,0,1: the tile index is its own parity
                B->xmax[par * 4 + qt][row] = fmaxf(mx0, mx1);
                asm volatile("bar.sync 1, 512;" ::: "memory");                    // the 16 softmax warps
                const float mt = fmaxf(fmaxf(B->xmax[par * 4 + 0][row], B->xmax[par * 4 + 1][row]),
                                       fmaxf(B->xmax[par * 4 + 2][row], B->xmax[par * 4 + 3][row]));
                const float m_new = fmaxf(m_used[i], mt);
                const bool grow = (m_new > m_used[i]) && (j > 0);
                if (__any_sync(0xffffffffu, grow)) {
                    // rescale the running output of this warp's rows -- this thread's 16 of the 64 O' columns -- and
                    // its partial sum (O_i is quiescent here: every MMA issued before S_i(j) has completed, PV_i(j)
                    // is not issued until all 512 threads arrive on p_full)
                    const float f = grow ? ex2((m_used[i] - m_new) * LOG2E) : 1.f;
                    uint32_t orr[16];
                    tmem_ld16(tO, orr);
                    tmem_wait_ld();
#pragma unroll
                    for (int k = 0; k < 16; ++k) orr[k] = __float_as_uint(__uint_as_float(orr[k]) * f);
                    tmem_st16(tO, orr);
                    l0[i] *= f; l1[i] *= f;
                }
                m_used[i] = m_new;
                const float neg = m_new * LOG2E;
                // p = 2^(s*log2e - m*log2e) two scores per FFMA2, packed partial row sums (FADD2); fp16 split of P: hi = p
                // truncated to 11 significant bits (one LOP3; exactly representable in fp16 for p >= 2^-14), lo = p - hi exact
                // in fp32 -- no fp16 -> fp32 unpack: 9.5 instead of 14 instructions per score pair
                uint32_t ph[16], pl[16];
                const float negs = -neg;
                const uint64_t l2e2 = pk2(LOG2E, LOG2E), neg2 = pk2(negs, negs);
                uint64_t ls = pk2(0.f, 0.f);
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    float a0, a1;
                    upk2(fma2(pk2(__uint_as_float(sr[2 * t]), __uint_as_float(sr[2 * t + 1])), l2e2, neg2), a0, a1);
                    const float p0 = ex2(a0), p1 = ex2(a1);
                    const uint64_t p2 = pk2(p0, p1);
                    ls = add2(ls, p2);
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
                {
                    float s0, s1;
                    upk2(ls, s0, s1);
                    l0[i] += s0; l1[i] += s1;
                }
                tmem_st16(tS, ph);
                if (EXACT) tmem_st16(tPl, pl);
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive_warp(&B->p_full[i]);
            }
        }

        // ---- epilogue: per tile, this thread finishes output channels [8*qt, 8*qt+8) of its row
#pragma unroll
        for (int i = 0; i < 2; ++i) B->xsum[i * 4 + qt][row] = l0[i] + l1[i];
        asm volatile("bar.sync 1, 512;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float l = (B->xsum[i * 4 + 0][row] + B->xsum[i * 4 + 1][row]) + (B->xsum[i * 4 + 2][row] + B->xsum[i * 4 + 3][row]);
            const uint32_t tO = tmem + lane_addr + 256 + i * 64;
            const bool dump = a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && i == 0;
            const int q = q0 + i * BM + row;
            float o[8];
            if (T > 0) {
                mbar_wait(&B->o_final[i], 0);
                tc_fence_after();
                uint32_t o0[8], o1[8];
                tmem_ld8(tO + qt * 8, o0);
                tmem_ld8(tO + 32 + qt * 8, o1);
                tmem_wait_ld();
                if (dump) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        a.dbg[128 * 128 + row * 64 + qt * 8 + k] = __uint_as_float(o0[k]);
                        a.dbg[128 * 128 + row * 64 + 32 + qt * 8 + k] = __uint_as_float(o1[k]);
                    }
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = __uint_as_float(o0[k]) + __uint_as_float(o1[k]);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = 0.f;
            }
            if (q < a.N) {
                if (a.splits == 1) {
                    const float inv = 1.f / l;
                    float* dst = a.O + (size_t)q * a.ldo + h * 32 + qt * 8;
                    *reinterpret_cast<float4*>(dst) = make_float4(o[0] * inv, o[1] * inv, o[2] * inv, o[3] * inv);
                    *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4] * inv, o[5] * inv, o[6] * inv, o[7] * inv);
                } else {
                    float* dst = a.Opart + ((size_t)z * a.N + q) * (a.H * 32) + h * 32 + qt * 8;
                    *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
                    if (qt == 0) {
                        a.Mpart[((size_t)z * a.H + h) * a.N + q] = m_used[i];
                        a.Lpart[((size_t)z * a.H + h) * a.N + q] = l;
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------ "ahead" layout: scores one tile ahead of the softmax
// Same arithmetic as lt_attn_tc_kernel<EXACT, false>, different TMEM plan and issue order (round-1 trips 16-18 showed
// that the tile time of every two-buffer layout is the serial chain TMEM read -> max exchange -> ex2 pass, because the
// next score tile does not exist yet while the current one is processed: S_i aliases P_i, so S_i(j+1) queues behind
// PV_i(j)).  Here:
//   TMEM   S_0 | S_1 | S_2 (128 fp32 columns each) | O_0 | O_1 (64 each) = 512 columns.  Score tile n = 2j + i (key tile
//          j, query tile i) lives in buffer n % 3.  P_hi AND P_lo of the 32 keys a thread owns overwrite its own 32
//          score columns (hi: columns [32t, 32t+16), lo: [32t+16, 32t+32)), which frees the former Plo_0 | Plo_1.
//   MMA    S(0), S(1), S(2) up front; then for every n: wait P(n) -> PV(n) -> S(n+3) (same buffer, in pipe order).
//          While the softmax warps are on tile n, S(n+1) has long completed (it was issued behind PV(n-2)).
//   softmax 16 warps, 4 threads per row as in the default layout, but the tcgen05.ld of tile n+1 is issued before the
//          ex2 pass of tile n and completed after it (two register sets; P is produced 16 keys at a time to stay inside
//          the 96-register budget), so the TMEM read port works under the MUFU pass instead of in series with it.
//   O_i    PV(n) accumulates into O_(n % 2).  S(n) is no longer ordered behind PV(n-2), so the (rare) rescale of O_i
//          waits for o_done[i] (committed after every PV) instead of relying on s_full.
//   smem   Q 2 tiles + 4-stage K/V ring (S runs up to 1.5 key tiles ahead of PV) = 160 KB.
constexpr int STAGES3 = 4;

struct __align__(8) Barriers3 {
    uint64_t q_full;
    uint64_t kv_full[STAGES3];
    uint64_t kv_free[STAGES3];
    uint64_t s_full[3];     // S(n) complete in buffer n % 3; use k = n / 3 of a buffer completes phase k
    uint64_t p_full[3];     // 512 arrivals: P(n) written over S(n)
    uint64_t o_done[2];     // PV(n) complete, n % 2 == i; the j-th PV of query tile i completes phase j (rescale guard: at
                            // tile (j, i) the softmax knows PV(j - 2, i) is complete, so it is at most one phase behind)
    uint64_t o_final[2];    // last PV into O_i complete (single phase: the epilogue may be two o_done phases behind)
    uint32_t tmem_base;
    float xmax[8][BM];      // [(n & 1) * 4 + column quarter][row]
    float xsum[8][BM];      // [query tile * 4 + column quarter][row]
};

template <bool EXACT>
__global__ void __launch_bounds__(NTHREADS, 1)
lt_attn_tc3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const LtArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;                                  // 2 tiles
    uint8_t* sK = sQ + 2 * TILE_BYTES;                   // STAGES3 tiles
    uint8_t* sV = sK + STAGES3 * TILE_BYTES;             // STAGES3 tiles
    Barriers3* B = reinterpret_cast<Barriers3*>(sV + STAGES3 * TILE_BYTES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * (2 * BM), h = blockIdx.y, z = blockIdx.z;
    pdl_trigger();

    if (tid == 0) {
        mbar_init(&B->q_full, 1);
        for (int s = 0; s < STAGES3; ++s) { mbar_init(&B->kv_full[s], 1); mbar_init(&B->kv_free[s], 1); }
        for (int b = 0; b < 3; ++b) { mbar_init(&B->s_full[b], 1); mbar_init(&B->p_full[b], 4 * BM / 32); }
        for (int i = 0; i < 2; ++i) { mbar_init(&B->o_done[i], 1); mbar_init(&B->o_final[i], 1); }
        fence_mbar_init();
    }
    if (warp == MMA_WARP) tmem_alloc<512>(&B->tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = B->tmem_base;
    pdl_wait();
    const int Tk = a.Tk_dev ? *a.Tk_dev : a.Tk;
    const int tiles_total = (Tk + BN - 1) / BN;
    const int per = (tiles_total + a.splits - 1) / a.splits;
    const int tb = z * per;
    int T = tiles_total - tb;
    T = T < 0 ? 0 : (T > per ? per : T);
    const int nT = 2 * T;                                 // score tiles of this CTA: n = 2 j + i

    if (warp == TMA_WARP) {
        if (elect_one() && T > 0) {
            tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
            mbar_arrive_expect_tx(&B->q_full, 2 * TILE_BYTES);
            tma_load_3d(sQ, &tmQ, &B->q_full, 0, q0, h);
            tma_load_3d(sQ + TILE_BYTES, &tmQ, &B->q_full, 0, q0 + BM, h);
            for (int j = 0; j < T; ++j) {
                const int s = j % STAGES3;
                if (j >= STAGES3) mbar_wait(&B->kv_free[s], ((j / STAGES3) - 1) & 1);
                mbar_arrive_expect_tx(&B->kv_full[s], 2 * TILE_BYTES);
                tma_load_3d(sK + s * TILE_BYTES, &tmK, &B->kv_full[s], 0, (tb + j) * BN, h);
                tma_load_3d(sV + s * TILE_BYTES, &tmV, &B->kv_full[s], 0, (tb + j) * BN, h);
            }
        }
    } else if (warp == MMA_WARP) {
        if (elect_one() && T > 0) {
            constexpr uint32_t IDESC_S = idesc_f16(128, 128, 0, 0);
            constexpr uint32_t IDESC_O = idesc_f16(128, 64, 0, 1);
            const uint64_t dQ0 = smem_desc_sw128(smem_u32(sQ)), dQ1 = smem_desc_sw128(smem_u32(sQ) + TILE_BYTES);
            const uint64_t dK = smem_desc_sw128(smem_u32(sK)), dV = smem_desc_sw128(smem_u32(sV));
            int kv_ready = -1;                              // highest key tile whose K/V have been waited for
            auto issue_S = [&](int n) {
                const int i = n & 1, j = n >> 1, s = j % STAGES3;
                if (j > kv_ready) {
                    mbar_wait_cp(&B->kv_full[s], (j / STAGES3) & 1, a.spin);
                    tc_fence_after();
                    kv_ready = j;
                }
                const uint64_t q = i ? dQ1 : dQ0;
                const uint64_t k = dK + (uint64_t)(s * (TILE_BYTES >> 4));
                const uint32_t d = tmem + (n % 3) * 128;
                mma_ss(d, q, k, IDESC_S, 0);
                mma_ss(d, q + 2, k + 2, IDESC_S, 1);
                if (EXACT) {
                    mma_ss(d, q + 4, k, IDESC_S, 1);          // Ql Kh
                    mma_ss(d, q + 6, k + 2, IDESC_S, 1);
                    mma_ss(d, q, k + 4, IDESC_S, 1);          // Qh Kl
                    mma_ss(d, q + 2, k + 6, IDESC_S, 1);
                }
                mma_commit(&B->s_full[n % 3]);
            };
            auto issue_PV = [&](int n) {
                const int i = n & 1, j = n >> 1, s = j % STAGES3;
                const uint64_t v = dV + (uint64_t)(s * (TILE_BYTES >> 4));
                const uint32_t d = tmem + 384 + i * 64;
                const uint32_t p = tmem + (n % 3) * 128;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)      // P_hi of keys [16 kk, +16) at columns 32 (kk / 2) + 8 (kk % 2)
                    mma_ts(d, p + 32 * (kk >> 1) + 8 * (kk & 1), v + 128 * kk, IDESC_O, (kk > 0 || j > 0) ? 1u : 0u);
                if (EXACT) {
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)  // P_lo 16 columns further up in the same thread's score columns
                        mma_ts(d, p + 32 * (kk >> 1) + 16 + 8 * (kk & 1), v + 128 * kk, IDESC_O, 1);
                }
                mma_commit(&B->o_done[i]);
                if (n + 2 >= nT) mma_commit(&B->o_final[i]);        // the last PV into O_i
                if (i == 1) mma_commit(&B->kv_free[s]);
            };
            mbar_wait(&B->q_full, 0);
            tc_fence_after();
            for (int n = 0; n < 3 && n < nT; ++n) issue_S(n);
            for (int n = 0; n < nT; ++n) {
                mbar_wait_cp(&B->p_full[n % 3], (n / 3) & 1, a.spin);
                tc_fence_after();
                issue_PV(n);
                if (n + 3 < nT) issue_S(n + 3);
            }
        }
    } else {
        const int qt = warp >> 2, wq = warp & 3;           // column quarter, TMEM lane quadrant
        const int row = wq * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
        float m_used[2] = {-INFINITY, -INFINITY}, l0[2] = {0.f, 0.f}, l1[2] = {0.f, 0.f};
        uint32_t srA[32], srB[32];
        // one score tile: `sr` holds this thread's 32 scores of tile n, `srn` receives those of tile n + 1
        // b / bn: score buffers of tile n and n + 1 (n % 3 kept as a rotating counter); next_par: phase parity of S(n + 1)
        auto tile = [&](uint32_t (&sr)[32], uint32_t (&srn)[32], const int i, const int j, const int b, const int bn,
                        const bool has_next, const uint32_t next_par) {
            const uint32_t tS = tmem + lane_addr + b * 128 + qt * 32;
            const uint32_t tO = tmem + lane_addr + 384 + i * 64 + qt * 16;
            const int key0 = (tb + j) * BN + qt * 32;
            if (a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && i == 0 && j == 0) {
#pragma unroll
                for (int k = 0; k < 32; ++k) a.dbg[row * 128 + qt * 32 + k] = __uint_as_float(sr[k]);
            }
            if (key0 + 32 > Tk) {
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
            B->xmax[i * 4 + qt][row] = fmaxf(mx0, mx1);
            asm volatile("bar.sync 1, 512;" ::: "memory");
            const float mt = fmaxf(fmaxf(B->xmax[i * 4 + 0][row], B->xmax[i * 4 + 1][row]),
                                   fmaxf(B->xmax[i * 4 + 2][row], B->xmax[i * 4 + 3][row]));
            const float m_new = fmaxf(m_used[i], mt);
            const bool grow = (m_new > m_used[i]) && (j > 0);
            if (__any_sync(0xffffffffu, grow)) {
                // O_i must be quiescent: PV(n - 2) complete (o_done), PV(n) not issued before all threads arrive on p_full
                mbar_wait(&B->o_done[i], (uint32_t)((j - 1) & 1));
                tc_fence_after();
                const float f = grow ? ex2((m_used[i] - m_new) * LOG2E) : 1.f;
                uint32_t orr[16];
                tmem_ld16(tO, orr);
                tmem_wait_ld();
#pragma unroll
                for (int k = 0; k < 16; ++k) orr[k] = __float_as_uint(__uint_as_float(orr[k]) * f);
                tmem_st16(tO, orr);
                l0[i] *= f; l1[i] *= f;
            }
            m_used[i] = m_new;
            const float neg = m_new * LOG2E;
            const uint32_t tSn = tmem + lane_addr + bn * 128 + qt * 32;
            if (has_next) {                 // S(n + 1) was issued behind PV(n - 2): complete long ago in steady state
                mbar_wait_cp(&B->s_full[bn], next_par, a.spin);
                tc_fence_after();
                tmem_ld16(tSn, srn);        // first half now; second half once sr[0..15] are dead (register budget)
            }
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                if (hf == 1 && has_next) tmem_ld16(tSn + 16, srn + 16);
                uint32_t ph[8], pl[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const float p0 = ex2(fmaf(__uint_as_float(sr[16 * hf + 2 * t]), LOG2E, -neg));
                    const float p1 = ex2(fmaf(__uint_as_float(sr[16 * hf + 2 * t + 1]), LOG2E, -neg));
                    s0 += p0; s1 += p1;
                    const __half2 hi = __floats2half2_rn(p0, p1);
                    ph[t] = *reinterpret_cast<const uint32_t*>(&hi);
                    if (EXACT) {
                        const __half2 lo = __floats2half2_rn(p0 - __low2float(hi), p1 - __high2float(hi));
                        pl[t] = *reinterpret_cast<const uint32_t*>(&lo);
                    }
                }
                tmem_st8(tS + 8 * hf, ph);               // keys [32 qt + 16 hf, +16) -> columns [32 qt + 8 hf, +8)
                if (EXACT) tmem_st8(tS + 16 + 8 * hf, pl);
            }
            l0[i] += s0; l1[i] += s1;
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive_warp(&B->p_full[b]);
            if (has_next) tmem_wait_ld32(srn);
        };
        if (T > 0) {
            mbar_wait(&B->s_full[0], 0);
            tc_fence_after();
            tmem_ld32(tmem + lane_addr + qt * 32, srA);
            tmem_wait_ld32(srA);
        }
        // buffer b of tile n and the parity of its s_full phase rotate with n: bit b of `par` = uses of buffer b so far (mod 2)
        int b = 0;
        uint32_t par = 1u;                                  // the prologue consumed phase 0 of s_full[0]
        for (int j = 0; j < T; ++j) {
            int bn = b == 2 ? 0 : b + 1;
            tile(srA, srB, 0, j, b, bn, true, (par >> bn) & 1u);
            par ^= 1u << bn;
            b = bn;
            bn = b == 2 ? 0 : b + 1;
            tile(srB, srA, 1, j, b, bn, j + 1 < T, (par >> bn) & 1u);
            par ^= 1u << bn;                                // (harmless after the last tile)
            b = bn;
        }

        // ---- epilogue: per query tile, this thread finishes output channels [8*qt, 8*qt+8) of its row
#pragma unroll
        for (int i = 0; i < 2; ++i) B->xsum[i * 4 + qt][row] = l0[i] + l1[i];
        asm volatile("bar.sync 1, 512;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float l = (B->xsum[i * 4 + 0][row] + B->xsum[i * 4 + 1][row]) + (B->xsum[i * 4 + 2][row] + B->xsum[i * 4 + 3][row]);
            const uint32_t tO = tmem + lane_addr + 384 + i * 64;
            const bool dump = a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && i == 0;
            const int q = q0 + i * BM + row;
            float o[8];
            if (T > 0) {
                mbar_wait(&B->o_final[i], 0);                            // the last PV into O_i
                tc_fence_after();
                uint32_t o0[8], o1[8];
                tmem_ld8(tO + qt * 8, o0);
                tmem_ld8(tO + 32 + qt * 8, o1);
                tmem_wait_ld();
                if (dump) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        a.dbg[128 * 128 + row * 64 + qt * 8 + k] = __uint_as_float(o0[k]);
                        a.dbg[128 * 128 + row * 64 + 32 + qt * 8 + k] = __uint_as_float(o1[k]);
                    }
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = __uint_as_float(o0[k]) + __uint_as_float(o1[k]);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = 0.f;
            }
            if (q < a.N) {
                if (a.splits == 1) {
                    const float inv = 1.f / l;
                    float* dst = a.O + (size_t)q * a.ldo + h * 32 + qt * 8;
                    *reinterpret_cast<float4*>(dst) = make_float4(o[0] * inv, o[1] * inv, o[2] * inv, o[3] * inv);
                    *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4] * inv, o[5] * inv, o[6] * inv, o[7] * inv);
                } else {
                    float* dst = a.Opart + ((size_t)z * a.N + q) * (a.H * 32) + h * 32 + qt * 8;
                    *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
                    if (qt == 0) {
                        a.Mpart[((size_t)z * a.H + h) * a.N + q] = m_used[i];
                        a.Lpart[((size_t)z * a.H + h) * a.N + q] = l;
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------ operand packing
// src fp32 [rows][ld] (head h at columns h*32) -> dst halfs [H][cap][64] at row offset: [hi(32) | lo(32)]
__global__ void pack_rows64_kernel(const float* __restrict__ src, int ld, __half* __restrict__ dst, int cap, int rows,
                                   int H, int row_off, const int* __restrict__ row_off_dev, float div) {
    pdl_sync();
    const int off = row_off_dev ? *row_off_dev : row_off;
    const size_t total = (size_t)rows * H * 8;  // 4 channels per thread
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c4 = i % 8;
        const int hh = (i / 8) % H;
        const int r = i / (8 * (size_t)H);
        float4 v = *reinterpret_cast<const float4*>(src + (size_t)r * ld + hh * 32 + c4 * 4);
        if (div != 1.f) { v.x = v.x / div; v.y = v.y / div; v.z = v.z / div; v.w = v.w / div; }
        const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
        const __half2 l0 = __floats2half2_rn(v.x - __low2float(h0), v.y - __high2float(h0));
        const __half2 l1 = __floats2half2_rn(v.z - __low2float(h1), v.w - __high2float(h1));
        __half* d = dst + ((size_t)hh * cap + off + r) * 64 + c4 * 4;
        *reinterpret_cast<__half2*>(d) = h0;
        *reinterpret_cast<__half2*>(d + 2) = h1;
        *reinterpret_cast<__half2*>(d + 32) = l0;
        *reinterpret_cast<__half2*>(d + 34) = l1;
    }
}

}  // namespace tc
}  // namespace aotb

namespace aotb { namespace tc {
int launch_lt_attn_pair(const void* Qp, int Nq_cap, const void* Kp, const void* Vp, int kv_cap, int N, int Tk,
                        const int* Tk_dev, int H, float* O, int ldo, float* Opart, float* Mpart, float* Lpart, int splits,
                        int exact, int spin, cudaStream_t st);      // lt_attn_tc2.cu
} }

using namespace aotb;

// Pack fp32 rows into the split-fp16 operand layout of the tensor-core attention kernel.
// src [rows][ld] fp32, dst [H][cap][64] fp16; written at rows [row_off, row_off+rows) (device counter optional);
// values are divided by `div` first (T for Q -- attention.py:82 -- 1 for K/V).
extern "C" int aotb_tc_pack_rows_f16x2(const float* src, int ld, void* dst, int cap, int rows, int H, int row_off,
                                       const int* row_off_dev, float div, void* stream) {
    AOTB_REQUIRE(src && dst && rows > 0 && H > 0 && ld % 4 == 0 && cap > 0, "aotb_tc_pack_rows_f16x2: bad args");
    AOTB_REQUIRE(row_off_dev || row_off + rows <= cap, "aotb_tc_pack_rows_f16x2: rows exceed capacity");
    const size_t total = (size_t)rows * H * 8;
    int g = (int)((total + 255) / 256);
    if (g > 148 * 8) g = 148 * 8;
    launch(tc::pack_rows64_kernel, dim3(g), dim3(256), 0, (cudaStream_t)stream, src, ld, (__half*)dst, cap, rows, H, row_off,
                                                                row_off_dev, div);
    return check_launch("aotb_tc_pack_rows_f16x2");
}

extern "C" size_t aotb_lt_attn_tc_smem_bytes(void) {
    return (size_t)(2 + 2 * tc::STAGES) * tc::TILE_BYTES + sizeof(tc::Barriers) + 1024;
}

static size_t lt3_smem_bytes() { return (size_t)(2 + 2 * tc::STAGES3) * tc::TILE_BYTES + sizeof(tc::Barriers3) + 1024; }

// Qp [H][Nq_cap][64], Kp/Vp [H][kv_cap][64] packed fp16x2 operands (zero-filled beyond the live rows);
// O [N][ldo] fp32 (head h at columns h*32).  splits > 1 writes un-normalised partials
// (Opart [splits][N][H*32], Mpart/Lpart [splits][H][N]) for aotb_attn_merge_f32.
extern "C" int aotb_lt_attn_tc_f16x2(const void* Qp, int Nq_cap, const void* Kp, const void* Vp, int kv_cap, int N,
                                     int Tk, const int* Tk_dev, int H, float* O, int ldo, float* Opart, float* Mpart,
                                     float* Lpart, int splits, int exact, float* dbg, void* stream) {
    AOTB_REQUIRE(Qp && Kp && Vp && N > 0 && H > 0 && (Tk > 0 || Tk_dev) && splits >= 1,
                 "aotb_lt_attn_tc_f16x2: bad args");
    AOTB_REQUIRE(Nq_cap >= ((N + 255) / 256) * 256, "aotb_lt_attn_tc_f16x2: Q buffer must be padded to 256 rows");
    AOTB_REQUIRE(splits == 1 ? (O != nullptr && ldo % 4 == 0) : (Opart && Mpart && Lpart),
                 "aotb_lt_attn_tc_f16x2: output buffers");
    AOTB_REQUIRE(((uintptr_t)Qp | (uintptr_t)Kp | (uintptr_t)Vp) % 128 == 0, "aotb_lt_attn_tc_f16x2: alignment");
    if (exact & 16)           // bit 4: "pair" layout (two co-resident CTAs per SM, 64-key tiles; lt_attn_tc2.cu)
        return tc::launch_lt_attn_pair(Qp, Nq_cap, Kp, Vp, kv_cap, N, Tk, Tk_dev, H, O, ldo, Opart, Mpart, Lpart, splits,
                                       exact & 1, (exact >> 2) & 1, (cudaStream_t)stream);
    CUtensorMap tq, tk, tv;
    int rc;
    if ((rc = tc::make_tmap_rows64(&tq, Qp, Nq_cap, H)) != AOTB_OK) return rc;
    if ((rc = tc::make_tmap_rows64(&tk, Kp, kv_cap, H)) != AOTB_OK) return rc;
    if ((rc = tc::make_tmap_rows64(&tv, Vp, kv_cap, H)) != AOTB_OK) return rc;
    const size_t smem = aotb_lt_attn_tc_smem_bytes();
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tc::lt_attn_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(tc::lt_attn_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(tc::lt_attn_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(tc::lt_attn_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(tc::lt_attn_tc3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lt3_smem_bytes());
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(tc::lt_attn_tc3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lt3_smem_bytes());
        if (e != cudaSuccess) {
            set_error("aotb_lt_attn_tc_f16x2: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return AOTB_ERR_CUDA;
        }
        configured = true;
    }
    tc::LtArgs a;
    a.N = N; a.Tk = Tk; a.Tk_dev = Tk_dev; a.H = H; a.O = O; a.ldo = ldo;
    a.Opart = Opart; a.Mpart = Mpart; a.Lpart = Lpart; a.splits = splits; a.exact = exact & 1; a.spin = (exact >> 2) & 1; a.dbg = dbg;
    dim3 grid(cdiv(N, 2 * tc::BM), H, splits);
    const dim3 block(tc::NTHREADS);
    cudaStream_t st = (cudaStream_t)stream;
    if (exact & 8) {          // bit 3: "ahead" layout (three score buffers, TMEM read under the ex2 pass)
        if (exact & 1) launch(tc::lt_attn_tc3_kernel<true>, dim3(grid), block, lt3_smem_bytes(), st, tq, tk, tv, a);
        else launch(tc::lt_attn_tc3_kernel<false>, dim3(grid), block, lt3_smem_bytes(), st, tq, tk, tv, a);
        return check_launch("aotb_lt_attn_tc_f16x2");
    }
    switch (exact & 3) {      // bit 0: fp16x2 "exact" operands; bit 1: two-group softmax layout
        case 3: launch(tc::lt_attn_tc_kernel<true, true>, dim3(grid), block, smem, st, tq, tk, tv, a); break;
        case 2: launch(tc::lt_attn_tc_kernel<false, true>, dim3(grid), block, smem, st, tq, tk, tv, a); break;
        case 1: launch(tc::lt_attn_tc_kernel<true, false>, dim3(grid), block, smem, st, tq, tk, tv, a); break;
        default: launch(tc::lt_attn_tc_kernel<false, false>, dim3(grid), block, smem, st, tq, tk, tv, a);
    }
    return check_launch("aotb_lt_attn_tc_f16x2");
}
