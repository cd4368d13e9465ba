"""Functional (address-level) model of the tcgen05 attention kernels -- no GPU needed.

What it models, at the granularity the kernels' address arithmetic works at:
  * shared memory as bytes; a TMA box load of [rows x 64 halfs] lands as rows of 128 B (the 128B swizzle permutes 16-byte
    chunks inside 1 KB groups and is undone by the MMA unit, so it is transparent at this level);
  * UMMA shared-memory descriptors as (byte address >> 4): K-major operand = `rows` rows starting at the tile row the address
    points to, 16 halfs (32 B) of K starting at the address's offset inside the 128 B row; MN-major B operand = 16 rows of
    K starting at the addressed row, N halfs taken from the start of each row;
  * tensor memory as [128 lanes][512 columns] of 32-bit words; an MMA accumulates fp32 into N columns from the D address;
    a TMEM A operand is 8 columns of packed half2 (K = 16) per lane;
  * tcgen05.ld / st .32x32b.xN as N consecutive columns of the issuing thread's lane.
These are exactly the conventions lt_attn_tc_kernel (validated on B200) relies on; `model_lt_tile` re-executes THAT kernel's
address arithmetic as a calibration of the model, `model_lt_ahead` and `model_gp` then execute the two kernels that have not
run on a GPU yet (lt_attn_tc3_kernel, gp_attn_tc_kernel).  Each transcription follows the .cu source statement by statement
(same constants, same offsets) in dependency order; the result must equal softmax(Q K^T) V computed directly.

Run: python scripts/tc_layout_model.py
"""
import numpy as np


def split16(x):
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float32)).astype(np.float16)
    return hi, lo


def pack_rows(x, heads, cap, div=1.0):
    """aotb_tc_pack_rows_f16x2: fp32 [rows][heads*32] -> halfs [heads][cap][64] = [hi(32) | lo(32)]."""
    rows = x.shape[0]
    out = np.zeros((heads, cap, 64), np.float16)
    xs = (x / np.float32(div)).astype(np.float32).reshape(rows, heads, 32).transpose(1, 0, 2)
    hi, lo = split16(xs)
    out[:, :rows, :32] = hi
    out[:, :rows, 32:] = lo
    return out


class Machine:
    def __init__(self, smem_bytes):
        self.smem = np.zeros(smem_bytes // 2, np.float16)           # halfs
        self.tmem = np.zeros((128, 512), np.uint32)

    # ---- TMA: box [rows x 64 halfs] of packed[head] starting at row r0 -> smem byte offset dst
    def tma_load(self, dst, packed, head, r0, rows):
        src = np.zeros((rows, 64), np.float16)
        avail = max(0, min(rows, packed.shape[1] - r0))
        src[:avail] = packed[head, r0:r0 + avail]                    # out-of-bounds rows are zero-filled
        self.smem[dst // 2: dst // 2 + rows * 64] = src.reshape(-1)

    # ---- operands from descriptors (desc = byte address >> 4, relative to the start of dynamic smem)
    def kmajor(self, desc, rows):
        addr = desc << 4
        row0, colb = addr // 128, addr % 128
        m = self.smem[row0 * 64: (row0 + rows) * 64].reshape(rows, 64)
        return m[:, colb // 2: colb // 2 + 16].astype(np.float32)    # [rows][16]

    def mnmajor(self, desc, n):
        addr = desc << 4
        assert addr % 128 == 0
        row0 = addr // 128
        m = self.smem[row0 * 64: (row0 + 16) * 64].reshape(16, 64)
        return m[:, :n].astype(np.float32)                           # [16 (K)][n]

    def tmem_a(self, col):                                           # A operand from TMEM: 8 columns of half2
        w = self.tmem[:, col: col + 8].copy()
        return w.view(np.float16).reshape(128, 16).astype(np.float32)

    def acc(self, d_col, n, prod, accumulate):
        cur = self.tmem[:, d_col: d_col + n].view(np.float32)
        new = (cur + prod) if accumulate else prod
        self.tmem[:, d_col: d_col + n] = new.astype(np.float32).view(np.uint32)

    def mma_ss(self, d_col, a_desc, b_desc, n, accumulate):          # D[128 x n] (+)= A[128 x 16] B[n x 16]^T
        self.acc(d_col, n, self.kmajor(a_desc, 128) @ self.kmajor(b_desc, n).T, accumulate)

    def mma_ts(self, d_col, a_col, b_desc, n, accumulate):           # D[128 x n] (+)= A_tmem[128 x 16] B[16 x n]
        self.acc(d_col, n, self.tmem_a(a_col) @ self.mnmajor(b_desc, n), accumulate)

    def ld(self, lane, col, n):
        return self.tmem[lane, col: col + n].view(np.float32).copy()

    def st_words(self, lane, col, words):
        self.tmem[lane, col: col + len(words)] = words

    def st_f32(self, lane, col, vals):
        self.tmem[lane, col: col + len(vals)] = np.asarray(vals, np.float32).view(np.uint32)


def half2_words(p):                                                  # pairs (p[2t], p[2t+1]) -> packed hi words, lo words
    hi = p.astype(np.float16)
    lo = (p - hi.astype(np.float32)).astype(np.float16)
    return hi.view(np.uint32), lo.view(np.uint32)


def reference(Q, K, V, T):
    s = (Q / np.float32(T)).astype(np.float64) @ K.astype(np.float64).T
    p = np.exp(s - s.max(1, keepdims=True))
    return (p / p.sum(1, keepdims=True)) @ V.astype(np.float64)


# --------------------------------------------------------------------------------------------------------------------
# calibration: lt_attn_tc_kernel<exact, tile layout> (validated on B200), one head, one CTA (256 queries), one split
# --------------------------------------------------------------------------------------------------------------------
def model_lt_tile(Q, K, V, ahead=False):
    """Q [<=256][32], K/V [Tk][32].  `ahead` switches to lt_attn_tc3_kernel's TMEM plan / issue order / P aliasing."""
    N, Tk = Q.shape[0], K.shape[0]
    TILE, STAGES = 128 * 128, (4 if ahead else 3)
    Qp, Kp, Vp = pack_rows(Q, 1, 256, np.sqrt(32.0)), pack_rows(K, 1, Tk + 256), pack_rows(V, 1, Tk + 256)
    sQ, sK = 0, 2 * TILE
    sV = sK + STAGES * TILE
    M = Machine(sV + STAGES * TILE)
    M.tma_load(sQ, Qp, 0, 0, 128)
    M.tma_load(sQ + TILE, Qp, 0, 128, 128)
    T = (Tk + 127) // 128
    dQ = [sQ >> 4, (sQ + TILE) >> 4]
    dK, dV = sK >> 4, sV >> 4
    O_base = 384 if ahead else 256
    m_used = np.full((2, 128), -np.inf, np.float32)
    l = np.zeros((2, 128), np.float32)

    def issue_S(j, i):
        s = j % STAGES
        q, k = dQ[i], dK + s * (TILE >> 4)
        d = (((2 * j + i) % 3) if ahead else i) * 128
        M.mma_ss(d, q, k, 128, False)
        M.mma_ss(d, q + 2, k + 2, 128, True)
        M.mma_ss(d, q + 4, k, 128, True)
        M.mma_ss(d, q + 6, k + 2, 128, True)
        M.mma_ss(d, q, k + 4, 128, True)
        M.mma_ss(d, q + 2, k + 6, 128, True)

    def issue_PV(j, i):
        s = j % STAGES
        v = dV + s * (TILE >> 4)
        d = O_base + i * 64
        p = (((2 * j + i) % 3) if ahead else i) * 128
        for kk in range(8):
            M.mma_ts(d, p + 32 * (kk >> 1) + 8 * (kk & 1), v + 128 * kk, 64, kk > 0 or j > 0)
        for kk in range(8):
            a = (p + 32 * (kk >> 1) + 16 + 8 * (kk & 1)) if ahead else (384 + i * 64 + 8 * kk)
            M.mma_ts(d, a, v + 128 * kk, 64, True)

    def softmax(j, i):
        buf = ((2 * j + i) % 3) if ahead else i
        S = np.stack([M.ld(r, buf * 128, 128) for r in range(128)])          # every thread's 32 columns, all quarters
        key = j * 128 + np.arange(128)
        S[:, key >= Tk] = -np.inf
        m_new = np.maximum(m_used[i], S.max(1))
        f = np.where(np.isfinite(m_used[i]), np.exp2((m_used[i] - m_new) * np.float32(1.4426950408889634)), 1.0).astype(np.float32)
        if j > 0:
            for r in range(128):
                M.st_f32(r, O_base + i * 64, M.ld(r, O_base + i * 64, 64) * f[r])
            l[i] *= f
        m_used[i] = m_new
        P = np.exp2((S - m_new[:, None]) * np.float32(1.4426950408889634)).astype(np.float32)
        l[i] += P.sum(1)
        for r in range(128):
            for qt in range(4):
                hi, lo = half2_words(P[r, 32 * qt: 32 * qt + 32])
                tS = buf * 128 + qt * 32
                if ahead:
                    for hf in range(2):
                        M.st_words(r, tS + 8 * hf, hi[8 * hf: 8 * hf + 8])
                        M.st_words(r, tS + 16 + 8 * hf, lo[8 * hf: 8 * hf + 8])
                else:
                    M.st_words(r, tS, hi)
                    M.st_words(r, 384 + i * 64 + qt * 16, lo)

    for j in range(T):
        s = j % STAGES
        M.tma_load(sK + s * TILE, Kp, 0, j * 128, 128)
        M.tma_load(sV + s * TILE, Vp, 0, j * 128, 128)
        for i in range(2):
            issue_S(j, i)
            softmax(j, i)
            issue_PV(j, i)
    out = np.zeros((256, 32), np.float32)
    for i in range(2):
        for r in range(128):
            o = M.ld(r, O_base + i * 64, 64)
            out[i * 128 + r] = (o[:32] + o[32:]) / l[i][r]
    return out[:N]


# --------------------------------------------------------------------------------------------------------------------
# gp_attn_tc_kernel: one query tile (128), one d_v slice `vs` of 128 channels, one split
# --------------------------------------------------------------------------------------------------------------------
def model_gp(Q, K, V, vs):
    """Q [<=128][128], K [Tk][128], V [Tk][dv]; returns the [N][128] slice vs of softmax(Q K^T / sqrt(128)) V."""
    N, Tk, dv = Q.shape[0], K.shape[0], V.shape[1]
    QTILE, KVTILE = 128 * 128, 64 * 128
    STAGE = 4 * KVTILE
    Qp, Kp, Vp = pack_rows(Q, 4, 128, np.sqrt(128.0)), pack_rows(K, 4, Tk + 64), pack_rows(V, dv // 32, Tk + 64)
    sQ = 0
    sK = sQ + 4 * QTILE
    sV = sK + 2 * STAGE
    M = Machine(sV + 2 * STAGE)
    for c in range(4):
        M.tma_load(sQ + c * QTILE, Qp, c, 0, 128)
    T = (Tk + 63) // 64
    dQ, dK, dV = sQ >> 4, sK >> 4, sV >> 4
    m_used = np.full(128, -np.inf, np.float32)
    l = np.zeros(128, np.float32)
    for n in range(T):
        s = n & 1
        for c in range(4):
            M.tma_load(sK + s * STAGE + c * KVTILE, Kp, c, n * 64, 64)
            M.tma_load(sV + s * STAGE + c * KVTILE, Vp, vs * 4 + c, n * 64, 64)
        # ---- issue_S(n)
        d = (n % 3) * 64
        for c in range(4):
            q = dQ + c * (QTILE >> 4)
            k = dK + s * (STAGE >> 4) + c * (KVTILE >> 4)
            M.mma_ss(d, q, k, 64, c > 0)
            M.mma_ss(d, q + 2, k + 2, 64, True)
            M.mma_ss(d, q + 4, k, 64, True)
            M.mma_ss(d, q + 6, k + 2, 64, True)
            M.mma_ss(d, q, k + 4, 64, True)
            M.mma_ss(d, q + 2, k + 6, 64, True)
        # ---- softmax tile n (16 warps: quadrant wq = row / 32, quarter qt owns 16 columns)
        b = n % 3
        S = np.stack([np.concatenate([M.ld(r, b * 64 + qt * 16, 16) for qt in range(4)]) for r in range(128)])
        key = n * 64 + np.arange(64)
        S[:, key >= Tk] = -np.inf
        m_new = np.maximum(m_used, S.max(1))
        if n > 0:
            f = np.exp2((m_used - m_new) * np.float32(1.4426950408889634)).astype(np.float32)
            for r in range(128):
                for qt in range(4):
                    tO = 192 + qt * 64
                    for cc in range(4):
                        M.st_f32(r, tO + 16 * cc, M.ld(r, tO + 16 * cc, 16) * f[r])
            l *= f
        m_used = m_new
        P = np.exp2((S - m_new[:, None]) * np.float32(1.4426950408889634)).astype(np.float32)
        l += P.sum(1)
        for r in range(128):
            for qt in range(4):
                hi, lo = half2_words(P[r, 16 * qt: 16 * qt + 16])
                tS = b * 64 + qt * 16
                M.st_words(r, tS, hi)
                M.st_words(r, tS + 8, lo)
        # ---- issue_PV(n)
        p = (n % 3) * 64
        for c in range(4):
            d = 192 + c * 64
            v = dV + s * (STAGE >> 4) + c * (KVTILE >> 4)
            for kk in range(4):
                M.mma_ts(d, p + 16 * kk, v + 128 * kk, 64, kk > 0 or n > 0)
            for kk in range(4):
                M.mma_ts(d, p + 16 * kk + 8, v + 128 * kk, 64, True)
    out = np.zeros((128, 128), np.float32)
    for r in range(128):
        for qt in range(4):
            tO = 192 + qt * 64
            for hb in range(2):
                o0, o1 = M.ld(r, tO + 16 * hb, 16), M.ld(r, tO + 32 + 16 * hb, 16)
                out[r, qt * 32 + hb * 16: qt * 32 + hb * 16 + 16] = (o0 + o1) / l[r]
    return out[:N]


def main():
    rng = np.random.default_rng(0)
    worst = 0.0
    for (N, Tk) in ((256, 128), (200, 300), (256, 385)):
        Q = (rng.standard_normal((N, 32)) * 3).astype(np.float32)
        K = rng.standard_normal((Tk, 32)).astype(np.float32)
        V = rng.standard_normal((Tk, 32)).astype(np.float32)
        ref = reference(Q, K, V, np.sqrt(32.0))
        for ahead in (False, True):
            err = np.abs(model_lt_tile(Q, K, V, ahead) - ref).max()
            worst = max(worst, err)
            print(f"lt_attn {'ahead (tc3)' if ahead else 'tile (validated kernel: calibration)'}: N={N} Tk={Tk} max err {err:.2e}")
            assert err < 1e-4
    for (N, Tk, dv) in ((128, 64, 256), (100, 200, 256), (128, 333, 384)):
        Q = (rng.standard_normal((N, 128)) * 2).astype(np.float32)
        K = rng.standard_normal((Tk, 128)).astype(np.float32)
        V = rng.standard_normal((Tk, dv)).astype(np.float32)
        ref = reference(Q, K, V, np.sqrt(128.0))
        for vs in range(dv // 128):
            err = np.abs(model_gp(Q, K, V, vs) - ref[:, vs * 128:(vs + 1) * 128]).max()
            worst = max(worst, err)
            print(f"gp_attn_tc: N={N} Tk={Tk} dv={dv} slice {vs} max err {err:.2e}")
            assert err < 1e-4
    print(f"tc layout model: all address arithmetic consistent (worst error {worst:.2e})")


if __name__ == "__main__":
    main()
