// DeAOT long-term attention on the tensor cores as two GEMMs around a row softmax (AOTB_DEAOT_LT=gemm).
//
// GatedPropagation.forward, networks/layers/attention.py:672-704, with the DeAOT head shape (1 head, d_qk = 128,
// d_v = 1024: transformer.py:541-548): attn = softmax((Q / T) K^T) over the memory bank, out = attn @ V.  With d_v = 1024
// the contraction is PV-dominated (2 * N * Tk * 1152 FLOPs, 2304 per exponential), so unlike the 8 x 32 AOT heads it is
// tensor-bound rather than MUFU-bound; the fp32 CUDA-core flash kernel (attention_simt.cu) runs it at a few per cent of
// what the tensor pipe offers.  This first tensor-core formulation reuses the fp32-faithful split-fp16 GEMM of conv_tc.cu:
//
//   S = Q K^T            aotb_conv2d_nhwc_tc with the bank's keys as the "weights"  [Tk_cap][128]  (hi / lo fp16)
//   P = softmax(S / T)   row_softmax_kernel, in place; columns >= the live key count (device counter) become 0
//   O = P V              aotb_conv2d_nhwc_tc with V^T as the weights                [1024][Tk_cap] (hi / lo fp16)
//
// i.e. the reference's own three steps (attention.py:686-704) with the score matrix materialised ([N][Tk_cap] fp32,
// 269 MB at a 24-frame capacity: ~0.1 ms of HBM traffic per pass) and all shapes fixed by the bank CAPACITY, so one
// captured CUDA graph serves the growing bank (the live count only enters through the device counter in the softmax).
// The kernels here maintain the split-fp16 operand copies of the bank at append time and do the softmax:
//   split_rows_kernel   fp32 rows -> hi / lo fp16 rows at a (device-resident) row offset            (keys)
//   split_cols_kernel   fp32 rows -> hi / lo fp16 COLUMNS of the transposed bank at a column offset  (values)
//   row_softmax_kernel  three passes over a row that stays in L2 (max, sum of exp, normalised write)
// A fused QK^T-softmax-PV kernel for this head shape (O needs 1024 fp32 TMEM columns, i.e. a d_v split over CTAs) is the
// round-2 item; this path is its parity baseline.
#include "common.cuh"
#include <cuda_fp16.h>

namespace aotb {

// src [rows][lds] fp32 (C columns, C % 4 == 0) -> hi / lo [.][ldw] fp16 rows [off, off + rows)
__global__ void split_rows_kernel(const float* __restrict__ src, int lds, __half* __restrict__ hi,
                                  __half* __restrict__ lo, int ldw, int rows, int C, int off,
                                  const int* __restrict__ off_dev) {
    pdl_sync();
    const int o = off_dev ? *off_dev : off;
    const int C4 = C >> 2;
    const size_t total = (size_t)rows * C4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / C4), c = (int)(i - (size_t)r * C4) * 4;
        const float4 v = *reinterpret_cast<const float4*>(src + (size_t)r * lds + c);
        const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
        const __half2 l0 = __floats2half2_rn(v.x - __low2float(h0), v.y - __high2float(h0));
        const __half2 l1 = __floats2half2_rn(v.z - __low2float(h1), v.w - __high2float(h1));
        __half* dh = hi + (size_t)(o + r) * ldw + c;
        __half* dl = lo + (size_t)(o + r) * ldw + c;
        *reinterpret_cast<__half2*>(dh) = h0;
        *reinterpret_cast<__half2*>(dh + 2) = h1;
        *reinterpret_cast<__half2*>(dl) = l0;
        *reinterpret_cast<__half2*>(dl + 2) = l1;
    }
}

// src [rows][lds] fp32 (C columns) -> hiT / loT [C][ldt] fp16, columns [off, off + rows): 32 x 32 tiles through shared
// memory so that both the reads (along C) and the writes (along rows) are contiguous
__global__ void split_cols_kernel(const float* __restrict__ src, int lds, __half* __restrict__ hiT,
                                  __half* __restrict__ loT, int ldt, int rows, int C, int off,
                                  const int* __restrict__ off_dev) {
    __shared__ float tile[32][33];
    pdl_sync();
    const int o = off_dev ? *off_dev : off;
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < C) ? src[(size_t)r * lds + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < C && r < rows) {
            const float v = tile[threadIdx.x][i];
            const __half h = __float2half_rn(v);
            hiT[(size_t)c * ldt + o + r] = h;
            loT[(size_t)c * ldt + o + r] = __float2half_rn(v - __half2float(h));
        }
    }
}

__device__ __forceinline__ float block_max(float v, float* sh) {
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = sh[0];
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = fmaxf(r, sh[i]);
    __syncthreads();
    return r;
}
__device__ __forceinline__ float block_sum(float v, float* sh) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = sh[0];
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r += sh[i];      // fixed order: deterministic
    __syncthreads();
    return r;
}

// S [N][ld]: row r becomes softmax(scale * S[r][0:valid]) in columns [0, valid) and 0 in [valid, cols); one CTA per row
__global__ void __launch_bounds__(256) row_softmax_kernel(float* __restrict__ S, int ld, int cols, int Tk,
                                                          const int* __restrict__ Tk_dev, float scale) {
    __shared__ float sh[8];
    pdl_sync();
    int valid = Tk_dev ? *Tk_dev : Tk;
    valid = valid < cols ? valid : cols;
    float* row = S + (size_t)blockIdx.x * ld;
    const int v4 = valid & ~3;
    float m = -INFINITY;
    for (int c = threadIdx.x * 4; c < v4; c += 1024) {
        const float4 x = *reinterpret_cast<const float4*>(row + c);
        m = fmaxf(m, fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w)));
    }
    if (threadIdx.x < valid - v4) m = fmaxf(m, row[v4 + threadIdx.x]);
    m = block_max(m, sh);
    float s = 0.f;
    for (int c = threadIdx.x * 4; c < v4; c += 1024) {
        const float4 x = *reinterpret_cast<const float4*>(row + c);
        s += (expf((x.x - m) * scale) + expf((x.y - m) * scale)) + (expf((x.z - m) * scale) + expf((x.w - m) * scale));
    }
    if (threadIdx.x < valid - v4) s += expf((row[v4 + threadIdx.x] - m) * scale);
    s = block_sum(s, sh);
    const float inv = 1.f / s;
    for (int c = threadIdx.x * 4; c < cols; c += 1024) {       // cols % 4 == 0
        float4 x = *reinterpret_cast<const float4*>(row + c);
        x.x = c + 0 < valid ? expf((x.x - m) * scale) * inv : 0.f;
        x.y = c + 1 < valid ? expf((x.y - m) * scale) * inv : 0.f;
        x.z = c + 2 < valid ? expf((x.z - m) * scale) * inv : 0.f;
        x.w = c + 3 < valid ? expf((x.w - m) * scale) * inv : 0.f;
        *reinterpret_cast<float4*>(row + c) = x;
    }
}

}  // namespace aotb

using namespace aotb;

extern "C" int aotb_split_rows_f16x2(const float* src, int lds, void* hi, void* lo, int ldw, int rows, int C, int row_off,
                                     const int* row_off_dev, void* stream) {
    AOTB_REQUIRE(src && hi && lo && rows > 0 && C > 0 && C % 4 == 0 && lds % 4 == 0 && ldw % 4 == 0 && ldw >= C,
                 "aotb_split_rows_f16x2: bad args");
    const size_t total = (size_t)rows * (C / 4);
    size_t g = (total + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    launch(split_rows_kernel, dim3((unsigned)g), dim3(256), 0, (cudaStream_t)stream, src, lds, (__half*)hi, (__half*)lo, ldw,
           rows, C, row_off, row_off_dev);
    return check_launch("aotb_split_rows_f16x2");
}

extern "C" int aotb_split_cols_f16x2(const float* src, int lds, void* hiT, void* loT, int ldt, int rows, int C, int col_off,
                                     const int* col_off_dev, void* stream) {
    AOTB_REQUIRE(src && hiT && loT && rows > 0 && C > 0 && ldt > 0, "aotb_split_cols_f16x2: bad args");
    AOTB_REQUIRE(col_off_dev || col_off + rows <= ldt, "aotb_split_cols_f16x2: columns exceed the row length");
    dim3 grid(cdiv(rows, 32), cdiv(C, 32)), block(32, 8);
    launch(split_cols_kernel, grid, block, 0, (cudaStream_t)stream, src, lds, (__half*)hiT, (__half*)loT, ldt, rows, C,
           col_off, col_off_dev);
    return check_launch("aotb_split_cols_f16x2");
}

extern "C" int aotb_row_softmax_f32(float* S, int ld, int N, int cols, int Tk, const int* Tk_dev, float scale,
                                    void* stream) {
    AOTB_REQUIRE(S && N > 0 && cols > 0 && cols % 4 == 0 && ld % 4 == 0 && ld >= cols && (Tk > 0 || Tk_dev) &&
                     ((uintptr_t)S % 16 == 0),
                 "aotb_row_softmax_f32: bad args");
    launch(row_softmax_kernel, dim3(N), dim3(256), 0, (cudaStream_t)stream, S, ld, cols, Tk, Tk_dev, scale);
    return check_launch("aotb_row_softmax_f32");
}
