// Frame I/O around the hot path on the GPU: the reference prepares every frame on the host
// (dataloaders/eval_datasets.py:60-61 float copy of the cv2 image, dataloaders/video_transforms.py:594-715 MultiRestrictSize =
// cv2.resize(INTER_CUBIC) + MultiToTensor = / 255, - mean, / std, HWC -> CHW) and converts every predicted mask on the host
// (utils/image.py:103-105).  Here the uint8 frame is uploaded as it is (3 bytes per pixel instead of 12) and one kernel does
// resize + normalisation + layout change; the label map leaves the device as uint8.
#include "common.cuh"
#include <cstdint>

namespace aotb {

// One thread per output pixel: 4 x 4 Keys-cubic taps (A = -0.75; tap indices and weights precomputed per axis on the host
// exactly as cv2 derives them), horizontal pass first, then vertical, each left to right in fp32 without contraction -- the
// order of cv2's HResizeCubic / VResizeCubic -- then (v / 255 [fp32] - mean [fp64]) / std [fp64] as numpy evaluates
// MultiToTensor's three statements.  ix == nullptr: no resize (the frame already has the network size).
__global__ void preprocess_bgr_u8_kernel(const uint8_t* __restrict__ img, int H, int W, const int* __restrict__ ix,
                                         const float* __restrict__ cx, const int* __restrict__ iy,
                                         const float* __restrict__ cy, float* __restrict__ out, int Ho, int Wo, int flip) {
    pdl_sync();
    const double mean[3] = {0.485, 0.456, 0.406}, stdv[3] = {0.229, 0.224, 0.225};
    const int total = Ho * Wo;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int oy = i / Wo, ox = i - oy * Wo;
        const int sx = flip ? Wo - 1 - ox : ox;              // flip is applied after the resize (video_transforms.py:677-688)
        float v[3];
        if (ix) {
            int xs[4], ys[4];
            float ax[4], ay[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { xs[k] = ix[sx * 4 + k]; ax[k] = cx[sx * 4 + k]; ys[k] = iy[oy * 4 + k]; ay[k] = cy[oy * 4 + k]; }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float acc = 0.f;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const uint8_t* row = img + ((size_t)ys[r] * W) * 3 + c;
                    float hsum = __fmul_rn((float)row[xs[0] * 3], ax[0]);
                    hsum = __fadd_rn(hsum, __fmul_rn((float)row[xs[1] * 3], ax[1]));
                    hsum = __fadd_rn(hsum, __fmul_rn((float)row[xs[2] * 3], ax[2]));
                    hsum = __fadd_rn(hsum, __fmul_rn((float)row[xs[3] * 3], ax[3]));
                    const float t = __fmul_rn(hsum, ay[r]);
                    acc = r == 0 ? t : __fadd_rn(acc, t);
                }
                v[c] = acc;
            }
        } else {
            const uint8_t* p = img + ((size_t)oy * W + sx) * 3;
            v[0] = (float)p[0]; v[1] = (float)p[1]; v[2] = (float)p[2];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float a = __fdiv_rn(v[c], 255.f);
            const float b = (float)((double)a - mean[c]);
            out[(size_t)c * total + i] = (float)((double)b / stdv[c]);
        }
    }
}

__global__ void label_to_u8_kernel(const float* __restrict__ in, uint8_t* __restrict__ out, int n) {
    pdl_sync();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = (uint8_t)(int)in[i];
}

}  // namespace aotb

using namespace aotb;

// img: uint8 [H][W][3] (device, channel order as decoded); taps: ix / cx [Wo][4], iy / cy [Ho][4] (all null = no resize);
// out: fp32 [3][Ho][Wo] = the tensor MultiRestrictSize + MultiToTensor produce for `current_img`.
extern "C" int aotb_preprocess_bgr_u8(const void* img, int H, int W, const int* ix, const float* cx, const int* iy,
                                      const float* cy, float* out, int Ho, int Wo, int flip, void* stream) {
    AOTB_REQUIRE(img && out && H > 0 && W > 0 && Ho > 0 && Wo > 0, "aotb_preprocess_bgr_u8: bad args");
    const bool taps = ix && cx && iy && cy;
    AOTB_REQUIRE(taps || (!ix && !cx && !iy && !cy && Ho == H && Wo == W),
                 "aotb_preprocess_bgr_u8: pass all four tap tables, or none when the size is unchanged");
    int g = cdiv(Ho * Wo, 256);
    if (g > 148 * 16) g = 148 * 16;
    launch(preprocess_bgr_u8_kernel, dim3(g), dim3(256), 0, (cudaStream_t)stream, (const uint8_t*)img, H, W, taps ? ix : nullptr,
           cx, iy, cy, out, Ho, Wo, flip);
    return check_launch("aotb_preprocess_bgr_u8");
}

extern "C" int aotb_label_to_u8(const float* label, void* out_u8, int n, void* stream) {
    AOTB_REQUIRE(label && out_u8 && n > 0, "aotb_label_to_u8: bad args");
    int g = cdiv(n, 256);
    if (g > 148 * 8) g = 148 * 8;
    launch(label_to_u8_kernel, dim3(g), dim3(256), 0, (cudaStream_t)stream, label, (uint8_t*)out_u8, n);
    return check_launch("aotb_label_to_u8");
}
