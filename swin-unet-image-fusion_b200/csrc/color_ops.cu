// Input / output edges of the reference's inference script, on the device and batched:
//   sf_bgr_to_ycrcb : a015_dataset.py:86-93 (cv2 BGR -> YCrCb on uint8) + a015:56-60 (ToImage, ToDtype scale)
//                     + the channel split of a017_test.py:68
//   sf_ycrcb_to_rgb : a017_test.py:83-88 (clamp, concat with CrCb, cv2 YCrCb -> RGB on float32)
// Byte / integer arithmetic restated from OpenCV's 8-bit fixed-point path and bit-exact with it; the float
// path uses the same fused multiply-adds as OpenCV's vector code.  Pure streaming kernels: one thread per
// pixel quad, 128-bit stores, grid = a multiple of the SM count.
#include "common.cuh"

namespace sf {

static constexpr int kCT = 256;
static inline int color_grid(long long n) {
    long long b = (n + kCT - 1) / kCT;
    const long long cap = (long long)sm_count() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

__device__ __forceinline__ void ycrcb_u8(int b, int g, int r, int& y, int& cr, int& cb) {
    constexpr int S = 14, HALF = 1 << (S - 1), DELTA = 128 << S;
    y = (r * 4899 + g * 9617 + b * 1868 + HALF) >> S;
    cr = ((r - y) * 11682 + DELTA + HALF) >> S;
    cb = ((b - y) * 9241 + DELTA + HALF) >> S;
    y = min(max(y, 0), 255); cr = min(max(cr, 0), 255); cb = min(max(cb, 0), 255);
}

// bgr (B,H,W,3) uint8 -> y (B,1,H,W), crcb (B,2,H,W) fp32 = u8 * fl32(1/255)
__global__ void k_bgr_to_ycrcb(const uint8_t* __restrict__ bgr, float* __restrict__ y, float* __restrict__ crcb, int B, long long HW) {
    const float inv255 = (float)(1.0 / 255);
    const long long nq = (HW + 3) / 4;   // pixel quads per image
    const long long total = (long long)B * nq;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long img = i / nq, q = i - img * nq;
        const long long p0 = q * 4;
        const int n = (int)min(4LL, HW - p0);
        const uint8_t* src = bgr + (img * HW + p0) * 3;
        float yy[4], cr[4], cb[4];
        uint8_t raw[12];
        if (n == 4 && ((reinterpret_cast<uintptr_t>(src) & 3) == 0)) {
            const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
            uint32_t w0 = __ldg(s32), w1 = __ldg(s32 + 1), w2 = __ldg(s32 + 2);
            *reinterpret_cast<uint32_t*>(raw) = w0; *reinterpret_cast<uint32_t*>(raw + 4) = w1; *reinterpret_cast<uint32_t*>(raw + 8) = w2;
        } else {
            for (int k = 0; k < 12; k++) raw[k] = k < 3 * n ? src[k] : 0;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int a, b2, c;
            ycrcb_u8(raw[3 * k], raw[3 * k + 1], raw[3 * k + 2], a, b2, c);
            yy[k] = (float)a * inv255; cr[k] = (float)b2 * inv255; cb[k] = (float)c * inv255;
        }
        float* yo = y + img * HW + p0;
        float* cro = crcb + img * 2 * HW + p0;
        float* cbo = cro + HW;
        if (n == 4 && (HW & 3) == 0) {
            *reinterpret_cast<float4*>(yo) = make_float4(yy[0], yy[1], yy[2], yy[3]);
            *reinterpret_cast<float4*>(cro) = make_float4(cr[0], cr[1], cr[2], cr[3]);
            *reinterpret_cast<float4*>(cbo) = make_float4(cb[0], cb[1], cb[2], cb[3]);
        } else {
            for (int k = 0; k < n; k++) { yo[k] = yy[k]; cro[k] = cr[k]; cbo[k] = cb[k]; }
        }
    }
}

// fus_y (B,1,H,W), crcb (B,2,H,W) -> rgb (B,3,H,W) fp32
__global__ void k_ycrcb_to_rgb(const float* __restrict__ fy, const float* __restrict__ crcb, float* __restrict__ rgb, int B, long long HW) {
    const long long total = (long long)B * HW;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long img = i / HW, p = i - img * HW;
        const float y = fminf(fmaxf(__ldg(fy + i), 0.f), 1.f);
        const float cr = __ldg(crcb + img * 2 * HW + p) - 0.5f, cb = __ldg(crcb + img * 2 * HW + HW + p) - 0.5f;
        float* o = rgb + img * 3 * HW + p;
        o[0] = __fmaf_rn(cr, 1.403f, y);
        o[HW] = __fmaf_rn(cr, -0.714f, __fmaf_rn(cb, -0.344f, y));
        o[2 * HW] = __fmaf_rn(cb, 1.773f, y);
    }
}

}  // namespace sf

extern "C" int sf_bgr_to_ycrcb(const uint8_t* bgr, float* y, float* crcb, int B, int H, int W, void* stream) {
    using namespace sf;
    SF_CHECK_ARG(bgr && y && crcb && B > 0 && H > 0 && W > 0, "sf_bgr_to_ycrcb: null pointer or empty image (%d,%d,%d)", B, H, W);
    const long long HW = (long long)H * W;
    ProfScope ps("bgr_to_ycrcb", 0.0, 15.0 * B * HW, as_stream(stream));
    k_bgr_to_ycrcb<<<color_grid((long long)B * ((HW + 3) / 4)), kCT, 0, as_stream(stream)>>>(bgr, y, crcb, B, HW);
    SF_CHECK_LAUNCH("sf_bgr_to_ycrcb");
    return SF_OK;
}

extern "C" int sf_ycrcb_to_rgb(const float* fus_y, const float* crcb, float* rgb, int B, int H, int W, void* stream) {
    using namespace sf;
    SF_CHECK_ARG(fus_y && crcb && rgb && B > 0 && H > 0 && W > 0, "sf_ycrcb_to_rgb: null pointer or empty image (%d,%d,%d)", B, H, W);
    const long long HW = (long long)H * W;
    ProfScope ps("ycrcb_to_rgb", 0.0, 24.0 * B * HW, as_stream(stream));
    k_ycrcb_to_rgb<<<color_grid((long long)B * HW), kCT, 0, as_stream(stream)>>>(fus_y, crcb, rgb, B, HW);
    SF_CHECK_LAUNCH("sf_ycrcb_to_rgb");
    return SF_OK;
}
