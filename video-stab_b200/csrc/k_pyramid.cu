// k_pyramid.cu — analysis-image build: cv::resize(INTER_LINEAR) + cv::cvtColor(BGR2GRAY) +
// the cv::pyrDown levels cv::calcOpticalFlowPyrLK builds internally.
// Reference call sites: Stabilizer.cpp:304-305 (first frame, 480x270), :449-450 (960x540),
// :602 (prevGray up-sampling on frame 1), :611 (pyramid inside PyrLK).
// Arithmetic specification: oracle/cv_models.py resize_linear / bgr2gray / pyr_down (pinned
// bit-exact against cv2 4.13).  Pure integer work, HBM-bound: one pass over the BGR frame.
#include "kernels.h"

// ---------------------------------------------------------------- cv::resize INTER_LINEAR tables
struct AxisTap {
    int s0, s1;      // tap indices (already clamped)
    int a0, a1;      // 11-bit coefficients
};

// horizontal semantics: index clamped AND fraction zeroed at both ends
static __device__ __forceinline__ AxisTap tap_h(int d, int src, double scale) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= src - 1) { s = src - 1; f = 0.f; }
    AxisTap t;
    t.s0 = s;
    t.s1 = min(s + 1, src - 1);
    t.a0 = __float2int_rn((1.f - f) * 2048.f);
    t.a1 = __float2int_rn(f * 2048.f);
    return t;
}
// vertical semantics: coefficients from the unclamped fraction, row indices clipped
static __device__ __forceinline__ AxisTap tap_v(int d, int src, double scale) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    AxisTap t;
    t.s0 = min(max(s, 0), src - 1);
    t.s1 = min(max(s + 1, 0), src - 1);
    t.a0 = __float2int_rn((1.f - f) * 2048.f);
    t.a1 = __float2int_rn(f * 2048.f);
    return t;
}
static __device__ __forceinline__ int vres(int h0, int h1, int b0, int b1) {
    return (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
}
static __device__ __forceinline__ int gray_of(int b, int g, int r) {
    return (3735 * b + 19235 * g + 9798 * r + 16384) >> 15;
}

// One analysis pixel (ox,oy) of the (aw x ah) gray image from the full-resolution BGR frame.
template <int MODE>   // 0: exact 2x (INTER_AREA fast path), 1: generic 11-bit bilinear
static __device__ __forceinline__ int analysis_pixel(const uint8_t* __restrict__ src, int w, int h, size_t stride,
                                                     int aw, int ah, int ox, int oy, double sx, double sy) {
    if (MODE == 0) {
        const uint8_t* r0 = src + (size_t)(2 * oy) * stride + 6 * ox;
        const uint8_t* r1 = r0 + stride;
        int b = (r0[0] + r0[3] + r1[0] + r1[3] + 2) >> 2;
        int g = (r0[1] + r0[4] + r1[1] + r1[4] + 2) >> 2;
        int r = (r0[2] + r0[5] + r1[2] + r1[5] + 2) >> 2;
        return gray_of(b, g, r);
    } else {
        AxisTap tx = tap_h(ox, w, sx);
        AxisTap ty = tap_v(oy, h, sy);
        const uint8_t* r0 = src + (size_t)ty.s0 * stride;
        const uint8_t* r1 = src + (size_t)ty.s1 * stride;
        int c[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            int h0 = r0[3 * tx.s0 + k] * tx.a0 + r0[3 * tx.s1 + k] * tx.a1;
            int h1 = r1[3 * tx.s0 + k] * tx.a0 + r1[3 * tx.s1 + k] * tx.a1;
            c[k] = vres(h0, h1, ty.a0, ty.a1);
        }
        return gray_of(c[0], c[1], c[2]);
    }
}

// grid: (ceil((aw+2P)/128), ah+2P, lanes).  Writes the level INCLUDING its reflect-101 frame, so no
// later kernel needs border logic.  The 2x path reads 6 bytes x 2 rows per pixel: every sector of
// the source is touched exactly once (interior) and the reads of a warp cover 2 x 192 contiguous bytes.
template <int MODE>
__global__ void __launch_bounds__(128) k_gray_resize(const LaneDev* __restrict__ lanes, PtrPack src, int w, int h,
                                                      size_t stride, int slot, double sx, double sy) {
    const LaneDev& L = lanes[blockIdx.z];
    const GrayLevel lv = slot < 0 ? L.small0 : L.pyr[slot].lv[0];
    int px = blockIdx.x * blockDim.x + threadIdx.x - VS_PAD;
    int py = blockIdx.y - VS_PAD;
    if (px >= lv.w + VS_PAD) return;
    int ox = reflect101(px, lv.w), oy = reflect101(py, lv.h);
    int v = analysis_pixel<MODE>(src.p[blockIdx.z], w, h, stride, lv.w, lv.h, ox, oy, sx, sy);
    lv.base[(ptrdiff_t)py * lv.pitch + px] = (uint8_t)v;
}

void launch_gray_resize(const LaneDev* lanes, int n_lanes, const PtrPack& src, int w, int h, size_t stride,
                        int slot, cudaStream_t st) {
    int aw = slot < 0 ? VS_FW : VS_AW, ah = slot < 0 ? VS_FH : VS_AH;
    dim3 grid((aw + 2 * VS_PAD + 127) / 128, ah + 2 * VS_PAD, n_lanes);
    double sx = 1.0 / ((double)aw / (double)w), sy = 1.0 / ((double)ah / (double)h);
    if (w == 2 * aw && h == 2 * ah)
        k_gray_resize<0><<<grid, 128, 0, st>>>(lanes, src, w, h, stride, slot, sx, sy);
    else
        k_gray_resize<1><<<grid, 128, 0, st>>>(lanes, src, w, h, stride, slot, sx, sy);
}

// ---------------------------------------------------------------- 480x270 gray -> 960x540 (frame 1 only)
__global__ void __launch_bounds__(128) k_upsample_small(const LaneDev* __restrict__ lanes, int slot, double sx, double sy) {
    const LaneDev& L = lanes[blockIdx.z];
    const GrayLevel s = L.small0;
    const GrayLevel d = L.pyr[slot].lv[0];
    int px = blockIdx.x * blockDim.x + threadIdx.x - VS_PAD;
    int py = blockIdx.y - VS_PAD;
    if (px >= d.w + VS_PAD) return;
    int ox = reflect101(px, d.w), oy = reflect101(py, d.h);
    AxisTap tx = tap_h(ox, s.w, sx);
    AxisTap ty = tap_v(oy, s.h, sy);
    const uint8_t* r0 = s.base + (ptrdiff_t)ty.s0 * s.pitch;
    const uint8_t* r1 = s.base + (ptrdiff_t)ty.s1 * s.pitch;
    int h0 = r0[tx.s0] * tx.a0 + r0[tx.s1] * tx.a1;
    int h1 = r1[tx.s0] * tx.a0 + r1[tx.s1] * tx.a1;
    d.base[(ptrdiff_t)py * d.pitch + px] = (uint8_t)vres(h0, h1, ty.a0, ty.a1);
}

void launch_upsample_small(const LaneDev* lanes, int n_lanes, int slot, cudaStream_t st) {
    dim3 grid((VS_AW + 2 * VS_PAD + 127) / 128, VS_AH + 2 * VS_PAD, n_lanes);
    k_upsample_small<<<grid, 128, 0, st>>>(lanes, slot, 1.0 / ((double)VS_AW / VS_FW), 1.0 / ((double)VS_AH / VS_FH));
}

// ---------------------------------------------------------------- cv::pyrDown
// out(x,y) = (sum_{i,j} k_i k_j src(2x+i-2, 2y+j-2) + 128) >> 8, k = [1 4 6 4 1]; the source level's
// materialised reflect-101 frame supplies the out-of-image taps.  Padded output domain again.
__global__ void __launch_bounds__(128) k_pyrdown(const LaneDev* __restrict__ lanes, int slot, int level) {
    const LaneDev& L = lanes[blockIdx.z];
    const GrayLevel s = L.pyr[slot].lv[level - 1];
    const GrayLevel d = L.pyr[slot].lv[level];
    int px = blockIdx.x * blockDim.x + threadIdx.x - VS_PAD;
    int py = blockIdx.y - VS_PAD;
    if (px >= d.w + VS_PAD) return;
    int ox = reflect101(px, d.w), oy = reflect101(py, d.h);
    const uint8_t* c = s.base + (ptrdiff_t)(2 * oy) * s.pitch + 2 * ox;
    int acc = 0;
#pragma unroll
    for (int j = -2; j <= 2; ++j) {
        const uint8_t* r = c + (ptrdiff_t)j * s.pitch;
        int row = r[-2] + 4 * r[-1] + 6 * r[0] + 4 * r[1] + r[2];
        const int kj = (j == 0) ? 6 : ((j == -1 || j == 1) ? 4 : 1);
        acc += kj * row;
    }
    d.base[(ptrdiff_t)py * d.pitch + px] = (uint8_t)((acc + 128) >> 8);
}

void launch_pyrdown(const LaneDev* lanes, int n_lanes, int slot, cudaStream_t st) {
    int w = VS_AW, h = VS_AH;
    for (int l = 1; l < VS_LEVELS; ++l) {
        w = (w + 1) / 2;
        h = (h + 1) / 2;
        dim3 grid((w + 2 * VS_PAD + 127) / 128, h + 2 * VS_PAD, n_lanes);
        k_pyrdown<<<grid, 128, 0, st>>>(lanes, slot, l);
    }
}

// ---------------------------------------------------------------- generic cv::resize INTER_LINEAR (8UC1/8UC3)
template <int CH>
__global__ void __launch_bounds__(128) k_resize_linear(const uint8_t* __restrict__ src, int sw, int sh, size_t sstride,
                                                        uint8_t* __restrict__ dst, int dw, int dh, size_t dstride,
                                                        double sx, double sy, int area2) {
    int ox = blockIdx.x * blockDim.x + threadIdx.x;
    int oy = blockIdx.y;
    if (ox >= dw) return;
    uint8_t* o = dst + (size_t)oy * dstride + (size_t)ox * CH;
    if (area2) {
        const uint8_t* r0 = src + (size_t)(2 * oy) * sstride + (size_t)(2 * ox) * CH;
        const uint8_t* r1 = r0 + sstride;
#pragma unroll
        for (int k = 0; k < CH; ++k) o[k] = (uint8_t)((r0[k] + r0[CH + k] + r1[k] + r1[CH + k] + 2) >> 2);
        return;
    }
    AxisTap tx = tap_h(ox, sw, sx);
    AxisTap ty = tap_v(oy, sh, sy);
    const uint8_t* r0 = src + (size_t)ty.s0 * sstride;
    const uint8_t* r1 = src + (size_t)ty.s1 * sstride;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        int h0 = r0[CH * tx.s0 + k] * tx.a0 + r0[CH * tx.s1 + k] * tx.a1;
        int h1 = r1[CH * tx.s0 + k] * tx.a0 + r1[CH * tx.s1 + k] * tx.a1;
        o[k] = (uint8_t)vres(h0, h1, ty.a0, ty.a1);
    }
}

void launch_resize_linear(const uint8_t* src, int sw, int sh, size_t sstride, int ch,
                          uint8_t* dst, int dw, int dh, size_t dstride, cudaStream_t st) {
    dim3 grid((dw + 127) / 128, dh, 1);
    double sx = 1.0 / ((double)dw / (double)sw), sy = 1.0 / ((double)dh / (double)sh);
    int area2 = (sw == 2 * dw && sh == 2 * dh);
    if (ch == 3)
        k_resize_linear<3><<<grid, 128, 0, st>>>(src, sw, sh, sstride, dst, dw, dh, dstride, sx, sy, area2);
    else
        k_resize_linear<1><<<grid, 128, 0, st>>>(src, sw, sh, sstride, dst, dw, dh, dstride, sx, sy, area2);
}

// ---------------------------------------------------------------- tight <-> padded level copies (tests)
__global__ void k_pack_level(const uint8_t* __restrict__ src, GrayLevel d) {
    int px = blockIdx.x * blockDim.x + threadIdx.x - VS_PAD;
    int py = blockIdx.y - VS_PAD;
    if (px >= d.w + VS_PAD) return;
    d.base[(ptrdiff_t)py * d.pitch + px] = src[(size_t)reflect101(py, d.h) * d.w + reflect101(px, d.w)];
}
__global__ void k_unpack_level(GrayLevel s, uint8_t* __restrict__ dst) {
    int px = blockIdx.x * blockDim.x + threadIdx.x;
    int py = blockIdx.y;
    if (px >= s.w) return;
    dst[(size_t)py * s.w + px] = s.base[(ptrdiff_t)py * s.pitch + px];
}
void launch_pack_level(const uint8_t* src, GrayLevel dst, cudaStream_t st) {
    dim3 grid((dst.w + 2 * VS_PAD + 127) / 128, dst.h + 2 * VS_PAD, 1);
    k_pack_level<<<grid, 128, 0, st>>>(src, dst);
}
void launch_unpack_level(GrayLevel src, uint8_t* dst, cudaStream_t st) {
    dim3 grid((src.w + 127) / 128, src.h, 1);
    k_unpack_level<<<grid, 128, 0, st>>>(src, dst);
}
