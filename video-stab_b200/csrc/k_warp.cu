// k_warp.cu — the output stage: cv::copyMakeBorder (Stabilizer.cpp:981-990) + cv::warpAffine
// INTER_LINEAR / BORDER_CONSTANT (:1056-1060) + crop-and-zoom (:1108-1124), writing the output once.
// Specification: oracle/cv_models.py warp_affine / copy_make_border / resize_linear (bit-exact vs
// cv2 4.13).  Pure integer arithmetic: 10-bit fixed-point source coordinates, 5-bit sub-pixel
// position, 15-bit bilinear weights.  HBM-bound: algorithmic traffic is one read + one write of the
// frame (2*3*W*H bytes).
#include "kernels.h"

// Source coordinate of one output pixel, exactly as cv::warpAffine computes it:
//   adelta[x] = rint(M0*x*1024), X0 = rint((M1*y+M2)*1024) + 16, X = (X0 + adelta[x]) >> 5
struct FixedCoord {
    int sx, sy, ax, ay;
};
static __device__ __forceinline__ int sat_short(int v) { return min(max(v, -32768), 32767); }
static __device__ __forceinline__ int sat_int(double v) {
    // cv::saturate_cast<int>(double) == cvRound (round half to even), saturating
    return __double2int_rn(v);
}
static __device__ __forceinline__ FixedCoord warp_coord(const double* __restrict__ m, int x, int y) {
    int ad = sat_int(m[0] * (double)x * 1024.0);
    int bd = sat_int(m[3] * (double)x * 1024.0);
    int X0 = sat_int((m[1] * (double)y + m[2]) * 1024.0) + 16;
    int Y0 = sat_int((m[4] * (double)y + m[5]) * 1024.0) + 16;
    int X = (X0 + ad) >> 5, Y = (Y0 + bd) >> 5;
    FixedCoord c;
    c.sx = sat_short(X >> 5);
    c.sy = sat_short(Y >> 5);
    c.ax = X & 31;
    c.ay = Y & 31;
    return c;
}

static __device__ __forceinline__ int border_map(int p, int len, int mode) {
    // cv::borderInterpolate; returns -1 for BORDER_CONSTANT outside
    if ((unsigned)p < (unsigned)len) return p;
    if (mode == 0) return -1;                       // BORDER_CONSTANT
    if (mode == 1) return p < 0 ? 0 : len - 1;      // BORDER_REPLICATE
    if (mode == 3) {                                // BORDER_WRAP
        int q = p % len;
        return q < 0 ? q + len : q;
    }
    if (len == 1) return 0;
    const int delta = (mode == 4) ? 1 : 0;          // REFLECT_101 : REFLECT
    do {
        if (p < 0) p = -p - 1 + delta;
        else p = len - 1 - (p - len) - delta;
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

// One output pixel (3 channels).  BORDER: the source is the virtual (w+2b)x(h+2b) bordered frame.
template <bool BORDER>
static __device__ __forceinline__ void warp_pixel(const uint8_t* __restrict__ src, int w, int h, size_t stride,
                                                  const double* __restrict__ m, int b, int bmode, int x, int y,
                                                  uint8_t* __restrict__ out) {
    FixedCoord c = warp_coord(m, x, y);
    const int vw = BORDER ? w + 2 * b : w, vh = BORDER ? h + 2 * b : h;
    const int w00 = (32 - c.ax) * (32 - c.ay), w01 = c.ax * (32 - c.ay), w10 = (32 - c.ax) * c.ay, w11 = c.ax * c.ay;
    int acc0 = 0, acc1 = 0, acc2 = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        int tx = c.sx + (t & 1), ty = c.sy + (t >> 1);
        int wt = t == 0 ? w00 : t == 1 ? w01 : t == 2 ? w10 : w11;
        if ((unsigned)tx >= (unsigned)vw || (unsigned)ty >= (unsigned)vh) continue;   // BORDER_CONSTANT 0
        if (BORDER) {
            tx = border_map(tx - b, w, bmode);
            ty = border_map(ty - b, h, bmode);
            if (tx < 0 || ty < 0) continue;
        }
        const uint8_t* p = src + (size_t)ty * stride + 3 * tx;
        acc0 += wt * p[0];
        acc1 += wt * p[1];
        acc2 += wt * p[2];
    }
    // weights are (..)*32 with sum 32768: (acc*32 + 16384) >> 15 == (acc + 512) >> 10
    out[0] = (uint8_t)((acc0 + 512) >> 10);
    out[1] = (uint8_t)((acc1 + 512) >> 10);
    out[2] = (uint8_t)((acc2 + 512) >> 10);
}

template <bool BORDER>
__global__ void __launch_bounds__(256) k_warp_lanes(const LaneDev* __restrict__ lanes, PtrPack src, MutPtrPack dst, WarpGeom g) {
    const WarpParams* wp = lanes[blockIdx.z].wp;
    int x = blockIdx.x * 64 + (threadIdx.x & 63);
    int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= g.out_w || y >= g.out_h) return;
    uint8_t* o = dst.p[blockIdx.z] + (size_t)y * g.out_stride + 3 * x;
    warp_pixel<BORDER>(src.p[blockIdx.z], g.src_w, g.src_h, g.src_stride, wp->m, g.border, g.border_mode, x, y, o);
}

__global__ void __launch_bounds__(256) k_warp_frames(const uint8_t* __restrict__ src, int sw, int sh, size_t sstride, size_t sframe,
                                                      uint8_t* __restrict__ dst, int dw, int dh, size_t dstride, size_t dframe,
                                                      const WarpParams* __restrict__ wps) {
    int x = blockIdx.x * 64 + (threadIdx.x & 63);
    int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= dw || y >= dh) return;
    uint8_t* o = dst + blockIdx.z * dframe + (size_t)y * dstride + 3 * x;
    warp_pixel<false>(src + blockIdx.z * sframe, sw, sh, sstride, wps[blockIdx.z].m, 0, 0, x, y, o);
}

void launch_warp(const LaneDev* lanes, int n_lanes, const PtrPack& src, const MutPtrPack& dst, WarpGeom g,
                 uint8_t* const* scratch, cudaStream_t st) {
    if (g.mode == 2) {
        // crop+zoom, two passes for now: warp into the lane's scratch frame, then cv::resize the
        // (b,b,w-2b,h-2b) crop back to w x h.
        MutPtrPack tmp;
        for (int i = 0; i < n_lanes; ++i) tmp.p[i] = scratch[i];
        WarpGeom g1 = g;
        g1.mode = 0; g1.out_w = g.src_w; g1.out_h = g.src_h; g1.out_stride = (size_t)g.src_w * 3;
        dim3 grid((g1.out_w + 63) / 64, (g1.out_h + 3) / 4, n_lanes);
        k_warp_lanes<false><<<grid, 256, 0, st>>>(lanes, src, tmp, g1);
        int b = g.border, cw = g.src_w - 2 * b, ch = g.src_h - 2 * b;
        for (int i = 0; i < n_lanes; ++i)
            launch_resize_linear(scratch[i] + (size_t)b * g1.out_stride + 3 * b, cw, ch, g1.out_stride, 3,
                                 dst.p[i], g.out_w, g.out_h, g.out_stride, st);
        return;
    }
    dim3 grid((g.out_w + 63) / 64, (g.out_h + 3) / 4, n_lanes);
    if (g.mode == 1)
        k_warp_lanes<true><<<grid, 256, 0, st>>>(lanes, src, dst, g);
    else
        k_warp_lanes<false><<<grid, 256, 0, st>>>(lanes, src, dst, g);
}

void launch_warp_matrices(const uint8_t* src, int sw, int sh, size_t sstride, size_t sframe,
                          uint8_t* dst, int dw, int dh, size_t dstride, size_t dframe,
                          const WarpParams* d_wp, int n_frames, cudaStream_t st) {
    dim3 grid((dw + 63) / 64, (dh + 3) / 4, n_frames);
    k_warp_frames<<<grid, 256, 0, st>>>(src, sw, sh, sstride, sframe, dst, dw, dh, dstride, dframe, d_wp);
}
