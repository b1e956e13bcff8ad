// k_warp.cu — the output stage: cv::copyMakeBorder (Stabilizer.cpp:981-990) + cv::warpAffine
// INTER_LINEAR / BORDER_CONSTANT (:1056-1060) + crop-and-zoom (:1108-1124), writing the output once.
// Specification: oracle/cv_models.py warp_affine / copy_make_border / resize_linear (bit-exact vs
// cv2 4.13).  Pure integer arithmetic: 10-bit fixed-point source coordinates, 5-bit sub-pixel
// position, 15-bit bilinear weights.  HBM-bound: algorithmic traffic is one read + one write of the
// frame (2*3*W*H bytes).
#include "kernels.h"
#include <climits>

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

// ------------------------------------------------------------------------------------------------
// Tiled fast path (plain warp, mode 0).  One CTA = one 128x32 output tile:
//   1. the exact source bounding box of the tile is derived from its four corners (the fixed-point
//      coordinate is monotone in x and in y), widened to 4-pixel (12-byte) groups;
//   2. the box is staged into shared memory with coalesced 32-bit loads (a warp reads 384 contiguous
//      bytes per 3 instructions) and re-packed on the fly to 4-byte BGRx pixels, one conflict-free
//      STS.128 per lane; rows/columns outside the frame are staged as zeros, which IS
//      cv::BORDER_CONSTANT(0), so the inner loop has no border logic at all;
//   3. lane l produces pixels x0+l+32j (consecutive lanes -> consecutive shared-memory words, no bank
//      conflicts): 4 aligned LDS.32 taps, byte gathers with PRMT, horizontal pass on DP4A (weights
//      32-ax, ax), vertical pass on IMAD with the weight pre-scaled by 64 so the rounded 8-bit result
//      sits in byte 2 of the accumulator ((acc+512)>>10 without a shift);
//   4. the 128x32x3 output tile is staged in shared memory and written once with 128-bit stores.
// Tiles whose source box does not fit (large rotations) or whose coordinates approach the int16
// saturation of cv::remap fall back to the per-pixel path, CTA-uniformly.
#define WT_W 128
#define WT_H 32
#define WT_PITCH 144                 // fixed source-tile row pitch in BGRx words (128 + rotation slack + alignment)
#define WT_ROWS 42                   // source-tile row capacity  (144*42*4 = 24 KB)
#define WT_THREADS 256
#define WT_GRPS 36                   // 4-pixel column groups of the staged box (= WT_PITCH / 4)
#define WT_SROWS 7                   // staging rows in flight: 36 x 7 = 252 threads
#define WT_STAGE_IT 6                // staging tasks per thread (42 rows / 7)

struct WarpTileSmem {
    uint32_t src[WT_PITCH * WT_ROWS];
    uint32_t out[(WT_THREADS / 32) * WT_W];   // one BGRx output row per warp (4 KB)
    int2 rowXY[WT_H];
    int2 colAB[WT_W];
    int box[6];                          // ax0, by0, ngrp, nrows, ok, unused
};

static __device__ __forceinline__ void warp_tile(WarpTileSmem& S, const uint8_t* __restrict__ src, int sw, int sh,
                                                 size_t sstride, uint8_t* __restrict__ dst, int dw, int dh,
                                                 size_t dstride, const double* __restrict__ m, bool src_vec, bool dst_vec) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * WT_W, y0 = blockIdx.y * WT_H;
    const double m0 = m[0], m1 = m[1], m2 = m[2], m3 = m[3], m4 = m[4], m5 = m[5];

    // fixed-point row/column terms of cv::warpAffine, one per thread; columns beyond the frame reuse the
    // last valid column so that their (discarded) taps stay inside the staged box
    if (tid < WT_W) {
        double xd = (double)min(x0 + tid, dw - 1);
        S.colAB[tid] = make_int2(sat_int(m0 * xd * 1024.0), sat_int(m3 * xd * 1024.0));
    } else if (tid < WT_W + WT_H) {
        double yd = (double)min(y0 + tid - WT_W, dh - 1);
        S.rowXY[tid - WT_W] = make_int2(sat_int((m1 * yd + m2) * 1024.0) + 16, sat_int((m4 * yd + m5) * 1024.0) + 16);
    } else if (tid == WT_W + WT_H) {
        const int xa = x0, xb = min(x0 + WT_W, dw) - 1, ya = y0, yb = min(y0 + WT_H, dh) - 1;
        int minx = INT_MAX, maxx = INT_MIN, miny = INT_MAX, maxy = INT_MIN;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double xd = (double)((c & 1) ? xb : xa), yd = (double)((c & 2) ? yb : ya);
            int X = (sat_int((m1 * yd + m2) * 1024.0) + 16 + sat_int(m0 * xd * 1024.0)) >> 10;
            int Y = (sat_int((m4 * yd + m5) * 1024.0) + 16 + sat_int(m3 * xd * 1024.0)) >> 10;
            minx = min(minx, X); maxx = max(maxx, X); miny = min(miny, Y); maxy = max(maxy, Y);
        }
        int ax0 = minx & ~3;                                   // floor to a 4-pixel group (also for negatives)
        int ngrp = (maxx + 1 - ax0) / 4 + 1;
        int nrows = maxy + 2 - miny;
        bool ok = minx > -30000 && maxx < 30000 && miny > -30000 && maxy < 30000 &&
                  ngrp * 4 <= WT_PITCH && nrows <= WT_ROWS && ngrp > 0 && nrows > 0;
        S.box[0] = ax0; S.box[1] = miny; S.box[2] = ngrp; S.box[3] = nrows; S.box[4] = ok ? 1 : 0;
    }
    __syncthreads();
    const int ax0 = S.box[0], by0 = S.box[1], ngrp = S.box[2], nrows = S.box[3];
    if (!S.box[4]) {
        // generic per-pixel path for this tile
        for (int i = tid; i < WT_W * WT_H; i += WT_THREADS) {
            int x = x0 + (i & (WT_W - 1)), y = y0 + i / WT_W;
            if (x < dw && y < dh) warp_pixel<false>(src, sw, sh, sstride, m, 0, 0, x, y, dst + (size_t)y * dstride + 3 * x);
        }
        return;
    }
    // ---- stage the source box as BGRx: one task = 4 pixels = 12 source bytes -> one 16-byte store.
    //      Fixed 2-D thread map: 36 column groups x 7 rows (252 of 256 threads), each thread walks down
    //      its column 7 rows at a time; all its loads are issued before the first is consumed.
    {
        const bool interior = src_vec && ax0 >= 0 && ax0 + 4 * ngrp <= sw && by0 >= 0 && by0 + nrows <= sh;
        const int r7 = tid / WT_GRPS, q = tid - r7 * WT_GRPS;    // constant divisor
        const bool colok = q < ngrp && r7 < WT_SROWS;
        if (interior) {
            uint32_t w0[WT_STAGE_IT], w1[WT_STAGE_IT], w2[WT_STAGE_IT];
            const uint32_t* gw = reinterpret_cast<const uint32_t*>(src + (size_t)(by0 + r7) * sstride + (size_t)(3 * ax0)) + 3 * q;
            const size_t gstep = (size_t)WT_SROWS * sstride / 4;
#pragma unroll
            for (int k = 0; k < WT_STAGE_IT; ++k) {
                if (colok && r7 + k * WT_SROWS < nrows) {
                    w0[k] = __ldg(gw); w1[k] = __ldg(gw + 1); w2[k] = __ldg(gw + 2);
                }
                gw += gstep;
            }
            uint32_t* d = S.src + r7 * WT_PITCH + 4 * q;
#pragma unroll
            for (int k = 0; k < WT_STAGE_IT; ++k) {
                if (colok && r7 + k * WT_SROWS < nrows) {
                    uint4 o;
                    o.x = w0[k];                                    // [B0 G0 R0 --]
                    o.y = __byte_perm(w0[k], w1[k], 0x0543);        // [B1 G1 R1 --]
                    o.z = __byte_perm(w1[k], w2[k], 0x0432);        // [B2 G2 R2 --]
                    o.w = w2[k] >> 8;                               // [B3 G3 R3 --]
                    *reinterpret_cast<uint4*>(d) = o;
                }
                d += WT_SROWS * WT_PITCH;
            }
        } else if (colok) {
            // tiles touching the frame border (or unaligned frames): per-pixel, zero outside = BORDER_CONSTANT
            for (int r = r7; r < nrows; r += WT_SROWS) {
                const int gx = ax0 + 4 * q, gy = by0 + r;
                uint32_t v[4] = {0u, 0u, 0u, 0u};
                if ((unsigned)gy < (unsigned)sh) {
                    const uint8_t* g = src + (size_t)gy * sstride + 3 * gx;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if ((unsigned)(gx + c) < (unsigned)sw) {
                            const uint8_t* pp = g + 3 * c;
                            v[c] = (uint32_t)pp[0] | ((uint32_t)pp[1] << 8) | ((uint32_t)pp[2] << 16);
                        }
                    }
                }
                *reinterpret_cast<uint4*>(S.src + r * WT_PITCH + 4 * q) = make_uint4(v[0], v[1], v[2], v[3]);
            }
        }
    }
    __syncthreads();
    // ---- compute: warp -> rows, lane -> pixels x0 + lane + 32 j
    int ad[4], bd[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int2 c = S.colAB[lane + 32 * j];
        ad[j] = c.x - (ax0 << 10);                               // fold the box origin into the fixed-point terms
        bd[j] = c.y - (by0 << 10);
    }
    const int tw = min(WT_W, dw - x0);
    const bool vec_out = dst_vec && tw == WT_W;
    uint32_t* const orow = S.out + warp * WT_W;
    for (int rr = warp; rr < WT_H; rr += WT_THREADS / 32) {
        if (y0 + rr >= dh) break;
        const int2 xy = S.rowXY[rr];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int t1 = xy.x + ad[j], t2 = xy.y + bd[j];
            const int sx = t1 >> 10, sy = t2 >> 10;              // box-relative integer source coordinates
            const uint32_t ax = (uint32_t)(t1 >> 5) & 31u;
            const uint32_t wy1 = ((uint32_t)t2 << 1) & 0x7c0u;   // ay * 64
            const uint32_t wy0 = 2048u - wy1;                    // (32 - ay) * 64
            const uint32_t* p = S.src + sy * WT_PITCH + sx;
            const uint32_t t00 = p[0], t01 = p[1], t10 = p[WT_PITCH], t11 = p[WT_PITCH + 1];
            const uint32_t Wa = __byte_perm(32u - ax, ax, 0x7740);   // bytes [32-ax, ax, 0, 0]
            const uint32_t Wb = Wa << 16;                            // bytes [0, 0, 32-ax, ax]
            const uint32_t u0 = __byte_perm(t00, t01, 0x5140);   // [b00 b01 g00 g01]
            const uint32_t u1 = __byte_perm(t00, t01, 0x5162);   // [r00 r01 ..]
            const uint32_t l0 = __byte_perm(t10, t11, 0x5140);
            const uint32_t l1 = __byte_perm(t10, t11, 0x5162);
            const uint32_t hb0 = __dp4a(u0, Wa, 0u), hg0 = __dp4a(u0, Wb, 0u), hr0 = __dp4a(u1, Wa, 0u);
            const uint32_t hb1 = __dp4a(l0, Wa, 0u), hg1 = __dp4a(l0, Wb, 0u), hr1 = __dp4a(l1, Wa, 0u);
            const uint32_t vb = hb0 * wy0 + (hb1 * wy1 + 32768u);     // (acc + 512) << 6 : result in byte 2
            const uint32_t vg = hg0 * wy0 + (hg1 * wy1 + 32768u);
            const uint32_t vr = hr0 * wy0 + (hr1 * wy1 + 32768u);
            orow[lane + 32 * j] = __byte_perm(__byte_perm(vb, vg, 0x0062), vr, 0x0610);   // [B G R --]
        }
        __syncwarp();
        // ---- the warp writes its finished row once: lane -> 4 pixels -> 12 packed bytes, a warp stores
        //      384 contiguous bytes
        uint8_t* grow = dst + (size_t)(y0 + rr) * dstride + (size_t)x0 * 3;
        if (vec_out) {
            const uint4 v = *reinterpret_cast<const uint4*>(orow + 4 * lane);
            uint32_t* g = reinterpret_cast<uint32_t*>(grow) + 3 * lane;
            g[0] = __byte_perm(v.x, v.y, 0x4210);            // B0 G0 R0 B1
            g[1] = __byte_perm(v.y, v.z, 0x5421);            // G1 R1 B2 G2
            g[2] = __byte_perm(v.z, v.w, 0x6542);            // R2 B3 G3 R3
        } else {
            for (int c = lane; c < tw; c += 32) {
                const uint32_t v = orow[c];
                grow[3 * c] = (uint8_t)v; grow[3 * c + 1] = (uint8_t)(v >> 8); grow[3 * c + 2] = (uint8_t)(v >> 16);
            }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(WT_THREADS, 6) k_warp_tiled_lanes(const LaneDev* __restrict__ lanes, PtrPack src, MutPtrPack dst,
                                                                  WarpGeom g, int src_vec, int dst_vec) {
    __shared__ WarpTileSmem S;
    warp_tile(S, src.p[blockIdx.z], g.src_w, g.src_h, g.src_stride, dst.p[blockIdx.z], g.out_w, g.out_h, g.out_stride,
              lanes[blockIdx.z].wp->m, src_vec != 0, dst_vec != 0);
}

__global__ void __launch_bounds__(WT_THREADS, 6) k_warp_tiled_frames(const uint8_t* __restrict__ src, int sw, int sh, size_t sstride,
                                                                   size_t sframe, uint8_t* __restrict__ dst, int dw, int dh,
                                                                   size_t dstride, size_t dframe,
                                                                   const WarpParams* __restrict__ wps, int src_vec, int dst_vec) {
    __shared__ WarpTileSmem S;
    warp_tile(S, src + blockIdx.z * sframe, sw, sh, sstride, dst + blockIdx.z * dframe, dw, dh, dstride,
              wps[blockIdx.z].m, src_vec != 0, dst_vec != 0);
}

static inline bool vec_ok(const void* p, size_t stride, int a) { return ((uintptr_t)p % a == 0) && (stride % a == 0); }

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

static void launch_warp_plain(const LaneDev* lanes, int n_lanes, const PtrPack& src, const MutPtrPack& dst, const WarpGeom& g,
                              cudaStream_t st) {
    bool sv = true, dv = true;
    for (int i = 0; i < n_lanes; ++i) {
        sv = sv && vec_ok(src.p[i], g.src_stride, 4);
        dv = dv && vec_ok(dst.p[i], g.out_stride, 4);
    }
    dim3 grid((g.out_w + WT_W - 1) / WT_W, (g.out_h + WT_H - 1) / WT_H, n_lanes);
    k_warp_tiled_lanes<<<grid, WT_THREADS, 0, st>>>(lanes, src, dst, g, sv ? 1 : 0, dv ? 1 : 0);
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
        launch_warp_plain(lanes, n_lanes, src, tmp, g1, st);
        int b = g.border, cw = g.src_w - 2 * b, ch = g.src_h - 2 * b;
        for (int i = 0; i < n_lanes; ++i)
            launch_resize_linear(scratch[i] + (size_t)b * g1.out_stride + 3 * b, cw, ch, g1.out_stride, 3,
                                 dst.p[i], g.out_w, g.out_h, g.out_stride, st);
        return;
    }
    if (g.mode == 1) {
        dim3 grid((g.out_w + 63) / 64, (g.out_h + 3) / 4, n_lanes);
        k_warp_lanes<true><<<grid, 256, 0, st>>>(lanes, src, dst, g);
    } else {
        launch_warp_plain(lanes, n_lanes, src, dst, g, st);
    }
}

void launch_warp_matrices(const uint8_t* src, int sw, int sh, size_t sstride, size_t sframe,
                          uint8_t* dst, int dw, int dh, size_t dstride, size_t dframe,
                          const WarpParams* d_wp, int n_frames, cudaStream_t st) {
    const bool sv = vec_ok(src, sstride, 4) && sframe % 4 == 0, dv = vec_ok(dst, dstride, 4) && dframe % 4 == 0;
    dim3 grid((dw + WT_W - 1) / WT_W, (dh + WT_H - 1) / WT_H, n_frames);
    k_warp_tiled_frames<<<grid, WT_THREADS, 0, st>>>(src, sw, sh, sstride, sframe, dst, dw, dh, dstride, dframe, d_wp,
                                                     sv ? 1 : 0, dv ? 1 : 0);
}

__global__ void __launch_bounds__(256) k_warp_frames_border(const uint8_t* __restrict__ src, int sw, int sh, size_t sstride,
                                                             size_t sframe, uint8_t* __restrict__ dst, int dw, int dh,
                                                             size_t dstride, size_t dframe, const WarpParams* __restrict__ wps,
                                                             int border, int border_mode) {
    int x = blockIdx.x * 64 + (threadIdx.x & 63);
    int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= dw || y >= dh) return;
    uint8_t* o = dst + blockIdx.z * dframe + (size_t)y * dstride + 3 * x;
    warp_pixel<true>(src + blockIdx.z * sframe, sw, sh, sstride, wps[blockIdx.z].m, border, border_mode, x, y, o);
}

void launch_warp_frames_mode(const uint8_t* src, int sw, int sh, size_t sstride, size_t sframe, uint8_t* dst,
                             size_t dstride, size_t dframe, const WarpParams* d_wp, int n_frames, int mode, int border,
                             int border_mode, uint8_t* scratch, cudaStream_t st) {
    if (n_frames <= 0) return;
    if (mode == 0) {
        launch_warp_matrices(src, sw, sh, sstride, sframe, dst, sw, sh, dstride, dframe, d_wp, n_frames, st);
    } else if (mode == 1) {
        const int dw = sw + 2 * border, dh = sh + 2 * border;
        dim3 grid((dw + 63) / 64, (dh + 3) / 4, n_frames);
        k_warp_frames_border<<<grid, 256, 0, st>>>(src, sw, sh, sstride, sframe, dst, dw, dh, dstride, dframe, d_wp, border, border_mode);
    } else {
        // crop+zoom: warp each frame into the scratch frame, then cv::resize the (b,b,w-2b,h-2b) crop
        const size_t tight = (size_t)sw * 3;
        for (int i = 0; i < n_frames; ++i) {
            launch_warp_matrices(src + i * sframe, sw, sh, sstride, sframe, scratch, sw, sh, tight, tight * sh, d_wp + i, 1, st);
            launch_resize_linear(scratch + (size_t)border * tight + 3 * border, sw - 2 * border, sh - 2 * border, tight, 3,
                                 dst + i * dframe, sw, sh, dstride, st);
        }
    }
}
