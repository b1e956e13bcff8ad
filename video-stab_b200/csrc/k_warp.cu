// k_warp.cu — the output stage: cv::copyMakeBorder (Stabilizer.cpp:981-990) + cv::warpAffine
// INTER_LINEAR / BORDER_CONSTANT (:1056-1060) + crop-and-zoom (:1108-1124), writing the output once.
// Specification: oracle/cv_models.py warp_affine / copy_make_border / resize_linear (bit-exact vs
// cv2 4.13).  Pure integer arithmetic: 10-bit fixed-point source coordinates, 5-bit sub-pixel
// position, 15-bit bilinear weights.  HBM-bound: algorithmic traffic is one read + one write of the
// frame (2*3*W*H bytes).
#include "kernels.h"
#include <climits>
#include <cuda.h>

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

// One output pixel (3 channels) as a word [B G R 0].  BORDER: the source is the virtual (w+2b)x(h+2b) bordered frame.
template <bool BORDER>
static __device__ __forceinline__ uint32_t warp_pixel_word(const uint8_t* __restrict__ src, int w, int h, size_t stride,
                                                           const double* __restrict__ m, int b, int bmode, int x, int y) {
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
    return (uint32_t)((acc0 + 512) >> 10) | ((uint32_t)((acc1 + 512) >> 10) << 8) | ((uint32_t)((acc2 + 512) >> 10) << 16);
}
template <bool BORDER>
static __device__ __forceinline__ void warp_pixel(const uint8_t* __restrict__ src, int w, int h, size_t stride,
                                                  const double* __restrict__ m, int b, int bmode, int x, int y,
                                                  uint8_t* __restrict__ out) {
    const uint32_t v = warp_pixel_word<BORDER>(src, w, h, stride, m, b, bmode, x, y);
    out[0] = (uint8_t)v; out[1] = (uint8_t)(v >> 8); out[2] = (uint8_t)(v >> 16);
}

// ------------------------------------------------------------------------------------------------
// Tiled fast path (plain warp, mode 0).  One CTA walks a vertical strip of up to WT_MAXT 128x32 output
// tiles.  The kernel is ISSUE-bound, not HBM-bound (profiles/r01_b_summary.md), so everything below is
// arranged to cut instructions per output pixel:
//   0. per strip: the fixed-point column terms (adelta,bdelta) and row terms (X0,Y0) of cv::warpAffine are
//      computed once, in double, one entry per thread, and kept in shared memory / registers;
//   1. per tile: one thread derives the exact source bounding box from the four corner coordinates (the
//      fixed-point coordinate is monotone in x and in y), widened to 4-pixel (12-byte) groups;
//   2. the box is staged with coalesced 32-bit loads and re-packed to 4-byte words [B G R R'] where R' is
//      the red of the NEXT pixel (so the red pair of a bilinear tap is already adjacent), one conflict-free
//      STS.128 per 4 pixels; positions outside the frame are staged as zeros, which IS cv::BORDER_CONSTANT(0).
//      The shared row pitch is exactly 1024 bytes, so the row offset of a tap is (t2 & ~1023): no multiply;
//   3. lane l produces pixels x0+l+32j: 4 aligned LDS.32 taps, two PRMT, the horizontal pass as six
//      IDP.2A with ONE 16-bit weight pair (16384-512ax | 512ax), the vertical pass as six IMAD with weights
//      (1024-32ay, 32ay).  Total scale 2^24, so the rounded 8-bit result is byte 3 of the accumulator:
//      ((acc + 512) >> 10 without a shift; exactness argument in DESIGN.md);
//   4. each warp transposes its finished row through 512 bytes of shared memory and writes it once with
//      coalesced 32-bit stores (a warp stores 384 contiguous bytes).
// Tiles whose source box does not fit (large rotations / scales) or whose coordinates approach the int16
// saturation of cv::remap fall back to the per-pixel path, CTA-uniformly.
#define WT_W 128
#define WT_H 32
#define WT_PITCHW 256                // shared row pitch in words (1024 bytes)
#define WT_ROWS 42                   // source-tile row capacity (42 KB)
#define WT_THREADS 256
#define WT_GRPS 36                   // 4-pixel column groups of the staged box (box width <= 144)
#define WT_SROWS 7                   // staging rows in flight: 36 x 7 = 252 threads
#define WT_STAGE_IT 6                // staging tasks per thread (7 * 6 = 42 = WT_ROWS)
#define WT_MAXT 4                    // tiles per CTA strip (TMA kernel; 8 measured no faster: tail imbalance)
#define WT_MAXT_F 4                  // tiles per CTA strip (fallback kernel: static shared memory limit)

struct WarpTileSmem {
    uint32_t src[WT_ROWS * WT_PITCHW];        // must stay first: 1024-byte aligned (the row offset trick)
    uint32_t out[(WT_THREADS / 32) * WT_W];   // one BGRx output row per warp (4 KB); aliased by colAB in the strip prologue
    int2 rowXY[WT_H * WT_MAXT_F];
    int box[2][8];                            // ax0, by0, ngrp, nrows, ok, interior
};

// shared-memory accesses with explicit 32-bit addresses (the tap address is built with integer tricks)
template <int OFF>
static __device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF>
static __device__ __forceinline__ void sts32(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0+%1], %2;" :: "r"(a), "n"(OFF), "r"(v) : "memory");
}
template <int OFF>
static __device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.u32 [%0+%1], {%2, %3, %4, %5};" :: "r"(a), "n"(OFF), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// 16 source bytes (4 pixels + the next pixel's red) -> 4 words [B G R R']
//   bytes: w0 = B0 G0 R0 B1 | w1 = G1 R1 B2 G2 | w2 = R2 B3 G3 R3 | w3 = B4 G4 R4 ..
static __device__ __forceinline__ uint4 repack_bgrr(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
    uint4 o;
    o.x = __byte_perm(w0, w1, 0x5210);                                   // [B0 G0 R0 R1]
    o.y = __byte_perm(__byte_perm(w0, w1, 0x0543), w2, 0x4210);          // [B1 G1 R1 R2]
    o.z = __byte_perm(w1, w2, 0x7432);                                   // [B2 G2 R2 R3]
    o.w = __byte_perm(w2, w3, 0x6321);                                   // [B3 G3 R3 R4]
    return o;
}

// one staging task of a tile that touches the frame border: positions gx0..gx0+3 of source row gy
static __device__ __noinline__ uint4 stage_edge_task(const uint8_t* __restrict__ src, int sw, int sh, size_t sstride,
                                                     int gx0, int gy, bool src_vec) {
    if ((unsigned)gy >= (unsigned)sh || gx0 + 4 < 0 || gx0 >= sw) return make_uint4(0u, 0u, 0u, 0u);   // BORDER_CONSTANT 0
    const uint8_t* row = src + (size_t)gy * sstride;
    // the 16-byte read may spill into the neighbouring row (masked below); it must stay inside the frame buffer
    const bool legal = src_vec && !(gy == 0 && gx0 < 0) && !(gy == sh - 1 && gx0 + 6 > sw);
    if (legal) {
        const uint32_t* gw = reinterpret_cast<const uint32_t*>(row + 3 * gx0);
        uint4 o = repack_bgrr(__ldg(gw), __ldg(gw + 1), __ldg(gw + 2), __ldg(gw + 3));
        if (gx0 < 0 || gx0 + 5 > sw) {
            uint32_t v[5];
#pragma unroll
            for (int c = 0; c < 5; ++c) v[c] = ((unsigned)(gx0 + c) < (unsigned)sw) ? 0xFFFFFFFFu : 0u;
            o.x &= (v[0] & 0x00FFFFFFu) | (v[1] & 0xFF000000u);
            o.y &= (v[1] & 0x00FFFFFFu) | (v[2] & 0xFF000000u);
            o.z &= (v[2] & 0x00FFFFFFu) | (v[3] & 0xFF000000u);
            o.w &= (v[3] & 0x00FFFFFFu) | (v[4] & 0xFF000000u);
        }
        return o;
    }
    uint32_t v[5] = {0u, 0u, 0u, 0u, 0u};
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        if ((unsigned)(gx0 + c) < (unsigned)sw) {
            const uint8_t* pp = row + 3 * (gx0 + c);
            v[c] = (uint32_t)pp[0] | ((uint32_t)pp[1] << 8) | ((uint32_t)pp[2] << 16);
        }
    }
    return make_uint4(v[0] | ((v[1] >> 16) << 24), v[1] | ((v[2] >> 16) << 24), v[2] | ((v[3] >> 16) << 24),
                      v[3] | ((v[4] >> 16) << 24));
}

static __device__ __forceinline__ void warp_strip(WarpTileSmem& S, const uint8_t* __restrict__ src, int sw, int sh,
                                                  size_t sstride, uint8_t* __restrict__ dst, int dw, int dh,
                                                  size_t dstride, const double* __restrict__ m, int rows_per_cta,
                                                  bool src_vec, bool dst_vec) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * WT_W, ys = blockIdx.y * rows_per_cta;
    const int ye = min(ys + rows_per_cta, dh);
    int2* const colAB = reinterpret_cast<int2*>(S.out);

    // ---- 0. strip prologue: fixed-point column/row terms of cv::warpAffine, one per thread.  Columns beyond
    //         the frame reuse the last valid column so that their (discarded) taps stay inside the staged box.
    if (tid < WT_W) {
        const double xd = (double)min(x0 + tid, dw - 1);
        colAB[tid] = make_int2(sat_int(m[0] * xd * 1024.0), sat_int(m[3] * xd * 1024.0));
    } else {
        for (int r = tid - WT_W; r < ye - ys; r += WT_THREADS - WT_W) {
            const double yd = (double)(ys + r);
            S.rowXY[r] = make_int2(sat_int((m[1] * yd + m[2]) * 1024.0) + 16, sat_int((m[4] * yd + m[5]) * 1024.0) + 16);
        }
    }
    __syncthreads();
    uint32_t adT[4], bd[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int2 c = colAB[lane + 32 * j];
        adT[j] = (uint32_t)c.x * 16u;                            // 14 fractional bits: 512*ax sits at bits 9..13
        bd[j] = (uint32_t)c.y;
    }
    const int tw = min(WT_W, dw - x0);
    const int2 cA = colAB[0], cB = colAB[tw - 1];
    const bool col_ok = max(max(abs(cA.x), abs(cB.x)), max(abs(cA.y), abs(cB.y))) < (1 << 26);
    const bool vec_out = dst_vec && tw == WT_W;
    const uint32_t s_src = (uint32_t)__cvta_generic_to_shared(S.src);          // multiple of 1024
    const uint32_t s_px = (uint32_t)__cvta_generic_to_shared(S.out + warp * WT_W + lane);        // pixel lane + 32 j at +128 j bytes
    const uint32_t s_v4 = (uint32_t)__cvta_generic_to_shared(S.out + warp * WT_W + 4 * lane);    // pixels 4 lane .. 4 lane + 3
    // staging thread map: 36 column groups x 7 rows
    const int r7 = tid / WT_GRPS, q = tid - r7 * WT_GRPS;
    const uint32_t s_stage = s_src + (uint32_t)(r7 * WT_PITCHW + 4 * q) * 4u;

    int buf = 0;
    for (int y0 = ys; y0 < ye; y0 += WT_H, buf ^= 1) {
        // ---- 1. source box of this tile (thread 0); the barrier also orders the previous tile's reads of
        //         S.src / colAB before this tile's staging
        if (tid == 0) {
            const int yb = min(y0 + WT_H, ye) - 1;
            const int2 rA = S.rowXY[y0 - ys], rB = S.rowXY[yb - ys];
            int minx = INT_MAX, maxx = INT_MIN, miny = INT_MAX, maxy = INT_MIN;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int2 cc = (c & 1) ? cB : cA, rr = (c & 2) ? rB : rA;
                const int X = (rr.x + cc.x) >> 10, Y = (rr.y + cc.y) >> 10;
                minx = min(minx, X); maxx = max(maxx, X); miny = min(miny, Y); maxy = max(maxy, Y);
            }
            const int ax0 = minx & ~3;                           // floor to a 4-pixel group (also for negatives)
            const int ngrp = (maxx + 1 - ax0) / 4 + 1;
            const int nrows = maxy + 2 - miny;
            const bool row_ok = max(max(abs(rA.x), abs(rB.x)), max(abs(rA.y), abs(rB.y))) < (1 << 26);
            const bool ok = col_ok && row_ok && minx > -30000 && maxx < 30000 && miny > -30000 && maxy < 30000 &&
                            ngrp <= WT_GRPS && nrows <= WT_ROWS && ngrp > 0 && nrows > 0;
            // interior: every 16-byte task of the box (12 bytes + the next pixel's red) lies inside its frame row
            const bool interior = src_vec && ax0 >= 0 && ax0 + 4 * ngrp + 2 <= sw && miny >= 0 && miny + nrows <= sh;
            int* bx = S.box[buf];
            bx[0] = ax0; bx[1] = miny; bx[2] = ngrp; bx[3] = nrows; bx[4] = ok ? 1 : 0; bx[5] = interior ? 1 : 0;
        }
        __syncthreads();
        const int ax0 = S.box[buf][0], by0 = S.box[buf][1], ngrp = S.box[buf][2], nrows = S.box[buf][3];
        if (!S.box[buf][4]) {
            // generic per-pixel path for this tile
            for (int i = tid; i < WT_W * WT_H; i += WT_THREADS) {
                const int x = x0 + (i & (WT_W - 1)), y = y0 + i / WT_W;
                if (x < dw && y < ye) warp_pixel<false>(src, sw, sh, sstride, m, 0, 0, x, y, dst + (size_t)y * dstride + 3 * x);
            }
            continue;
        }
        // ---- 2. stage the source box: one task = 4 positions = 16 source bytes -> one 16-byte store.  Each
        //         thread walks down its column group 7 rows at a time; all its loads are issued before the
        //         first is consumed.
        {
            const bool colok = q < ngrp && r7 < WT_SROWS;
            if (S.box[buf][5]) {
                // Loads are unconditional from clamped (always valid) addresses, so only the store is predicated:
                // idle lanes re-read the last group / last row of the box (L1 hits).  Two batches of three rows:
                // 12 loads in flight per thread, then their three 16-byte stores.
                const uint8_t* tbase = src + (size_t)by0 * sstride + (size_t)(3 * ax0 + 12 * min(q, ngrp - 1));
                const unsigned stride32 = (unsigned)sstride, rlast = (unsigned)(nrows - 1);
#define WT_BATCH(k0)                                                                                           \
                if ((k0) * WT_SROWS < nrows) {                          /* CTA-uniform */                      \
                    uint32_t w0[3], w1[3], w2[3], w3[3];                                                       \
                    _Pragma("unroll") for (int k = 0; k < 3; ++k) {                                            \
                        const unsigned r = min((unsigned)(r7 + ((k0) + k) * WT_SROWS), rlast);                 \
                        const uint32_t* gw = reinterpret_cast<const uint32_t*>(tbase + (size_t)r * stride32);  \
                        w0[k] = __ldg(gw); w1[k] = __ldg(gw + 1); w2[k] = __ldg(gw + 2); w3[k] = __ldg(gw + 3); \
                    }                                                                                          \
                    _Pragma("unroll") for (int k = 0; k < 3; ++k) {                                            \
                        const uint4 o = repack_bgrr(w0[k], w1[k], w2[k], w3[k]);                               \
                        if (colok && r7 + ((k0) + k) * WT_SROWS < nrows) {                                     \
                            if (k == 0) sts128<((k0) + 0) * WT_SROWS * WT_PITCHW * 4>(s_stage, o.x, o.y, o.z, o.w); \
                            if (k == 1) sts128<((k0) + 1) * WT_SROWS * WT_PITCHW * 4>(s_stage, o.x, o.y, o.z, o.w); \
                            if (k == 2) sts128<((k0) + 2) * WT_SROWS * WT_PITCHW * 4>(s_stage, o.x, o.y, o.z, o.w); \
                        }                                                                                      \
                    }                                                                                          \
                }
                WT_BATCH(0) WT_BATCH(3)
#undef WT_BATCH
            } else if (colok) {
                // tiles touching the frame border: zeros outside = BORDER_CONSTANT
                for (int r = r7; r < nrows; r += WT_SROWS) {
                    const uint4 o = stage_edge_task(src, sw, sh, sstride, ax0 + 4 * q, by0 + r, src_vec);
                    *reinterpret_cast<uint4*>(S.src + r * WT_PITCHW + 4 * q) = o;
                }
            }
        }
        __syncthreads();
        // ---- 3. compute: warp -> rows, lane -> pixels x0 + lane + 32 j
        const uint32_t bx14 = (uint32_t)ax0 << 14, by10 = ((uint32_t)by0 << 10) - s_src;
        uint8_t* grow = dst + (size_t)(y0 + warp) * dstride + (size_t)x0 * 3 + (vec_out ? 12 * lane : 0);
#pragma unroll
        for (int rr = warp; rr < WT_H; rr += WT_THREADS / 32) {
            const int y = y0 + rr;
            if (y >= ye) break;
            const int2 xy = S.rowXY[y - ys];
            const uint32_t rxT = (uint32_t)xy.x * 16u - bx14;
            const uint32_t ryS = (uint32_t)xy.y - by10;          // box-relative y (10 fractional bits) + shared base of S.src
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t T1 = rxT + adT[j];                     // box-relative x, 14 fractional bits
                const uint32_t t2 = ryS + bd[j];
                uint32_t a;                                           // &S.src[sy][sx] = (t2 & ~1023) + 4 * (T1 >> 14)
                asm("mad.lo.u32 %0, %1, 4, %2;" : "=r"(a) : "r"(T1 >> 14), "r"(t2 & 0xFFFFFC00u));
                const uint32_t A16 = T1 & 0x3E00u;                    // 512 * ax
                const uint32_t W16 = A16 * 0xFFFFu + 16384u;          // u16 pair (16384 - 512 ax | 512 ax)
                const uint32_t B = t2 & 0x3E0u, Bc = 1024u - B;       // 32 * ay, 32 * (32 - ay)
                const uint32_t t00 = lds32<0>(a), t01 = lds32<4>(a), t10 = lds32<WT_PITCHW * 4>(a), t11 = lds32<WT_PITCHW * 4 + 4>(a);
                const uint32_t u0 = __byte_perm(t00, t01, 0x5140);    // [b00 b01 g00 g01]; the red pair is bytes 2,3 of t00
                const uint32_t l0 = __byte_perm(t10, t11, 0x5140);
                const uint32_t hb0 = __dp2a_lo(W16, u0, 0u), hg0 = __dp2a_hi(W16, u0, 0u), hr0 = __dp2a_hi(W16, t00, 0u);
                const uint32_t hb1 = __dp2a_lo(W16, l0, 0u), hg1 = __dp2a_hi(W16, l0, 0u), hr1 = __dp2a_hi(W16, t10, 0u);
                const uint32_t vb = hb0 * Bc + (hb1 * B + 0x800000u);  // (acc + 512) << 14 : result in byte 3
                const uint32_t vg = hg0 * Bc + (hg1 * B + 0x800000u);
                const uint32_t vr = hr0 * Bc + (hr1 * B + 0x800000u);
                const uint32_t px = __byte_perm(__byte_perm(vb, vg, 0x0073), vr, 0x0710);   // [B G R --]
                if (j == 0) sts32<0>(s_px, px);
                else if (j == 1) sts32<128>(s_px, px);
                else if (j == 2) sts32<256>(s_px, px);
                else sts32<384>(s_px, px);
            }
            __syncwarp();
            // ---- 4. the warp writes its finished row once: lane -> 4 pixels -> 12 packed bytes
            if (vec_out) {
                uint32_t vx, vy, vz, vw;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(vx), "=r"(vy), "=r"(vz), "=r"(vw) : "r"(s_v4));
                uint32_t* g = reinterpret_cast<uint32_t*>(grow);
                g[0] = __byte_perm(vx, vy, 0x4210);              // B0 G0 R0 B1
                g[1] = __byte_perm(vy, vz, 0x5421);              // G1 R1 B2 G2
                g[2] = __byte_perm(vz, vw, 0x6542);              // R2 B3 G3 R3
            } else {
                const uint32_t* orow = S.out + warp * WT_W;
                for (int c = lane; c < tw; c += 32) {
                    const uint32_t v = orow[c];
                    grow[3 * c] = (uint8_t)v; grow[3 * c + 1] = (uint8_t)(v >> 8); grow[3 * c + 2] = (uint8_t)(v >> 16);
                }
            }
            grow += (size_t)(WT_THREADS / 32) * dstride;
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// TMA variant of the tiled fast path (the default whenever the frame geometry allows a tensor map: 16-byte
// aligned base and strides, width % 4 == 0).  Differences from warp_strip above:
//   * the raw source box (packed BGR, 480 bytes x 42 rows) is fetched by ONE cp.async.bulk.tensor.3d issued by
//     thread 0 into shared memory, completion on an mbarrier; the tensor map is over {row words, rows, frames}
//     and TMA zero-fills everything outside the frame, which IS cv::BORDER_CONSTANT(0) - there is no edge path;
//   * the fetch of tile t+1 is issued right after tile t has been re-packed, so it overlaps tile t's compute;
//   * the re-pack to [B G R R'] words reads the raw box from shared memory (stride-12-byte LDS.32, conflict
//     free), so no global load instruction, address arithmetic or predicate is left in the kernel.
#define WT_RAW_PITCH 480                 // bytes per raw row: 3 * (12 + 4 * WT_GRPS + 1) = 471 rounded to 16; TMA needs a 16-byte
                                         // aligned start, i.e. a 16-pixel aligned box origin (48 bytes), hence the 12 extra pixels
#define WT_RAW_WORDS (WT_RAW_PITCH / 4)
#define WT_TPITCH 640                    // shared row pitch of the re-packed box in the TMA kernel (160 words; 4 CTAs/SM)
#define WT_TMA_SMEM (WT_ROWS * WT_TPITCH + WT_ROWS * WT_RAW_PITCH + (WT_THREADS / 32) * WT_W * 4 + WT_H * WT_MAXT * 8 + 64 + 16)
#define WT_TMA_MAXPACK 8
#define WT_TGRPS 40                      // re-pack thread map: threads per row (WT_GRPS of them active)
#define WT_TSROWS 6                      // rows in flight: 40 x 6 = 240 threads, 7 iterations cover WT_ROWS

struct TmapPack {
    CUtensorMap m[WT_TMA_MAXPACK];
};

static __device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(mbar), "r"(parity) : "memory");
}

// lanes_mode: one 2-D-like map per lane (maps[z], frame coordinate 0); else one map, frame coordinate z
// BORDER: the source is the virtual cv::copyMakeBorder frame (border_ pixels of margin, mode bmode_)
template <bool BORDER>
__global__ void __launch_bounds__(WT_THREADS, 4)
k_warp_tma(const __grid_constant__ TmapPack pack, const CUtensorMap* __restrict__ dmaps, int lanes_mode,
           const LaneDev* __restrict__ lanes, const WarpParams* __restrict__ wps,
           PtrPack srcp, const uint8_t* __restrict__ src0, size_t sframe, size_t sstride, int sw, int sh,
           MutPtrPack dstp, uint8_t* __restrict__ dst0, size_t dframe, int dw, int dh, size_t dstride,
           int rows_per_cta, int dst_vec, int wp_slot, int border_, int bmode_) {
    const int border = BORDER ? border_ : 0, bmode = BORDER ? bmode_ : 0;
    extern __shared__ __align__(1024) unsigned char wt_smem[];
    uint32_t* const S_src = reinterpret_cast<uint32_t*>(wt_smem);
    unsigned char* const S_raw = wt_smem + WT_ROWS * WT_TPITCH;
    uint32_t* const S_out = reinterpret_cast<uint32_t*>(S_raw + WT_ROWS * WT_RAW_PITCH);
    int2* const S_rowXY = reinterpret_cast<int2*>(S_out + (WT_THREADS / 32) * WT_W);
    int (*S_box)[8] = reinterpret_cast<int (*)[8]>(S_rowXY + WT_H * WT_MAXT);
    unsigned long long* const S_mbar = reinterpret_cast<unsigned long long*>(S_box + 2);

    const int z = blockIdx.z;
    const double* __restrict__ m = lanes_mode ? lanes[z].wpb[wp_slot]->m : wps[z].m;
    uint8_t* __restrict__ dst = lanes_mode ? dstp.p[z] : dst0 + (size_t)z * dframe;
    const CUtensorMap* tmap = dmaps ? dmaps + (lanes_mode ? z : 0) : &pack.m[lanes_mode ? z : 0];
    const int zc = lanes_mode ? 0 : z;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * WT_W, ys = blockIdx.y * rows_per_cta;
    const int ye = min(ys + rows_per_cta, dh);
    int2* const colAB = reinterpret_cast<int2*>(S_out);
    const uint32_t s_src = (uint32_t)__cvta_generic_to_shared(S_src);          // multiple of 1024
    const uint32_t s_raw = (uint32_t)__cvta_generic_to_shared(S_raw);
    const uint32_t s_mbar = (uint32_t)__cvta_generic_to_shared(S_mbar);

    // ---- 0. strip prologue
    if (tid < WT_W) {
        const double xd = (double)min(x0 + tid, dw - 1);
        colAB[tid] = make_int2(sat_int(m[0] * xd * 1024.0), sat_int(m[3] * xd * 1024.0));
    } else {
        for (int r = tid - WT_W; r < ye - ys; r += WT_THREADS - WT_W) {
            const double yd = (double)(ys + r);
            S_rowXY[r] = make_int2(sat_int((m[1] * yd + m[2]) * 1024.0) + 16, sat_int((m[4] * yd + m[5]) * 1024.0) + 16);
        }
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s_mbar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (dmaps) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" :: "l"(tmap) : "memory");
    }
    __syncthreads();
    uint32_t adT[4], bd[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int2 c = colAB[lane + 32 * j];
        adT[j] = (uint32_t)c.x * 16u;
        bd[j] = (uint32_t)c.y;
    }
    const int tw = min(WT_W, dw - x0);
    const int2 cA = colAB[0], cB = colAB[tw - 1];
    const bool col_ok = max(max(abs(cA.x), abs(cB.x)), max(abs(cA.y), abs(cB.y))) < (1 << 26);
    const bool vec_out = dst_vec && tw == WT_W;
    const uint32_t s_px = (uint32_t)__cvta_generic_to_shared(S_out + warp * WT_W + lane);
    const uint32_t s_v4 = (uint32_t)__cvta_generic_to_shared(S_out + warp * WT_W + 4 * lane);
    // re-pack thread map: 40 threads per row (36 active column groups) x 6 rows.  40 = 8 * 5 keeps every
    // quarter-warp of the 16-byte stores inside one row, and 3 q + 120 r (the raw word bank of lane (q, r), raw
    // pitch 480 bytes) is collision-free across the two rows a warp can straddle: both accesses are conflict free.
    const int r7 = tid / WT_TGRPS, q = tid - r7 * WT_TGRPS;
    const uint32_t s_stage = s_src + (uint32_t)(r7 * WT_TPITCH + 16 * q);
    const uint32_t s_rawt = s_raw + (uint32_t)(r7 * WT_RAW_PITCH + 12 * q);

    // source box of the tile starting at row y0 -> S_box[b]; issues its fetch when the box fits (thread 0 only)
    auto box_and_fetch = [&](int y0, int b) {
        const int yb = min(y0 + WT_H, ye) - 1;
        const int2 rA = S_rowXY[y0 - ys], rB = S_rowXY[yb - ys];
        int minx = INT_MAX, maxx = INT_MIN, miny = INT_MAX, maxy = INT_MIN;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int2 cc = (c & 1) ? cB : cA, rr = (c & 2) ? rB : rA;
            const int X = (rr.x + cc.x) >> 10, Y = (rr.y + cc.y) >> 10;
            minx = min(minx, X); maxx = max(maxx, X); miny = min(miny, Y); maxy = max(maxy, Y);
        }
        // (minx..maxx, miny..maxy) are coordinates in the bordered frame (cv::copyMakeBorder, Stabilizer.cpp:981-990);
        // the frame itself starts at (border, border).  A zero margin is what TMA fills in anyway; for the other
        // border modes only tiles whose taps all lie inside the frame take the staged path.
        const int rminx = minx - border, rmaxx = maxx - border, rminy = miny - border, rmaxy = maxy - border;
        const bool inside = border == 0 || bmode == 0 || (rminx >= 0 && rmaxx + 1 < sw && rminy >= 0 && rmaxy + 1 < sh);
        const int ax0 = rminx & ~3;
        const int ngrp = (rmaxx + 1 - ax0) / 4 + 1;
        const int nrows = maxy + 2 - miny;
        const bool row_ok = max(max(abs(rA.x), abs(rB.x)), max(abs(rA.y), abs(rB.y))) < (1 << 26);
        const bool ok = col_ok && row_ok && inside && minx > -30000 && maxx < 30000 && miny > -30000 && maxy < 30000 &&
                        ngrp <= WT_GRPS && nrows <= WT_ROWS && ngrp > 0 && nrows > 0;
        const int axT = rminx & ~15;                         // TMA box origin: 16 pixels = 48 bytes = 12 words
        int* bx = S_box[b];
        bx[0] = ax0 + border; bx[1] = miny; bx[2] = ngrp; bx[3] = nrows; bx[4] = ok ? 1 : 0; bx[5] = 3 * (ax0 - axT);
        if (ok) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s_mbar), "r"(WT_ROWS * WT_RAW_PITCH) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         :: "r"(s_raw), "l"(tmap), "r"(s_mbar), "r"(3 * (axT >> 2)), "r"(rminy), "r"(zc) : "memory");
        }
    };
    if (tid == 0) box_and_fetch(ys, 0);
    __syncthreads();

    int buf = 0;
    uint32_t phase = 0;
    for (int y0 = ys; y0 < ye; y0 += WT_H, buf ^= 1) {
        const int ax0 = S_box[buf][0], by0 = S_box[buf][1], nrows = S_box[buf][3];
        const bool ok = S_box[buf][4] != 0;
        const bool more = y0 + WT_H < ye;
        if (ok) {
            // ---- 2. wait for the raw box, re-pack it: one task = 16 raw bytes -> 4 words [B G R R']
            mbar_wait(s_mbar, phase);
            phase ^= 1;
            // No per-thread row / column test: TMA always delivers the full 42 x 480-byte box, so rows beyond nrows
            // and column groups beyond ngrp are re-packed too (never read as taps) — cheaper than the predicates.
            const uint32_t s_rawq = s_rawt + (uint32_t)S_box[buf][5];
            if (tid < WT_TGRPS * WT_TSROWS) {
#define WT_REPACK(k)                                                                                           \
                if ((k) * WT_TSROWS < nrows) {                           /* CTA-uniform */                      \
                    const uint32_t w0 = lds32<(k) * WT_TSROWS * WT_RAW_PITCH>(s_rawq);                          \
                    const uint32_t w1 = lds32<(k) * WT_TSROWS * WT_RAW_PITCH + 4>(s_rawq);                      \
                    const uint32_t w2 = lds32<(k) * WT_TSROWS * WT_RAW_PITCH + 8>(s_rawq);                      \
                    const uint32_t w3 = lds32<(k) * WT_TSROWS * WT_RAW_PITCH + 12>(s_rawq);                     \
                    const uint4 o = repack_bgrr(w0, w1, w2, w3);                                               \
                    sts128<(k) * WT_TSROWS * WT_TPITCH>(s_stage, o.x, o.y, o.z, o.w);                       \
                }
                WT_REPACK(0) WT_REPACK(1) WT_REPACK(2) WT_REPACK(3) WT_REPACK(4) WT_REPACK(5) WT_REPACK(6)
#undef WT_REPACK
            }
        }
        __syncthreads();                                   // S_src ready, raw box free
        if (tid == 0 && more) box_and_fetch(y0 + WT_H, buf ^ 1);      // overlaps this tile's compute
        if (!ok) {
            // generic per-pixel path for this tile
            for (int i = tid; i < WT_W * WT_H; i += WT_THREADS) {
                const int x = x0 + (i & (WT_W - 1)), y = y0 + i / WT_W;
                if (x < dw && y < ye) {
                    const uint8_t* sp = lanes_mode ? srcp.p[z] : src0 + (size_t)z * sframe;
                    warp_pixel<BORDER>(sp, sw, sh, sstride, m, border, bmode, x, y, dst + (size_t)y * dstride + 3 * x);
                }
            }
        } else {
            // ---- 3. compute: warp -> rows, lane -> pixels x0 + lane + 32 j
            const uint32_t bx14 = (uint32_t)ax0 << 14, by10 = (uint32_t)by0 << 10;
            uint8_t* grow = dst + (size_t)(y0 + warp) * dstride + (size_t)x0 * 3 + (vec_out ? 12 * lane : 0);
#pragma unroll
            for (int rr = warp; rr < WT_H; rr += WT_THREADS / 32) {
                const int y = y0 + rr;
                if (y >= ye) break;
                const int2 xy = S_rowXY[y - ys];
                const uint32_t rxT = (uint32_t)xy.x * 16u - bx14;
                const uint32_t ryS = (uint32_t)xy.y - by10;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t T1 = rxT + adT[j];
                    const uint32_t t2 = ryS + bd[j];
                    const uint32_t a = s_src + 4u * ((t2 >> 10) * (WT_TPITCH / 4) + (T1 >> 14));   // &S_src[sy][sx]
                    const uint32_t A16 = T1 & 0x3E00u;
                    const uint32_t W16 = A16 * 0xFFFFu + 16384u;
                    const uint32_t B = t2 & 0x3E0u, Bc = 1024u - B;
                    const uint32_t t00 = lds32<0>(a), t01 = lds32<4>(a), t10 = lds32<WT_TPITCH>(a), t11 = lds32<WT_TPITCH + 4>(a);
                    const uint32_t u0 = __byte_perm(t00, t01, 0x5140);
                    const uint32_t l0 = __byte_perm(t10, t11, 0x5140);
                    const uint32_t hb0 = __dp2a_lo(W16, u0, 0u), hg0 = __dp2a_hi(W16, u0, 0u), hr0 = __dp2a_hi(W16, t00, 0u);
                    const uint32_t hb1 = __dp2a_lo(W16, l0, 0u), hg1 = __dp2a_hi(W16, l0, 0u), hr1 = __dp2a_hi(W16, t10, 0u);
                    const uint32_t vb = hb0 * Bc + (hb1 * B + 0x800000u);
                    const uint32_t vg = hg0 * Bc + (hg1 * B + 0x800000u);
                    const uint32_t vr = hr0 * Bc + (hr1 * B + 0x800000u);
                    const uint32_t px = __byte_perm(__byte_perm(vb, vg, 0x0073), vr, 0x0710);
                    if (j == 0) sts32<0>(s_px, px);
                    else if (j == 1) sts32<128>(s_px, px);
                    else if (j == 2) sts32<256>(s_px, px);
                    else sts32<384>(s_px, px);
                }
                __syncwarp();
                if (vec_out) {
                    uint32_t vx, vy, vz, vw;
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(vx), "=r"(vy), "=r"(vz), "=r"(vw) : "r"(s_v4));
                    uint32_t* g = reinterpret_cast<uint32_t*>(grow);
                    g[0] = __byte_perm(vx, vy, 0x4210);
                    g[1] = __byte_perm(vy, vz, 0x5421);
                    g[2] = __byte_perm(vz, vw, 0x6542);
                } else {
                    const uint32_t* orow = S_out + warp * WT_W;
                    for (int c = lane; c < tw; c += 32) {
                        const uint32_t v = orow[c];
                        grow[3 * c] = (uint8_t)v; grow[3 * c + 1] = (uint8_t)(v >> 8); grow[3 * c + 2] = (uint8_t)(v >> 16);
                    }
                }
                grow += (size_t)(WT_THREADS / 32) * dstride;
                __syncwarp();
            }
        }
        if (more) __syncthreads();                         // S_src (and S_box[buf]) free for the next tile
    }
}

// ------------------------------------------------------------------------------------------------
// k_warp_quad: the tiled warp without the re-pack stage.  ncu on k_warp_tma (profiles/r01_e_warp_ncu.json) shows the
// shared-memory pipe (11.1 wavefronts per 32 output pixels) and the issue slots (48.7 instructions) saturating together;
// a third of both goes into turning the packed 3-byte pixels of the TMA box into 4-byte words and into carrying finished
// pixels through shared memory so that a warp can write whole rows.  Here a lane owns FOUR ADJACENT output pixels:
//   * their taps are 15 contiguous bytes per source row (5 pixels x 3 channels), i.e. five aligned 32-bit words of the RAW
//     box whatever the byte phase o = (3 sx) & 3; the channel pairs come out of those words with constant-selector PRMTs
//     (one code variant per o, chosen by a warp-uniform switch), so the box is used as TMA delivers it;
//   * the horizontal weight pair and the tap address are derived once per quad, the two middle taps of each row are shared
//     between neighbouring pixels, and the 12 packed output bytes leave as three 32-bit stores straight from registers;
//   * a warp covers 8 quads x 4 consecutive rows per step: raw word index 3 q + 120 row takes 32 distinct banks
//     (120 = 24 mod 32), so the ten tap loads of a step are conflict free without any padding.
// Per step (4 pixels per lane): 10 LDS + 16 PRMT + 24 IDP.2A + 24 IMAD + 8 PRMT (pack) + 3 PRMT + 3 STG + ~45 for the
// coordinates, votes and loop, against 4 x 36 in k_warp_tma plus its re-pack; measured 42.1 warp-instructions and 4.3
// shared-memory wavefronts per 32 output pixels against 48.7 and 11.1 (profiles/r02_b_warp_ncu.json).
// The vertical terms are lane constants while Y0 advances by exactly 4096 per 4 rows (checked per step, recomputed when
// not).  Exactness: every coordinate is the integer cv::warpAffine computes (X0(y) + adelta(x), Y0(y) + bdelta(x)); the
// quad path only requires what it then uses - px 0 and px 3 give consecutive sx (so the five words are the taps of all
// four pixels), one source row pair for the quad, one byte phase for the warp - and a warp whose step fails a test takes
// the per-pixel path for that step (same integers, byte loads from the same box).
#define WQ_BOXB (WT_ROWS * WT_RAW_PITCH)                 // bytes one TMA box delivers
#define WQ_RAWB ((WQ_BOXB + 127) / 128 * 128)            // raw buffer stride: TMA destinations are 128-byte aligned
#define WQ_SMEM (2 * WQ_RAWB + WT_W * 8 + WT_H * 8 * 8 + 2 * 32 + 2 * 8 + 8 * 4 + 64)

template <int OFF>
static __device__ __forceinline__ uint2 lds64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(a), "n"(OFF));
    return v;
}
static __device__ __forceinline__ uint32_t lds8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
template <int OFF>
static __device__ __forceinline__ uint32_t lds32v(uint32_t a) {            // not volatile: free to be scheduled with its neighbours
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
    return v;
}

// PRMT selectors for the pixel whose first byte sits at byte P of a run of aligned words: [b0 b1 g0 g1] and [. . r0 r1]
template <int P> struct QuadSel {
    static constexpr int s = P & 3, m = P >> 2;
    static constexpr int sr = (P + 2) & 3, mr = (P + 2) >> 2;
    static constexpr uint32_t bg = (uint32_t)(s | ((s + 3) << 4) | ((s + 1) << 8) | ((s + 4) << 12));
    static constexpr uint32_t r = (uint32_t)(sr | ((sr + 3) << 4) | (sr << 8) | ((sr + 3) << 12));
};

template <int P>
static __device__ __forceinline__ uint32_t quad_pixel(const uint32_t (&wt)[5], const uint32_t (&wb)[5], uint32_t W16, uint32_t B, uint32_t Bc) {
    typedef QuadSel<P> S;
    const uint32_t u0 = __byte_perm(wt[S::m], wt[S::m + 1], S::bg), r0 = __byte_perm(wt[S::mr], wt[S::mr + 1], S::r);
    const uint32_t l0 = __byte_perm(wb[S::m], wb[S::m + 1], S::bg), r1 = __byte_perm(wb[S::mr], wb[S::mr + 1], S::r);
    const uint32_t hb0 = __dp2a_lo(W16, u0, 0u), hg0 = __dp2a_hi(W16, u0, 0u), hr0 = __dp2a_hi(W16, r0, 0u);
    const uint32_t hb1 = __dp2a_lo(W16, l0, 0u), hg1 = __dp2a_hi(W16, l0, 0u), hr1 = __dp2a_hi(W16, r1, 0u);
    const uint32_t vb = hb0 * Bc + (hb1 * B + 0x800000u);
    const uint32_t vg = hg0 * Bc + (hg1 * B + 0x800000u);
    const uint32_t vr = hr0 * Bc + (hr1 * B + 0x800000u);
    return __byte_perm(__byte_perm(vb, vg, 0x0073), vr, 0x0710);
}

// four adjacent pixels from the five words of each of the two source rows at shared address a (byte phase O)
template <int O>
static __device__ __forceinline__ void quad_step(uint32_t a, const uint32_t (&W)[4], const uint32_t (&B)[4], const uint32_t (&Bc)[4], uint32_t (&px)[4]) {
    uint32_t wt[5], wb[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) wt[i] = lds32v<0>(a + 4u * i);
#pragma unroll
    for (int i = 0; i < 5; ++i) wb[i] = lds32v<0>(a + (uint32_t)WT_RAW_PITCH + 4u * i);
    px[0] = quad_pixel<O + 0>(wt, wb, W[0], B[0], Bc[0]);
    px[1] = quad_pixel<O + 3>(wt, wb, W[1], B[1], Bc[1]);
    px[2] = quad_pixel<O + 6>(wt, wb, W[2], B[2], Bc[2]);
    px[3] = quad_pixel<O + 9>(wt, wb, W[3], B[3], Bc[3]);
}

// A quad that straddles a change of source row (px 0..k-1 read rows R, R+1; px k..3 read rows R+1, R+2, or the other way
// round): every pixel loads its own two or three words per row from its own row pair; nothing is shared, the blend is the same.
template <int P>
static __device__ __forceinline__ uint32_t quad_pixel_own(uint32_t a, uint32_t W16, uint32_t B) {
    typedef QuadSel<P> S;
    uint32_t wt[5], wb[5];
#pragma unroll
    for (int i = S::m; i <= S::mr + 1; ++i) {
        wt[i] = lds32v<0>(a + 4u * i);
        wb[i] = lds32v<0>(a + (uint32_t)WT_RAW_PITCH + 4u * i);
    }
    return quad_pixel<P>(wt, wb, W16, B, 1024u - B);
}
template <int O>
static __device__ __forceinline__ void quad_step_split(uint32_t a, const uint32_t (&W)[4], const uint32_t (&B)[4], uint32_t smask, uint32_t (&px)[4]) {
    px[0] = quad_pixel_own<O + 0>(a + ((smask >> 0) & 1u) * (uint32_t)WT_RAW_PITCH, W[0], B[0]);
    px[1] = quad_pixel_own<O + 3>(a + ((smask >> 1) & 1u) * (uint32_t)WT_RAW_PITCH, W[1], B[1]);
    px[2] = quad_pixel_own<O + 6>(a + ((smask >> 2) & 1u) * (uint32_t)WT_RAW_PITCH, W[2], B[2]);
    px[3] = quad_pixel_own<O + 9>(a + ((smask >> 3) & 1u) * (uint32_t)WT_RAW_PITCH, W[3], B[3]);
}
// bitwise c ? b : a
static __device__ __forceinline__ uint32_t bitsel(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

#define WQ_MAXT 8                         // tiles per CTA strip
template <bool BORDER>
__global__ void __launch_bounds__(WT_THREADS, 4)
k_warp_quad(const __grid_constant__ TmapPack pack, const CUtensorMap* __restrict__ dmaps, int lanes_mode,
            const LaneDev* __restrict__ lanes, const WarpParams* __restrict__ wps,
            PtrPack srcp, const uint8_t* __restrict__ src0, size_t sframe, size_t sstride, int sw, int sh,
            MutPtrPack dstp, uint8_t* __restrict__ dst0, size_t dframe, int dw, int dh, size_t dstride,
            int rows_per_cta, int dst_vec, int wp_slot, int border_, int bmode_) {
    const int border = BORDER ? border_ : 0, bmode = BORDER ? bmode_ : 0;
    extern __shared__ __align__(1024) unsigned char wt_smem[];
    unsigned char* const S_raw = wt_smem;                                              // [2][WT_ROWS][WT_RAW_PITCH]
    int2* const colAB = reinterpret_cast<int2*>(wt_smem + 2 * WQ_RAWB);               // [WT_W] (adelta, bdelta)
    int2* const S_rowXY = colAB + WT_W;                                               // [WT_H * WQ_MAXT] (X0, Y0), rows past the frame included
    int (*S_box)[8] = reinterpret_cast<int (*)[8]>(S_rowXY + WT_H * WQ_MAXT);          // [2]
    unsigned long long* const S_mbar = reinterpret_cast<unsigned long long*>(S_box + 2);   // [2]
    int* const S_misc = reinterpret_cast<int*>(S_mbar + 2);                            // [tile of the strip]: irregular-step mask

    // CTAs are dispatched in linear block order.  When the frame height is not a whole number of strips the last strip of every
    // frame is short (1080 rows = 4 x 256 + 56): those CTAs are moved to the END of the order, so that they fill the slots the
    // last round of full strips leaves idle instead of leaving a quarter-filled extra round (longest work first).
    int bx = blockIdx.x, by = blockIdx.y, z = blockIdx.z;
    {
        const int nx = gridDim.x, ny = gridDim.y, nz = gridDim.z;
        if (ny > 1 && dh - (ny - 1) * rows_per_cta < rows_per_cta) {
            const unsigned L = blockIdx.x + nx * (blockIdx.y + ny * blockIdx.z), F = (unsigned)nx * (ny - 1) * nz;
            if (L < F) { bx = L % nx; const unsigned r = L / nx; by = r % (ny - 1); z = r / (ny - 1); }
            else { const unsigned r = L - F; bx = r % nx; z = r / nx; by = ny - 1; }
        }
    }
    const double* __restrict__ m = lanes_mode ? lanes[z].wpb[wp_slot]->m : wps[z].m;
    uint8_t* __restrict__ dst = lanes_mode ? dstp.p[z] : dst0 + (size_t)z * dframe;
    const CUtensorMap* tmap = dmaps ? dmaps + (lanes_mode ? z : 0) : &pack.m[lanes_mode ? z : 0];
    const int zc = lanes_mode ? 0 : z;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = bx * WT_W, ys = by * rows_per_cta;
    const int ye = min(ys + rows_per_cta, dh);
    const uint32_t s_raw = (uint32_t)__cvta_generic_to_shared(S_raw);
    const uint32_t s_mbar = (uint32_t)__cvta_generic_to_shared(S_mbar);
    const uint32_t s_rowXY = (uint32_t)__cvta_generic_to_shared(S_rowXY);
    const uint32_t s_colAB = (uint32_t)__cvta_generic_to_shared(colAB);

    // ---- 0. strip prologue: column terms and row terms of cv::warpAffine.  A row thread's warp holds the 32 rows of one
    //         tile; it also notes which 4-row steps cannot keep the vertical terms of the step the same lane did before
    //         (4 rows up; 20 rows up for the first step of a 16-row half): Y0 must have advanced by exactly 1024 per row.
    if (tid < WT_W) {
        const double xd = (double)min(x0 + tid, dw - 1);     // columns beyond the frame repeat the last one: their taps stay in the box
        colAB[tid] = make_int2(sat_int(m[0] * xd * 1024.0), sat_int(m[3] * xd * 1024.0));
    } else {
        for (int r = tid - WT_W; r < rows_per_cta; r += WT_THREADS - WT_W) {      // rows past the frame are computed too, never stored
            const int back = (r & 15) < 4 ? 20 : 4;
            const double yd = (double)(ys + r), yp = (double)(ys + r - back);
            const int Y0 = sat_int((m[4] * yd + m[5]) * 1024.0) + 16;
            S_rowXY[r] = make_int2(sat_int((m[1] * yd + m[2]) * 1024.0) + 16, Y0);
            const unsigned irr = __ballot_sync(0xFFFFFFFFu, Y0 - (sat_int((m[4] * yp + m[5]) * 1024.0) + 16) != 1024 * back);
            if (lane == 0) {
                unsigned steps = 0u;
#pragma unroll
                for (int s8 = 0; s8 < 8; ++s8) steps |= ((irr >> (4 * s8)) & 0xFu) ? (1u << s8) : 0u;
                S_misc[r >> 5] = (int)steps;
            }
        }
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s_mbar));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s_mbar + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (dmaps) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" :: "l"(tmap) : "memory");
    }
    __syncthreads();
    const int tw = min(WT_W, dw - x0);
    const int2 cA = colAB[0], cB = colAB[tw - 1];
    const bool col_ok = max(max(abs(cA.x), abs(cB.x)), max(abs(cA.y), abs(cB.y))) < (1 << 26);
    // lane constants: quad q of segment seg; the lane walks rows 16 half + 4 t + j of every tile
    const int seg = warp & 3, half = warp >> 2, q = lane & 7, j = lane >> 3;
    const int c0 = 32 * seg + 4 * q;
    const int npx = min(max(tw - c0, 0), 4);
    const int cq = npx > 0 ? c0 : 32 * seg;                 // a quad wholly beyond the frame recomputes the segment's first quad (not stored)
    const bool seg_on = 32 * seg < tw;
    uint32_t adT0, dT3, M1, M2;
    bool lane_bad;
    {
        const int2 k0 = colAB[cq], k1 = colAB[cq + 1], k2 = colAB[cq + 2], k3 = colAB[cq + 3];
        adT0 = (uint32_t)k0.x * 16u;
        // e_k - e_0 with e = adelta - 1024 x: all four equal, or one step of +-1 somewhere inside the quad
        const int d1 = k1.x - k0.x - 1024, d2 = k2.x - k0.x - 2048, d3 = k3.x - k0.x - 3072;
        lane_bad = !((d3 == 0 && d1 == 0 && d2 == 0) || ((d3 == 1 || d3 == -1) && (d1 == 0 || d1 == d3) && (d2 == d1 || d2 == d3)));
        dT3 = (uint32_t)(d3 * 16);
        M1 = (d1 != 0) ? 0xFFFFFFFFu : 0u;
        M2 = (d2 != 0) ? 0xFFFFFFFFu : 0u;
    }
    uint32_t s_bd = s_colAB + 8u * (uint32_t)cq;             // this quad's four (adelta, bdelta) entries
    asm volatile("" : "+r"(s_bd), "+r"(adT0), "+r"(dT3), "+r"(M1), "+r"(M2));

    // source box of the tile starting at row y0 -> S_box[b]; issues its fetch into raw buffer b when the box fits
    auto box_and_fetch = [&](int y0, int b) {
        const int yb = min(y0 + WT_H, ye) - 1;
        const int2 rA = S_rowXY[y0 - ys], rB = S_rowXY[yb - ys];
        int minx = INT_MAX, maxx = INT_MIN, miny = INT_MAX, maxy = INT_MIN;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int2 cc = (c & 1) ? cB : cA, rr = (c & 2) ? rB : rA;
            const int X = (rr.x + cc.x) >> 10, Y = (rr.y + cc.y) >> 10;
            minx = min(minx, X); maxx = max(maxx, X); miny = min(miny, Y); maxy = max(maxy, Y);
        }
        const int rminx = minx - border, rmaxx = maxx - border, rminy = miny - border, rmaxy = maxy - border;
        const bool inside = border == 0 || bmode == 0 || (rminx >= 0 && rmaxx + 1 < sw && rminy >= 0 && rmaxy + 1 < sh);
        const int axT = rminx & ~15;                         // TMA box origin: 16 pixels = 48 bytes = 12 words
        const int nrows = maxy + 2 - miny;
        const bool row_ok = max(max(abs(rA.x), abs(rB.x)), max(abs(rA.y), abs(rB.y))) < (1 << 26);
        const bool ok = col_ok && row_ok && inside && minx > -30000 && maxx < 30000 && miny > -30000 && maxy < 30000 &&
                        rmaxx + 2 - axT <= WT_RAW_PITCH / 3 && nrows <= WT_ROWS && nrows > 0;
        int* bx = S_box[b];
        bx[0] = axT + border; bx[1] = miny; bx[4] = ok ? 1 : 0;        // box origin in the (bordered) coordinates the taps use
        if (ok) {
            const uint32_t mb = s_mbar + 8u * (uint32_t)b;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mb), "r"(WQ_BOXB) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         :: "r"(s_raw + (uint32_t)(b * WQ_RAWB)), "l"(tmap), "r"(mb), "r"(3 * (axT >> 2)), "r"(rminy), "r"(zc) : "memory");
        }
    };
    if (tid == 0) {
        box_and_fetch(ys, 0);
        if (ys + WT_H < ye) box_and_fetch(ys + WT_H, 1);
    }
    __syncthreads();

    // per-lane walk state, carried from step to step and from tile to tile
    uint32_t rowp = s_rowXY + 8u * (uint32_t)(16 * half + j);                  // this lane's (X0, Y0) entry
    // entries from rowp_endq on are rows past the frame (a quad beyond the frame's right edge stores nothing at all)
    uint32_t rowp_endq = npx == 4 ? s_rowXY + 8u * (uint32_t)(ye - ys) : 0u;
    asm volatile("" : "+r"(rowp_endq));                                        // (kept in a register, not re-derived every step)
    uint32_t rowp_endj = s_rowXY + 8u * (uint32_t)(ye - ys + j);               // the warp's step is inside the frame while rowp < rowp_endj
    asm volatile("" : "+r"(rowp_endj));
    uint8_t* g = dst + (size_t)(ys + 16 * half + j) * dstride + (size_t)(x0 + c0) * 3;
    const size_t gstep = 4 * dstride;
    uint32_t Bk[4] = {0u, 0u, 0u, 0u}, Bck[4] = {0u, 0u, 0u, 0u}, rowA = 0u, smask = 0u;
    bool ybad = true, anysplit = false, stale = true;        // stale: no vertical terms carried into the next tile
    int oy_prev = 0;

    int buf = 0;
    uint32_t phase = 0;                                     // bit b: parity of raw buffer b's barrier
    for (int y0 = ys; y0 < ye; y0 += WT_H, buf ^= 1) {
        const int ox = S_box[buf][0], oy = S_box[buf][1];
        const bool ok = S_box[buf][4] != 0;
        if (!ok) {
            for (int i = tid; i < WT_W * WT_H; i += WT_THREADS) {
                const int x = x0 + (i & (WT_W - 1)), y = y0 + i / WT_W;
                if (x < dw && y < ye) {
                    const uint8_t* sp = lanes_mode ? srcp.p[z] : src0 + (size_t)z * sframe;
                    warp_pixel<BORDER>(sp, sw, sh, sstride, m, border, bmode, x, y, dst + (size_t)y * dstride + 3 * x);
                }
            }
            stale = true;
            rowp += 8u * WT_H;
            g += 8 * gstep;
        } else {
            mbar_wait(s_mbar + 8u * (uint32_t)buf, (phase >> buf) & 1u);
            phase ^= 1u << buf;
            if (seg_on) {                                    // (a warp whose 16 rows lie past the frame finds rowp >= rowp_stop below)
                const uint32_t s_box = s_raw + (uint32_t)(buf * WQ_RAWB);
                const uint32_t kx = adT0 - ((uint32_t)ox << 14);
                uint32_t irrbits = ((uint32_t)S_misc[(y0 - ys) >> 5] >> (4 * half)) | (stale ? 1u : 0u);   // bit t: derive the vertical terms at step t
                // carried row address -> this tile's box (the regular step below adds the usual 4 rows)
                rowA += (uint32_t)((buf ? WQ_RAWB : -WQ_RAWB) + (oy_prev - oy + 16) * WT_RAW_PITCH);
                const uint32_t rowp_stop = min(rowp + 128u, rowp_endj);
#pragma unroll 1
                for (; rowp < rowp_stop; rowp += 32u, g += gstep, irrbits >>= 1) {
                    const uint2 xy = lds64<0>(rowp);
                    if (irrbits & 1u) {
                        // vertical terms of the four columns: weights, source rows (bdelta is monotone in x, so the rows of
                        // px 1, 2 lie between those of px 0 and px 3)
                        uint32_t b0, b1, b2, b3, u0, u1, u2, u3;
                        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(u0), "=r"(b0), "=r"(u1), "=r"(b1) : "r"(s_bd));
                        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+16];" : "=r"(u2), "=r"(b2), "=r"(u3), "=r"(b3) : "r"(s_bd));
                        const uint32_t yb = xy.y - ((uint32_t)oy << 10);
                        const uint32_t t20 = yb + b0, t21 = yb + b1, t22 = yb + b2, t23 = yb + b3;
                        Bk[0] = t20 & 0x3E0u; Bk[1] = t21 & 0x3E0u; Bk[2] = t22 & 0x3E0u; Bk[3] = t23 & 0x3E0u;
                        Bck[0] = 1024u - Bk[0]; Bck[1] = 1024u - Bk[1]; Bck[2] = 1024u - Bk[2]; Bck[3] = 1024u - Bk[3];
                        const uint32_t r0 = t20 >> 10, r3 = t23 >> 10, rb = min(r0, r3);
                        smask = (r0 - rb) | (((t21 >> 10) - rb) << 1) | (((t22 >> 10) - rb) << 2) | ((r3 - rb) << 3);
                        ybad = lane_bad || max(r0, r3) - rb > 1u;
                        anysplit = __any_sync(0xFFFFFFFFu, r0 != r3);
                        rowA = s_box + rb * (uint32_t)WT_RAW_PITCH;
                    } else {
                        rowA += 4u * (uint32_t)WT_RAW_PITCH;
                    }
                    // horizontal terms of px 0 and px 3 (px 1, 2 follow one of them: M1, M2)
                    const uint32_t T0 = xy.x * 16u + kx, T3 = T0 + dT3;
                    const uint32_t c = T0 >> 14, c3 = 3u * c;                      // px 0's column / byte in the box
                    const uint32_t o = c3 & 3u;
                    const uint32_t o0 = __shfl_sync(0xFFFFFFFFu, o, 0);
                    const bool bad = ybad || (T3 >> 14) != c || o != o0;
                    uint32_t px[4];
                    if (!__any_sync(0xFFFFFFFFu, bad)) {
                        const uint32_t a = rowA + (c3 & ~3u);
                        const uint32_t W0 = (T0 & 0x3E00u) * 0xFFFFu + 16384u, W3 = (T3 & 0x3E00u) * 0xFFFFu + 16384u;
                        const uint32_t W[4] = {W0, bitsel(W0, W3, M1), bitsel(W0, W3, M2), W3};
                        if (!anysplit) {
                            if (o0 < 2u) { if (o0 == 0u) quad_step<0>(a, W, Bk, Bck, px); else quad_step<1>(a, W, Bk, Bck, px); }
                            else { if (o0 == 2u) quad_step<2>(a, W, Bk, Bck, px); else quad_step<3>(a, W, Bk, Bck, px); }
                        } else {
                            if (o0 < 2u) { if (o0 == 0u) quad_step_split<0>(a, W, Bk, smask, px); else quad_step_split<1>(a, W, Bk, smask, px); }
                            else { if (o0 == 2u) quad_step_split<2>(a, W, Bk, smask, px); else quad_step_split<3>(a, W, Bk, smask, px); }
                        }
                    } else {
                        // per-pixel path: the same integers, byte loads from the box
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int2 ck = colAB[cq + k];
                            const uint32_t T1 = xy.x * 16u - ((uint32_t)ox << 14) + (uint32_t)ck.x * 16u;
                            const uint32_t t2 = xy.y - ((uint32_t)oy << 10) + (uint32_t)ck.y;
                            const uint32_t ab = s_box + (t2 >> 10) * (uint32_t)WT_RAW_PITCH + 3u * (T1 >> 14);
                            const uint32_t ax = (T1 >> 9) & 31u, ay = (t2 >> 5) & 31u;
                            const uint32_t w00 = (32u - ax) * (32u - ay), w01 = ax * (32u - ay), w10 = (32u - ax) * ay, w11 = ax * ay;
                            uint32_t v = 0u;
#pragma unroll
                            for (int ch = 0; ch < 3; ++ch) {
                                const uint32_t acc = w00 * lds8(ab + ch) + w01 * lds8(ab + 3 + ch) +
                                                     w10 * lds8(ab + WT_RAW_PITCH + ch) + w11 * lds8(ab + WT_RAW_PITCH + 3 + ch);
                                v |= ((acc + 512u) >> 10) << (8 * ch);
                            }
                            px[k] = v;
                        }
                    }
                    // three aligned words per quad, predicated on the row being inside the frame (and the quad inside its width)
                    asm volatile("{\n.reg .pred p;\nsetp.lt.u32 p, %4, %5;\n@p st.global.u32 [%0], %1;\n@p st.global.u32 [%0+4], %2;\n@p st.global.u32 [%0+8], %3;\n}\n"
                                 :: "l"(g), "r"(__byte_perm(px[0], px[1], 0x4210)), "r"(__byte_perm(px[1], px[2], 0x5421)),
                                    "r"(__byte_perm(px[2], px[3], 0x6542)), "r"(rowp), "r"(rowp_endq) : "memory");
                }
                stale = false;
                oy_prev = oy;
                rowp += 8u * 16u;                            // on to the same rows of the next tile (a short last tile ends the strip)
                g += 4 * gstep;
            }
        }
        __syncthreads();                                   // every warp is done with raw buffer buf
        if (tid == 0 && y0 + 2 * WT_H < ye) box_and_fetch(y0 + 2 * WT_H, buf);
    }
}

// ------------------------------------------------------------------------------------------------
// Crop-and-zoom fused into the warp (Stabilizer.cpp:1056-1060 + :1108-1124): cv::warpAffine, the (b, b, W-2b, H-2b)
// crop and cv::resize back to W x H, reading the source once and writing the output once.  One CTA walks a strip
// of 120x30 OUTPUT tiles.  The resize taps of such a tile cover at most 122x32 pixels of the warped image (the
// zoom factor is > 1), so per tile
//   A. the 128x32 warped tile that contains them is produced exactly like a k_warp_tma tile (same TMA box, same
//      re-pack, same fixed-point inner loop) but lands in shared memory as [B G R -] words, and
//   B. the output tile is cv::resize's 11-bit bilinear of that shared tile: horizontal pass as IDP.2A with the
//      coefficient pair (a0 | a1 << 16), vertical pass ((b * (h >> 4)) >> 16 as IMAD.HI with b << 16), rounding
//      (v + 2) >> 2 — term for term VResizeLinear / HResizeLinear of OpenCV (oracle/cv_models.py resize_linear).
// The warped frame never exists in HBM.
#define ZT_W 120
#define ZT_H 30
#define ZT_MAXT 4
#define ZT_SMEM (WT_ROWS * WT_TPITCH + WT_ROWS * WT_RAW_PITCH + WT_H * WT_W * 4 + (WT_THREADS / 32) * WT_W * 4 + \
                 2 * WT_H * 8 + 2 * WT_H * 16 + 64 + 16)

__global__ void __launch_bounds__(WT_THREADS, 3)
k_warp_zoom_tma(const __grid_constant__ TmapPack pack, const CUtensorMap* __restrict__ dmaps, int lanes_mode,
                const LaneDev* __restrict__ lanes, const WarpParams* __restrict__ wps,
                PtrPack srcp, const uint8_t* __restrict__ src0, size_t sframe, size_t sstride, int sw, int sh,
                MutPtrPack dstp, uint8_t* __restrict__ dst0, size_t dframe, size_t dstride,
                int rows_per_cta, int dst_vec, int wp_slot, int border, double zsx, double zsy) {
    extern __shared__ __align__(1024) unsigned char wt_smem[];
    uint32_t* const S_src = reinterpret_cast<uint32_t*>(wt_smem);
    unsigned char* const S_raw = wt_smem + WT_ROWS * WT_TPITCH;
    uint32_t* const S_w = reinterpret_cast<uint32_t*>(S_raw + WT_ROWS * WT_RAW_PITCH);       // warped tile, 128 words x 32 rows
    uint32_t* const S_out = S_w + WT_H * WT_W;
    int2* const S_rowXY = reinterpret_cast<int2*>(S_out + (WT_THREADS / 32) * WT_W);         // [2][WT_H]
    int4* const S_vtap = reinterpret_cast<int4*>(S_rowXY + 2 * WT_H);                        // [2][WT_H]
    int (*S_box)[8] = reinterpret_cast<int (*)[8]>(S_vtap + 2 * WT_H);
    unsigned long long* const S_mbar = reinterpret_cast<unsigned long long*>(S_box + 2);

    const int z = blockIdx.z;
    const double* __restrict__ m = lanes_mode ? lanes[z].wpb[wp_slot]->m : wps[z].m;
    uint8_t* __restrict__ dst = lanes_mode ? dstp.p[z] : dst0 + (size_t)z * dframe;
    const CUtensorMap* tmap = dmaps ? dmaps + (lanes_mode ? z : 0) : &pack.m[lanes_mode ? z : 0];
    const int zc = lanes_mode ? 0 : z;
    const int dw = sw, dh = sh, cw = sw - 2 * border, ch = sh - 2 * border;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * ZT_W, ys = blockIdx.y * rows_per_cta;
    const int ye = min(ys + rows_per_cta, dh);
    int2* const colAB = reinterpret_cast<int2*>(S_out);              // 128 entries (1 KB)
    int2* const colTap = colAB + WT_W;                               // 120 entries
    const uint32_t s_src = (uint32_t)__cvta_generic_to_shared(S_src);
    const uint32_t s_raw = (uint32_t)__cvta_generic_to_shared(S_raw);
    const uint32_t s_w = (uint32_t)__cvta_generic_to_shared(S_w);
    const uint32_t s_mbar = (uint32_t)__cvta_generic_to_shared(S_mbar);

    // first crop column any tap of this tile column range touches; warped column of tile column c is wx0 + c
    const int c0 = tap_h(x0, cw, zsx).s0;
    const int wx0 = border + c0;

    // row terms of the 32 warped rows and vertical taps of the 30 output rows of the tile starting at output row y0
    auto prep_tile = [&](int y0, int b, int i) {
        const int r0 = tap_v(y0, ch, zsy).s0;
        if (i < WT_H) {
            const double yd = (double)min(border + r0 + i, sh - 1);
            S_rowXY[b * WT_H + i] = make_int2(sat_int((m[1] * yd + m[2]) * 1024.0) + 16, sat_int((m[4] * yd + m[5]) * 1024.0) + 16);
        } else if (i < WT_H + ZT_H) {
            const int k = i - WT_H;
            const AxisTap t = tap_v(min(y0 + k, dh - 1), ch, zsy);
            S_vtap[b * WT_H + k] = make_int4((t.s0 - r0) * (WT_W * 4), (t.s1 - r0) * (WT_W * 4), t.a0 << 16, t.a1 << 16);
        }
    };

    // ---- 0. strip prologue
    if (tid < WT_W) {
        const double xd = (double)min(wx0 + tid, sw - 1);
        colAB[tid] = make_int2(sat_int(m[0] * xd * 1024.0), sat_int(m[3] * xd * 1024.0));
        if (tid < ZT_W) {
            const AxisTap t = tap_h(min(x0 + tid, dw - 1), cw, zsx);
            colTap[tid] = make_int2(4 * (t.s0 - c0), t.a0 | (t.a1 << 16));
        }
    } else if (tid < WT_W + WT_H + ZT_H) {
        prep_tile(ys, 0, tid - WT_W);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s_mbar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (dmaps) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" :: "l"(tmap) : "memory");
    }
    __syncthreads();
    uint32_t adT[4], bd[4], colS[4], colW[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int2 c = colAB[lane + 32 * j];
        adT[j] = (uint32_t)c.x * 16u;
        bd[j] = (uint32_t)c.y;
        const int2 t = colTap[min(lane + 32 * j, ZT_W - 1)];
        colS[j] = (uint32_t)t.x;
        colW[j] = (uint32_t)t.y;
    }
    const int tw = min(ZT_W, dw - x0);
    const int2 cA = colAB[0], cB = colAB[WT_W - 1];
    const bool col_ok = max(max(abs(cA.x), abs(cB.x)), max(abs(cA.y), abs(cB.y))) < (1 << 26);
    const bool vec_out = dst_vec && tw == ZT_W;
    const uint32_t s_px = (uint32_t)__cvta_generic_to_shared(S_out + warp * WT_W + lane);
    const uint32_t s_v4 = (uint32_t)__cvta_generic_to_shared(S_out + warp * WT_W + 4 * lane);
    const uint32_t s_wpx = s_w + 4u * (uint32_t)lane;
    const int r7 = tid / WT_TGRPS, q = tid - r7 * WT_TGRPS;
    const uint32_t s_stage = s_src + (uint32_t)(r7 * WT_TPITCH + 16 * q);
    const uint32_t s_rawt = s_raw + (uint32_t)(r7 * WT_RAW_PITCH + 12 * q);

    // source box of the warped tile whose row terms are (rA, rB) -> S_box[b]; issues its fetch when the box fits
    auto box_and_fetch = [&](int2 rA, int2 rB, int b) {
        int minx = INT_MAX, maxx = INT_MIN, miny = INT_MAX, maxy = INT_MIN;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int2 cc = (c & 1) ? cB : cA, rr = (c & 2) ? rB : rA;
            const int X = (rr.x + cc.x) >> 10, Y = (rr.y + cc.y) >> 10;
            minx = min(minx, X); maxx = max(maxx, X); miny = min(miny, Y); maxy = max(maxy, Y);
        }
        const int ax0 = minx & ~3;
        const int ngrp = (maxx + 1 - ax0) / 4 + 1;
        const int nrows = maxy + 2 - miny;
        const bool row_ok = max(max(abs(rA.x), abs(rB.x)), max(abs(rA.y), abs(rB.y))) < (1 << 26);
        const bool ok = col_ok && row_ok && minx > -30000 && maxx < 30000 && miny > -30000 && maxy < 30000 &&
                        ngrp <= WT_GRPS && nrows <= WT_ROWS && ngrp > 0 && nrows > 0;
        const int axT = minx & ~15;
        int* bx = S_box[b];
        bx[0] = ax0; bx[1] = miny; bx[2] = ngrp; bx[3] = nrows; bx[4] = ok ? 1 : 0; bx[5] = 3 * (ax0 - axT);
        if (ok) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s_mbar), "r"(WT_ROWS * WT_RAW_PITCH) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         :: "r"(s_raw), "l"(tmap), "r"(s_mbar), "r"(3 * (axT >> 2)), "r"(miny), "r"(zc) : "memory");
        }
    };
    // the row terms of a tile's first and last warped row, recomputed by the issuing thread (the table entries of
    // the NEXT tile are being written by other threads at that moment)
    auto row_terms = [&](int y0, int i) {
        const int r0 = tap_v(y0, ch, zsy).s0;
        const double yd = (double)min(border + r0 + i, sh - 1);
        return make_int2(sat_int((m[1] * yd + m[2]) * 1024.0) + 16, sat_int((m[4] * yd + m[5]) * 1024.0) + 16);
    };
    if (tid == 0) box_and_fetch(S_rowXY[0], S_rowXY[WT_H - 1], 0);
    __syncthreads();

    int buf = 0;
    uint32_t phase = 0;
    for (int y0 = ys; y0 < ye; y0 += ZT_H, buf ^= 1) {
        const int ax0 = S_box[buf][0], by0 = S_box[buf][1], nrows = S_box[buf][3];
        const bool ok = S_box[buf][4] != 0;
        const bool more = y0 + ZT_H < ye;
        const int orows = min(ZT_H, ye - y0);                                   // output rows of this tile
        const int wrows = (S_vtap[buf * WT_H + orows - 1].y >> 9) + 1;          // warped rows its taps touch
        if (ok) {
            mbar_wait(s_mbar, phase);
            phase ^= 1;
            // No per-thread row / column test: TMA always delivers the full 42 x 480-byte box, so rows beyond nrows
            // and column groups beyond ngrp are re-packed too (never read as taps) — cheaper than the predicates.
            const uint32_t s_rawq = s_rawt + (uint32_t)S_box[buf][5];
            if (tid < WT_TGRPS * WT_TSROWS) {
#define WT_REPACK(k)                                                                                           \
                if ((k) * WT_TSROWS < nrows) {                           /* CTA-uniform */                      \
                    const uint32_t w0 = lds32<(k) * WT_TSROWS * WT_RAW_PITCH>(s_rawq);                          \
                    const uint32_t w1 = lds32<(k) * WT_TSROWS * WT_RAW_PITCH + 4>(s_rawq);                      \
                    const uint32_t w2 = lds32<(k) * WT_TSROWS * WT_RAW_PITCH + 8>(s_rawq);                      \
                    const uint32_t w3 = lds32<(k) * WT_TSROWS * WT_RAW_PITCH + 12>(s_rawq);                     \
                    const uint4 o = repack_bgrr(w0, w1, w2, w3);                                               \
                    sts128<(k) * WT_TSROWS * WT_TPITCH>(s_stage, o.x, o.y, o.z, o.w);                       \
                }
                WT_REPACK(0) WT_REPACK(1) WT_REPACK(2) WT_REPACK(3) WT_REPACK(4) WT_REPACK(5) WT_REPACK(6)
#undef WT_REPACK
            }
        }
        __syncthreads();                                   // S_src ready, raw box free, previous tile's stage B done
        if (more) {
            if (tid == 0) box_and_fetch(row_terms(y0 + ZT_H, 0), row_terms(y0 + ZT_H, WT_H - 1), buf ^ 1);
            else if (tid >= WT_W && tid < WT_W + WT_H + ZT_H) prep_tile(y0 + ZT_H, buf ^ 1, tid - WT_W);
        }
        // ---- A. the warped tile -> S_w
        if (!ok) {
            const uint8_t* sp = lanes_mode ? srcp.p[z] : src0 + (size_t)z * sframe;
            const int r0 = tap_v(y0, ch, zsy).s0;
            for (int i = tid; i < WT_W * wrows; i += WT_THREADS) {
                const int c = i & (WT_W - 1), r = i / WT_W;
                S_w[r * WT_W + c] = warp_pixel_word<false>(sp, sw, sh, sstride, m, 0, 0, min(wx0 + c, sw - 1), min(border + r0 + r, sh - 1));
            }
        } else {
            const uint32_t bx14 = (uint32_t)ax0 << 14, by10 = (uint32_t)by0 << 10;
#pragma unroll
            for (int rr = warp; rr < WT_H; rr += WT_THREADS / 32) {
                if (rr >= wrows) break;
                const int2 xy = S_rowXY[buf * WT_H + rr];
                const uint32_t rxT = (uint32_t)xy.x * 16u - bx14;
                const uint32_t ryS = (uint32_t)xy.y - by10;
                const uint32_t s_row = s_wpx + (uint32_t)rr * (WT_W * 4);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t T1 = rxT + adT[j];
                    const uint32_t t2 = ryS + bd[j];
                    const uint32_t a = s_src + 4u * ((t2 >> 10) * (WT_TPITCH / 4) + (T1 >> 14));
                    const uint32_t A16 = T1 & 0x3E00u;
                    const uint32_t W16 = A16 * 0xFFFFu + 16384u;
                    const uint32_t B = t2 & 0x3E0u, Bc = 1024u - B;
                    const uint32_t t00 = lds32<0>(a), t01 = lds32<4>(a), t10 = lds32<WT_TPITCH>(a), t11 = lds32<WT_TPITCH + 4>(a);
                    const uint32_t u0 = __byte_perm(t00, t01, 0x5140);
                    const uint32_t l0 = __byte_perm(t10, t11, 0x5140);
                    const uint32_t hb0 = __dp2a_lo(W16, u0, 0u), hg0 = __dp2a_hi(W16, u0, 0u), hr0 = __dp2a_hi(W16, t00, 0u);
                    const uint32_t hb1 = __dp2a_lo(W16, l0, 0u), hg1 = __dp2a_hi(W16, l0, 0u), hr1 = __dp2a_hi(W16, t10, 0u);
                    const uint32_t vb = hb0 * Bc + (hb1 * B + 0x800000u);
                    const uint32_t vg = hg0 * Bc + (hg1 * B + 0x800000u);
                    const uint32_t vr = hr0 * Bc + (hr1 * B + 0x800000u);
                    const uint32_t px = __byte_perm(__byte_perm(vb, vg, 0x0073), vr, 0x0710);
                    if (j == 0) sts32<0>(s_row, px);
                    else if (j == 1) sts32<128>(s_row, px);
                    else if (j == 2) sts32<256>(s_row, px);
                    else sts32<384>(s_row, px);
                }
            }
        }
        __syncthreads();                                   // S_w complete (and S_src free for the next re-pack)
        // ---- B. cv::resize of the shared warped tile: warp -> output rows, lane -> pixels x0 + lane + 32 j
        uint8_t* grow = dst + (size_t)(y0 + warp) * dstride + (size_t)x0 * 3 + (vec_out ? 12 * lane : 0);
#pragma unroll
        for (int rr = warp; rr < ZT_H; rr += WT_THREADS / 32) {
            if (rr >= orows) break;
            const int4 vt = S_vtap[buf * WT_H + rr];
            const uint32_t B0 = (uint32_t)vt.z, B1 = (uint32_t)vt.w;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j == 3 && lane >= ZT_W - 96) break;
                const uint32_t a0 = s_w + colS[j] + (uint32_t)vt.x, a1 = s_w + colS[j] + (uint32_t)vt.y;
                const uint32_t t00 = lds32<0>(a0), t01 = lds32<4>(a0), t10 = lds32<0>(a1), t11 = lds32<4>(a1);
                const uint32_t W = colW[j];
                const uint32_t u0 = __byte_perm(t00, t01, 0x5140), r0p = __byte_perm(t00, t01, 0x0062);
                const uint32_t u1 = __byte_perm(t10, t11, 0x5140), r1p = __byte_perm(t10, t11, 0x0062);
                const uint32_t hb0 = __dp2a_lo(W, u0, 0u) >> 4, hg0 = __dp2a_hi(W, u0, 0u) >> 4, hr0 = __dp2a_lo(W, r0p, 0u) >> 4;
                const uint32_t hb1 = __dp2a_lo(W, u1, 0u) >> 4, hg1 = __dp2a_hi(W, u1, 0u) >> 4, hr1 = __dp2a_lo(W, r1p, 0u) >> 4;
                const uint32_t vb = __umulhi(B1, hb1) + (__umulhi(B0, hb0) + 2u);
                const uint32_t vg = __umulhi(B1, hg1) + (__umulhi(B0, hg0) + 2u);
                const uint32_t vr = __umulhi(B1, hr1) + (__umulhi(B0, hr0) + 2u);
                const uint32_t px = (vb >> 2) | ((vg << 6) & 0xFF00u) | ((vr << 14) & 0xFF0000u);
                if (j == 0) sts32<0>(s_px, px);
                else if (j == 1) sts32<128>(s_px, px);
                else if (j == 2) sts32<256>(s_px, px);
                else sts32<384>(s_px, px);
            }
            __syncwarp();
            if (vec_out) {
                if (lane < ZT_W / 4) {
                    uint32_t vx, vy, vz, vw;
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(vx), "=r"(vy), "=r"(vz), "=r"(vw) : "r"(s_v4));
                    uint32_t* g = reinterpret_cast<uint32_t*>(grow);
                    g[0] = __byte_perm(vx, vy, 0x4210);
                    g[1] = __byte_perm(vy, vz, 0x5421);
                    g[2] = __byte_perm(vz, vw, 0x6542);
                }
            } else {
                const uint32_t* orow = S_out + warp * WT_W;
                for (int c = lane; c < tw; c += 32) {
                    const uint32_t v = orow[c];
                    grow[3 * c] = (uint8_t)v; grow[3 * c + 1] = (uint8_t)(v >> 8); grow[3 * c + 2] = (uint8_t)(v >> 16);
                }
            }
            grow += (size_t)(WT_THREADS / 32) * dstride;
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(WT_THREADS, 4) k_warp_tiled_lanes(const LaneDev* __restrict__ lanes, PtrPack src, MutPtrPack dst,
                                                                  WarpGeom g, int rows_per_cta, int src_vec, int dst_vec) {
    __shared__ __align__(16) WarpTileSmem S;
    warp_strip(S, src.p[blockIdx.z], g.src_w, g.src_h, g.src_stride, dst.p[blockIdx.z], g.out_w, g.out_h, g.out_stride,
               lanes[blockIdx.z].wpb[g.wp_slot]->m, rows_per_cta, src_vec != 0, dst_vec != 0);
}

__global__ void __launch_bounds__(WT_THREADS, 4) k_warp_tiled_frames(const uint8_t* __restrict__ src, int sw, int sh, size_t sstride,
                                                                   size_t sframe, uint8_t* __restrict__ dst, int dw, int dh,
                                                                   size_t dstride, size_t dframe,
                                                                   const WarpParams* __restrict__ wps, int rows_per_cta,
                                                                   int src_vec, int dst_vec) {
    __shared__ __align__(16) WarpTileSmem S;
    warp_strip(S, src + blockIdx.z * sframe, sw, sh, sstride, dst + blockIdx.z * dframe, dw, dh, dstride,
               wps[blockIdx.z].m, rows_per_cta, src_vec != 0, dst_vec != 0);
}

// tiles per CTA strip: enough CTAs to fill the machine first (148 SMs x 4 resident CTAs), then longer strips
static inline int strip_rows(int dw, int dh, int n_frames, int max_tiles) {
    const long tiles = (long)((dw + WT_W - 1) / WT_W) * ((dh + WT_H - 1) / WT_H) * n_frames;
    int t = 1;
    while (t < max_tiles && tiles / (t * 2) >= 148L * 4 * 4) t *= 2;
    return t * WT_H;
}

static inline bool vec_ok(const void* p, size_t stride, int a) { return ((uintptr_t)p % a == 0) && (stride % a == 0); }

template <bool BORDER>
__global__ void __launch_bounds__(256) k_warp_lanes(const LaneDev* __restrict__ lanes, PtrPack src, MutPtrPack dst, WarpGeom g) {
    const WarpParams* wp = lanes[blockIdx.z].wpb[g.wp_slot];
    int x = blockIdx.x * 64 + (threadIdx.x & 63);
    int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= g.out_w || y >= g.out_h) return;
    uint8_t* o = dst.p[blockIdx.z] + (size_t)y * g.out_stride + 3 * x;
    warp_pixel<BORDER>(src.p[blockIdx.z], g.src_w, g.src_h, g.src_stride, wp->m, g.border, g.border_mode, x, y, o);
}

// ---- host side of the TMA path ------------------------------------------------------------------------
typedef CUresult (*TmaEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TmaEncodeFn tma_encode_fn() {
    static TmaEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (TmaEncodeFn)p;
        cudaGetLastError();
    }
    return fn;
}
static inline bool tma_geometry_ok(const void* base, int w, size_t stride, size_t frame_bytes, int n) {
    return ((uintptr_t)base % 16 == 0) && stride % 16 == 0 && (n <= 1 || frame_bytes % 16 == 0) && w % 4 == 0 &&
           (size_t)w * 3 <= stride && w >= 16;
}
// tensor map over {row words (u32), rows, frames}; box = one raw source box (WT_RAW_WORDS x WT_ROWS x 1)
static bool tma_encode_map(CUtensorMap* m, const uint8_t* base, int w, int h, size_t stride, size_t frame_bytes, int n);
// Frames come from a ring or from a handful of decoder surfaces, so the same (address, geometry) recurs every few frames:
// a small direct-mapped, per-thread cache saves the driver's encode call on the per-frame enqueue path.
static bool tma_make_map(CUtensorMap* m, const uint8_t* base, int w, int h, size_t stride, size_t frame_bytes, int n) {
    struct Entry { const uint8_t* base; int w, h, n; size_t stride, frame_bytes; CUtensorMap map; };
    static thread_local Entry cache[256] = {};
    Entry& e = cache[((uintptr_t)base >> 12) * 2654435761u >> 24 & 255];
    if (e.base == base && e.w == w && e.h == h && e.n == n && e.stride == stride && e.frame_bytes == frame_bytes) {
        *m = e.map;
        return true;
    }
    if (!tma_encode_map(m, base, w, h, stride, frame_bytes, n)) return false;
    e.base = base; e.w = w; e.h = h; e.n = n; e.stride = stride; e.frame_bytes = frame_bytes; e.map = *m;
    return true;
}
static bool tma_encode_map(CUtensorMap* m, const uint8_t* base, int w, int h, size_t stride, size_t frame_bytes, int n) {
    TmaEncodeFn enc = tma_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)w * 3 / 4, (cuuint64_t)h, (cuuint64_t)(n > 0 ? n : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)stride, (cuuint64_t)(n > 1 ? frame_bytes : stride * (size_t)h)};
    cuuint32_t box[3] = {WT_RAW_WORDS, WT_ROWS, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static bool warp_v1() {
    static const bool v = [] { const char* e = getenv("VS_WARP_V1"); return e && *e == '1'; }();
    return v;
}
static bool tma_kernel_ready() {
    static int state[64] = {0};                   // per device: 0 unknown, 1 ready, -1 unavailable
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    if (state[dev] == 0) {
        cudaError_t e = cudaFuncSetAttribute(k_warp_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_TMA_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_warp_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_TMA_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_warp_quad<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WQ_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_warp_quad<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WQ_SMEM);
        state[dev] = (e == cudaSuccess && tma_encode_fn()) ? 1 : -1;
        cudaGetLastError();
    }
    return state[dev] == 1;
}

#define WT_LAUNCH(QUAD, BORDER, ...)                                                            \
    do {                                                                                        \
        if (QUAD) k_warp_quad<BORDER><<<grid, WT_THREADS, WQ_SMEM, st>>>(__VA_ARGS__);           \
        else k_warp_tma<BORDER><<<grid, WT_THREADS, WT_TMA_SMEM, st>>>(__VA_ARGS__);             \
    } while (0)

static void launch_warp_plain(const LaneDev* lanes, int n_lanes, const PtrPack& src, const MutPtrPack& dst, const WarpGeom& g,
                              cudaStream_t st) {
    bool sv = true, dv = true, tma = tma_kernel_ready() && (n_lanes <= WT_TMA_MAXPACK || g.d_tmaps != nullptr);
    for (int i = 0; i < n_lanes; ++i) {
        sv = sv && vec_ok(src.p[i], g.src_stride, 4);
        dv = dv && vec_ok(dst.p[i], g.out_stride, 4);
        tma = tma && tma_geometry_ok(src.p[i], g.src_w, g.src_stride, 0, 1);
    }
    if (tma) {
        const bool quad = !warp_v1() && dv && g.out_w % 4 == 0;     // k_warp_quad writes packed quads with aligned 32-bit stores
        const int rows = strip_rows(g.out_w, g.out_h, n_lanes, quad ? WQ_MAXT : WT_MAXT);
        dim3 grid((g.out_w + WT_W - 1) / WT_W, (g.out_h + rows - 1) / rows, n_lanes);
        TmapPack pack;
        CUtensorMap big[VS_MAX_GROUP];
        CUtensorMap* maps = n_lanes <= WT_TMA_MAXPACK ? pack.m : big;
        for (int i = 0; i < n_lanes && tma; ++i) tma = tma_make_map(maps + i, src.p[i], g.src_w, g.src_h, g.src_stride, 0, 1);
        if (tma) {
            const CUtensorMap* dmaps = nullptr;
            if (n_lanes > WT_TMA_MAXPACK) {
                cudaMemcpyAsync(g.d_tmaps, big, sizeof(CUtensorMap) * n_lanes, cudaMemcpyHostToDevice, st);
                dmaps = (const CUtensorMap*)g.d_tmaps;
            }
            if (g.mode == 1 && g.border > 0)
                WT_LAUNCH(quad, true, pack, dmaps, 1, lanes, nullptr, src, nullptr, 0, g.src_stride, g.src_w, g.src_h, dst, nullptr, 0,
                          g.out_w, g.out_h, g.out_stride, rows, dv ? 1 : 0, g.wp_slot, g.border, g.border_mode);
            else
                WT_LAUNCH(quad, false, pack, dmaps, 1, lanes, nullptr, src, nullptr, 0, g.src_stride, g.src_w, g.src_h, dst, nullptr, 0,
                          g.out_w, g.out_h, g.out_stride, rows, dv ? 1 : 0, g.wp_slot, 0, 0);
            return;
        }
    }
    if (g.mode == 1) {                                    // no TMA for this geometry: per-pixel kernel
        dim3 grid((g.out_w + 63) / 64, (g.out_h + 3) / 4, n_lanes);
        k_warp_lanes<true><<<grid, 256, 0, st>>>(lanes, src, dst, g);
        return;
    }
    const int rows = strip_rows(g.out_w, g.out_h, n_lanes, WT_MAXT_F);
    dim3 grid((g.out_w + WT_W - 1) / WT_W, (g.out_h + rows - 1) / rows, n_lanes);
    k_warp_tiled_lanes<<<grid, WT_THREADS, 0, st>>>(lanes, src, dst, g, rows, sv ? 1 : 0, dv ? 1 : 0);
}

static bool zoom_kernel_ready() {
    static int state[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    if (state[dev] == 0) {
        cudaError_t e = cudaFuncSetAttribute(k_warp_zoom_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, ZT_SMEM);
        state[dev] = (e == cudaSuccess && tma_encode_fn()) ? 1 : -1;
        cudaGetLastError();
    }
    return state[dev] == 1;
}
static inline int zoom_strip_rows(int dw, int dh, int n_frames) {
    const long tiles = (long)((dw + ZT_W - 1) / ZT_W) * ((dh + ZT_H - 1) / ZT_H) * n_frames;
    int t = 1;
    while (t < ZT_MAXT && tiles / (t * 2) >= 148L * 3 * 4) t *= 2;
    return t * ZT_H;
}
// cv::resize's inverse scale factors for the (w-2b, h-2b) -> (w, h) zoom, as launch_resize_linear forms them
static inline void zoom_scales(int w, int h, int b, double* sx, double* sy) {
    *sx = 1.0 / ((double)w / (double)(w - 2 * b));
    *sy = 1.0 / ((double)h / (double)(h - 2 * b));
}

// fused crop+zoom for per-lane frames; false when the TMA path is not available for this geometry
static bool launch_warp_zoom_lanes(const LaneDev* lanes, int n_lanes, const PtrPack& src, const MutPtrPack& dst, const WarpGeom& g,
                                   cudaStream_t st) {
    if (!zoom_kernel_ready() || !(n_lanes <= WT_TMA_MAXPACK || g.d_tmaps != nullptr)) return false;
    if (g.out_w != g.src_w || g.out_h != g.src_h || g.src_w - 2 * g.border <= 0 || g.src_h - 2 * g.border <= 0) return false;
    bool dv = true;
    for (int i = 0; i < n_lanes; ++i) {
        if (!tma_geometry_ok(src.p[i], g.src_w, g.src_stride, 0, 1)) return false;
        dv = dv && vec_ok(dst.p[i], g.out_stride, 4);
    }
    TmapPack pack;
    CUtensorMap big[VS_MAX_GROUP];
    CUtensorMap* maps = n_lanes <= WT_TMA_MAXPACK ? pack.m : big;
    for (int i = 0; i < n_lanes; ++i)
        if (!tma_make_map(maps + i, src.p[i], g.src_w, g.src_h, g.src_stride, 0, 1)) return false;
    const CUtensorMap* dmaps = nullptr;
    if (n_lanes > WT_TMA_MAXPACK) {
        cudaMemcpyAsync(g.d_tmaps, big, sizeof(CUtensorMap) * n_lanes, cudaMemcpyHostToDevice, st);
        dmaps = (const CUtensorMap*)g.d_tmaps;
    }
    const int rows = zoom_strip_rows(g.out_w, g.out_h, n_lanes);
    dim3 grid((g.out_w + ZT_W - 1) / ZT_W, (g.out_h + rows - 1) / rows, n_lanes);
    double zsx, zsy;
    zoom_scales(g.src_w, g.src_h, g.border, &zsx, &zsy);
    k_warp_zoom_tma<<<grid, WT_THREADS, ZT_SMEM, st>>>(pack, dmaps, 1, lanes, nullptr, src, nullptr, 0, g.src_stride, g.src_w, g.src_h,
                                                       dst, nullptr, 0, g.out_stride, rows, dv ? 1 : 0, g.wp_slot, g.border, zsx, zsy);
    return true;
}

// fused crop+zoom for contiguous frames with device-resident warp set-ups
static bool launch_warp_zoom_frames(const uint8_t* src, int sw, int sh, size_t sstride, size_t sframe, uint8_t* dst, size_t dstride,
                                    size_t dframe, const WarpParams* d_wp, int n_frames, int border, cudaStream_t st) {
    if (!zoom_kernel_ready() || sw - 2 * border <= 0 || sh - 2 * border <= 0) return false;
    if (!tma_geometry_ok(src, sw, sstride, sframe, n_frames)) return false;
    TmapPack pack;
    if (!tma_make_map(pack.m, src, sw, sh, sstride, sframe, n_frames)) return false;
    const bool dv = vec_ok(dst, dstride, 4) && dframe % 4 == 0;
    const int rows = zoom_strip_rows(sw, sh, n_frames);
    dim3 grid((sw + ZT_W - 1) / ZT_W, (sh + rows - 1) / rows, n_frames);
    double zsx, zsy;
    zoom_scales(sw, sh, border, &zsx, &zsy);
    PtrPack sp{};
    MutPtrPack dp{};
    k_warp_zoom_tma<<<grid, WT_THREADS, ZT_SMEM, st>>>(pack, nullptr, 0, nullptr, d_wp, sp, src, sframe, sstride, sw, sh, dp, dst, dframe,
                                                       dstride, rows, dv ? 1 : 0, 0, border, zsx, zsy);
    return true;
}

int launch_warp(const LaneDev* lanes, int n_lanes, const PtrPack& src, const MutPtrPack& dst, WarpGeom g,
                uint8_t* const* scratch, cudaStream_t st) {
    if (g.mode == 2) {
        if (launch_warp_zoom_lanes(lanes, n_lanes, src, dst, g, st)) return 1;
        // no TMA for this geometry: two passes — warp into the lane's scratch frame, then cv::resize the
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
        return 1 + n_lanes;
    }
    launch_warp_plain(lanes, n_lanes, src, dst, g, st);      // mode 0, or mode 1 (copyMakeBorder folded into the source box)
    return 1;
}

// border > 0: the source is the virtual copyMakeBorder frame (dw, dh = sw + 2 border, sh + 2 border); false when the
// TMA path is not available (the caller then uses the per-pixel border kernel)
static bool launch_warp_matrices_border(const uint8_t* src, int sw, int sh, size_t sstride, size_t sframe,
                                        uint8_t* dst, int dw, int dh, size_t dstride, size_t dframe,
                                        const WarpParams* d_wp, int n_frames, int border, int bmode, cudaStream_t st) {
    const bool dv = vec_ok(dst, dstride, 4) && dframe % 4 == 0;
    if (!(tma_kernel_ready() && tma_geometry_ok(src, sw, sstride, sframe, n_frames))) return false;
    const bool quad = !warp_v1() && dv && dw % 4 == 0;
    const int rows = strip_rows(dw, dh, n_frames, quad ? WQ_MAXT : WT_MAXT);
    dim3 grid((dw + WT_W - 1) / WT_W, (dh + rows - 1) / rows, n_frames);
    TmapPack pack;
    if (!tma_make_map(pack.m, src, sw, sh, sstride, sframe, n_frames)) return false;
    PtrPack sp{};
    MutPtrPack dp{};
    if (border > 0)
        WT_LAUNCH(quad, true, pack, nullptr, 0, nullptr, d_wp, sp, src, sframe, sstride, sw, sh, dp, dst, dframe, dw, dh, dstride, rows,
                  dv ? 1 : 0, 0, border, bmode);
    else
        WT_LAUNCH(quad, false, pack, nullptr, 0, nullptr, d_wp, sp, src, sframe, sstride, sw, sh, dp, dst, dframe, dw, dh, dstride, rows,
                  dv ? 1 : 0, 0, 0, 0);
    return true;
}

void launch_warp_matrices(const uint8_t* src, int sw, int sh, size_t sstride, size_t sframe,
                          uint8_t* dst, int dw, int dh, size_t dstride, size_t dframe,
                          const WarpParams* d_wp, int n_frames, cudaStream_t st) {
    const bool sv = vec_ok(src, sstride, 4) && sframe % 4 == 0, dv = vec_ok(dst, dstride, 4) && dframe % 4 == 0;
    if (launch_warp_matrices_border(src, sw, sh, sstride, sframe, dst, dw, dh, dstride, dframe, d_wp, n_frames, 0, 0, st)) return;
    const int rows = strip_rows(dw, dh, n_frames, WT_MAXT_F);
    dim3 grid((dw + WT_W - 1) / WT_W, (dh + rows - 1) / rows, n_frames);
    k_warp_tiled_frames<<<grid, WT_THREADS, 0, st>>>(src, sw, sh, sstride, sframe, dst, dw, dh, dstride, dframe, d_wp, rows,
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

int launch_warp_frames_mode(const uint8_t* src, int sw, int sh, size_t sstride, size_t sframe, uint8_t* dst,
                            size_t dstride, size_t dframe, const WarpParams* d_wp, int n_frames, int mode, int border,
                            int border_mode, uint8_t* scratch, cudaStream_t st) {
    if (n_frames <= 0) return 0;
    if (mode == 0) {
        launch_warp_matrices(src, sw, sh, sstride, sframe, dst, sw, sh, dstride, dframe, d_wp, n_frames, st);
    } else if (mode == 1) {
        const int dw = sw + 2 * border, dh = sh + 2 * border;
        if (launch_warp_matrices_border(src, sw, sh, sstride, sframe, dst, dw, dh, dstride, dframe, d_wp, n_frames, border, border_mode, st))
            return 1;
        dim3 grid((dw + 63) / 64, (dh + 3) / 4, n_frames);
        k_warp_frames_border<<<grid, 256, 0, st>>>(src, sw, sh, sstride, sframe, dst, dw, dh, dstride, dframe, d_wp, border, border_mode);
    } else {
        if (launch_warp_zoom_frames(src, sw, sh, sstride, sframe, dst, dstride, dframe, d_wp, n_frames, border, st)) return 1;
        // no TMA for this geometry: warp each frame into the scratch frame, then cv::resize the (b,b,w-2b,h-2b) crop
        const size_t tight = (size_t)sw * 3;
        for (int i = 0; i < n_frames; ++i) {
            launch_warp_matrices(src + i * sframe, sw, sh, sstride, sframe, scratch, sw, sh, tight, tight * sh, d_wp + i, 1, st);
            launch_resize_linear(scratch + (size_t)border * tight + 3 * border, sw - 2 * border, sh - 2 * border, tight, 3,
                                 dst + i * dframe, sw, sh, dstride, st);
        }
        return 2 * n_frames;
    }
    return 1;
}

// ------------------------------------------------------------------------------------------------
// border_type "fade" (Stabilizer.cpp:914-978, 1070-1106).  The reference's border mask ends up all 255 (its second
// cv::rectangle repaints the whole mask), so the history frame is blended into, and updated from, the WHOLE bordered
// frame.  Three byte-wise passes around the plain warp; `bordered(x, y)` is the frame with a constant-0 margin.
//   init   : history = bordered                                   (first output only)
//   blend  : src     = cv::addWeighted(history, alpha, bordered, 1-alpha)   (SIMD path: rint(fma(h, alpha, s*beta)))
//   update : history = (uchar)((1.0f-0.1f)*history + 0.1f*stabilized)        (float32, truncation)
static __device__ __forceinline__ uint8_t bordered_byte(const uint8_t* __restrict__ f, int w, int h, size_t stride, int b, int xb, int y) {
    const int x3 = xb - 3 * b, yy = y - b;          // xb: byte column of the bordered row
    return (x3 >= 0 && x3 < 3 * w && yy >= 0 && yy < h) ? f[(size_t)yy * stride + x3] : (uint8_t)0;
}
__global__ void __launch_bounds__(256) k_fade_blend(PtrPack frames, int w, int h, size_t stride, int b, uint8_t* __restrict__ hist,
                                                     uint8_t* __restrict__ blend, size_t lane_bytes, float alpha, float beta, int init) {
    const int bw3 = 3 * (w + 2 * b), bh = h + 2 * b;
    const int xb = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (xb >= bw3 || y >= bh) return;
    const size_t o = blockIdx.z * lane_bytes + (size_t)y * bw3 + xb;
    const uint8_t s = bordered_byte(frames.p[blockIdx.z], w, h, stride, b, xb, y);
    uint8_t hv = hist[o];
    if (init) { hv = s; hist[o] = s; }
    const float t = __fmaf_rn((float)hv, alpha, __fmul_rn((float)s, beta));
    blend[o] = (uint8_t)min(max(__float2int_rn(t), 0), 255);
}
__global__ void __launch_bounds__(256) k_fade_update(uint8_t* __restrict__ hist, size_t lane_bytes, MutPtrPack outs, size_t out_stride,
                                                      int bw3, int bh) {
    const int xb = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (xb >= bw3 || y >= bh) return;
    const size_t o = blockIdx.z * lane_bytes + (size_t)y * bw3 + xb;
    const float hv = (float)hist[o], sv = (float)outs.p[blockIdx.z][(size_t)y * out_stride + xb];
    const float keep = __fsub_rn(1.0f, 0.1f);
    hist[o] = (uint8_t)(int)__fadd_rn(__fmul_rn(keep, hv), __fmul_rn(0.1f, sv));
}
void launch_fade_blend(const PtrPack& frames, int n_lanes, int w, int h, size_t stride, int b, uint8_t* hist, uint8_t* blend,
                       float alpha, float beta, bool init, cudaStream_t st) {
    const int bw3 = 3 * (w + 2 * b), bh = h + 2 * b;
    dim3 grid((bw3 + 255) / 256, bh, n_lanes);
    k_fade_blend<<<grid, 256, 0, st>>>(frames, w, h, stride, b, hist, blend, (size_t)bw3 * bh, alpha, beta, init ? 1 : 0);
}
void launch_fade_update(uint8_t* hist, const MutPtrPack& outs, size_t out_stride, int n_lanes, int w, int h, int b, cudaStream_t st) {
    const int bw3 = 3 * (w + 2 * b), bh = h + 2 * b;
    dim3 grid((bw3 + 255) / 256, bh, n_lanes);
    k_fade_update<<<grid, 256, 0, st>>>(hist, (size_t)bw3 * bh, outs, out_stride, bw3, bh);
}
