// k_pyramid.cu — analysis-image build: cv::resize(INTER_LINEAR) + cv::cvtColor(BGR2GRAY) +
// the cv::pyrDown levels cv::calcOpticalFlowPyrLK builds internally.
// Reference call sites: Stabilizer.cpp:304-305 (first frame, 480x270), :449-450 (960x540),
// :602 (prevGray up-sampling on frame 1), :611 (pyramid inside PyrLK).
// Arithmetic specification: oracle/cv_models.py resize_linear / bgr2gray / pyr_down (pinned
// bit-exact against cv2 4.13).  Pure integer work, HBM-bound: one pass over the BGR frame.
#include "kernels.h"
#include <climits>

// cv::resize INTER_LINEAR tap tables (AxisTap, tap_h, tap_v, vres): common.cuh
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

// ---------------------------------------------------------------- exact-2x fast path (1080p -> 960x540)
// One thread = 4 output pixels = one 32-bit store, from 24 source bytes x 2 rows (three aligned 64-bit loads per
// row; a warp reads 768 contiguous bytes per row).  The 2x2 box sums are taken straight off the packed BGR words
// with DP4A byte-select masks (byte k of a 12-byte group belongs to pixel k/3, channel k%3), so there is no byte
// unpacking at all.  A thread also writes the BORDER_REFLECT_101 images of its pixels into the level's
// materialised frame (rows: whole words; columns: bytes), so the padded level is complete after one launch.
static __device__ __forceinline__ uint32_t gray4_from_sums(const uint32_t* sb, const uint32_t* sg, const uint32_t* sr) {
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int b = (int)((sb[k] + 2u) >> 2), g = (int)((sg[k] + 2u) >> 2), r = (int)((sr[k] + 2u) >> 2);
        out |= (uint32_t)gray_of(b, g, r) << (8 * k);
    }
    return out;
}

// four horizontally adjacent level-0 pixels (x % 4 == 0) as one 32-bit store, plus their BORDER_REFLECT_101 images in the
// level's materialised frame (rows: whole words; columns: bytes)
static __device__ __forceinline__ void store4_with_mirrors(const GrayLevel& lv, int x, int y, uint32_t v) {
    uint8_t* row_ptr = lv.base + (ptrdiff_t)y * lv.pitch;
    *reinterpret_cast<uint32_t*>(row_ptr + x) = v;
    int ym = INT_MIN;
    if (y >= 1 && y <= VS_PAD) ym = -y;
    else if (y >= lv.h - 1 - VS_PAD && y <= lv.h - 2) ym = 2 * (lv.h - 1) - y;
    uint8_t* mrow_ptr = lv.base + (ptrdiff_t)(ym == INT_MIN ? y : ym) * lv.pitch;
    if (ym != INT_MIN) *reinterpret_cast<uint32_t*>(mrow_ptr + x) = v;
    if (x <= VS_PAD || x + 3 >= lv.w - 1 - VS_PAD) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int xx = x + k;
            int xm = INT_MIN;
            if (xx >= 1 && xx <= VS_PAD) xm = -xx;
            else if (xx >= lv.w - 1 - VS_PAD && xx <= lv.w - 2) xm = 2 * (lv.w - 1) - xx;
            if (xm != INT_MIN) {
                const uint8_t b = (uint8_t)(v >> (8 * k));
                row_ptr[xm] = b;
                if (ym != INT_MIN) mrow_ptr[xm] = b;
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_gray_half(const LaneDev* __restrict__ lanes, PtrPack src, size_t stride, int slot) {
    const LaneDev& L = lanes[blockIdx.z];
    const GrayLevel lv = L.pyr[slot].lv[0];
    const int per_row = lv.w >> 2;                                   // threads per output row
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= per_row * lv.h) return;
    const int y = id / per_row, x = (id - y * per_row) << 2;
    const uint2* r0 = reinterpret_cast<const uint2*>(src.p[blockIdx.z] + (size_t)(2 * y) * stride + 6 * x);
    const uint2* r1 = reinterpret_cast<const uint2*>(src.p[blockIdx.z] + (size_t)(2 * y + 1) * stride + 6 * x);
    uint32_t w[2][6];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const uint2 a = __ldg(r0 + k), b = __ldg(r1 + k);
        w[0][2 * k] = a.x; w[0][2 * k + 1] = a.y; w[1][2 * k] = b.x; w[1][2 * k + 1] = b.y;
    }
    // bytes of a 12-byte group (3 words): B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3 ; output pixel A = source pixels 0,1, B = 2,3
    uint32_t sb[4] = {0u, 0u, 0u, 0u}, sg[4] = {0u, 0u, 0u, 0u}, sr[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int row = 0; row < 2; ++row) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const uint32_t w0 = w[row][3 * g], w1 = w[row][3 * g + 1], w2 = w[row][3 * g + 2];
            sb[2 * g] = __dp4a(w0, 0x01000001u, sb[2 * g]);                                   // B0 + B1
            sg[2 * g] = __dp4a(w1, 0x00000001u, __dp4a(w0, 0x00000100u, sg[2 * g]));          // G0 + G1
            sr[2 * g] = __dp4a(w1, 0x00000100u, __dp4a(w0, 0x00010000u, sr[2 * g]));          // R0 + R1
            sb[2 * g + 1] = __dp4a(w2, 0x00000100u, __dp4a(w1, 0x00010000u, sb[2 * g + 1]));  // B2 + B3
            sg[2 * g + 1] = __dp4a(w2, 0x00010000u, __dp4a(w1, 0x01000000u, sg[2 * g + 1]));  // G2 + G3
            sr[2 * g + 1] = __dp4a(w2, 0x01000001u, sr[2 * g + 1]);                           // R2 + R3
        }
    }
    store4_with_mirrors(lv, x, y, gray4_from_sums(sb, sg, sr));
}

// ---------------------------------------------------------------- exact-4x fast path (2160p -> 960x540)
// cv::resize INTER_LINEAR at scale 4 samples at 4 x + 1.5: coefficients (1024, 1024) on source pixels 4x+1, 4x+2 of rows
// 4y+1, 4y+2, i.e. the rounded mean of the centre 2x2 of every 4x4 block (oracle/cv_models.py resize_linear; the
// >>4, >>16 truncations of the vertical pass are exact for these coefficients).  One thread = 4 output pixels = 48
// source bytes x 2 rows as three 128-bit loads per row (a warp reads 1536 contiguous bytes per row); the channel
// sums come off the packed words with DP4A byte-select masks as in k_gray_half.
__global__ void __launch_bounds__(256) k_gray_quarter(const LaneDev* __restrict__ lanes, PtrPack src, size_t stride, int slot) {
    const LaneDev& L = lanes[blockIdx.z];
    const GrayLevel lv = L.pyr[slot].lv[0];
    const int per_row = lv.w >> 2;
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= per_row * lv.h) return;
    const int y = id / per_row, x = (id - y * per_row) << 2;
    const uint4* r0 = reinterpret_cast<const uint4*>(src.p[blockIdx.z] + (size_t)(4 * y + 1) * stride + 12 * x);
    const uint4* r1 = reinterpret_cast<const uint4*>(src.p[blockIdx.z] + (size_t)(4 * y + 2) * stride + 12 * x);
    uint4 a[3], b[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { a[k] = __ldg(r0 + k); b[k] = __ldg(r1 + k); }
    uint32_t sb[4], sg[4], sr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        // output pixel k = 12 bytes = words 3k .. 3k+2 of the 12-word row segment: B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
        uint32_t w0[2], w1[2], w2[2];
#pragma unroll
        for (int row = 0; row < 2; ++row) {
            const uint4* v = row ? b : a;
            const uint32_t words[12] = {v[0].x, v[0].y, v[0].z, v[0].w, v[1].x, v[1].y, v[1].z, v[1].w, v[2].x, v[2].y, v[2].z, v[2].w};
            w0[row] = words[3 * k]; w1[row] = words[3 * k + 1]; w2[row] = words[3 * k + 2];
        }
        sb[k] = __dp4a(w1[1], 0x00010000u, __dp4a(w0[1], 0x01000000u, __dp4a(w1[0], 0x00010000u, __dp4a(w0[0], 0x01000000u, 0u))));   // B1 + B2
        sg[k] = __dp4a(w1[1], 0x01000001u, __dp4a(w1[0], 0x01000001u, 0u));                                                             // G1 + G2
        sr[k] = __dp4a(w2[1], 0x00000001u, __dp4a(w1[1], 0x00000100u, __dp4a(w2[0], 0x00000001u, __dp4a(w1[0], 0x00000100u, 0u))));   // R1 + R2
    }
    store4_with_mirrors(lv, x, y, gray4_from_sums(sb, sg, sr));
}

void launch_gray_resize(const LaneDev* lanes, int n_lanes, const PtrPack& src, int w, int h, size_t stride,
                        int slot, cudaStream_t st, int aw_full, int ah_full) {
    int aw = slot < 0 ? VS_FW : aw_full, ah = slot < 0 ? VS_FH : ah_full;
    dim3 grid((aw + 2 * VS_PAD + 127) / 128, ah + 2 * VS_PAD, n_lanes);
    double sx = 1.0 / ((double)aw / (double)w), sy = 1.0 / ((double)ah / (double)h);
    bool aligned = slot >= 0 && stride % 8 == 0 && aw % 4 == 0;
    for (int i = 0; i < n_lanes; ++i) aligned = aligned && ((uintptr_t)src.p[i] % 8 == 0);
    bool aligned16 = aligned && stride % 16 == 0;
    for (int i = 0; i < n_lanes; ++i) aligned16 = aligned16 && ((uintptr_t)src.p[i] % 16 == 0);
    if (w == 2 * aw && h == 2 * ah && aligned) {
        dim3 g2(((aw / 4) * ah + 255) / 256, 1, n_lanes);
        k_gray_half<<<g2, 256, 0, st>>>(lanes, src, stride, slot);
    } else if (w == 4 * aw && h == 4 * ah && aligned16) {
        dim3 g2(((aw / 4) * ah + 255) / 256, 1, n_lanes);
        k_gray_quarter<<<g2, 256, 0, st>>>(lanes, src, stride, slot);
    } else if (w == 2 * aw && h == 2 * ah)
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

void launch_upsample_small(const LaneDev* lanes, int n_lanes, int slot, cudaStream_t st, int aw, int ah) {
    dim3 grid((aw + 2 * VS_PAD + 127) / 128, ah + 2 * VS_PAD, n_lanes);
    k_upsample_small<<<grid, 128, 0, st>>>(lanes, slot, 1.0 / ((double)aw / VS_FW), 1.0 / ((double)ah / VS_FH));
}

// ---------------------------------------------------------------- cv::pyrDown
// out(x,y) = (sum_{i,j} k_i k_j src(2x+i-2, 2y+j-2) + 128) >> 8, k = [1 4 6 4 1]; the source level's
// materialised reflect-101 frame supplies the out-of-image taps.  Padded output domain again.
// Both pyrDown levels in one launch.  One CTA = one 16x16 tile of level 2; it needs a 36x36 region of level 1
// (recomputed per CTA, 27 % redundancy) and that a 76x76 region of level 0, read once from the padded level-0
// plane into shared memory.  Both passes are separable ([1 4 6 4 1] rows then columns, exact integers).  The CTA
// owns (and stores, with their reflect-101 images) the 32x32 level-1 pixels under its tile and the tile itself.
#define PD_T2 16
#define PD_R1 36
#define PD_R0 76

// two horizontally adjacent pixels (x even, both inside the level) as one 16-bit store, plus their reflect-101 images
static __device__ __forceinline__ void store2_with_mirrors(const GrayLevel& lv, int x, int y, unsigned short v2) {
    uint8_t* r = lv.base + (ptrdiff_t)y * lv.pitch;
    // A level of odd width ends in half a pair: only its first pixel exists.  The column behind it belongs to the frame and is
    // written by the owner of its reflect-101 source (x = w - 2); storing the pair there would race with that store.
    const bool pair = x + 1 < lv.w;
    if (pair) *reinterpret_cast<unsigned short*>(r + x) = v2; else r[x] = (uint8_t)v2;
    int ym = INT_MIN;
    if (y >= 1 && y <= VS_PAD) ym = -y;
    else if (y >= lv.h - 1 - VS_PAD && y <= lv.h - 2) ym = 2 * (lv.h - 1) - y;
    uint8_t* rm = lv.base + (ptrdiff_t)(ym == INT_MIN ? y : ym) * lv.pitch;
    if (ym != INT_MIN) { if (pair) *reinterpret_cast<unsigned short*>(rm + x) = v2; else rm[x] = (uint8_t)v2; }
    if (x <= VS_PAD || x + 1 >= lv.w - 1 - VS_PAD) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int xx = x + k;
            if (k == 1 && !pair) break;
            int xm = INT_MIN;
            if (xx >= 1 && xx <= VS_PAD) xm = -xx;
            else if (xx >= lv.w - 1 - VS_PAD && xx <= lv.w - 2) xm = 2 * (lv.w - 1) - xx;
            if (xm != INT_MIN) {
                const uint8_t b = (uint8_t)(v2 >> (8 * k));
                r[xm] = b;
                if (ym != INT_MIN) rm[xm] = b;
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_pyrdown2(const LaneDev* __restrict__ lanes, int slot) {
    __shared__ uint32_t L0w[PD_R0][PD_R0 / 4 + 1];    // 80-byte rows: columns x0o-2 .. x0o+77 (word aligned)
    __shared__ uint32_t H1w[PD_R0][PD_R1 / 2];        // row-pass sums of two adjacent columns per word (each <= 16 * 255)
    __shared__ __align__(4) uint8_t L1[PD_R1][PD_R1];
    __shared__ uint32_t H2w[PD_R1][PD_T2 / 2];
    const LaneDev& L = lanes[blockIdx.z];
    const GrayLevel g0 = L.pyr[slot].lv[0], g1 = L.pyr[slot].lv[1], g2 = L.pyr[slot].lv[2];
    const int tid = threadIdx.x;
    const int x2o = blockIdx.x * PD_T2, y2o = blockIdx.y * PD_T2;     // level-2 tile origin
    const int x1o = 2 * x2o - 2, y1o = 2 * y2o - 2;                   // level-1 region origin
    const int x0o = 2 * x1o - 2, y0o = 2 * y1o - 2;                   // level-0 region origin (>= -6: inside the frame)
    // 1. level-0 region: 76 rows x 20 aligned words (x0o - 2 is a multiple of 4 and >= -8, the right end stays inside
    //    the 16-pixel frame), rows clamped into the padded plane (clamped rows are never used).  Six independent
    //    loads per thread, all issued before the first store.
    {
        constexpr int RW = PD_R0 / 4 + 1, NW = PD_R0 * RW, PER = (NW + 255) / 256;
        uint32_t v[PER];
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int i = tid + 256 * k;
            const int r = min(i / RW, PD_R0 - 1), cw = i - (i / RW) * RW;
            const int gy = min(max(y0o + r, -VS_PAD), g0.h + VS_PAD - 1);
            v[k] = __ldg(reinterpret_cast<const uint32_t*>(g0.base + (ptrdiff_t)gy * g0.pitch + (x0o - 2)) + cw);
        }
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int i = tid + 256 * k;
            if (i < NW) L0w[i / RW][i - (i / RW) * RW] = v[k];
        }
    }
    __syncthreads();
    // Both passes work on PAIRS of adjacent output columns.  Row pass: the five taps of column 2m and of column 2m+1 sit in
    // three aligned words, so each sum is two DP4A with the [1 4 6 4 1] weights as byte masks.  Column pass: the two
    // 16-bit sums of a pair share a word and 1+4+6+4+1 = 16 row sums of <= 4080 stay below 65536, so one word-wide
    // multiply-add chain filters both columns; (acc + 128) >> 8 per half with one add, one shift and one mask.
    // 2. level 1, row pass: H1[r][c] = sum_i k_i L0[r][2c+2+i]   (pair m = columns 2m, 2m+1 = bytes 4m+2 .. 4m+8)
    for (int i = tid; i < PD_R0 * (PD_R1 / 2); i += 256) {
        const int r = i / (PD_R1 / 2), m = i - r * (PD_R1 / 2);
        const uint32_t wa = L0w[r][m], wb = L0w[r][m + 1], wc = L0w[r][m + 2];
        const uint32_t h0 = __dp4a(wa, 0x04010000u, __dp4a(wb, 0x00010406u, 0u));
        const uint32_t h1 = __dp4a(wb, 0x04060401u, wc & 255u);
        H1w[r][m] = h0 | (h1 << 16);
    }
    __syncthreads();
    // 3. level 1, column pass (in-image positions), stored by the owner
    for (int i = tid; i < PD_R1 * (PD_R1 / 2); i += 256) {
        const int r = i / (PD_R1 / 2), m = i - r * (PD_R1 / 2), c = 2 * m;
        const int x1 = x1o + c, y1 = y1o + r;                 // x1 is even; the second pixel of the last pair of an odd-width level lies outside (3b, store2_with_mirrors)
        if ((unsigned)x1 < (unsigned)g1.w && (unsigned)y1 < (unsigned)g1.h) {
            const uint32_t acc = H1w[2 * r][m] + 4u * H1w[2 * r + 1][m] + 6u * H1w[2 * r + 2][m] + 4u * H1w[2 * r + 3][m] + H1w[2 * r + 4][m];
            const uint32_t v = ((acc + 0x00800080u) >> 8) & 0x00FF00FFu;
            const unsigned short v2 = (unsigned short)(v | (v >> 8));           // [v(2m), v(2m+1)]
            *reinterpret_cast<unsigned short*>(&L1[r][c]) = v2;
            if (r >= 2 && r < 2 + 2 * PD_T2 && c >= 2 && c < 2 + 2 * PD_T2) store2_with_mirrors(g1, x1, y1, v2);
        }
    }
    __syncthreads();
    // 3b. positions outside the level-1 image take their BORDER_REFLECT_101 source (always inside this region)
    for (int i = tid; i < PD_R1 * PD_R1; i += 256) {
        const int r = i / PD_R1, c = i - r * PD_R1;
        const int x1 = x1o + c, y1 = y1o + r;
        if (!((unsigned)x1 < (unsigned)g1.w && (unsigned)y1 < (unsigned)g1.h)) {
            const int rx = reflect101(x1, g1.w) - x1o, ry = reflect101(y1, g1.h) - y1o;
            if ((unsigned)rx < PD_R1 && (unsigned)ry < PD_R1) L1[r][c] = L1[ry][rx];
        }
    }
    __syncthreads();
    // 4. level 2: row pass over the 36 region rows (pair m = columns 2m, 2m+1 = bytes 4m .. 4m+6 of the L1 row), then column pass
    const uint32_t (*L1w)[PD_R1 / 4] = reinterpret_cast<const uint32_t (*)[PD_R1 / 4]>(&L1[0][0]);
    for (int i = tid; i < PD_R1 * (PD_T2 / 2); i += 256) {
        const int r = i / (PD_T2 / 2), m = i - r * (PD_T2 / 2);
        const uint32_t wa = L1w[r][m], wb = L1w[r][m + 1];
        const uint32_t h0 = __dp4a(wa, 0x04060401u, wb & 255u);
        const uint32_t h1 = __dp4a(wa, 0x04010000u, __dp4a(wb, 0x00010406u, 0u));
        H2w[r][m] = h0 | (h1 << 16);
    }
    __syncthreads();
    if (tid < PD_T2 * (PD_T2 / 2)) {
        const int r = tid / (PD_T2 / 2), m = tid - r * (PD_T2 / 2);
        const int x2 = x2o + 2 * m, y2 = y2o + r;
        if (x2 < g2.w && y2 < g2.h) {
            const uint32_t acc = H2w[2 * r][m] + 4u * H2w[2 * r + 1][m] + 6u * H2w[2 * r + 2][m] + 4u * H2w[2 * r + 3][m] + H2w[2 * r + 4][m];
            const uint32_t v = ((acc + 0x00800080u) >> 8) & 0x00FF00FFu;
            store2_with_mirrors(g2, x2, y2, (unsigned short)(v | (v >> 8)));
        }
    }
}

void launch_pyrdown(const LaneDev* lanes, int n_lanes, int slot, cudaStream_t st, int aw, int ah) {
    const int w2 = ((aw + 1) / 2 + 1) / 2, h2 = ((ah + 1) / 2 + 1) / 2;
    dim3 grid((w2 + PD_T2 - 1) / PD_T2, (h2 + PD_T2 - 1) / PD_T2, n_lanes);
    k_pyrdown2<<<grid, 256, 0, st>>>(lanes, slot);
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
