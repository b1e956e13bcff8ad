// k_roll.cu — vs::RollCorrection::autoCorrectRoll (reference src/RollCorrection.cpp:16-155) as
// sm_100a kernels: the frame is downscaled and turned to gray, Canny edges and a standard Hough transform find the long
// straight lines, their mean angle drives an exponentially smoothed roll estimate, and the full-resolution frame is rotated by
// it.  Everything, including the scalar angle logic, stays on the device: a call is a fixed launch sequence on one stream.
//
// The reference calls cv::cuda:: functions (resize, cvtColor, CannyEdgeDetector, HoughLinesDetector, buildWarpAffineMaps,
// remap).  Parity is defined against the CPU functions of the same OpenCV — cv::resize, cv::cvtColor, cv::Canny, cv::HoughLines,
// cv::remap — driven by the reference's own RollCorrection.cpp (oracle/ref_stages.py); specification of each step:
//   resize + gray   11-bit fixed-point bilinear (oracle/cv_models.py resize_linear) then (3735 B + 19235 G + 9798 R + 16384) >> 15
//   Canny           Sobel 3x3 (BORDER_REPLICATE), L1 magnitude, NMS with the TG22 fixed-point sector test, hysteresis
//   HoughLines      float tables sin/cos(theta) / rho, r = cvRound(x cos + y sin), 4-neighbour local maxima above the threshold,
//                   ordered by (votes desc, accumulator index asc)
//   rotate          maps = invertAffine(getRotationMatrix2D) narrowed to float, evaluated in float, then cv::remap INTER_LINEAR
//                   BORDER_REPLICATE: coordinates rounded to 1/32 px, 15-bit weights, (sum + 16384) >> 15
// Integer / index work is bit-exact; the one floating-point residual is cos/sin of the roll angle in double (CUDA's libdevice vs
// glibc, <= 2 ulp), which can move a map coefficient by one float ulp.
#include "kernels.h"
#include "roll.h"

#include <cmath>
#include <cstdio>
#include <vector>

#define RL_MAX_LINES 4096

// ---------------------------------------------------------------------------------------------- 1. small gray image
// cv::resize(INTER_LINEAR) of the BGR frame to (sw, sh), then cv::cvtColor(BGR2GRAY)
__global__ void __launch_bounds__(256) k_roll_small_gray(const uint8_t* __restrict__ src, int w, int h, size_t stride,
                                                         uint8_t* __restrict__ gray, int sw, int sh, double scx, double scy) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= sw || y >= sh) return;
    const AxisTap tx = tap_h(x, w, scx), ty = tap_v(y, h, scy);
    const uint8_t* r0 = src + (size_t)ty.s0 * stride;
    const uint8_t* r1 = src + (size_t)ty.s1 * stride;
    int v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int h0 = r0[3 * tx.s0 + c] * tx.a0 + r0[3 * tx.s1 + c] * tx.a1;
        const int h1 = r1[3 * tx.s0 + c] * tx.a0 + r1[3 * tx.s1 + c] * tx.a1;
        v[c] = vres(h0, h1, ty.a0, ty.a1);
    }
    gray[(size_t)y * sw + x] = (uint8_t)((3735 * v[0] + 19235 * v[1] + 9798 * v[2] + 16384) >> 15);
}

// ---------------------------------------------------------------------------------------------- 2. Canny: gradient + NMS
// map: 1 = not an edge, 0 = weak candidate, 2 = strong edge (OpenCV's convention, canny.cpp).  Strong pixels are queued.
#define CN_TX 32
#define CN_TY 8
__global__ void __launch_bounds__(CN_TX * CN_TY) k_roll_canny_nms(const uint8_t* __restrict__ gray, int w, int h, int low, int high,
                                                                   unsigned int* __restrict__ map, int* __restrict__ queue,
                                                                   int* __restrict__ counters) {
    __shared__ int sg[CN_TY + 4][CN_TX + 4];       // gray, 2-pixel halo, BORDER_REPLICATE
    __shared__ int sm[CN_TY + 2][CN_TX + 2];       // |dx| + |dy|, 1-pixel halo, 0 outside the image
    const int tid = threadIdx.y * CN_TX + threadIdx.x;
    const int x0 = blockIdx.x * CN_TX, y0 = blockIdx.y * CN_TY;
    for (int i = tid; i < (CN_TY + 4) * (CN_TX + 4); i += CN_TX * CN_TY) {
        const int ly = i / (CN_TX + 4), lx = i - ly * (CN_TX + 4);
        const int gx = min(max(x0 + lx - 2, 0), w - 1), gy = min(max(y0 + ly - 2, 0), h - 1);
        sg[ly][lx] = gray[(size_t)gy * w + gx];
    }
    __syncthreads();
    for (int i = tid; i < (CN_TY + 2) * (CN_TX + 2); i += CN_TX * CN_TY) {
        const int ly = i / (CN_TX + 2), lx = i - ly * (CN_TX + 2);
        const int gx = x0 + lx - 1, gy = y0 + ly - 1;
        int m = 0;
        if (gx >= 0 && gx < w && gy >= 0 && gy < h) {
            // the gradient of pixel (gx, gy) uses replicate-clamped neighbours of THAT pixel: at the image border the halo
            // column / row of sg already holds the clamped value
            const int cy = ly + 1, cx = lx + 1;
            const int dx = (sg[cy - 1][cx + 1] - sg[cy - 1][cx - 1]) + 2 * (sg[cy][cx + 1] - sg[cy][cx - 1]) + (sg[cy + 1][cx + 1] - sg[cy + 1][cx - 1]);
            const int dy = (sg[cy + 1][cx - 1] - sg[cy - 1][cx - 1]) + 2 * (sg[cy + 1][cx] - sg[cy - 1][cx]) + (sg[cy + 1][cx + 1] - sg[cy - 1][cx + 1]);
            m = abs(dx) + abs(dy);
        }
        sm[ly][lx] = m;
    }
    __syncthreads();
    const int gx = x0 + threadIdx.x, gy = y0 + threadIdx.y;
    if (gx >= w || gy >= h) return;
    const int cy = threadIdx.y + 2, cx = threadIdx.x + 2;
    const int xs = (sg[cy - 1][cx + 1] - sg[cy - 1][cx - 1]) + 2 * (sg[cy][cx + 1] - sg[cy][cx - 1]) + (sg[cy + 1][cx + 1] - sg[cy + 1][cx - 1]);
    const int ys = (sg[cy + 1][cx - 1] - sg[cy - 1][cx - 1]) + 2 * (sg[cy + 1][cx] - sg[cy - 1][cx]) + (sg[cy + 1][cx + 1] - sg[cy - 1][cx + 1]);
    const int my = threadIdx.y + 1, mx = threadIdx.x + 1;
    const int m = sm[my][mx];
    unsigned int res = 1;
    if (m > low) {
        const int TG22 = 13573;                   // tan(22.5 deg) in Q15
        const int ax = abs(xs), ay = abs(ys) << 15;
        const int tg22x = ax * TG22;
        bool keep;
        if (ay < tg22x) keep = m > sm[my][mx - 1] && m >= sm[my][mx + 1];
        else {
            const int tg67x = tg22x + (ax << 16);
            if (ay > tg67x) keep = m > sm[my - 1][mx] && m >= sm[my + 1][mx];
            else {
                const int s = ((xs ^ ys) < 0) ? -1 : 1;
                keep = m > sm[my - 1][mx - s] && m > sm[my + 1][mx + s];
            }
        }
        if (keep) res = m > high ? 2u : 0u;
    }
    const int p = gy * w + gx;
    map[p] = res;
    if (res == 2u) queue[atomicAdd(&counters[0], 1)] = p;      // counters: [0] produced (tail), [1] claimed (head), [2] completed
}

// ---------------------------------------------------------------------------------------------- 3. Canny: hysteresis
// Work queue over the strong pixels: an item's thread walks its weak neighbourhood depth-first, claiming pixels with an atomic
// compare-and-swap (each pixel is claimed exactly once, so the final map does not depend on the order).  Threads that find the
// queue momentarily empty wait only for items produced by threads that are running; the kernel ends when every produced item is
// complete.
// head, reservation cursor and seed count all start at the number of strong pixels the NMS kernel queued
__global__ void k_roll_seed_counters(int* __restrict__ counters) {
    if (threadIdx.x == 0) { const int n0 = counters[0]; counters[1] = n0; counters[3] = n0; counters[6] = n0; }
}
#define HY_STACK 24
// counters: [0] tail (items published), [1] head (next spilled item to claim), [2] items completed, [3] slot reservation cursor,
//           [6] number of strong seeds.  [1], [3], [6] start at the seed count (copied on the stream after the NMS kernel).
static __device__ __forceinline__ void hyst_walk(unsigned int* __restrict__ map, int w, int h, int* __restrict__ queue, int* __restrict__ counters, int seed) {
    int stack[HY_STACK];
    int sp = 0;
    stack[sp++] = seed;
    while (sp) {
        const int q = stack[--sp];
        const int y = q / w, x = q - y * w;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int dx = (k == 0 || k == 3 || k == 5) ? -1 : (k == 1 || k == 6) ? 0 : 1;
            const int dy = k < 3 ? -1 : (k < 5 ? 0 : 1);
            const int nx = x + dx, ny = y + dy;
            if ((unsigned)nx >= (unsigned)w || (unsigned)ny >= (unsigned)h) continue;
            const int n = ny * w + nx;
            if (map[n] == 0u && atomicCAS(&map[n], 0u, 2u) == 0u) {
                if (sp < HY_STACK) stack[sp++] = n;
                else {
                    const int t = atomicAdd(&counters[3], 1);                  // reserve a slot, fill it, then extend the tail in order
                    ((volatile int*)queue)[t] = n;
                    __threadfence();
                    while (atomicCAS(&counters[0], t, t + 1) != t) {}
                }
            }
        }
    }
}
__global__ void __launch_bounds__(128) k_roll_hysteresis(unsigned int* __restrict__ map, int w, int h, int* __restrict__ queue,
                                                          int* __restrict__ counters, int capacity) {
    volatile int* vc = counters;
    // phase 1: the strong seeds, statically shared out
    const int n0 = counters[6];
    int mine = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n0; i += gridDim.x * blockDim.x) {
        hyst_walk(map, w, h, queue, counters, queue[i]);
        ++mine;
    }
    if (mine) { __threadfence(); atomicAdd(&counters[2], mine); }
    // phase 2: items spilled by walks whose stack overflowed (rare), claimed dynamically until every published item is complete
    while (true) {
        const int idx = atomicAdd(&counters[1], 1);
        if (idx >= capacity) return;
        while (true) {
            const int done = vc[2];
            __threadfence();
            const int tail = vc[0];
            if (idx < tail) break;
            if (done == tail) return;
            __nanosleep(500);
        }
        __threadfence();
        hyst_walk(map, w, h, queue, counters, ((volatile int*)queue)[idx]);
        __threadfence();
        atomicAdd(&counters[2], 1);
    }
}

// ---------------------------------------------------------------------------------------------- 4. edge image + edge list
__global__ void __launch_bounds__(256) k_roll_edge_list(const unsigned int* __restrict__ map, int n, int w, uint8_t* __restrict__ edges,
                                                        unsigned int* __restrict__ list, int* __restrict__ counters) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool e = p < n && map[p] == 2u;
    if (p < n) edges[p] = e ? 255 : 0;
    const unsigned b = __ballot_sync(0xFFFFFFFFu, e);
    int base = 0;
    const int lane = threadIdx.x & 31;
    if (lane == 0 && b) base = atomicAdd(&counters[4], __popc(b));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (e) {
        const int y = p / w, x = p - y * w;
        list[base + __popc(b & ((1u << lane) - 1))] = (unsigned)x | ((unsigned)y << 16);
    }
}

// ---------------------------------------------------------------------------------------------- 5. Hough accumulation
// one CTA per angle: the rho histogram lives in shared memory
__global__ void __launch_bounds__(256) k_roll_hough(const unsigned int* __restrict__ list, const int* __restrict__ counters,
                                                    const float* __restrict__ tab_sin, const float* __restrict__ tab_cos, int numrho,
                                                    int* __restrict__ accum) {
    extern __shared__ int hist[];
    const int n = blockIdx.x;
    for (int i = threadIdx.x; i < numrho + 2; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const int count = counters[4];
    const float c = tab_cos[n], s = tab_sin[n];
    const int half = (numrho - 1) / 2;
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        const unsigned v = list[i];
        const float x = (float)(v & 0xFFFFu), y = (float)(v >> 16);
        const int r = __float2int_rn(__fadd_rn(__fmul_rn(x, c), __fmul_rn(y, s))) + half;
        atomicAdd(&hist[r + 1], 1);
    }
    __syncthreads();
    int* row = accum + (size_t)(n + 1) * (numrho + 2);
    for (int i = threadIdx.x; i < numrho + 2; i += blockDim.x) row[i] = hist[i];
}

// ---------------------------------------------------------------------------------------------- 6. lines, angle, rotation set-up
// 6a. local maxima of the accumulator above the threshold, one CTA per angle (rows n - 1 and n + 1 come from L2): keys
//     (votes << 32) | ~accumulator-index, so that one descending sort gives cv::HoughLines' order (votes desc, index asc)
__global__ void __launch_bounds__(256) k_roll_maxima(const int* __restrict__ accum, int numrho, int threshold,
                                                     unsigned long long* __restrict__ cand, int* __restrict__ counters) {
    const int n = blockIdx.x, pitch = numrho + 2;
    const int* row = accum + (size_t)(n + 1) * pitch;
    for (int r = threadIdx.x; r < numrho; r += blockDim.x) {
        const int v = row[r + 1];
        if (v > threshold && v > row[r] && v >= row[r + 2] && v > row[r + 1 - pitch] && v >= row[r + 1 + pitch]) {
            const int k = atomicAdd(&counters[5], 1);
            const unsigned base = (unsigned)((n + 1) * pitch + r + 1);
            if (k < RL_MAX_LINES) cand[k] = ((unsigned long long)(unsigned)v << 32) | (unsigned)(0xFFFFFFFFu - base);
        }
    }
}

// 6b. lines ordered like cv::HoughLines -> the reference's angle statistics and exponential smoothing (RollCorrection.cpp:92-140,
//     double arithmetic in source order) -> rotation matrix, its inverse, float coefficients.  One CTA.
__global__ void __launch_bounds__(1024) k_roll_lines(const unsigned long long* __restrict__ cand, const int* __restrict__ counters,
                                                     int numangle, int numrho, int threshold, float rho,
                                                     float theta, RollParamsDev prm, int w, int h, RollState* __restrict__ st,
                                                     float* __restrict__ lines_out) {
    __shared__ unsigned long long keys[RL_MAX_LINES];
    const int pitch = numrho + 2;
    const int n_found = counters[5];
    for (int i = threadIdx.x; i < min(n_found, RL_MAX_LINES); i += blockDim.x) keys[i] = cand[i];
    __syncthreads();
    const int found = n_found;
    const int nl = min(found, RL_MAX_LINES);
    int np2 = 1;
    while (np2 < nl) np2 <<= 1;
    for (int i = nl + threadIdx.x; i < np2; i += blockDim.x) keys[i] = 0ull;
    __syncthreads();
    // bitonic sort, descending: (votes desc, accumulator index asc) - hough_cmp_gt
    for (int k = 2; k <= np2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < np2; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long a = keys[i], b = keys[l];
                    const bool desc = (i & k) == 0;
                    if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[l] = a; }
                }
            }
            __syncthreads();
        }
    if (threadIdx.x != 0) return;
    // ---- RollCorrection.cpp:92-140
    double sum = 0.0;
    int count = 0;
    const float scale = 0.5f;
    for (int i = 0; i < nl; ++i) {
        const unsigned base = 0xFFFFFFFFu - (unsigned)(keys[i] & 0xFFFFFFFFull);
        const int n = (int)(base / (unsigned)pitch) - 1, r = (int)(base - (unsigned)(n + 1) * (unsigned)pitch) - 1;
        const float lrho = __fmul_rn(__fsub_rn((float)r, __fmul_rn((float)(numrho - 1), scale)), rho);
        const float ltheta = __fadd_rn(0.0f, __fmul_rn((float)n, theta));
        if (lines_out && i < RL_MAX_LINES) { lines_out[3 * i] = lrho; lines_out[3 * i + 1] = ltheta; lines_out[3 * i + 2] = (float)(unsigned)(keys[i] >> 32); }
        const double deg = __dsub_rn(__ddiv_rn(__dmul_rn((double)ltheta, 180.0), 3.1415926535897932384626433832795), 90.0);
        if (deg >= prm.angle_filter_min && deg <= prm.angle_filter_max) { sum = __dadd_rn(sum, deg); ++count; }
    }
    // The reference reads linesMat.total() entries of the 2 x N matrix cv::cuda::HoughLinesDetector returns: after the N lines come
    // the N vote cells, which decode to theta = 0, i.e. -90 degrees each (mini_cv_cuda.hpp).  They only count when the filter
    // band reaches -90.
    if (nl > 0 && -90.0 >= prm.angle_filter_min && -90.0 <= prm.angle_filter_max)
        for (int i = 0; i < nl; ++i) { sum = __dadd_rn(sum, -90.0); ++count; }
    double ang = st->first ? 0.0 : st->smoothed_angle;
    st->first = 0;
    if (count == 0) ang = __dmul_rn(ang, prm.angle_decay);
    else {
        const double detected = __ddiv_rn(sum, (double)count);
        double na = __dadd_rn(__dmul_rn(prm.angle_smoothing_alpha, detected), __dmul_rn(__dsub_rn(1.0, prm.angle_smoothing_alpha), ang));
        double diff = __dsub_rn(na, ang);
        if (fabs(diff) > prm.max_angle_change_deg && prm.max_angle_change_deg > 0.0) {
            diff = diff > 0 ? prm.max_angle_change_deg : -prm.max_angle_change_deg;
            na = __dadd_rn(ang, diff);
        }
        ang = na;
    }
    st->smoothed_angle = ang;
    st->n_lines = found;
    // cv::getRotationMatrix2D(center = (w / 2.0f, h / 2.0f), angle, 1.0), cv::invertAffineTransform, narrowing to float
    const double cx = (double)((float)w / 2.0f), cy = (double)((float)h / 2.0f);
    const double a = __dmul_rn(ang, __ddiv_rn(3.1415926535897932384626433832795, 180.0));
    const double alpha = cos(a), beta = sin(a);
    const double m0 = alpha, m1 = beta, m2 = __dsub_rn(__dmul_rn(__dsub_rn(1.0, alpha), cx), __dmul_rn(beta, cy));
    const double m3 = -beta, m4 = alpha, m5 = __dadd_rn(__dmul_rn(beta, cx), __dmul_rn(__dsub_rn(1.0, alpha), cy));
    double D = __dsub_rn(__dmul_rn(m0, m4), __dmul_rn(m1, m3));
    D = D != 0 ? __ddiv_rn(1.0, D) : 0.0;
    const double A11 = __dmul_rn(m4, D), A22 = __dmul_rn(m0, D), A12 = __dmul_rn(-m1, D), A21 = __dmul_rn(-m3, D);
    const double b1 = __dsub_rn(__dmul_rn(-A11, m2), __dmul_rn(A12, m5));
    const double b2 = __dsub_rn(__dmul_rn(-A21, m2), __dmul_rn(A22, m5));
    st->coef[0] = (float)A11; st->coef[1] = (float)A12; st->coef[2] = (float)b1;
    st->coef[3] = (float)A21; st->coef[4] = (float)A22; st->coef[5] = (float)b2;
}

// ---------------------------------------------------------------------------------------------- 7. rotation (cv::remap)
// map(x, y) = (c0 x + c1 y) + c2 in float; coordinates rounded to 1/32 px; INTER_LINEAR with 15-bit weights; BORDER_REPLICATE.
// One thread = four output pixels = twelve output bytes.
static __device__ __forceinline__ uint32_t roll_pixel(const uint8_t* __restrict__ src, int w, int h, size_t stride, const float* c, int x, int y) {
    const float fx = __fadd_rn(__fadd_rn(__fmul_rn(c[0], (float)x), __fmul_rn(c[1], (float)y)), c[2]);
    const float fy = __fadd_rn(__fadd_rn(__fmul_rn(c[3], (float)x), __fmul_rn(c[4], (float)y)), c[5]);
    const int X = __float2int_rn(__fmul_rn(fx, 32.f)), Y = __float2int_rn(__fmul_rn(fy, 32.f));
    const int sx = min(max(X >> 5, -32768), 32767), sy = min(max(Y >> 5, -32768), 32767);
    const int ax = X & 31, ay = Y & 31;
    const int x0 = min(max(sx, 0), w - 1), x1 = min(max(sx + 1, 0), w - 1);
    const int y0 = min(max(sy, 0), h - 1), y1 = min(max(sy + 1, 0), h - 1);
    const uint8_t* r0 = src + (size_t)y0 * stride;
    const uint8_t* r1 = src + (size_t)y1 * stride;
    const int w00 = (32 - ax) * (32 - ay), w01 = ax * (32 - ay), w10 = (32 - ax) * ay, w11 = ax * ay;
    uint32_t out = 0;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const int v = r0[3 * x0 + ch] * w00 + r0[3 * x1 + ch] * w01 + r1[3 * x0 + ch] * w10 + r1[3 * x1 + ch] * w11;
        out |= (uint32_t)((v + 512) >> 10) << (8 * ch);          // (v * 32 + 16384) >> 15
    }
    return out;
}
__global__ void __launch_bounds__(256) k_roll_remap(const uint8_t* __restrict__ src, int w, int h, size_t stride, uint8_t* __restrict__ dst,
                                                    size_t dstride, const RollState* __restrict__ st, int vec) {
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x >= w) return;
    float c[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) c[i] = st->coef[i];
    uint8_t* o = dst + (size_t)y * dstride + 3 * (size_t)x;
    if (vec && x + 3 < w) {
        const uint32_t p0 = roll_pixel(src, w, h, stride, c, x, y), p1 = roll_pixel(src, w, h, stride, c, x + 1, y);
        const uint32_t p2 = roll_pixel(src, w, h, stride, c, x + 2, y), p3 = roll_pixel(src, w, h, stride, c, x + 3, y);
        uint32_t* ow = reinterpret_cast<uint32_t*>(o);
        ow[0] = p0 | (p1 << 24);
        ow[1] = (p1 >> 8) | (p2 << 16);
        ow[2] = (p2 >> 16) | (p3 << 8);
    } else {
        for (int k = 0; k < 4 && x + k < w; ++k) {
            const uint32_t p = roll_pixel(src, w, h, stride, c, x + k, y);
            o[3 * k] = (uint8_t)p; o[3 * k + 1] = (uint8_t)(p >> 8); o[3 * k + 2] = (uint8_t)(p >> 16);
        }
    }
}

// ============================================================================================== host side
#define RCUDA(x)                                                                       \
    do {                                                                               \
        cudaError_t e__ = (x);                                                         \
        if (e__ != cudaSuccess) return vs_set_cuda_error(e__, #x, __FILE__, __LINE__); \
    } while (0)

vs_status RollCorrector::create(const vs_roll_params& p, int device, RollCorrector** out) {
    *out = nullptr;
    if (p.canny_aperture != 3) return vs_set_error(VS_ERR_UNSUPPORTED, "roll correction: only canny_aperture 3 (the reference default) is built");
    if (!(p.hough_rho > 0.f) || !(p.hough_theta > 0.f)) return vs_set_error(VS_ERR_INVALID_ARG, "roll correction: hough_rho and hough_theta must be positive");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return vs_set_error(VS_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
    }
    if (device < 0 || device >= count) return vs_set_error(VS_ERR_INVALID_ARG, "bad device ordinal");
    cudaDeviceProp prop;
    RCUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return vs_set_error(VS_ERR_NO_DEVICE, "kernels are built for sm_100a (B200) only");
    RollCorrector* r = new RollCorrector();
    r->p_ = p;
    r->device_ = device;
    // cv::HoughLines: angle table (computeNumangle / createTrigTable, hough.cpp) - derived from the parameters only
    const double min_theta = 0.0, max_theta = 3.1415926535897932384626433832795, step = (double)p.hough_theta;
    int numangle = (int)std::floor((max_theta - min_theta) / step) + 1;
    if (numangle > 1 && std::fabs(3.1415926535897932384626433832795 - (numangle - 1) * step) < step / 2) --numangle;
    r->numangle_ = numangle;
    const float irho = 1.0f / p.hough_rho;
    std::vector<float> ts(numangle), tc(numangle);
    float ang = (float)min_theta;
    for (int n = 0; n < numangle; ++n) {
        ts[n] = (float)(std::sin((double)ang) * irho);
        tc[n] = (float)(std::cos((double)ang) * irho);
        ang += p.hough_theta;
    }
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc((void**)&r->d_tab_, sizeof(float) * 2 * numangle);
    if (e == cudaSuccess) e = cudaMemcpy(r->d_tab_, ts.data(), sizeof(float) * numangle, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(r->d_tab_ + numangle, tc.data(), sizeof(float) * numangle, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc((void**)&r->d_state_, sizeof(RollState));
    if (e == cudaSuccess) e = cudaMalloc((void**)&r->d_lines_, sizeof(float) * 3 * RL_MAX_LINES);
    if (e == cudaSuccess) e = cudaMalloc((void**)&r->d_cand_, sizeof(unsigned long long) * RL_MAX_LINES);
    RollState init{};
    init.first = 1;
    if (e == cudaSuccess) e = cudaMemcpy(r->d_state_, &init, sizeof(init), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { delete r; return vs_set_cuda_error(e, "roll correction set-up", __FILE__, __LINE__); }
    *out = r;
    return VS_OK;
}

RollCorrector::~RollCorrector() {
    cudaSetDevice(device_);
    cudaDeviceSynchronize();
    for (void* p : {(void*)d_tab_, (void*)d_state_, (void*)d_lines_, (void*)d_cand_, (void*)d_gray_, (void*)d_map_, (void*)d_queue_, (void*)d_list_,
                    (void*)d_edges_, (void*)d_counters_, (void*)d_accum_, (void*)d_in_, (void*)d_out_})
        if (p) cudaFree(p);
}

vs_status RollCorrector::ensure(int w, int h) {
    if (w == w_ && h == h_) return VS_OK;
    for (void** p : {(void**)&d_gray_, (void**)&d_map_, (void**)&d_queue_, (void**)&d_list_, (void**)&d_edges_, (void**)&d_counters_,
                     (void**)&d_accum_, (void**)&d_in_, (void**)&d_out_})
        if (*p) { cudaFree(*p); *p = nullptr; }
    w_ = w; h_ = h;
    sw_ = (int)(w * p_.scale_factor);                      // static_cast<int>(input.cols * params.scaleFactor), RollCorrection.cpp:35-38
    sh_ = (int)(h * p_.scale_factor);
    if (sw_ <= 0 || sh_ <= 0) { sw_ = w; sh_ = h; }        // "scaleFactor might be 1 or 0? Edge case => skip" :43-46
    if (sw_ > 65535 || sh_ > 65535) return vs_set_error(VS_ERR_INVALID_ARG, "roll correction: analysis image too large");
    const size_t n = (size_t)sw_ * sh_;
    numrho_ = (int)std::lround(((double)(2 * (sw_ + sh_)) + 1) / (double)p_.hough_rho);     // cvRound(((max_rho - min_rho) + 1) / rho)
    RCUDA(cudaMalloc((void**)&d_gray_, n));
    RCUDA(cudaMalloc((void**)&d_map_, n * sizeof(unsigned int)));
    RCUDA(cudaMalloc((void**)&d_queue_, n * sizeof(int)));
    RCUDA(cudaMalloc((void**)&d_list_, n * sizeof(unsigned int)));
    RCUDA(cudaMalloc((void**)&d_edges_, n));
    RCUDA(cudaMalloc((void**)&d_counters_, 8 * sizeof(int)));
    RCUDA(cudaMalloc((void**)&d_accum_, (size_t)(numangle_ + 2) * (numrho_ + 2) * sizeof(int)));
    RCUDA(cudaMemset(d_accum_, 0, (size_t)(numangle_ + 2) * (numrho_ + 2) * sizeof(int)));
    if ((size_t)(numrho_ + 2) * sizeof(int) > 200 * 1024) return vs_set_error(VS_ERR_UNSUPPORTED, "roll correction: rho histogram does not fit in shared memory");
    static bool attr[64] = {};
    if (device_ < 64 && !attr[device_]) {
        RCUDA(cudaFuncSetAttribute(k_roll_hough, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr[device_] = true;
    }
    return VS_OK;
}

// the launch sequence; d_src / d_dst are device frames (tight or strided rows)
vs_status RollCorrector::correct_device(const uint8_t* d_src, int w, int h, size_t stride, uint8_t* d_dst, size_t dstride, cudaStream_t st) {
    if (!d_src || !d_dst || w < 4 || h < 4) return vs_set_error(VS_ERR_INVALID_ARG, "roll correction: bad frame");
    if (stride == 0) stride = (size_t)w * 3;
    if (dstride == 0) dstride = (size_t)w * 3;
    if (stride < (size_t)w * 3 || dstride < (size_t)w * 3) return vs_set_error(VS_ERR_INVALID_ARG, "stride smaller than a row");
    RCUDA(cudaSetDevice(device_));
    VS_TRY_ROLL(ensure(w, h));
    const int n = sw_ * sh_;
    const double scx = 1.0 / ((double)sw_ / (double)w), scy = 1.0 / ((double)sh_ / (double)h);
    RCUDA(cudaMemsetAsync(d_counters_, 0, 8 * sizeof(int), st));
    k_roll_small_gray<<<dim3((sw_ + 255) / 256, sh_), 256, 0, st>>>(d_src, w, h, stride, d_gray_, sw_, sh_, scx, scy);
    // cv::Canny: low = cvFloor(low_thresh), high = cvFloor(high_thresh) for the L1 norm
    const int low = (int)std::floor(std::min(p_.canny_threshold_low, p_.canny_threshold_high));
    const int high = (int)std::floor(std::max(p_.canny_threshold_low, p_.canny_threshold_high));
    k_roll_canny_nms<<<dim3((sw_ + CN_TX - 1) / CN_TX, (sh_ + CN_TY - 1) / CN_TY), dim3(CN_TX, CN_TY), 0, st>>>(d_gray_, sw_, sh_, low, high, d_map_, d_queue_, d_counters_);
    k_roll_seed_counters<<<1, 32, 0, st>>>(d_counters_);
    k_roll_hysteresis<<<32, 128, 0, st>>>(d_map_, sw_, sh_, d_queue_, d_counters_, n);
    k_roll_edge_list<<<(n + 255) / 256, 256, 0, st>>>(d_map_, n, sw_, d_edges_, d_list_, d_counters_);
    k_roll_hough<<<numangle_, 256, (numrho_ + 2) * sizeof(int), st>>>(d_list_, d_counters_, d_tab_, d_tab_ + numangle_, numrho_, d_accum_);
    RollParamsDev prm{p_.angle_filter_min, p_.angle_filter_max, p_.angle_smoothing_alpha, p_.angle_decay, p_.max_angle_change_deg};
    k_roll_maxima<<<numangle_, 256, 0, st>>>(d_accum_, numrho_, p_.hough_threshold, d_cand_, d_counters_);
    k_roll_lines<<<1, 1024, 0, st>>>(d_cand_, d_counters_, numangle_, numrho_, p_.hough_threshold, p_.hough_rho, p_.hough_theta, prm, w, h, d_state_, d_lines_);
    const int vec = ((uintptr_t)d_dst % 4 == 0 && dstride % 4 == 0) ? 1 : 0;
    k_roll_remap<<<dim3(((w + 3) / 4 + 255) / 256, h), 256, 0, st>>>(d_src, w, h, stride, d_dst, dstride, d_state_, vec);
    RCUDA(cudaGetLastError());
    launches_ += 9;
    return VS_OK;
}

vs_status RollCorrector::correct_host(const uint8_t* src, int w, int h, size_t stride, uint8_t* dst, size_t dstride) {
    if (!src || !dst || w < 4 || h < 4) return vs_set_error(VS_ERR_INVALID_ARG, "roll correction: bad frame");
    if (stride == 0) stride = (size_t)w * 3;
    if (dstride == 0) dstride = (size_t)w * 3;
    RCUDA(cudaSetDevice(device_));
    VS_TRY_ROLL(ensure(w, h));
    const size_t tight = (size_t)w * 3;
    if (!d_in_) RCUDA(cudaMalloc((void**)&d_in_, tight * h));
    if (!d_out_) RCUDA(cudaMalloc((void**)&d_out_, tight * h));
    RCUDA(cudaMemcpy2DAsync(d_in_, tight, src, stride, tight, h, cudaMemcpyHostToDevice, 0));
    VS_TRY_ROLL(correct_device(d_in_, w, h, tight, d_out_, tight, 0));
    RCUDA(cudaMemcpy2DAsync(dst, dstride, d_out_, tight, tight, h, cudaMemcpyDeviceToHost, 0));
    RCUDA(cudaStreamSynchronize(0));
    return VS_OK;
}

vs_status RollCorrector::reset() {
    RollState init{};
    init.first = 1;
    RCUDA(cudaSetDevice(device_));
    RCUDA(cudaDeviceSynchronize());
    RCUDA(cudaMemcpy(d_state_, &init, sizeof(init), cudaMemcpyHostToDevice));
    return VS_OK;
}

vs_status RollCorrector::state(RollState* out, int* n_edges) {
    RCUDA(cudaSetDevice(device_));
    RCUDA(cudaDeviceSynchronize());
    RCUDA(cudaMemcpy(out, d_state_, sizeof(RollState), cudaMemcpyDeviceToHost));
    if (n_edges) {
        *n_edges = 0;
        if (d_counters_) RCUDA(cudaMemcpy(n_edges, d_counters_ + 4, sizeof(int), cudaMemcpyDeviceToHost));
    }
    return VS_OK;
}

vs_status RollCorrector::debug(uint8_t* gray, uint8_t* edges, float* lines, int cap_lines) {
    RCUDA(cudaSetDevice(device_));
    RCUDA(cudaDeviceSynchronize());
    if (!d_gray_) return vs_set_error(VS_ERR_INVALID_ARG, "roll correction: no frame processed yet");
    if (gray) RCUDA(cudaMemcpy(gray, d_gray_, (size_t)sw_ * sh_, cudaMemcpyDeviceToHost));
    if (edges) RCUDA(cudaMemcpy(edges, d_edges_, (size_t)sw_ * sh_, cudaMemcpyDeviceToHost));
    if (lines && cap_lines > 0) RCUDA(cudaMemcpy(lines, d_lines_, sizeof(float) * 3 * (cap_lines < RL_MAX_LINES ? cap_lines : RL_MAX_LINES), cudaMemcpyDeviceToHost));
    return VS_OK;
}
