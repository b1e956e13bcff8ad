// k_lk.cu — pyramidal Lucas-Kanade tracking, one warp per key point.
// Replaces cv::calcOpticalFlowPyrLK(prev, cur, pts, winSize 15x15, maxLevel 2,
// TermCriteria(COUNT+EPS, 20, 0.03)) at Stabilizer.cpp:611-619.
// Specification: oracle/cv_models.py lk_track (bit-exact against cv2 4.13, including the float32
// accumulation ORDER of OpenCV's 128-bit SIMD loop).
//
// Per pyramid level a warp stages, with aligned 32-bit loads, the 18x18 template patch of the previous
// frame and a 32x32 search region of the current frame into shared memory (one memory round trip per
// level; iterations then run entirely out of shared memory and only re-stage if the window leaves the
// region).  The 2x2 normal equations are reduced with warp shuffles.  The reference accumulates its
// sums in float32 in a fixed order; that order only matters when a partial sum can leave the exactly
// representable integer range, so the kernel first reduces sum(|term|) with integer shuffles and
//   - if it is <= 2^24, every float add in the reference is exact and the result equals the exact
//     integer sum: one warp-wide integer reduction (fast path, the common case near convergence);
//   - otherwise it replays the reference's ordered chains (4 SIMD lanes + scalar tail) from terms
//     pre-converted in parallel (slow path).
// Latency/occupancy-bound: ~200 warps per frame per lane.
#include "kernels.h"
#include <climits>
#include <cstdlib>

#define LK_WARPS 4
#define LK_NPIX (VS_WIN * VS_WIN)        // 225
#define LK_PP 18                         // template patch edge: 15 + 1 (bilinear) + 2 (Scharr)
#define LK_PW 24                         // staged template row: 18 + up to 3 bytes of alignment slack
#define LK_JR 32                         // search region edge
#define LK_PER_LANE 8                    // window pixels per lane: ceil(225 / 32)

struct LkSmem {
    uint32_t P[LK_PP][LK_PW / 4];        // template patch rows (bytes), origin (px0, ipy-1)
    short2 D[16][16];                    // Scharr (Ix,Iy) at (ipx+c, ipy+r); zero outside the image
    uint32_t J[LK_JR][LK_JR / 4];        // search region rows (bytes), origin (jx0, jy0)
    int term[3][LK_NPIX];                // per-pixel products (int, or float bits for the scalar tails)
};

static __device__ __forceinline__ void lk_weights(float a, float b, int& w00, int& w01, int& w10, int& w11) {
    const float s = 16384.f;
    w00 = __float2int_rn(__fmul_rn(__fmul_rn(1.f - a, 1.f - b), s));
    w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, 1.f - b), s));
    w10 = __float2int_rn(__fmul_rn(__fmul_rn(1.f - a, b), s));
    w11 = 16384 - w00 - w01 - w10;
}

// IDP.2A with SIGNED 16-bit weights and unsigned bytes: w11 = 16384 - w00 - w01 - w10 can come out as -1
static __device__ __forceinline__ int dp2a_lo_su(int a, unsigned b, int c) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
static __device__ __forceinline__ int dp2a_hi_su(int a, unsigned b, int c) {
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

static __device__ __forceinline__ int clamp_abs_sum(unsigned v) { return (int)min(v, 1u << 25); }

// Ordered replay of one float32 accumulation of the reference over the 15x15 window.
//   COV = true : terms are single products; SIMD lane j sees x=j then x=j+4 of each row
//   COV = false: terms are pair sums of pixels x and x+4 (x = 0..3), one per row
// `t` holds ints for x < 8 and float bits for the scalar tail x >= 8.  Lanes 0..3 run the four SIMD
// chains, lane 4 the tail; the result (tail + ((q0+q2)+(q1+q3))) is returned on every lane.
template <bool COV>
static __device__ __forceinline__ float ordered_sum(const int* t, int lane) {
    const unsigned FULL = 0xffffffffu;
    float acc = 0.f;
    if (lane < 4) {
#pragma unroll 5
        for (int y = 0; y < VS_WIN; ++y) {
            const int a = t[y * VS_WIN + lane], b = t[y * VS_WIN + lane + 4];
            if (COV) {
                acc = __fadd_rn(__int2float_rn(a), acc);
                acc = __fadd_rn(__int2float_rn(b), acc);
            } else {
                acc = __fadd_rn(acc, __int2float_rn(a + b));
            }
        }
    } else if (lane == 4) {
#pragma unroll 5
        for (int y = 0; y < VS_WIN; ++y) {
#pragma unroll
            for (int x = 8; x < VS_WIN; ++x) acc = __fadd_rn(acc, __int_as_float(t[y * VS_WIN + x]));
        }
    }
    const float q0 = __shfl_sync(FULL, acc, 0), q1 = __shfl_sync(FULL, acc, 1);
    const float q2 = __shfl_sync(FULL, acc, 2), q3 = __shfl_sync(FULL, acc, 3);
    const float tl = __shfl_sync(FULL, acc, 4);
    return __fadd_rn(tl, __fadd_rn(__fadd_rn(q0, q2), __fadd_rn(q1, q3)));
}

// The mismatch vector has its SIMD chains interleaved differently (OpenCV zips Ix/Iy): for the x
// component chains are (pixels 0&4, 1&5 -> vector qb0 lanes 0,2 ; 2&6, 3&7 -> qb1 lanes 0,2) and the
// total is tail + ((qb0[0]+qb1[0]) + (qb0[2]+qb1[2])).  Lane l<4 accumulates pixel pair (l, l+4).
static __device__ __forceinline__ float ordered_sum_b(const int* t, int lane) {
    const unsigned FULL = 0xffffffffu;
    float acc = 0.f;
    if (lane < 4) {
#pragma unroll 5
        for (int y = 0; y < VS_WIN; ++y)
            acc = __fadd_rn(acc, __int2float_rn(t[y * VS_WIN + lane] + t[y * VS_WIN + lane + 4]));
    } else if (lane == 4) {
#pragma unroll 5
        for (int y = 0; y < VS_WIN; ++y) {
#pragma unroll
            for (int x = 8; x < VS_WIN; ++x) acc = __fadd_rn(acc, __int_as_float(t[y * VS_WIN + x]));
        }
    }
    // pixel pairs: lane0=(0,4) -> qb0[.], lane1=(1,5) -> qb0[.+2], lane2=(2,6) -> qb1[.], lane3=(3,7) -> qb1[.+2]
    const float p0 = __shfl_sync(FULL, acc, 0), p1 = __shfl_sync(FULL, acc, 1);
    const float p2 = __shfl_sync(FULL, acc, 2), p3 = __shfl_sync(FULL, acc, 3);
    const float tl = __shfl_sync(FULL, acc, 4);
    return __fadd_rn(tl, __fadd_rn(__fadd_rn(p0, p2), __fadd_rn(p1, p3)));
}

__global__ void __launch_bounds__(LK_WARPS * 32, 5) k_pyr_lk(const LaneDev* __restrict__ lanes, int prev, int cur, int kp_slot, int lk_slot) {
    __shared__ LkSmem smem[LK_WARPS];
    const LaneDev& L = lanes[blockIdx.z];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pidx = blockIdx.x * LK_WARPS + warp;
    const int npts = min(*L.kpc[kp_slot], L.kp_capacity);
    if (pidx >= npts) return;                         // warp-uniform
    LkSmem& S = smem[warp];
    const unsigned FULL = 0xffffffffu;
    const float FLT_SCALE = 1.f / (1 << 20);
    const uint8_t* Pb = reinterpret_cast<const uint8_t*>(&S.P[0][0]);
    const uint8_t* Jb = reinterpret_cast<const uint8_t*>(&S.J[0][0]);

    const float2 pt = L.kpb[kp_slot][pidx];
    float nx = 0.f, ny = 0.f;                          // nextPts[ptidx]
    int status = 1;

    for (int level = VS_LEVELS - 1; level >= 0; --level) {
        const GrayLevel I = L.pyr[prev].lv[level];
        const GrayLevel Jl = L.pyr[cur].lv[level];
        const float sc = 1.f / (float)(1 << level);
        float px = __fmul_rn(pt.x, sc), py = __fmul_rn(pt.y, sc);
        float qx, qy;                                  // nextPt
        if (level == VS_LEVELS - 1) { qx = px; qy = py; }
        else { qx = __fmul_rn(nx, 2.f); qy = __fmul_rn(ny, 2.f); }
        nx = qx; ny = qy;
        px -= 7.f; py -= 7.f;
        const int ipx = (int)floorf(px), ipy = (int)floorf(py);
        if (ipx < -VS_WIN || ipx >= I.w || ipy < -VS_WIN || ipy >= I.h) {
            if (level == 0) status = 0;
            continue;
        }
        int w00, w01, w10, w11;
        lk_weights(px - (float)ipx, py - (float)ipy, w00, w01, w10, w11);
        qx -= 7.f; qy -= 7.f;

        __syncwarp();
        // ---- stage the template patch and the search region (aligned 32-bit loads, one round trip)
        const int px0 = (ipx - 1) & ~3, pofs = (ipx - 1) - px0;            // patch column offset 0..3
        int jx0 = 0, jy0 = 0;
        {
            const int inx = (int)floorf(qx), iny = (int)floorf(qy);
            jx0 = min(max((inx - 8) & ~3, -VS_PAD), Jl.w + VS_PAD - LK_JR);
            jy0 = min(max(iny - 8, -VS_PAD), Jl.h + VS_PAD - LK_JR);
            for (int i = lane; i < LK_PP * (LK_PW / 4); i += 32) {
                int r = i / (LK_PW / 4), c = i - r * (LK_PW / 4);
                S.P[r][c] = *reinterpret_cast<const uint32_t*>(I.base + (ptrdiff_t)(ipy - 1 + r) * I.pitch + px0 + 4 * c);
            }
            const uint32_t* jr = reinterpret_cast<const uint32_t*>(Jl.base + (ptrdiff_t)(jy0 + lane) * Jl.pitch + jx0);
#pragma unroll
            for (int c = 0; c < LK_JR / 4; ++c) S.J[lane][c] = jr[c];
        }
        __syncwarp();
        // ---- Scharr derivatives on the 16x16 support; the derivative plane is ZERO outside the image
        for (int i = lane; i < 256; i += 32) {
            int r = i >> 4, c = i & 15;
            int gx = 0, gy = 0;
            int ix = ipx + c, iy = ipy + r;
            if (ix >= 0 && ix < I.w && iy >= 0 && iy < I.h) {
                const uint8_t* p0 = Pb + r * LK_PW + c + pofs;
                int p00 = p0[0], p01 = p0[1], p02 = p0[2];
                int p10 = p0[LK_PW], p12 = p0[LK_PW + 2];
                int p20 = p0[2 * LK_PW], p21 = p0[2 * LK_PW + 1], p22 = p0[2 * LK_PW + 2];
                gx = 3 * (p02 - p00) + 10 * (p12 - p10) + 3 * (p22 - p20);
                gy = 3 * (p20 - p00) + 10 * (p21 - p01) + 3 * (p22 - p02);
            }
            S.D[r][c] = make_short2((short)gx, (short)gy);
        }
        __syncwarp();
        // ---- interpolated template window + covariance terms.  The lane's 8 window pixels (p = lane + 32 k) stay
        //      in registers for the whole level: Iw (5 fractional bits) and the packed (Ix, Iy) pair.
        unsigned abs11 = 0, abs12 = 0, abs22 = 0;
        int s11 = 0, s12 = 0, s22 = 0;
        int Iw_r[LK_PER_LANE], dI_r[LK_PER_LANE];
#pragma unroll
        for (int k = 0; k < LK_PER_LANE; ++k) {
            const int p = lane + 32 * k;
            Iw_r[k] = 0; dI_r[k] = 0;
            if (p < LK_NPIX) {
                const int y = p / VS_WIN, x = p - y * VS_WIN;
                const uint8_t* q0 = Pb + (y + 1) * LK_PW + x + 1 + pofs;
                const int iv = q0[0] * w00 + q0[1] * w01 + q0[LK_PW] * w10 + q0[LK_PW + 1] * w11;
                const short2 d00 = S.D[y][x], d01 = S.D[y][x + 1], d10 = S.D[y + 1][x], d11 = S.D[y + 1][x + 1];
                const int gx = (d00.x * w00 + d01.x * w01 + d10.x * w10 + d11.x * w11 + 8192) >> 14;
                const int gy = (d00.y * w00 + d01.y * w01 + d10.y * w10 + d11.y * w11 + 8192) >> 14;
                Iw_r[k] = (iv + 256) >> 9;
                dI_r[k] = (gx & 0xffff) | (gy << 16);
                const int t11 = gx * gx, t12 = gx * gy, t22 = gy * gy;
                s11 += t11; s12 += t12; s22 += t22;
                abs11 += (unsigned)t11; abs12 += (unsigned)abs(t12); abs22 += (unsigned)t22;
            }
        }
        float A[3];
        {
            const int b11 = __reduce_add_sync(FULL, clamp_abs_sum(abs11));
            const int b12 = __reduce_add_sync(FULL, clamp_abs_sum(abs12));
            const int b22 = __reduce_add_sync(FULL, clamp_abs_sum(abs22));
            const int e11 = __reduce_add_sync(FULL, s11), e12 = __reduce_add_sync(FULL, s12), e22 = __reduce_add_sync(FULL, s22);
            if (b11 <= (1 << 24) && b12 <= (1 << 24) && b22 <= (1 << 24)) {
                A[0] = (float)e11; A[1] = (float)e12; A[2] = (float)e22;
            } else {
                // a partial sum can leave the exact-integer range of float32: replay the reference's ordered chains
                __syncwarp();
#pragma unroll
                for (int k = 0; k < LK_PER_LANE; ++k) {
                    const int p = lane + 32 * k;
                    if (p < LK_NPIX) {
                        const int gx = (int)(short)(dI_r[k] & 0xffff), gy = dI_r[k] >> 16;
                        const int t11 = gx * gx, t12 = gx * gy, t22 = gy * gy;
                        const bool tail = (p - (p / VS_WIN) * VS_WIN) >= 8;
                        S.term[0][p] = tail ? __float_as_int(__int2float_rn(t11)) : t11;
                        S.term[1][p] = tail ? __float_as_int(__int2float_rn(t12)) : t12;
                        S.term[2][p] = tail ? __float_as_int(__int2float_rn(t22)) : t22;
                    }
                }
                __syncwarp();
                A[0] = (b11 <= (1 << 24)) ? (float)e11 : ordered_sum<true>(S.term[0], lane);
                A[1] = (b12 <= (1 << 24)) ? (float)e12 : ordered_sum<true>(S.term[1], lane);
                A[2] = (b22 <= (1 << 24)) ? (float)e22 : ordered_sum<true>(S.term[2], lane);
                __syncwarp();
            }
        }
        const float A11 = __fmul_rn(A[0], FLT_SCALE), A12 = __fmul_rn(A[1], FLT_SCALE), A22 = __fmul_rn(A[2], FLT_SCALE);
        float Dt = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        float dd = __fsub_rn(A11, A22);
        float minEig = __fdiv_rn(
            __fsub_rn(__fadd_rn(A22, A11),
                      __fsqrt_rn(__fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12)))),
            450.f);
        if ((double)minEig < 1e-4 || Dt < 1.1920928955078125e-7f) {
            if (level == 0) status = 0;
            continue;
        }
        Dt = __fdiv_rn(1.f, Dt);
        float pdx = 0.f, pdy = 0.f;
        uint32_t jw[LK_PER_LANE];
        int cur_inx = INT_MIN, cur_iny = INT_MIN;
#pragma unroll
        for (int k = 0; k < LK_PER_LANE; ++k) jw[k] = 0u;
        for (int j = 0; j < 20; ++j) {
            const int inx = (int)floorf(qx), iny = (int)floorf(qy);
            if (inx < -VS_WIN || inx >= Jl.w || iny < -VS_WIN || iny >= Jl.h) {
                if (level == 0) status = 0;
                break;
            }
            lk_weights(qx - (float)inx, qy - (float)iny, w00, w01, w10, w11);
            if (inx < jx0 || inx + 16 > jx0 + LK_JR || iny < jy0 || iny + 16 > jy0 + LK_JR) {
                // the window left the staged region: re-centre it (rare)
                __syncwarp();
                jx0 = min(max((inx - 8) & ~3, -VS_PAD), Jl.w + VS_PAD - LK_JR);
                jy0 = min(max(iny - 8, -VS_PAD), Jl.h + VS_PAD - LK_JR);
                const uint32_t* jr = reinterpret_cast<const uint32_t*>(Jl.base + (ptrdiff_t)(jy0 + lane) * Jl.pitch + jx0);
#pragma unroll
                for (int c = 0; c < LK_JR / 4; ++c) S.J[lane][c] = jr[c];
                __syncwarp();
                cur_inx = INT_MIN;
            }
            if (inx != cur_inx || iny != cur_iny) {
                // (re)build the lane's 2x2 tap words [q00 q01 q10 q11] for this integer window position
                const uint8_t* jb = Jb + (iny - jy0) * LK_JR + (inx - jx0);
#pragma unroll
                for (int k = 0; k < LK_PER_LANE; ++k) {
                    const int p = lane + 32 * k;
                    if (p < LK_NPIX) {
                        const int y = p / VS_WIN, x = p - y * VS_WIN;
                        const uint8_t* q0 = jb + y * LK_JR + x;
                        jw[k] = (uint32_t)q0[0] | ((uint32_t)q0[1] << 8) | ((uint32_t)q0[LK_JR] << 16) | ((uint32_t)q0[LK_JR + 1] << 24);
                    }
                }
                cur_inx = inx; cur_iny = iny;
            }
            // bilinear sample as two IDP.2A with signed 16-bit weight pairs and unsigned pixel bytes
            const int Wt = (w00 & 0xffff) | (w01 << 16), Wb = (w10 & 0xffff) | (w11 << 16);
            int sx = 0, sy = 0;
            unsigned absx = 0, absy = 0;
#pragma unroll
            for (int k = 0; k < LK_PER_LANE; ++k) {
                if (lane + 32 * k < LK_NPIX) {
                    const int jv = dp2a_hi_su(Wb, jw[k], dp2a_lo_su(Wt, jw[k], 256));
                    const int diff = (jv >> 9) - Iw_r[k];
                    const int tx = diff * (int)(short)(dI_r[k] & 0xffff), ty = diff * (dI_r[k] >> 16);
                    sx += tx; sy += ty;
                    absx += (unsigned)abs(tx); absy += (unsigned)abs(ty);
                }
            }
            const int bx = __reduce_add_sync(FULL, clamp_abs_sum(absx)), by = __reduce_add_sync(FULL, clamp_abs_sum(absy));
            const int ex = __reduce_add_sync(FULL, sx), ey = __reduce_add_sync(FULL, sy);
            float ib1, ib2;
            if (bx <= (1 << 24) && by <= (1 << 24)) {
                ib1 = (float)ex; ib2 = (float)ey;
            } else {
                __syncwarp();
#pragma unroll
                for (int k = 0; k < LK_PER_LANE; ++k) {
                    const int p = lane + 32 * k;
                    if (p < LK_NPIX) {
                        const int jv = dp2a_hi_su(Wb, jw[k], dp2a_lo_su(Wt, jw[k], 256));
                        const int diff = (jv >> 9) - Iw_r[k];
                        const int tx = diff * (int)(short)(dI_r[k] & 0xffff), ty = diff * (dI_r[k] >> 16);
                        const bool tail = (p - (p / VS_WIN) * VS_WIN) >= 8;
                        S.term[0][p] = tail ? __float_as_int(__int2float_rn(tx)) : tx;
                        S.term[1][p] = tail ? __float_as_int(__int2float_rn(ty)) : ty;
                    }
                }
                __syncwarp();
                ib1 = ordered_sum_b(S.term[0], lane);
                ib2 = ordered_sum_b(S.term[1], lane);
                __syncwarp();
            }
            float b1 = __fmul_rn(ib1, FLT_SCALE), b2 = __fmul_rn(ib2, FLT_SCALE);
            float ddx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), Dt);
            float ddy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), Dt);
            qx = __fadd_rn(qx, ddx); qy = __fadd_rn(qy, ddy);
            nx = __fadd_rn(qx, 7.f); ny = __fadd_rn(qy, 7.f);
            if ((double)ddx * (double)ddx + (double)ddy * (double)ddy <= 0.03 * 0.03) break;
            if (j > 0 && fabs((double)__fadd_rn(ddx, pdx)) < 0.01 && fabs((double)__fadd_rn(ddy, pdy)) < 0.01) {
                nx = __fsub_rn(nx, __fmul_rn(ddx, 0.5f));
                ny = __fsub_rn(ny, __fmul_rn(ddy, 0.5f));
                break;
            }
            pdx = ddx; pdy = ddy;
        }
    }
    // OpenCV re-checks the FINAL position when it computes the error measure (err is requested at Stabilizer.cpp:611-619):
    // a point whose window origin floor(nextPts - halfWin) left the level-0 image keeps its coordinates but loses its
    // status (lkpyramid.cpp, "if (status[ptidx] && err && level == 0 ...)").
    if (status) {
        const GrayLevel J0 = L.pyr[cur].lv[0];
        const int fx = (int)floorf(__fsub_rn(nx, 7.f)), fy = (int)floorf(__fsub_rn(ny, 7.f));
        if (fx < -VS_WIN || fx >= J0.w || fy < -VS_WIN || fy >= J0.h) status = 0;
    }
    if (lane == 0) {
        L.lkn[lk_slot][pidx] = make_float2(nx, ny);
        L.lks[lk_slot][pidx] = (uint8_t)status;
    }
}

// ------------------------------------------------------------------------------------------------
// TMA variant (the default): the template patch of EVERY level is fetched at the start of the kernel (they depend only on the
// key point) and the search region of a level is fetched when the level starts, each by ONE cp.async.bulk.tensor.2d issued by
// lane 0 on the padded gray plane of that level (box 48 bytes x 18 / 32 rows; the plane's materialised reflect-101 frame makes
// every needed byte in-bounds, and a 16-byte aligned box origin costs at most 15 extra columns).  Completion is an mbarrier per
// buffer.  The Scharr derivatives, the interpolated template window and the covariance sums only need the template, so they
// run while the search region is still in flight.  Arithmetic is unchanged from k_pyr_lk above.
#define LKT_PITCH 48
#define LKT_PBYTES 896                   // 18 x 48 = 864, padded to a multiple of 128
#define LKT_JBYTES (LK_JR * LKT_PITCH)   // 1536

struct LkSmemT {
    uint8_t P[VS_LEVELS][LKT_PBYTES];    // template patches, one per level; origin (pxa, ipy - 1) in plane coordinates
    uint8_t J[LKT_JBYTES];               // search region; origin (jxa, jy0)
    short2 D[16][16];
    int term[3][LK_NPIX];
    unsigned long long mbar[VS_LEVELS + 1];
    unsigned long long pad_[10];         // keep sizeof a multiple of 128
};
static_assert(sizeof(LkSmemT) % 128 == 0, "per-warp block must keep 128-byte alignment of the TMA destinations");

static __device__ __forceinline__ void lk_mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LKW_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LKD_%=;\n"
        "bra LKW_%=;\n"
        "LKD_%=:\n"
        "}\n" :: "r"(mbar), "r"(parity) : "memory");
}
static __device__ __forceinline__ void lk_tma_2d(uint32_t dst, const void* tmap, uint32_t mbar, int x, int y, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(dst), "l"(tmap), "r"(mbar), "r"(x), "r"(y) : "memory");
}

__global__ void __launch_bounds__(LK_WARPS * 32, 5) k_pyr_lk_tma(const LaneDev* __restrict__ lanes, int prev, int cur, int kp_slot, int lk_slot) {
    __shared__ __align__(128) LkSmemT smem[LK_WARPS];
    const LaneDev& L = lanes[blockIdx.z];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pidx = blockIdx.x * LK_WARPS + warp;
    const int npts = min(*L.kpc[kp_slot], L.kp_capacity);
    if (pidx >= npts) return;                         // warp-uniform
    LkSmemT& S = smem[warp];
    const unsigned FULL = 0xffffffffu;
    const float FLT_SCALE = 1.f / (1 << 20);
    const unsigned char* maps = reinterpret_cast<const unsigned char*>(L.lk_maps);
    const uint32_t s_mbar = (uint32_t)__cvta_generic_to_shared(&S.mbar[0]);
    const uint32_t s_J = (uint32_t)__cvta_generic_to_shared(&S.J[0]);

    const float2 pt = L.kpb[kp_slot][pidx];
    float nx = 0.f, ny = 0.f;                          // nextPts[ptidx]
    int status = 1;

    // ---- barriers, then the template patch of every level whose window origin is valid (the same test the level loop makes)
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k <= VS_LEVELS; ++k) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s_mbar + 8 * k));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int level = 0; level < VS_LEVELS; ++level) {
            const GrayLevel I = L.pyr[prev].lv[level];
            const float sc = 1.f / (float)(1 << level);
            const float px = __fmul_rn(pt.x, sc) - 7.f, py = __fmul_rn(pt.y, sc) - 7.f;
            const int ipx = (int)floorf(px), ipy = (int)floorf(py);
            if (ipx < -VS_WIN || ipx >= I.w || ipy < -VS_WIN || ipy >= I.h) continue;
            const void* mp = maps + (size_t)((prev * VS_LEVELS + level) * 2 + 0) * 128;
            asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" :: "l"(mp) : "memory");
            const int pxa = ((ipx - 1) + VS_PAD) & ~15;
            lk_tma_2d((uint32_t)__cvta_generic_to_shared(&S.P[level][0]), mp, s_mbar + 8 * level, pxa, ipy - 1 + VS_PAD, LK_PP * LKT_PITCH);
        }
    }
    __syncwarp();
    uint32_t j_phase = 0;

    for (int level = VS_LEVELS - 1; level >= 0; --level) {
        const GrayLevel I = L.pyr[prev].lv[level];
        const GrayLevel Jl = L.pyr[cur].lv[level];
        const float sc = 1.f / (float)(1 << level);
        float px = __fmul_rn(pt.x, sc), py = __fmul_rn(pt.y, sc);
        float qx, qy;                                  // nextPt
        if (level == VS_LEVELS - 1) { qx = px; qy = py; }
        else { qx = __fmul_rn(nx, 2.f); qy = __fmul_rn(ny, 2.f); }
        nx = qx; ny = qy;
        px -= 7.f; py -= 7.f;
        const int ipx = (int)floorf(px), ipy = (int)floorf(py);
        if (ipx < -VS_WIN || ipx >= I.w || ipy < -VS_WIN || ipy >= I.h) {
            if (level == 0) status = 0;
            continue;
        }
        int w00, w01, w10, w11;
        lk_weights(px - (float)ipx, py - (float)ipy, w00, w01, w10, w11);
        qx -= 7.f; qy -= 7.f;

        // ---- search region of this level: one bulk tensor copy, in flight while the template is processed
        const void* mj = maps + (size_t)((cur * VS_LEVELS + level) * 2 + 1) * 128;
        int jx0, jy0, jofs;
        auto fetch_J = [&](int inx, int iny) {
            jx0 = min(max((inx - 8) & ~3, -VS_PAD), Jl.w + VS_PAD - LK_JR);
            jy0 = min(max(iny - 8, -VS_PAD), Jl.h + VS_PAD - LK_JR);
            const int jxa = (jx0 + VS_PAD) & ~15;
            jofs = jx0 + VS_PAD - jxa;
            __syncwarp();                                             // every lane is done with the previous contents
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" :: "l"(mj) : "memory");
                lk_tma_2d(s_J, mj, s_mbar + 8 * VS_LEVELS, jxa, jy0 + VS_PAD, LKT_JBYTES);
            }
        };
        fetch_J((int)floorf(qx), (int)floorf(qy));
        // ---- template: wait for its patch (fetched at kernel start)
        lk_mbar_wait(s_mbar + 8 * level, 0);
        const uint8_t* Pb = &S.P[level][0];
        const int pofs = (ipx - 1 + VS_PAD) - (((ipx - 1) + VS_PAD) & ~15);        // 0..15
        // ---- Scharr derivatives on the 16x16 support; the derivative plane is ZERO outside the image
        for (int i = lane; i < 256; i += 32) {
            int r = i >> 4, c = i & 15;
            int gx = 0, gy = 0;
            int ix = ipx + c, iy = ipy + r;
            if (ix >= 0 && ix < I.w && iy >= 0 && iy < I.h) {
                const uint8_t* p0 = Pb + r * LKT_PITCH + c + pofs;
                int p00 = p0[0], p01 = p0[1], p02 = p0[2];
                int p10 = p0[LKT_PITCH], p12 = p0[LKT_PITCH + 2];
                int p20 = p0[2 * LKT_PITCH], p21 = p0[2 * LKT_PITCH + 1], p22 = p0[2 * LKT_PITCH + 2];
                gx = 3 * (p02 - p00) + 10 * (p12 - p10) + 3 * (p22 - p20);
                gy = 3 * (p20 - p00) + 10 * (p21 - p01) + 3 * (p22 - p02);
            }
            S.D[r][c] = make_short2((short)gx, (short)gy);
        }
        __syncwarp();
        unsigned abs11 = 0, abs12 = 0, abs22 = 0;
        int s11 = 0, s12 = 0, s22 = 0;
        int Iw_r[LK_PER_LANE], dI_r[LK_PER_LANE];
#pragma unroll
        for (int k = 0; k < LK_PER_LANE; ++k) {
            const int p = lane + 32 * k;
            Iw_r[k] = 0; dI_r[k] = 0;
            if (p < LK_NPIX) {
                const int y = p / VS_WIN, x = p - y * VS_WIN;
                const uint8_t* q0 = Pb + (y + 1) * LKT_PITCH + x + 1 + pofs;
                const int iv = q0[0] * w00 + q0[1] * w01 + q0[LKT_PITCH] * w10 + q0[LKT_PITCH + 1] * w11;
                const short2 d00 = S.D[y][x], d01 = S.D[y][x + 1], d10 = S.D[y + 1][x], d11 = S.D[y + 1][x + 1];
                const int gx = (d00.x * w00 + d01.x * w01 + d10.x * w10 + d11.x * w11 + 8192) >> 14;
                const int gy = (d00.y * w00 + d01.y * w01 + d10.y * w10 + d11.y * w11 + 8192) >> 14;
                Iw_r[k] = (iv + 256) >> 9;
                dI_r[k] = (gx & 0xffff) | (gy << 16);
                const int t11 = gx * gx, t12 = gx * gy, t22 = gy * gy;
                s11 += t11; s12 += t12; s22 += t22;
                abs11 += (unsigned)t11; abs12 += (unsigned)abs(t12); abs22 += (unsigned)t22;
            }
        }
        float A[3];
        {
            const int b11 = __reduce_add_sync(FULL, clamp_abs_sum(abs11));
            const int b12 = __reduce_add_sync(FULL, clamp_abs_sum(abs12));
            const int b22 = __reduce_add_sync(FULL, clamp_abs_sum(abs22));
            const int e11 = __reduce_add_sync(FULL, s11), e12 = __reduce_add_sync(FULL, s12), e22 = __reduce_add_sync(FULL, s22);
            if (b11 <= (1 << 24) && b12 <= (1 << 24) && b22 <= (1 << 24)) {
                A[0] = (float)e11; A[1] = (float)e12; A[2] = (float)e22;
            } else {
                __syncwarp();
#pragma unroll
                for (int k = 0; k < LK_PER_LANE; ++k) {
                    const int p = lane + 32 * k;
                    if (p < LK_NPIX) {
                        const int gx = (int)(short)(dI_r[k] & 0xffff), gy = dI_r[k] >> 16;
                        const int t11 = gx * gx, t12 = gx * gy, t22 = gy * gy;
                        const bool tail = (p - (p / VS_WIN) * VS_WIN) >= 8;
                        S.term[0][p] = tail ? __float_as_int(__int2float_rn(t11)) : t11;
                        S.term[1][p] = tail ? __float_as_int(__int2float_rn(t12)) : t12;
                        S.term[2][p] = tail ? __float_as_int(__int2float_rn(t22)) : t22;
                    }
                }
                __syncwarp();
                A[0] = (b11 <= (1 << 24)) ? (float)e11 : ordered_sum<true>(S.term[0], lane);
                A[1] = (b12 <= (1 << 24)) ? (float)e12 : ordered_sum<true>(S.term[1], lane);
                A[2] = (b22 <= (1 << 24)) ? (float)e22 : ordered_sum<true>(S.term[2], lane);
                __syncwarp();
            }
        }
        // the search region must have landed before this level is left by any path
        lk_mbar_wait(s_mbar + 8 * VS_LEVELS, j_phase);
        j_phase ^= 1;
        const uint8_t* Jb = &S.J[0];
        const float A11 = __fmul_rn(A[0], FLT_SCALE), A12 = __fmul_rn(A[1], FLT_SCALE), A22 = __fmul_rn(A[2], FLT_SCALE);
        float Dt = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        float dd = __fsub_rn(A11, A22);
        float minEig = __fdiv_rn(
            __fsub_rn(__fadd_rn(A22, A11),
                      __fsqrt_rn(__fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12)))),
            450.f);
        if ((double)minEig < 1e-4 || Dt < 1.1920928955078125e-7f) {
            if (level == 0) status = 0;
            continue;
        }
        Dt = __fdiv_rn(1.f, Dt);
        float pdx = 0.f, pdy = 0.f;
        uint32_t jw[LK_PER_LANE];
        int cur_inx = INT_MIN, cur_iny = INT_MIN;
#pragma unroll
        for (int k = 0; k < LK_PER_LANE; ++k) jw[k] = 0u;
        for (int j = 0; j < 20; ++j) {
            const int inx = (int)floorf(qx), iny = (int)floorf(qy);
            if (inx < -VS_WIN || inx >= Jl.w || iny < -VS_WIN || iny >= Jl.h) {
                if (level == 0) status = 0;
                break;
            }
            lk_weights(qx - (float)inx, qy - (float)iny, w00, w01, w10, w11);
            if (inx < jx0 || inx + 16 > jx0 + LK_JR || iny < jy0 || iny + 16 > jy0 + LK_JR) {
                // the window left the staged region: re-centre it (rare)
                fetch_J(inx, iny);
                lk_mbar_wait(s_mbar + 8 * VS_LEVELS, j_phase);
                j_phase ^= 1;
                cur_inx = INT_MIN;
            }
            if (inx != cur_inx || iny != cur_iny) {
                const uint8_t* jb = Jb + (iny - jy0) * LKT_PITCH + (inx - jx0) + jofs;
#pragma unroll
                for (int k = 0; k < LK_PER_LANE; ++k) {
                    const int p = lane + 32 * k;
                    if (p < LK_NPIX) {
                        const int y = p / VS_WIN, x = p - y * VS_WIN;
                        const uint8_t* q0 = jb + y * LKT_PITCH + x;
                        jw[k] = (uint32_t)q0[0] | ((uint32_t)q0[1] << 8) | ((uint32_t)q0[LKT_PITCH] << 16) | ((uint32_t)q0[LKT_PITCH + 1] << 24);
                    }
                }
                cur_inx = inx; cur_iny = iny;
            }
            const int Wt = (w00 & 0xffff) | (w01 << 16), Wb = (w10 & 0xffff) | (w11 << 16);
            int sx = 0, sy = 0;
            unsigned absx = 0, absy = 0;
#pragma unroll
            for (int k = 0; k < LK_PER_LANE; ++k) {
                if (lane + 32 * k < LK_NPIX) {
                    const int jv = dp2a_hi_su(Wb, jw[k], dp2a_lo_su(Wt, jw[k], 256));
                    const int diff = (jv >> 9) - Iw_r[k];
                    const int tx = diff * (int)(short)(dI_r[k] & 0xffff), ty = diff * (dI_r[k] >> 16);
                    sx += tx; sy += ty;
                    absx += (unsigned)abs(tx); absy += (unsigned)abs(ty);
                }
            }
            const int bx = __reduce_add_sync(FULL, clamp_abs_sum(absx)), by = __reduce_add_sync(FULL, clamp_abs_sum(absy));
            const int ex = __reduce_add_sync(FULL, sx), ey = __reduce_add_sync(FULL, sy);
            float ib1, ib2;
            if (bx <= (1 << 24) && by <= (1 << 24)) {
                ib1 = (float)ex; ib2 = (float)ey;
            } else {
                __syncwarp();
#pragma unroll
                for (int k = 0; k < LK_PER_LANE; ++k) {
                    const int p = lane + 32 * k;
                    if (p < LK_NPIX) {
                        const int jv = dp2a_hi_su(Wb, jw[k], dp2a_lo_su(Wt, jw[k], 256));
                        const int diff = (jv >> 9) - Iw_r[k];
                        const int tx = diff * (int)(short)(dI_r[k] & 0xffff), ty = diff * (dI_r[k] >> 16);
                        const bool tail = (p - (p / VS_WIN) * VS_WIN) >= 8;
                        S.term[0][p] = tail ? __float_as_int(__int2float_rn(tx)) : tx;
                        S.term[1][p] = tail ? __float_as_int(__int2float_rn(ty)) : ty;
                    }
                }
                __syncwarp();
                ib1 = ordered_sum_b(S.term[0], lane);
                ib2 = ordered_sum_b(S.term[1], lane);
                __syncwarp();
            }
            float b1 = __fmul_rn(ib1, FLT_SCALE), b2 = __fmul_rn(ib2, FLT_SCALE);
            float ddx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), Dt);
            float ddy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), Dt);
            qx = __fadd_rn(qx, ddx); qy = __fadd_rn(qy, ddy);
            nx = __fadd_rn(qx, 7.f); ny = __fadd_rn(qy, 7.f);
            if ((double)ddx * (double)ddx + (double)ddy * (double)ddy <= 0.03 * 0.03) break;
            if (j > 0 && fabs((double)__fadd_rn(ddx, pdx)) < 0.01 && fabs((double)__fadd_rn(ddy, pdy)) < 0.01) {
                nx = __fsub_rn(nx, __fmul_rn(ddx, 0.5f));
                ny = __fsub_rn(ny, __fmul_rn(ddy, 0.5f));
                break;
            }
            pdx = ddx; pdy = ddy;
        }
    }
    if (status) {
        const GrayLevel J0 = L.pyr[cur].lv[0];
        const int fx = (int)floorf(__fsub_rn(nx, 7.f)), fy = (int)floorf(__fsub_rn(ny, 7.f));
        if (fx < -VS_WIN || fx >= J0.w || fy < -VS_WIN || fy >= J0.h) status = 0;
    }
    if (lane == 0) {
        L.lkn[lk_slot][pidx] = make_float2(nx, ny);
        L.lks[lk_slot][pidx] = (uint8_t)status;
    }
}

// host: the tensor maps of one lane's pyramids, [slot][level][template box, search box], over the padded planes
#include <cuda.h>
typedef CUresult (*LkEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
bool lk_encode_maps(const LaneDev& hl, void* out_host) {
    static LkEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (LkEncodeFn)p;
        cudaGetLastError();
    }
    if (!fn || getenv("VS_LK_PLAIN")) return false;
    CUtensorMap* maps = reinterpret_cast<CUtensorMap*>(out_host);
    for (int s = 0; s < VS_PYR_SLOTS; ++s)
        for (int l = 0; l < VS_LEVELS; ++l) {
            const GrayLevel& g = hl.pyr[s].lv[l];
            void* plane = g.base - (ptrdiff_t)VS_PAD * g.pitch - VS_PAD;
            if ((uintptr_t)plane % 16 != 0 || g.pitch % 16 != 0) return false;
            cuuint64_t dims[2] = {(cuuint64_t)g.pitch, (cuuint64_t)(g.h + 2 * VS_PAD)};
            cuuint64_t strides[1] = {(cuuint64_t)g.pitch};
            cuuint32_t es[2] = {1, 1};
            for (int k = 0; k < 2; ++k) {
                cuuint32_t box[2] = {LKT_PITCH, (cuuint32_t)(k ? LK_JR : LK_PP)};
                if (fn(&maps[(s * VS_LEVELS + l) * 2 + k], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, plane, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                    return false;
            }
        }
    return true;
}

void launch_pyr_lk(const LaneDev* lanes, int n_lanes, int prev, int cur, int max_pts, int kp_slot, int lk_slot, cudaStream_t st, bool tma) {
    if (max_pts <= 0) return;
    dim3 grid((max_pts + LK_WARPS - 1) / LK_WARPS, 1, n_lanes);
    if (tma) k_pyr_lk_tma<<<grid, LK_WARPS * 32, 0, st>>>(lanes, prev, cur, kp_slot, lk_slot);
    else k_pyr_lk<<<grid, LK_WARPS * 32, 0, st>>>(lanes, prev, cur, kp_slot, lk_slot);
}
