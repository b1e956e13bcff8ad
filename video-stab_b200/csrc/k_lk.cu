// k_lk.cu — pyramidal Lucas-Kanade tracking, one warp per key point.
// Replaces cv::calcOpticalFlowPyrLK(prev, cur, pts, winSize 15x15, maxLevel 2,
// TermCriteria(COUNT+EPS, 20, 0.03)) at Stabilizer.cpp:611-619.
// Specification: oracle/cv_models.py lk_track (bit-exact against cv2 4.13, including the float32
// accumulation ORDER of OpenCV's 128-bit SIMD loop, which this kernel reproduces with ordered
// per-lane chains).  Patches are staged in shared memory; the 2x2 normal equations are reduced
// with warp shuffles.  Latency/occupancy-bound: ~200 warps per frame per lane.
#include "kernels.h"

#define LK_WARPS 4
#define LK_NPIX (VS_WIN * VS_WIN)        // 225
#define LK_PP 18                         // prev patch edge: 15 + 1 (bilinear) + 2 (Scharr)
#define LK_JP 16                         // next patch edge: 15 + 1

struct LkSmem {
    uint8_t P[LK_PP][LK_PP + 2];         // prev-level patch, origin (ipx-1, ipy-1)
    short2 D[LK_JP][LK_JP];              // Scharr (Ix,Iy) at (ipx+c, ipy+r); zero outside the image
    uint8_t J[LK_JP][LK_JP];             // next-level patch, origin (inx, iny)
    short Iw[LK_NPIX];                   // interpolated I window   (5 fractional bits)
    short2 dI[LK_NPIX];                  // interpolated derivative window
    int diff[LK_NPIX];                   // J - I per window pixel
};

static __device__ __forceinline__ void lk_weights(float a, float b, int& w00, int& w01, int& w10, int& w11) {
    const float s = 16384.f;
    w00 = __float2int_rn(__fmul_rn(__fmul_rn(1.f - a, 1.f - b), s));
    w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, 1.f - b), s));
    w10 = __float2int_rn(__fmul_rn(__fmul_rn(1.f - a, b), s));
    w11 = 16384 - w00 - w01 - w10;
}

__global__ void __launch_bounds__(LK_WARPS * 32) k_pyr_lk(const LaneDev* __restrict__ lanes, int prev, int cur) {
    __shared__ LkSmem smem[LK_WARPS];
    const LaneDev& L = lanes[blockIdx.z];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pidx = blockIdx.x * LK_WARPS + warp;
    const int npts = min(*L.kp_count, L.kp_capacity);
    if (pidx >= npts) return;                         // warp-uniform
    LkSmem& S = smem[warp];
    const unsigned FULL = 0xffffffffu;
    const float FLT_SCALE = 1.f / (1 << 20);

    const float2 pt = L.kp[pidx];
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

        __syncwarp();
        // ---- stage the prev patch (always inside image + reflect-101 frame)
        for (int i = lane; i < LK_PP * LK_PP; i += 32) {
            int r = i / LK_PP, c = i - r * LK_PP;
            S.P[r][c] = I.base[(ptrdiff_t)(ipy - 1 + r) * I.pitch + (ipx - 1 + c)];
        }
        __syncwarp();
        // ---- Scharr derivatives on the 16x16 support; the derivative plane is ZERO outside the image
        for (int i = lane; i < LK_JP * LK_JP; i += 32) {
            int r = i >> 4, c = i & 15;
            int gx = 0, gy = 0;
            int ix = ipx + c, iy = ipy + r;
            if (ix >= 0 && ix < I.w && iy >= 0 && iy < I.h) {
                int p00 = S.P[r][c], p01 = S.P[r][c + 1], p02 = S.P[r][c + 2];
                int p10 = S.P[r + 1][c], p12 = S.P[r + 1][c + 2];
                int p20 = S.P[r + 2][c], p21 = S.P[r + 2][c + 1], p22 = S.P[r + 2][c + 2];
                gx = 3 * (p02 - p00) + 10 * (p12 - p10) + 3 * (p22 - p20);
                gy = 3 * (p20 - p00) + 10 * (p21 - p01) + 3 * (p22 - p02);
            }
            S.D[r][c] = make_short2((short)gx, (short)gy);
        }
        __syncwarp();
        // ---- interpolated template window
        for (int p = lane; p < LK_NPIX; p += 32) {
            int y = p / VS_WIN, x = p - y * VS_WIN;
            int iv = S.P[y + 1][x + 1] * w00 + S.P[y + 1][x + 2] * w01 + S.P[y + 2][x + 1] * w10 + S.P[y + 2][x + 2] * w11;
            short2 d00 = S.D[y][x], d01 = S.D[y][x + 1], d10 = S.D[y + 1][x], d11 = S.D[y + 1][x + 1];
            int gx = d00.x * w00 + d01.x * w01 + d10.x * w10 + d11.x * w11;
            int gy = d00.y * w00 + d01.y * w01 + d10.y * w10 + d11.y * w11;
            S.Iw[p] = (short)((iv + 256) >> 9);
            S.dI[p] = make_short2((short)((gx + 8192) >> 14), (short)((gy + 8192) >> 14));
        }
        __syncwarp();
        // ---- covariance, ordered chains: lane c<15 -> sum k=c/5 (A11,A12,A22), chain j=c%5
        //      (j<4: SIMD lane j sees x=j then x=j+4 of each row; j==4: scalar tail x=8..14)
        float acc = 0.f;
        if (lane < 15) {
            const int k = lane / 5, j = lane - 5 * k;
            for (int y = 0; y < VS_WIN; ++y) {
                if (j < 4) {
#pragma unroll
                    for (int hh = 0; hh < 8; hh += 4) {
                        short2 d = S.dI[y * VS_WIN + j + hh];
                        int pr = (k == 0) ? d.x * d.x : (k == 1) ? d.x * d.y : d.y * d.y;
                        acc = __fadd_rn(__int2float_rn(pr), acc);
                    }
                } else {
#pragma unroll
                    for (int x = 8; x < VS_WIN; ++x) {
                        short2 d = S.dI[y * VS_WIN + x];
                        int pr = (k == 0) ? d.x * d.x : (k == 1) ? d.x * d.y : d.y * d.y;
                        acc = __fadd_rn(acc, __int2float_rn(pr));
                    }
                }
            }
        }
        float A[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float q0 = __shfl_sync(FULL, acc, 5 * k + 0), q1 = __shfl_sync(FULL, acc, 5 * k + 1);
            float q2 = __shfl_sync(FULL, acc, 5 * k + 2), q3 = __shfl_sync(FULL, acc, 5 * k + 3);
            float t = __shfl_sync(FULL, acc, 5 * k + 4);
            A[k] = __fmul_rn(__fadd_rn(t, __fadd_rn(__fadd_rn(q0, q2), __fadd_rn(q1, q3))), FLT_SCALE);
        }
        const float A11 = A[0], A12 = A[1], A22 = A[2];
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
        qx -= 7.f; qy -= 7.f;
        float pdx = 0.f, pdy = 0.f;
        for (int j = 0; j < 20; ++j) {
            const int inx = (int)floorf(qx), iny = (int)floorf(qy);
            if (inx < -VS_WIN || inx >= Jl.w || iny < -VS_WIN || iny >= Jl.h) {
                if (level == 0) status = 0;
                break;
            }
            lk_weights(qx - (float)inx, qy - (float)iny, w00, w01, w10, w11);
            __syncwarp();
            for (int i = lane; i < LK_JP * LK_JP; i += 32) {
                int r = i >> 4, c = i & 15;
                S.J[r][c] = Jl.base[(ptrdiff_t)(iny + r) * Jl.pitch + (inx + c)];
            }
            __syncwarp();
            for (int p = lane; p < LK_NPIX; p += 32) {
                int y = p / VS_WIN, x = p - y * VS_WIN;
                int jv = S.J[y][x] * w00 + S.J[y][x + 1] * w01 + S.J[y + 1][x] * w10 + S.J[y + 1][x + 1] * w11;
                S.diff[p] = ((jv + 256) >> 9) - (int)S.Iw[p];
            }
            __syncwarp();
            // ---- mismatch vector, ordered chains: lanes 0..7 = (v=lane>>2, l=lane&3) pair sums of
            //      pixels x=2v+(l>>1) and x+4, component l&1; lanes 8,9 = scalar tails (gx, gy)
            float bacc = 0.f;
            if (lane < 8) {
                const int x0 = 2 * (lane >> 2) + ((lane & 3) >> 1), comp = lane & 1;
                for (int y = 0; y < VS_WIN; ++y) {
                    int i0 = y * VS_WIN + x0;
                    short2 g0 = S.dI[i0], g1 = S.dI[i0 + 4];
                    int v = S.diff[i0] * (comp ? g0.y : g0.x) + S.diff[i0 + 4] * (comp ? g1.y : g1.x);
                    bacc = __fadd_rn(bacc, __int2float_rn(v));
                }
            } else if (lane < 10) {
                const int comp = lane - 8;
                for (int y = 0; y < VS_WIN; ++y) {
#pragma unroll
                    for (int x = 8; x < VS_WIN; ++x) {
                        int i0 = y * VS_WIN + x;
                        short2 g = S.dI[i0];
                        bacc = __fadd_rn(bacc, __int2float_rn(S.diff[i0] * (comp ? g.y : g.x)));
                    }
                }
            }
            float hi = __shfl_down_sync(FULL, bacc, 4);
            float qs = __fadd_rn(bacc, hi);                      // lanes 0..3: qb0 + qb1
            float qs0 = __shfl_sync(FULL, qs, 0), qs1 = __shfl_sync(FULL, qs, 1);
            float qs2 = __shfl_sync(FULL, qs, 2), qs3 = __shfl_sync(FULL, qs, 3);
            float s1 = __shfl_sync(FULL, bacc, 8), s2 = __shfl_sync(FULL, bacc, 9);
            float b1 = __fmul_rn(__fadd_rn(s1, __fadd_rn(qs0, qs2)), FLT_SCALE);
            float b2 = __fmul_rn(__fadd_rn(s2, __fadd_rn(qs1, qs3)), FLT_SCALE);
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
    if (lane == 0) {
        L.lk_next[pidx] = make_float2(nx, ny);
        L.lk_status[pidx] = (uint8_t)status;
    }
}

void launch_pyr_lk(const LaneDev* lanes, int n_lanes, int prev, int cur, int max_pts, cudaStream_t st) {
    if (max_pts <= 0) return;
    dim3 grid((max_pts + LK_WARPS - 1) / LK_WARPS, 1, n_lanes);
    k_pyr_lk<<<grid, LK_WARPS * 32, 0, st>>>(lanes, prev, cur);
}
