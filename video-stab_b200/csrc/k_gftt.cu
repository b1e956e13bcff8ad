// k_gftt.cu — Shi-Tomasi corner detection ("good features to track").
// Replaces cv::goodFeaturesToTrack(gray, maxCorners, quality, minDistance, noArray, blockSize=3)
// at Stabilizer.cpp:355-357 (first frame, user parameters) and :740-744 (every 2nd frame,
// hard-coded 200 / 0.02 / 15.0 / 3).
// Specification: oracle/cv_models.py min_eigen_map + gftt (ordered corner list bit-exact vs cv2 4.13
// on its baseline code path).  Two kernels:
//   k_eig_nms   Sobel -> products -> 3x3 box -> min eigenvalue -> 3x3 non-max suppression -> warp-ballot compaction
//               into 64-bit sort keys + global max, one launch, shared-memory tiled (the eigenvalue map never
//               reaches HBM)
//   k_select    per lane: keys <= quality * max dropped, radix-select of the strongest chunk, shared-memory bitonic
//               sort, and the order-dependent min-distance greedy pass done 32 candidates at a time by one warp;
//               leaves the detection counters zeroed for the next detection
// The float map must match OpenCV's op order exactly (no FMA contraction: explicit _rn intrinsics).
#include "kernels.h"

#define EIG_TW 64
#define EIG_TH 16
#define SEL_THREADS 1024
#define SEL_CHUNK_MAX 4096
#define SEL_CHUNK_FIRST 960

static __device__ __forceinline__ GrayLevel gftt_src(const LaneDev& L, int slot) {
    return slot < 0 ? L.small0 : L.pyr[slot].lv[0];
}

// ------------------------------------------------------------------------------------ k_eig_nms
// Sobel -> products -> 3x3 box -> min eigenvalue -> 3x3 non-max suppression -> candidate keys, one launch, the
// eigenvalue map never leaves shared memory.  One CTA = 64x16 candidate positions; it needs eigenvalues on 66x18,
// product maps on 68x20 and gray on 70x22 (staged with two aligned word loads per thread from the padded level).
// The box filter's BORDER_REFLECT_101 acts on the PRODUCT maps, so a product position outside the image takes the
// value computed AT its reflected coordinate (the cross term changes sign under reflection of the gray image, so the
// level's gray frame cannot stand in for it).  The quality threshold needs the global maximum, which is only known
// after this kernel: every strict-positive local maximum is emitted and k_select drops keys <= threshold
// ("zero everything <= thr, then dilate" and "dilate, then require > thr" select the same pixels).
#define EN_GW 72                       // staged gray row: x0-4 .. x0+67 (18 aligned words)
#define EN_GH (EIG_TH + 6)             // y0-3 .. y0+18
#define EN_PW (EIG_TW + 4)             // products: x0-2 .. x0+65
#define EN_PH (EIG_TH + 4)
#define EN_EW (EIG_TW + 2)             // eigenvalues: x0-1 .. x0+64
#define EN_EH (EIG_TH + 2)

__global__ void __launch_bounds__(256) k_eig_nms(const LaneDev* __restrict__ lanes, int slot, int gen) {   // gen: scratch set
    __shared__ uint32_t sgw[EN_GH][EN_GW / 4];
    __shared__ double sxx[EN_PH][EN_PW], sxy[EN_PH][EN_PW], syy[EN_PH][EN_PW];   // products, already widened: OpenCV's box filter sums them in double
    // horizontal Sobel passes (stage 2a) and the eigenvalue tile (stage 3 on) share storage: the former are dead by then
    __shared__ __align__(16) unsigned char s_u[EN_GH * EN_PW * 6];
    float (*const sT)[EN_PW] = reinterpret_cast<float (*)[EN_PW]>(s_u);                                   // [1 2 1] * scale row pass
    short (*const sD)[EN_PW] = reinterpret_cast<short (*)[EN_PW]>(s_u + EN_GH * EN_PW * 4);               // [-1 0 1] row pass
    float (*const se)[EN_EW] = reinterpret_cast<float (*)[EN_EW]>(s_u);
    static_assert(EN_EH * EN_EW * 4 <= EN_GH * EN_PW * 6 && EN_PW % 4 == 0, "shared eigenvalue tile");
    __shared__ unsigned int smax;
    __shared__ int s_count, s_base;
    __shared__ int s_warp_off[8];
    const LaneDev& L = lanes[blockIdx.z];
    const DetView D = det_view(L, gen, 0);
    const GrayLevel G = gftt_src(L, slot);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * EIG_TW, y0 = blockIdx.y * EIG_TH;
    const float f1 = (float)(1.0 / (4.0 * 3.0 * 255.0));
    const float f0 = 2.f * f1;
    if (tid == 0) { smax = 0u; s_count = 0; }

    // 1. gray (rows clamped into the padded plane: rows that far out are never used)
    {
        constexpr int NW = EN_GH * (EN_GW / 4);
        uint32_t v[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int i = min(tid + 256 * k, NW - 1);
            const int r = i / (EN_GW / 4), cw = i - r * (EN_GW / 4);
            const int gy = min(max(y0 - 3 + r, -VS_PAD), G.h + VS_PAD - 1);
            v[k] = __ldg(reinterpret_cast<const uint32_t*>(G.base + (ptrdiff_t)gy * G.pitch + (x0 - 4)) + cw);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int i = tid + 256 * k;
            if (i < NW) sgw[i / (EN_GW / 4)][i - (i / (EN_GW / 4)) * (EN_GW / 4)] = v[k];
        }
    }
    __syncthreads();
    // 2a. separable Sobel, row passes on every staged gray row (gray column of product column c is c + 2):
    //       sD = a[c+3] - a[c+1] (exact integer), sT = (a[c+1]*f1 + a[c+2]*f0) + a[c+3]*f1 (OpenCV's order, left to right).
    //     One task = 4 columns of one row from two aligned words; 8 int->float conversions instead of 12 per 4 pixels,
    //     and each row pass is computed once instead of once per vertical neighbour.
    for (int i = tid; i < EN_GH * (EN_PW / 4); i += 256) {
        const int r = i / (EN_PW / 4), g4 = i - r * (EN_PW / 4);
        const uint32_t w0 = sgw[r][g4], w1 = sgw[r][g4 + 1];
        const int b1 = (w0 >> 8) & 255, b2 = (w0 >> 16) & 255, b3 = w0 >> 24, b4 = w1 & 255, b5 = (w1 >> 8) & 255, b6 = (w1 >> 16) & 255;
        const float a1 = __fmul_rn((float)b1, f1), a3 = __fmul_rn((float)b3, f1), a4 = __fmul_rn((float)b4, f1), a6 = __fmul_rn((float)b6, f1);
        const float a2 = __fmul_rn((float)b2, f1), a5 = __fmul_rn((float)b5, f1);
        const float c2 = __fmul_rn((float)b2, f0), c3 = __fmul_rn((float)b3, f0), c4 = __fmul_rn((float)b4, f0), c5 = __fmul_rn((float)b5, f0);
        float4 t;
        t.x = __fadd_rn(__fadd_rn(a1, c2), a3);
        t.y = __fadd_rn(__fadd_rn(a2, c3), a4);
        t.z = __fadd_rn(__fadd_rn(a3, c4), a5);
        t.w = __fadd_rn(__fadd_rn(a4, c5), a6);
        *reinterpret_cast<float4*>(&sT[r][4 * g4]) = t;
        short4 d;
        d.x = (short)(b3 - b1); d.y = (short)(b4 - b2); d.z = (short)(b5 - b3); d.w = (short)(b6 - b4);
        *reinterpret_cast<short4*>(&sD[r][4 * g4]) = d;
    }
    __syncthreads();
    // 2b. products at in-image positions (x = x0-2+c, y = y0-2+r): column passes over rows r, r+1, r+2
    for (int i = tid; i < EN_PH * EN_PW; i += 256) {
        const int r = i / EN_PW, c = i - r * EN_PW;
        const int x = x0 - 2 + c, y = y0 - 2 + r;
        if ((unsigned)x >= (unsigned)G.w || (unsigned)y >= (unsigned)G.h) continue;
        // Dx: column pass [1 2 1]*scale as (S0+S2)*f1 + S1*f0 (S0+S2 is an exact small integer either way)
        const float dx = __fadd_rn(__fmul_rn((float)((int)sD[r][c] + (int)sD[r + 2][c]), f1), __fmul_rn((float)sD[r + 1][c], f0));
        // Dy: column pass [-1 0 1]
        const float dy = __fsub_rn(sT[r + 2][c], sT[r][c]);
        sxx[r][c] = (double)__fmul_rn(dx, dx);
        sxy[r][c] = (double)__fmul_rn(dx, dy);
        syy[r][c] = (double)__fmul_rn(dy, dy);
    }
    __syncthreads();
    // 2c. positions outside the image: BORDER_REFLECT_101 of the product maps (the source lies inside this tile)
    if (x0 == 0 || y0 == 0 || x0 + EIG_TW + 2 > G.w || y0 + EIG_TH + 2 > G.h) {
        for (int i = tid; i < EN_PH * EN_PW; i += 256) {
            const int r = i / EN_PW, c = i - r * EN_PW;
            const int x = x0 - 2 + c, y = y0 - 2 + r;
            if ((unsigned)x < (unsigned)G.w && (unsigned)y < (unsigned)G.h) continue;
            const int rc = reflect101(x, G.w) - (x0 - 2), rr = reflect101(y, G.h) - (y0 - 2);
            if ((unsigned)rc < EN_PW && (unsigned)rr < EN_PH) {
                sxx[r][c] = sxx[rr][rc]; sxy[r][c] = sxy[rr][rc]; syy[r][c] = syy[rr][rc];
            }
        }
        __syncthreads();
    }
    // 3. eigenvalues on the tile + 1 (x = x0-1+c, y = y0-1+r); outside the image: -1 (never a maximum, never >= a candidate).
    //    The 3x3 box sum is separable and exact in double (the sum of nine floats of similar magnitude), so any
    //    association gives OpenCV's value: one thread walks a column segment of EN_SEG rows with a sliding window of
    //    three row sums - 12 double adds per position instead of 27, and no float->double conversion here.
    unsigned int lmax = 0u;
    {
        constexpr int EN_SEG = 6, NSEG = EN_EH / EN_SEG;          // 18 rows = 3 segments; 66 columns x 3 = 198 threads
        static_assert(NSEG * EN_SEG == EN_EH && NSEG * EN_EW <= 256, "segment map");
        const int seg = tid / EN_EW, c = tid - seg * EN_EW;
        if (seg < NSEG) {
            const int x = x0 - 1 + c;
            double w0x = 0, w0y = 0, w0z = 0, w1x = 0, w1y = 0, w1z = 0;
#pragma unroll
            for (int rr = 0; rr < EN_SEG + 2; ++rr) {
                const int pr = seg * EN_SEG + rr;                 // product row
                const double rx = (sxx[pr][c] + sxx[pr][c + 1]) + sxx[pr][c + 2];
                const double ry = (sxy[pr][c] + sxy[pr][c + 1]) + sxy[pr][c + 2];
                const double rz = (syy[pr][c] + syy[pr][c + 1]) + syy[pr][c + 2];
                if (rr >= 2) {
                    const int r = pr - 2, y = y0 - 1 + r;
                    float e = -1.f;
                    if ((unsigned)x < (unsigned)G.w && (unsigned)y < (unsigned)G.h) {
                        const double bxx = (w0x + w1x) + rx, bxy = (w0y + w1y) + ry, byy = (w0z + w1z) + rz;
                        float a = __fmul_rn((float)bxx, 0.5f), b = (float)bxy, cc = __fmul_rn((float)byy, 0.5f);
                        float d = __fsub_rn(a, cc);
                        e = __fsub_rn(__fadd_rn(a, cc), __fsqrt_rn(__fadd_rn(__fmul_rn(d, d), __fmul_rn(b, b))));
                        if (e > 0.f) lmax = max(lmax, __float_as_uint(e));
                    }
                    se[r][c] = e;
                }
                w0x = w1x; w0y = w1y; w0z = w1z;
                w1x = rx; w1y = ry; w1z = rz;
            }
        }
    }
    lmax = __reduce_max_sync(0xffffffffu, lmax);
    if (lane == 0 && lmax) atomicMax(&smax, lmax);
    __syncthreads();
    if (tid == 0 && smax) atomicMax(D.eig_max, smax);
    // 4. 3x3 non-max suppression on the tile, warp-ballot compaction, one global atomic per CTA.  A thread owns four
    //    vertically adjacent positions (column c, rows 4q..4q+3): six rows of three eigenvalues give the row-wise maxima
    //    once, "e >= all eight neighbours" becomes e >= max(row above, row below, left, right) - no branches.
    bool is[4];
    float ev[4];
    int npos = 0;
    const int nc = tid & (EIG_TW - 1), nq = tid / EIG_TW;
    {
        float hm[6], side[6], ctr[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const float l = se[4 * nq + j][nc], m = se[4 * nq + j][nc + 1], rr = se[4 * nq + j][nc + 2];
            side[j] = fmaxf(l, rr);
            hm[j] = fmaxf(side[j], m);
            ctr[j] = m;
        }
        const int x = x0 + nc;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int y = y0 + 4 * nq + k;
            const float e = ctr[k + 1];
            ev[k] = e;
            is[k] = x >= 1 && x < G.w - 1 && y >= 1 && y < G.h - 1 && e > 0.f && e >= fmaxf(fmaxf(hm[k], hm[k + 2]), side[k + 1]);
            npos += is[k] ? 1 : 0;
        }
    }
    // exclusive prefix of npos inside the warp, then across warps
    int incl = npos;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31 && incl) s_warp_off[warp] = atomicAdd(&s_count, incl);
    __syncthreads();
    if (tid == 0 && s_count) s_base = atomicAdd(D.cand_count, s_count);
    __syncthreads();
    if (npos) {
        int pos = s_base + s_warp_off[warp] + incl - npos;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (is[k])
                D.cand[pos++] = ((unsigned long long)__float_as_uint(ev[k]) << 32) | (unsigned)((y0 + 4 * nq + k) * G.w + x0 + nc);
        }
    }
}

// ------------------------------------------------------------------------------------- k_select
#define SEL_BINS 4096
#define SEL_GRID_CELLS 2560          // min-distance grids up to this many cells live in shared memory

struct SelSmem {
    unsigned int hist[256];
    unsigned long long prefix, mask, T;
    int k, count, accepted, done;
    int warp_sum[32];
    int lo_bin, chunk_count, n_valid;
};

// K-th largest key strictly below U (keys are unique): MSD radix select, 8 bits per pass.
// Only used when one histogram bin alone overflows the sort buffer (massive ties).
static __device__ unsigned long long radix_select(const unsigned long long* __restrict__ keys, int n,
                                                  unsigned long long U, int K, SelSmem& S) {
    if (threadIdx.x == 0) { S.prefix = 0ull; S.mask = 0ull; S.k = K; }
    for (int shift = 56; shift >= 0; shift -= 8) {
        if (threadIdx.x < 256) S.hist[threadIdx.x] = 0u;
        __syncthreads();
        const unsigned long long prefix = S.prefix, mask = S.mask;
        for (int i = threadIdx.x; i < n; i += SEL_THREADS) {
            unsigned long long k = keys[i];
            if (k < U && (k & mask) == prefix) atomicAdd(&S.hist[(unsigned)(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int k = S.k, c = 0, d = 255;
            for (; d > 0; --d) {
                int hcount = (int)S.hist[d];
                if (c + hcount >= k) break;
                c += hcount;
            }
            S.k = k - c;
            S.prefix = prefix | ((unsigned long long)d << shift);
            S.mask = mask | (0xffull << shift);
        }
        __syncthreads();
    }
    return S.prefix;
}

// Descending bitonic sort of n keys (n a power of two >= 32) in shared memory.  Compare-exchange distances below 32 stay inside a
// warp: those stages run on registers with shuffles (no block barrier); only the distances >= 32 go through shared memory.
static __device__ __forceinline__ unsigned long long bitonic_warp_stages(unsigned long long v, int i, int k, int j) {
    for (; j > 0; j >>= 1) {
        const unsigned long long p = __shfl_xor_sync(0xffffffffu, v, j);
        const bool keep_max = (((i & j) == 0) == ((i & k) == 0));      // descending block: the lower index keeps the larger key
        v = keep_max ? (v > p ? v : p) : (v < p ? v : p);
    }
    return v;
}
static __device__ void bitonic_sort_desc(unsigned long long* a, int n) {
    // k = 2 .. 32: every stage is warp-local
    for (int i = threadIdx.x; i < n; i += SEL_THREADS) {
        unsigned long long v = a[i];
        for (int k = 2; k <= 32; k <<= 1) v = bitonic_warp_stages(v, i, k, k >> 1);
        a[i] = v;
    }
    __syncthreads();
    for (int k = 64; k <= n; k <<= 1) {
        for (int j = k >> 1; j >= 32; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += SEL_THREADS) {
                int ixj = i ^ j;
                if (ixj > i) {
                    unsigned long long x = a[i], y = a[ixj];
                    bool desc = (i & k) == 0;
                    if (desc ? (x < y) : (x > y)) { a[i] = y; a[ixj] = x; }
                }
            }
            __syncthreads();
        }
        for (int i = threadIdx.x; i < n; i += SEL_THREADS) a[i] = bitonic_warp_stages(a[i], i, k, 16);
        __syncthreads();
    }
}

// The order-dependent min-distance pass of cv::goodFeaturesToTrack over m sorted candidates, run by
// one warp 32 candidates at a time (a single warp is latency-bound, so the pass is written to be
// branch-free and shallow):
//   1. every lane tests its candidate against the corners accepted so far: 9 grid cells x 4 slots,
//      slots pre-filled with a far-away sentinel so no counts are read and nothing branches;
//   2. conflicts INSIDE the batch: each lane builds, once, the bit mask of earlier lanes closer than
//      minDistance; the rank-ordered greedy outcome is then the fixed point of a few ballot rounds
//      (a lane is accepted when no earlier undecided lane conflicts with it, rejected when an accepted
//      one does) - exactly the sequential result, without walking the lanes one by one;
//   3. accepted lanes append themselves in rank order (ballot prefix) and enter the grid.
// `sxy` / `scell` hold, per sorted candidate, (x | y << 16) and the grid cell index, computed in
// parallel by the whole CTA beforehand (all integer divisions live there).
#define GRID_EMPTY 0x40004000u      // sentinel slot: x = y = 16384, never within minDistance of a real pixel (no int overflow)
template <bool SMEM_GRID>
static __device__ bool greedy_pass(float2* kp_out, const unsigned* sxy, const int* scell, int m, int gw, int gh,
                                   bool use_grid, int imd2, int cap, unsigned int* gcount, unsigned int* gslot,
                                   int& accepted_io) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    const unsigned lt = (1u << lane) - 1u;
    int accepted = accepted_io;
    bool stop = false;
    for (int base = 0; base < m && !stop; base += 32) {
        bool ok = base + lane < m;
        const unsigned xy = ok ? sxy[base + lane] : GRID_EMPTY;
        const int ci = ok ? scell[base + lane] : 0;
        const int x = (int)(xy & 0xffffu), y = (int)(xy >> 16);
        unsigned cmask = 0u;
        if (use_grid) {
            const int cy = ci / gw, cx = ci - cy * gw;
            bool hit = false;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const int yy = min(max(cy + dy, 0), gh - 1);
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    const int xx = min(max(cx + dx, 0), gw - 1);
                    const int c2 = yy * gw + xx;
                    uint4 q;
                    if (SMEM_GRID) q = *reinterpret_cast<const uint4*>(gslot + c2 * VS_GRID_SLOTS);
                    else q = __ldcg(reinterpret_cast<const uint4*>(gslot + (size_t)c2 * VS_GRID_SLOTS));
                    const unsigned qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const int ddx = x - (int)(qq[s] & 0xffffu), ddy = y - (int)(qq[s] >> 16);
                        hit |= (ddx * ddx + ddy * ddy) < imd2;
                    }
                }
            }
            // earlier lanes of this batch closer than minDistance: all 32 candidates of the batch are read back from shared
            // memory (broadcast 128-bit loads, nothing depends on anything) and tested; only lanes that survived the grid test
            // can ever be accepted, so only their bits count
            ok = ok && !hit;
            const unsigned walk = __ballot_sync(FULL, ok);
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
                const uint4 q4 = *reinterpret_cast<const uint4*>(sxy + base + 4 * k4);
                const unsigned qv[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4) {
                    const int ddx = x - (int)(qv[s4] & 0xffffu), ddy = y - (int)(qv[s4] >> 16);
                    cmask |= ((ddx * ddx + ddy * ddy) < imd2) ? (1u << (4 * k4 + s4)) : 0u;
                }
            }
            cmask &= walk;
            cmask &= lt;
        }
        unsigned undecided = __ballot_sync(FULL, ok);
        unsigned acc = 0u;
        while (undecided) {
            const unsigned now = __ballot_sync(FULL, ok && (cmask & undecided) == 0u);
            acc |= now;
            if (ok && (((now >> lane) & 1u) || (cmask & acc))) ok = false;
            undecided = __ballot_sync(FULL, ok);
        }
        // cap: keep only the first (cap - accepted) of this batch in rank order
        int room = cap - accepted;
        int rank = __popc(acc & lt);
        const bool mine = ((acc >> lane) & 1u) && rank < room;
        if (mine) {
            kp_out[accepted + rank] = make_float2((float)x, (float)y);
            if (use_grid) {
                unsigned slot = atomicAdd(&gcount[ci], 1u);
                if (slot < VS_GRID_SLOTS) {
                    if (SMEM_GRID) gslot[ci * VS_GRID_SLOTS + slot] = xy;
                    else __stcg(&gslot[(size_t)ci * VS_GRID_SLOTS + slot], xy);
                }
            }
        }
        accepted += min(__popc(acc), room);
        if (accepted >= cap) stop = true;
        if (!SMEM_GRID) __threadfence_block();
        __syncwarp();
    }
    accepted_io = accepted;
    return stop;
}

// One CTA per lane.  Candidates arrive unordered (tens of thousands on textured frames) but the greedy
// pass usually stops after a few hundred, so the strongest ones are peeled off in value order chunk by
// chunk: a 4096-bin histogram of the float bits (one pass), a block-wide suffix scan to find the bin
// boundary that holds the next >= `target` candidates, a gather + shared-memory bitonic sort of just
// those, then the greedy pass.  Bins are value-ordered, so chunk order == global order.
__global__ void __launch_bounds__(SEL_THREADS) k_select(const LaneDev* __restrict__ lanes, int slot, int max_corners,
                                                         double quality, double min_dist, int record_frame_no, int gen, int kp_slot) {
    extern __shared__ unsigned long long sel_dyn[];
    unsigned long long* skeys = sel_dyn;                                        // SEL_CHUNK_MAX keys
    unsigned int* hist = reinterpret_cast<unsigned int*>(sel_dyn + SEL_CHUNK_MAX);   // SEL_BINS
    unsigned int* sgrid = hist + SEL_BINS;                                      // SEL_GRID_CELLS * (slots + 1)
    unsigned int* sxy = sgrid + SEL_GRID_CELLS * (1 + VS_GRID_SLOTS);           // SEL_CHUNK_MAX packed (x | y << 16)
    int* scell = reinterpret_cast<int*>(sxy + SEL_CHUNK_MAX);                   // SEL_CHUNK_MAX grid cell indices
    __shared__ SelSmem S;
    const LaneDev& L = lanes[blockIdx.z];
    const DetView D = det_view(L, gen, kp_slot);     // scratch generation read, key-point slot written (engine.cu)
    const GrayLevel G = gftt_src(L, slot);
    const int w = G.w, h = G.h;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = min(*D.cand_count, w * h);
    const int cap = (max_corners > 0) ? min(max_corners, L.kp_capacity) : L.kp_capacity;
    const bool use_grid = min_dist >= 1.0;
    const int cell = use_grid ? (int)rint(min_dist) : 1;
    const int gw = (w + cell - 1) / cell, gh = (h + cell - 1) / cell;
    const double md2 = min_dist * min_dist;
    const bool smem_grid = gw * gh <= SEL_GRID_CELLS;
    // grid layout: [slots: cells x 4 (16-byte aligned)] [counts: cells]
    unsigned int* gslot = smem_grid ? sgrid : D.grid;
    unsigned int* gcount = gslot + (size_t)(smem_grid ? SEL_GRID_CELLS : gw * gh) * VS_GRID_SLOTS;
    const int imd2 = (int)ceil(md2);             // integer d2 < md2  <=>  d2 < ceil(md2)
    if (use_grid)
        for (int i = tid; i < gw * gh; i += SEL_THREADS) {
            gcount[i] = 0u;
            if (smem_grid) *reinterpret_cast<uint4*>(gslot + i * VS_GRID_SLOTS) = make_uint4(GRID_EMPTY, GRID_EMPTY, GRID_EMPTY, GRID_EMPTY);
            else __stcg(reinterpret_cast<uint4*>(gslot + (size_t)i * VS_GRID_SLOTS), make_uint4(GRID_EMPTY, GRID_EMPTY, GRID_EMPTY, GRID_EMPTY));
        }
    for (int i = tid; i < SEL_BINS; i += SEL_THREADS) hist[i] = 0u;
    if (tid == 0) { S.accepted = 0; S.done = 0; }

    // bin = (float bits - bits(threshold)) >> shift, top bin holds the maximum
    const float mxv = __uint_as_float(*D.eig_max);
    const unsigned lo = __float_as_uint((float)((double)mxv * quality));
    const unsigned range = __float_as_uint(mxv) - lo;
    int shift = 0;
    while ((range >> shift) >= (unsigned)SEL_BINS) ++shift;
    __syncthreads();
    // every thread holds N and the maximum in registers now: leave this generation's counters zeroed for its next
    // detection (same stream, so ordered) — the engine needs no memset node per detection
    if (tid == 0) { *D.eig_max = 0u; *D.cand_count = 0; }
    for (int i0 = tid; i0 < N; i0 += 4 * SEL_THREADS) {          // four independent loads in flight per thread
        unsigned long long kk[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) kk[u] = (i0 + u * SEL_THREADS < N) ? __ldcg(D.cand + i0 + u * SEL_THREADS) : 0ull;
#pragma unroll
        for (int u = 0; u < 4; ++u)                               // keys <= threshold are not candidates (k_eig_nms emits all maxima)
            if (i0 + u * SEL_THREADS < N && (unsigned)(kk[u] >> 32) > lo) atomicAdd(&hist[((unsigned)(kk[u] >> 32) - lo) >> shift], 1u);
    }
    __syncthreads();
    // suffix scan (from the top bin down): thread r owns bins 4095-4r .. 4092-4r
    int own[4];
    int local = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { own[k] = (int)hist[SEL_BINS - 1 - 4 * tid - k]; local += own[k]; }
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) S.warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int v = S.warp_sum[lane], sc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, sc, o);
            if (lane >= o) sc += u;
        }
        S.warp_sum[lane] = sc - v;             // exclusive
    }
    __syncthreads();
    const int before = S.warp_sum[warp] + incl - local;    // candidates in bins above this thread's four
    if (tid == SEL_THREADS - 1) S.n_valid = before + local;  // all keys above the quality threshold
    __syncthreads();
    const int Nv = S.n_valid;

    int taken = 0;               // candidates already handed to the greedy pass (exactly those with key >= U)
    unsigned long long U = ~0ull;
    bool first = true;
    while (taken < Nv) {
        const int target = min(first ? SEL_CHUNK_FIRST : SEL_CHUNK_MAX / 2, Nv - taken);
        first = false;
        // lowest bin such that the not-yet-taken candidates in bins >= lo_bin number >= target
        {
            int cum = before;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int prev = cum;
                cum += own[k];
                if (prev - taken < target && cum - taken >= target) {
                    S.lo_bin = SEL_BINS - 1 - 4 * tid - k;
                    S.chunk_count = cum - taken;
                }
            }
        }
        __syncthreads();
        const int lo_bin = S.lo_bin;
        int m = S.chunk_count;
        // key lower bound of the chunk: bin >= lo_bin  <=>  float bits >= lo + (lo_bin << shift)
        unsigned long long T = (unsigned long long)((unsigned long long)lo + ((unsigned long long)lo_bin << shift)) << 32;
        if (m > SEL_CHUNK_MAX) {
            // one bin overflows the sort buffer (massive ties): split it exactly by key
            m = SEL_CHUNK_MAX;
            T = radix_select(D.cand, N, U, m, S);
        }
        int npad = 32;
        while (npad < m) npad <<= 1;
        if (tid == 0) S.count = 0;
        for (int i = tid; i < npad; i += SEL_THREADS) skeys[i] = 0ull;
        __syncthreads();
        for (int i0 = tid; i0 < N; i0 += 4 * SEL_THREADS) {
            unsigned long long kk[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) kk[u] = (i0 + u * SEL_THREADS < N) ? __ldcg(D.cand + i0 + u * SEL_THREADS) : 0ull;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i0 + u * SEL_THREADS < N && kk[u] < U && kk[u] >= T && (unsigned)(kk[u] >> 32) > lo)
                    skeys[atomicAdd(&S.count, 1)] = kk[u];
        }
        __syncthreads();
        bitonic_sort_desc(skeys, npad);
        for (int i = tid; i < m; i += SEL_THREADS) {       // unpack coordinates / cells in parallel
            const unsigned addr = (unsigned)(skeys[i] & 0xffffffffull);
            const int y = addr / w, x = addr - y * w;
            sxy[i] = (unsigned)x | ((unsigned)y << 16);
            scell[i] = use_grid ? (y / cell) * gw + (x / cell) : 0;
        }
        __syncthreads();
        if (tid < 32) {
            int accepted = S.accepted;
            bool stop = smem_grid ? greedy_pass<true>(D.kp, sxy, scell, m, gw, gh, use_grid, imd2, cap, gcount, gslot, accepted)
                                  : greedy_pass<false>(D.kp, sxy, scell, m, gw, gh, use_grid, imd2, cap, gcount, gslot, accepted);
            if (lane == 0) { S.accepted = accepted; S.done = stop ? 1 : 0; }
        }
        __syncthreads();
        if (S.done) break;
        taken += m;
        U = T;
        __syncthreads();
    }
    __syncthreads();
    const int n = S.accepted;
    if (tid == 0) *D.kp_count = n;
    if (slot < 0) {
        for (int i = tid; i < n; i += SEL_THREADS) L.first_corners[i] = D.kp[i];
        if (tid == 0) *L.first_count = n;
    }
    if (record_frame_no > 0) {
        if (tid == 0 && record_frame_no <= L.record_capacity) L.frec[record_frame_no - 1].n_detected = n;
        if (L.log_depth > 0) {
            float2* dst = L.log_detected + (size_t)((record_frame_no - 1) % L.log_depth) * L.kp_capacity;
            for (int i = tid; i < n; i += SEL_THREADS) dst[i] = D.kp[i];
        }
    }
}

#define SEL_DYN_BYTES (SEL_CHUNK_MAX * sizeof(unsigned long long) + SEL_BINS * sizeof(unsigned int) + \
                       SEL_GRID_CELLS * (1 + VS_GRID_SLOTS) * sizeof(unsigned int) + SEL_CHUNK_MAX * 2 * sizeof(unsigned int))

// ------------------------------------------------------------------------------------ any block size (first frame only)
// cv::goodFeaturesToTrack is called with params_.blockSize only on the very first frame (Stabilizer.cpp:355-357); every re-detection
// hard-codes 3 (:744).  For blockSize != 3 the eigenvalue map is computed per pixel, once per stream, by the two plain kernels below
// (the tiled k_eig_nms is specialised for 3x3): Sobel 3x3 scaled by 1 / (4 * blockSize * 255), products in float, the
// blockSize x blockSize box sum in double over BORDER_REFLECT_101 of the PRODUCT maps (window anchored at blockSize / 2), then
// (a + c) - sqrt((a - c)^2 + b^2) exactly as k_eig_nms, 3x3 non-max suppression and the same sort keys for k_select.
__global__ void __launch_bounds__(128) k_eig_generic(const LaneDev* __restrict__ lanes, int slot, int block, float* __restrict__ eig_all) {
    const LaneDev& L = lanes[blockIdx.z];
    const GrayLevel G = gftt_src(L, slot);
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= G.w) return;
    const float f1 = (float)(1.0 / (4.0 * (double)block * 255.0));
    const float f0 = 2.f * f1;
    const int a0 = block / 2;
    double sxx = 0, sxy = 0, syy = 0;
    for (int j = 0; j < block; ++j) {
        const int py = reflect101(y - a0 + j, G.h);
        double rx = 0, ry = 0, rz = 0;
        for (int i = 0; i < block; ++i) {
            const int px = reflect101(x - a0 + i, G.w);
            const uint8_t* r0 = G.base + (ptrdiff_t)(py - 1) * G.pitch + px;      // the level's materialised reflect frame supplies Sobel's border
            const uint8_t* r1 = r0 + G.pitch;
            const uint8_t* r2 = r1 + G.pitch;
            // row passes then column passes, in k_eig_nms's (= OpenCV's) operation order
            const float t0 = __fadd_rn(__fadd_rn(__fmul_rn((float)r0[-1], f1), __fmul_rn((float)r0[0], f0)), __fmul_rn((float)r0[1], f1));
            const float t2 = __fadd_rn(__fadd_rn(__fmul_rn((float)r2[-1], f1), __fmul_rn((float)r2[0], f0)), __fmul_rn((float)r2[1], f1));
            const int d0 = (int)r0[1] - (int)r0[-1], d1 = (int)r1[1] - (int)r1[-1], d2 = (int)r2[1] - (int)r2[-1];
            const float dx = __fadd_rn(__fmul_rn((float)(d0 + d2), f1), __fmul_rn((float)d1, f0));
            const float dy = __fsub_rn(t2, t0);
            rx += (double)__fmul_rn(dx, dx);
            ry += (double)__fmul_rn(dx, dy);
            rz += (double)__fmul_rn(dy, dy);
        }
        sxx += rx; sxy += ry; syy += rz;
    }
    const float a = __fmul_rn((float)sxx, 0.5f), b = (float)sxy, c = __fmul_rn((float)syy, 0.5f);
    const float d = __fsub_rn(a, c);
    const float e = __fsub_rn(__fadd_rn(a, c), __fsqrt_rn(__fadd_rn(__fmul_rn(d, d), __fmul_rn(b, b))));
    eig_all[(size_t)blockIdx.z * G.w * G.h + (size_t)y * G.w + x] = e;
}
__global__ void __launch_bounds__(128) k_nms_generic(const LaneDev* __restrict__ lanes, int slot, int gen, const float* __restrict__ eig_all) {
    const LaneDev& L = lanes[blockIdx.z];
    const DetView D = det_view(L, gen, 0);
    const GrayLevel G = gftt_src(L, slot);
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= G.w) return;
    const float* eig = eig_all + (size_t)blockIdx.z * G.w * G.h;
    const float e = eig[(size_t)y * G.w + x];
    if (e > 0.f) atomicMax(D.eig_max, __float_as_uint(e));
    if (x < 1 || x >= G.w - 1 || y < 1 || y >= G.h - 1 || !(e > 0.f)) return;
    bool is = true;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) is = is && e >= eig[(size_t)(y + dy) * G.w + x + dx];
    if (is) D.cand[atomicAdd(D.cand_count, 1)] = ((unsigned long long)__float_as_uint(e) << 32) | (unsigned)(y * G.w + x);
}

size_t gftt_grid_words(int w, int h, double min_dist) {
    if (min_dist < 1.0) return 8;
    int cell = (int)rint(min_dist);
    size_t cells = (size_t)((w + cell - 1) / cell) * ((h + cell - 1) / cell);
    return cells * (1 + VS_GRID_SLOTS);
}

void launch_good_features(const LaneDev* lanes, int n_lanes, int slot, int max_corners, double quality,
                          double min_dist, int record_frame_no, int gen, int kp_slot, cudaStream_t st, int block_size, float* eig_scratch, int aw, int ah) {
    const int w = slot < 0 ? VS_FW : aw, h = slot < 0 ? VS_FH : ah;
    // the attribute is per DEVICE (a process may hold handles on several GPUs): once per device, not once per process
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaFuncSetAttribute(k_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEL_DYN_BYTES);
        attr_set[dev] = true;
    }
    if (block_size != 3 && eig_scratch) {
        const dim3 gg((w + 127) / 128, h, n_lanes);
        k_eig_generic<<<gg, 128, 0, st>>>(lanes, slot, block_size, eig_scratch);
        k_nms_generic<<<gg, 128, 0, st>>>(lanes, slot, gen, eig_scratch);
    } else {
        dim3 g1((w + EIG_TW - 1) / EIG_TW, (h + EIG_TH - 1) / EIG_TH, n_lanes);
        k_eig_nms<<<g1, 256, 0, st>>>(lanes, slot, gen);
    }
    k_select<<<dim3(1, 1, n_lanes), SEL_THREADS, SEL_DYN_BYTES, st>>>(
        lanes, slot, max_corners, quality, min_dist, record_frame_no, gen, kp_slot);
}
