// k_motion.cu — everything between the tracker and the warp, fused into one launch per step:
//   status filter (Stabilizer.cpp:629-641)  ->  cv::estimateAffinePartial2D RANSAC 5.0/500/0.99 + refine
//   (:645-659)  ->  dx,dy,da decomposition (:660-662)  ->  trajectory accumulation (:681-688)  ->
//   path smoothing box|gaussian|kalman (:797-823, :1139-1172, :1364-1458, :1637-1673)  ->  motion intent
//   (:854-888, :1676-1780)  ->  2x3 matrix (:890-908) and cv::warpAffine's matrix inversion.
// One CTA of 32 warps per lane.  RANSAC is batched: 32 hypotheses are scored in parallel, one warp per
// hypothesis with ballot/popc inlier counting, then the reference's *sequential* adaptive-termination
// rule is replayed over the scores so the same hypothesis wins as on the CPU.
// Specification: oracle/cv_models.py estimate_affine_partial_2d + oracle/stabilizer_ref.py.
#include "kernels.h"

#define MO_THREADS 256                 // 8 warps: a 1024-thread CTA claims a whole SM (register file), which delays its
                                       // start while other streams' kernels occupy the SMs; the first RANSAC round has
                                       // 8 hypotheses (one per warp) and usually is the only one
#define MO_MAXP 2048                 // key-point capacity handled in shared memory
#define MO_BATCH 32                  // hypotheses per round
#define MO_TAIL 128                  // trajectory entries staged in shared memory for the sequential part

struct MoSmem {
    float2 from[MO_MAXP];
    float2 to[MO_MAXP];
    int warp_tot[MO_THREADS / 32];
    int idx[MO_BATCH][2];
    int good[MO_BATCH];
    double model[6];
    float bestF[6];
    double red[7][32];
    unsigned long long rng;
    int n, niters, iter, max_good, best_found, cont;
    float gk[512];                   // gaussian kernel taps
    float tail_path[3 * MO_TAIL];
    float tail_trf[3 * MO_TAIL];
    float tail_aux[2 * MO_TAIL];
    unsigned long long fastmod_M;
};

static __device__ __forceinline__ unsigned rng_next(unsigned long long& s) {
    s = (unsigned long long)(unsigned)s * 4164903690ull + (unsigned)(s >> 32);
    return (unsigned)s;
}

static __device__ void model_from_pair(float2 a1, float2 a2, float2 b1, float2 b2, double* M) {
    // AffinePartial2DEstimatorCallback::runKernel — exact 2-point similarity in double
    double x1 = a1.x, y1 = a1.y, x2 = a2.x, y2 = a2.y;
    double X1 = b1.x, Y1 = b1.y, X2 = b2.x, Y2 = b2.y;
    double d = 1. / ((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2));
    double S0 = d * ((X1 - X2) * (x1 - x2) + (Y1 - Y2) * (y1 - y2));
    double S1 = d * ((Y1 - Y2) * (x1 - x2) - (X1 - X2) * (y1 - y2));
    double S2 = d * ((Y1 - Y2) * (x1 * y2 - x2 * y1) - (X1 * y2 - X2 * y1) * (y1 - y2) - (X1 * x2 - X2 * x1) * (x1 - x2));
    double S3 = d * (-(X1 - X2) * (x1 * y2 - x2 * y1) - (Y1 * x2 - Y2 * x1) * (x1 - x2) - (Y1 * y2 - Y2 * y1) * (y1 - y2));
    M[0] = S0; M[1] = -S1; M[2] = S2; M[3] = S1; M[4] = S0; M[5] = S3;
}

static __device__ __forceinline__ bool is_inlier(const float* F, float2 f, float2 t) {
    // Affine2DEstimatorCallback::computeError in float32 (no contraction), threshold 5^2
    float a = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(F[0], f.x), __fmul_rn(F[1], f.y)), F[2]), t.x);
    float b = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(F[3], f.x), __fmul_rn(F[4], f.y)), F[5]), t.y);
    float e = __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
    return e <= 25.0f;
}

static __device__ int ransac_update_iters(double p, double ep, int max_iters) {
    // cv::RANSACUpdateNumIters(confidence, outlier ratio, modelPoints = 2, maxIters)
    p = fmax(p, 0.); p = fmin(p, 1.);
    ep = fmax(ep, 0.); ep = fmin(ep, 1.);
    double num = fmax(1. - p, 2.2250738585072014e-308);
    double denom = 1. - (1. - ep) * (1. - ep);
    if (denom < 2.2250738585072014e-308) return 0;
    num = log(num);
    denom = log(denom);
    return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : (int)rint(num / denom);
}

// Trajectory view used by the sequential smoothing code: element i of transforms_/path_ lives at
// p[3*(i-base)+c].  In the motion kernel the last MO_TAIL entries are staged in shared memory in
// parallel first, so the one-thread float32 replay never waits on a global-memory round trip.
struct Traj {
    const float* p;
    int base;
    __device__ __forceinline__ float at(int i, int c) const { return p[3 * (i - base) + c]; }
    __device__ __forceinline__ float at2(int i, int c) const { return p[2 * (i - base) + c]; }   // aux rows
};

// ---- scalar float32 helpers restating the reference's host arithmetic (one thread) -----------------
static __device__ float f_sqrt(float x) { return __fsqrt_rn(x); }

// std::cos / std::sin on float (Stabilizer.cpp:902-906) are glibc's cosf / sinf: NOT correctly rounded (sinf differs
// from (float)sin(double) for ~1 argument in 8000), and a last-bit difference in the matrix moves the fixed-point
// source coordinate of isolated pixels by 1/32 px.  These follow glibc's algorithm for |x| < pi/4 (the ARM optimized
// routines polynomial, evaluated in double without contraction; verified bit-identical against libm.so.6 on 600k
// arguments, oracle/ + tests/test_oracle_models.py); larger angles fall back to the correctly rounded value.
static __device__ float f_cos(float xf) {
    const float ax = fabsf(xf);
    if (ax < 2.44140625e-4f) return 1.0f;                                   // |x| < 2^-12
    if (!(ax < 0.78539816339744830962f)) return (float)cos((double)xf);
    const double x = (double)xf, x2 = __dmul_rn(x, x);
    const double c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10, c4 = 0x1.99343027bf8c3p-16;
    const double x4 = __dmul_rn(x2, x2);
    const double C2 = __dadd_rn(c3, __dmul_rn(x2, c4));
    const double C1 = __dadd_rn(c1, __dmul_rn(x2, c2));
    const double x6 = __dmul_rn(x4, x2);
    const double c = __dadd_rn(1.0, __dmul_rn(x2, C1));
    return (float)__dadd_rn(c, __dmul_rn(x6, C2));
}
static __device__ float f_sin(float xf) {
    const float ax = fabsf(xf);
    if (ax < 2.44140625e-4f) return xf;                                     // |x| < 2^-12
    if (!(ax < 0.78539816339744830962f)) return (float)sin((double)xf);
    const double x = (double)xf, x2 = __dmul_rn(x, x);
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    const double x3 = __dmul_rn(x, x2);
    const double S1 = __dadd_rn(s2, __dmul_rn(x2, s3));
    const double x7 = __dmul_rn(x3, x2);
    const double sv = __dadd_rn(x, __dmul_rn(x3, s1));
    return (float)__dadd_rn(sv, __dmul_rn(x7, S1));
}
// std::atan2 on float is glibc's atan2f (e_atan2f.c): atanf(|y/x|) with the quotient ROUNDED TO FLOAT first, which
// makes it differ from the correctly rounded atan2 in ~20 % of cases.  Same structure here; atanf follows glibc 2.39's
// s_atanf.c (fdlibm, float, no contraction) above 7/16 and the correctly rounded value below it (they agree there in
// all but ~1 case per 10^4 for the small angles this path produces).
static __device__ float f_atanf_pos(float x) {          // x >= 0
    const int ix = __float_as_int(x);
    if (ix < 0x3ee00000) return (float)atan((double)x);                     // |x| < 0.4375
    int id;
    if (ix < 0x3f980000) {                                                  // |x| < 1.1875
        if (ix < 0x3f300000) { id = 0; x = __fdiv_rn(__fsub_rn(__fmul_rn(2.0f, x), 1.0f), __fadd_rn(2.0f, x)); }
        else { id = 1; x = __fdiv_rn(__fsub_rn(x, 1.0f), __fadd_rn(x, 1.0f)); }
    } else if (ix < 0x401c0000) { id = 2; x = __fdiv_rn(__fsub_rn(x, 1.5f), __fadd_rn(1.0f, __fmul_rn(1.5f, x))); }
    else if (ix < 0x4c000000) { id = 3; x = __fdiv_rn(-1.0f, x); }
    else return __fadd_rn(__int_as_float(0x3fc90fda), __int_as_float(0x33a22168));
    const float atanhi[4] = {__int_as_float(0x3eed6338), __int_as_float(0x3f490fda), __int_as_float(0x3f7b985e), __int_as_float(0x3fc90fda)};
    const float atanlo[4] = {__int_as_float(0x31ac3769), __int_as_float(0x33222168), __int_as_float(0x33140fb4), __int_as_float(0x33a22168)};
    const float a0 = __int_as_float(0x3eaaaaaa), a1 = __int_as_float(0xbe4ccccd), a2 = __int_as_float(0x3e124925),
                a3 = __int_as_float(0xbde38e38), a4 = __int_as_float(0x3dba2e6e), a5 = __int_as_float(0xbd9d8795),
                a6 = __int_as_float(0x3d886b35), a7 = __int_as_float(0xbd6ef16b), a8 = __int_as_float(0x3d4bda59),
                a9 = __int_as_float(0xbd15a221), a10 = __int_as_float(0x3c8569d7);
    const float z = __fmul_rn(x, x), w = __fmul_rn(z, z);
#define VS_MA(a, b, c) __fadd_rn(a, __fmul_rn(b, c))
    const float s1 = __fmul_rn(z, VS_MA(a0, w, VS_MA(a2, w, VS_MA(a4, w, VS_MA(a6, w, VS_MA(a8, w, a10))))));
    const float s2 = __fmul_rn(w, VS_MA(a1, w, VS_MA(a3, w, VS_MA(a5, w, VS_MA(a7, w, a9)))));
#undef VS_MA
    return __fsub_rn(atanhi[id], __fsub_rn(__fsub_rn(__fmul_rn(x, __fadd_rn(s1, s2)), atanlo[id]), x));
}
static __device__ float f_atan2(float y, float x) {
    const int hx = __float_as_int(x), hy = __float_as_int(y);
    const int ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix >= 0x7f800000 || iy >= 0x7f800000) return (float)atan2((double)y, (double)x);     // inf / NaN
    const float pi = __int_as_float(0x40490fdb), pi_lo = __int_as_float(0xb3bbbd2e), pi_o_2 = __int_as_float(0x3fc90fdb);
    if (hx == 0x3f800000) { const float a = f_atanf_pos(fabsf(y)); return hy < 0 ? -a : a; }      // x == 1.0
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) return m < 2 ? y : (m == 2 ? pi : -pi);
    if (ix == 0) return hy < 0 ? -pi_o_2 : pi_o_2;
    const int k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = __fadd_rn(pi_o_2, __fmul_rn(0.5f, pi_lo));
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = f_atanf_pos(fabsf(__fdiv_rn(y, x)));
    switch (m) {
        case 0: return z;
        case 1: return -z;
        case 2: return __fsub_rn(pi, __fsub_rn(z, pi_lo));
        default: return __fsub_rn(__fsub_rn(z, pi_lo), pi);
    }
}

static __device__ int adaptive_radius(const Traj path, int n, int smoothing_radius) {
    // calculateAdaptiveRadius, Stabilizer.cpp:1637-1673
    if (n < 10) return smoothing_radius;
    int start = max(0, n - 20);
    float cnt = (float)(n - start);
    float mx = 0.f, my = 0.f, ma = 0.f;
    for (int i = start; i < n; ++i) {
        mx = __fadd_rn(mx, path.at(i, 0)); my = __fadd_rn(my, path.at(i, 1)); ma = __fadd_rn(ma, path.at(i, 2));
    }
    mx = __fdiv_rn(mx, cnt); my = __fdiv_rn(my, cnt); ma = __fdiv_rn(ma, cnt);
    float vx = 0.f, vy = 0.f, va = 0.f;
    for (int i = start; i < n; ++i) {
        float dx = __fsub_rn(path.at(i, 0), mx), dy = __fsub_rn(path.at(i, 1), my), da = __fsub_rn(path.at(i, 2), ma);
        vx = __fadd_rn(vx, __fmul_rn(dx, dx)); vy = __fadd_rn(vy, __fmul_rn(dy, dy)); va = __fadd_rn(va, __fmul_rn(da, da));
    }
    vx = __fdiv_rn(vx, cnt); vy = __fdiv_rn(vy, cnt); va = __fdiv_rn(va, cnt);
    float total = f_sqrt(__fadd_rn(__fadd_rn(vx, vy), __fmul_rn(va, 1000.f)));
    return (int)fmaxf(5.0f, fminf(25.0f, __fmul_rn(total, 2.0f)));
}

static __device__ float box_at(const Traj path, int comp, int n, int radius, int i, int drone) {
    // boxFilterConvolve, Stabilizer.cpp:1139-1172 — only element i is ever consumed
    int r = drone ? max(10, min(radius, 50)) : max(2, min(radius, 8));
    if (n <= r) return path.at(i, comp);
    int lo = max(0, i - r), hi = min(n - 1, i + r);
    float s = 0.f;
    for (int j = lo; j <= hi; ++j) s = __fadd_rn(s, path.at(j, comp));
    return __fdiv_rn(s, (float)(hi - lo + 1));
}

static __device__ float hf_median(const float* hist, int cnt, int comp) {
    // calculateMedianTranslation, Stabilizer.cpp:2531-2555 (cnt is 5..10)
    float v[10];
    for (int i = 0; i < 10; ++i) v[i] = i < cnt ? hist[2 * i + comp] : 0.f;
    for (int i = 1; i < 10; ++i) {
        if (i >= cnt) break;
        float x = v[i];
        int j = i - 1;
        while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; --j; }
        v[j + 1] = x;
    }
    const int mid = cnt / 2;
    float lo = 0.f, hi = 0.f;
    for (int i = 0; i < 10; ++i) { if (i == mid - 1) lo = v[i]; if (i == mid) hi = v[i]; }
    return (cnt & 1) ? hi : __fdiv_rn(__fadd_rn(lo, hi), 2.0f);
}

static __device__ void hf_filters(float* hf, const StepInfo& info, float* t) {
    // applyDeadZoneFreeze -> applyMicroShakeSuppression -> applyRotationLowPass -> updateTranslationHistory,
    // Stabilizer.cpp:666-671 with :2468-2529 and :2605-2681; float32 step for step
    const float dx = t[0], dy = t[1], da = t[2];
    const float thr = info.hf_dead_zone_threshold;
    const float mag = f_sqrt(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(__fmul_rn(da, da), 100.0f)));
    float acc = fmaxf(__fmul_rn(hf[VS_HF_ACCUMULATOR], info.hf_accumulator_decay), mag);
    acc = fminf(acc, __fmul_rn(thr, 5.0f));
    acc = fmaxf(0.0f, fminf(acc, 100.0f));
    int in_dz = __float_as_int(hf[VS_HF_IN_DEAD_ZONE]);
    int counter = __float_as_int(hf[VS_HF_FREEZE_COUNTER]);
    if (!in_dz && mag < thr) { in_dz = 1; counter = info.hf_freeze_duration; }
    if (in_dz) {
        --counter;
        if (counter <= 0 || mag > __fmul_rn(thr, 1.5f) || acc > __fmul_rn(thr, 1.2f)) { in_dz = 0; counter = 0; acc = 0.f; }
        else { t[0] = 0.f; t[1] = 0.f; t[2] = 0.f; }
    }
    hf[VS_HF_ACCUMULATOR] = acc;
    hf[VS_HF_IN_DEAD_ZONE] = __int_as_float(in_dz);
    hf[VS_HF_FREEZE_COUNTER] = __int_as_float(counter);

    int cnt = __float_as_int(hf[VS_HF_COUNT]);
    float* hist = hf + VS_HF_HIST;
    if (cnt >= 5) { hf[VS_HF_MEDIAN] = hf_median(hist, cnt, 0); hf[VS_HF_MEDIAN + 1] = hf_median(hist, cnt, 1); }
    const float mx = hf[VS_HF_MEDIAN], my = hf[VS_HF_MEDIAN + 1];
    const float ex = __fsub_rn(t[0], mx), ey = __fsub_rn(t[1], my);
    const float m2 = f_sqrt(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)));
    if (m2 < info.hf_shake_px) {
        t[0] = __fadd_rn(mx, __fmul_rn(ex, 0.01f)); t[1] = __fadd_rn(my, __fmul_rn(ey, 0.01f));
    } else if (m2 < __fmul_rn(info.hf_shake_px, 2.0f)) {
        t[0] = __fadd_rn(mx, __fmul_rn(ex, 0.05f)); t[1] = __fadd_rn(my, __fmul_rn(ey, 0.05f));
    }
    if (info.horizon_lock) {
        const float a = info.hf_rot_lp_alpha;
        const float lp = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, a), hf[VS_HF_ROT_LP]), __fmul_rn(a, t[2]));
        hf[VS_HF_ROT_LP] = lp;
        t[2] = lp;
    }
    if (cnt == 10) {
        for (int i = 0; i < 18; ++i) hist[i] = hist[i + 2];
        cnt = 9;
    }
    hist[2 * cnt] = t[0]; hist[2 * cnt + 1] = t[1];
    hf[VS_HF_COUNT] = __int_as_float(cnt + 1);
}

static __device__ float variance_f(const float* v, int n) {
    if (n == 0) return 0.f;
    float m = 0.f;
    for (int i = 0; i < n; ++i) m = __fadd_rn(m, v[i]);
    m = __fdiv_rn(m, (float)n);
    float var = 0.f;
    for (int i = 0; i < n; ++i) { float d = __fsub_rn(v[i], m); var = __fadd_rn(var, __fmul_rn(d, d)); }
    return __fdiv_rn(var, (float)n);
}
static __device__ float consistency_f(const float* v, int n) {
    if (n < 2) return 0.f;
    float var = variance_f(v, n);
    float m = 0.f;
    for (int i = 0; i < n; ++i) m = __fadd_rn(m, v[i]);
    m = __fdiv_rn(m, (float)n);
    if (m == 0.f) return 0.f;
    float c = __fdiv_rn(1.f, __fadd_rn(1.f, __fdiv_rn(var, __fmul_rn(m, m))));
    return fmaxf(0.f, fminf(1.f, c));
}

static __device__ int motion_intent(const Traj aux, int n_tr, const float* motion, int idx) {
    // analyzeMotionIntent, Stabilizer.cpp:1676-1719.  |t| and atan2(ty,tx) of every transform were
    // computed once when it was appended (pure functions of the transform), so this is loads + adds.
    float mag = aux.at2(idx, 0);
    float ang = (float)((double)__fmul_rn(fabsf(motion[2]), 180.0f) / 3.14159265358979323846 * (double)30.0f);
    if (n_tr >= 15) {
        float mags[15], dirs[15];
        int c = 0;
        for (int i = max(0, idx - 15); i < idx; ++i) {
            if (i < n_tr) {
                mags[c] = aux.at2(i, 0);
                dirs[c] = aux.at2(i, 1);
                ++c;
            }
        }
        if (c) {
            float dv = variance_f(dirs, c), mc = consistency_f(mags, c);
            if (dv < 0.5f && mc > 0.7f && mag > 5.0f) return 1;
            if (mag < 3.0f && mc < 0.3f && ang > 10.0f) return 2;
            if (mag > 3.0f && mag < 15.0f && dv > 0.5f) return 3;
        }
    }
    return 0;
}

// ---- warp-cooperative versions of the sequential float32 smoothing arithmetic ---------------------------------------
// The reference sums in a fixed order in float32, so the ORDER cannot change - but the loads can run in parallel and the
// ordered sum can be fed from registers: lane k holds sample k, every lane replays the same chain
// s = (...((0 + v0) + v1) ...) + v_{cnt-1} from broadcast shuffles.  The shuffles are independent of the adds, so they
// pipeline; one thread walking shared memory paid a load latency per term (6.6 us -> about 2 us for the whole section).
template <int MAXN>
static __device__ __forceinline__ float warp_seq_sum(float v, int cnt, float s = 0.f) {
#pragma unroll
    for (int k = 0; k < MAXN; ++k) {
        const float t = __shfl_sync(0xffffffffu, v, k);
        if (k < cnt) s = __fadd_rn(s, t);
    }
    return s;
}

static __device__ int adaptive_radius_warp(const Traj path, int n, int smoothing_radius, int lane) {
    // calculateAdaptiveRadius, Stabilizer.cpp:1637-1673 (same operations as adaptive_radius above)
    if (n < 10) return smoothing_radius;
    const int start = max(0, n - 20), c = n - start;
    const float cnt = (float)c;
    const bool on = lane < c;
    const float v0 = on ? path.at(start + lane, 0) : 0.f, v1 = on ? path.at(start + lane, 1) : 0.f, v2 = on ? path.at(start + lane, 2) : 0.f;
    const float mx = __fdiv_rn(warp_seq_sum<20>(v0, c), cnt), my = __fdiv_rn(warp_seq_sum<20>(v1, c), cnt), ma = __fdiv_rn(warp_seq_sum<20>(v2, c), cnt);
    const float dx = __fsub_rn(v0, mx), dy = __fsub_rn(v1, my), da = __fsub_rn(v2, ma);
    const float vx = __fdiv_rn(warp_seq_sum<20>(__fmul_rn(dx, dx), c), cnt), vy = __fdiv_rn(warp_seq_sum<20>(__fmul_rn(dy, dy), c), cnt),
                va = __fdiv_rn(warp_seq_sum<20>(__fmul_rn(da, da), c), cnt);
    const float total = f_sqrt(__fadd_rn(__fadd_rn(vx, vy), __fmul_rn(va, 1000.f)));
    return (int)fmaxf(5.0f, fminf(25.0f, __fmul_rn(total, 2.0f)));
}

static __device__ void box_at_warp(const Traj path, int n, int radius, int i, int drone, int lane, float* sm) {
    // boxFilterConvolve, Stabilizer.cpp:1139-1172, element i of all three components
    const int r = drone ? max(10, min(radius, 50)) : max(2, min(radius, 8));
    if (n <= r) { sm[0] = path.at(i, 0); sm[1] = path.at(i, 1); sm[2] = path.at(i, 2); return; }
    const int lo = max(0, i - r), hi = min(n - 1, i + r);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int base = lo; base <= hi; base += 32) {               // one pass unless drone mode widens the window
        const int c = min(32, hi - base + 1);
        const bool on = lane < c;
        const float v0 = on ? path.at(base + lane, 0) : 0.f, v1 = on ? path.at(base + lane, 1) : 0.f, v2 = on ? path.at(base + lane, 2) : 0.f;
        s0 = warp_seq_sum<32>(v0, c, s0); s1 = warp_seq_sum<32>(v1, c, s1); s2 = warp_seq_sum<32>(v2, c, s2);
    }
    const float w = (float)(hi - lo + 1);
    sm[0] = __fdiv_rn(s0, w); sm[1] = __fdiv_rn(s1, w); sm[2] = __fdiv_rn(s2, w);
}

static __device__ int motion_intent_warp(const Traj aux, int n_tr, float motion_da, int idx, int lane) {
    // analyzeMotionIntent, Stabilizer.cpp:1676-1719 (same operations as motion_intent above)
    const float mag = aux.at2(idx, 0);
    const float ang = (float)((double)__fmul_rn(fabsf(motion_da), 180.0f) / 3.14159265358979323846 * (double)30.0f);
    if (n_tr < 15) return 0;
    const int i0 = max(0, idx - 15);
    const int c = max(0, min(idx, n_tr) - i0);                  // the valid samples are a prefix of [i0, idx)
    if (c == 0) return 0;
    const bool on = lane < c;
    const float mg = on ? aux.at2(i0 + lane, 0) : 0.f, dr = on ? aux.at2(i0 + lane, 1) : 0.f;
    const float fc = (float)c;
    // variance_f(dirs)
    const float dm = __fdiv_rn(warp_seq_sum<15>(dr, c), fc);
    const float dd = __fsub_rn(dr, dm);
    const float dv = __fdiv_rn(warp_seq_sum<15>(__fmul_rn(dd, dd), c), fc);
    // consistency_f(mags)
    float mc = 0.f;
    if (c >= 2) {
        const float mm = __fdiv_rn(warp_seq_sum<15>(mg, c), fc);
        const float md = __fsub_rn(mg, mm);
        const float var = __fdiv_rn(warp_seq_sum<15>(__fmul_rn(md, md), c), fc);
        if (mm != 0.f) mc = fmaxf(0.f, fminf(1.f, __fdiv_rn(1.f, __fadd_rn(1.f, __fdiv_rn(var, __fmul_rn(mm, mm))))));
    }
    if (dv < 0.5f && mc > 0.7f && mag > 5.0f) return 1;
    if (mag < 3.0f && mc < 0.3f && ang > 10.0f) return 2;
    if (mag > 3.0f && mag < 15.0f && dv > 0.5f) return 3;
    return 0;
}

// cv::warpAffine's inversion of the float32 matrix promoted to double (imgwarp.cpp)
static __host__ __device__ void invert_affine(const float* T, double* m) {
    double M[6];
    for (int i = 0; i < 6; ++i) M[i] = (double)T[i];
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0. ? 1. / D : 0.;
    double A11 = M[4] * D, A22 = M[0] * D;
    M[0] = A11; M[1] *= -D; M[3] *= -D; M[4] = A22;
    double b1 = -M[0] * M[2] - M[1] * M[5];
    double b2 = -M[3] * M[2] - M[4] * M[5];
    M[2] = b1; M[5] = b2;
    for (int i = 0; i < 6; ++i) m[i] = M[i];
}
void warp_params_from_T(const float* T, WarpParams* wp) {
    invert_affine(T, wp->m);
    for (int i = 0; i < 6; ++i) wp->T[i] = T[i];
    wp->passthrough = 0;
    wp->da = 0.f;
}

// Smoothing + intent + matrix for the frame being emitted.  Runs on one thread (sequential float32
// arithmetic in the reference's order); S.gk is scratch for the gaussian taps.
// gaussianFilterConvolve's kernel (Stabilizer.cpp:1364-1383): size from sigma, taps exp(-x^2 / (2 sigma^2)) normalised by their
// float32 sum accumulated left to right.  Returns the size; taps are only written when it fits the scratch (<= 512).
static __device__ int gaussian_taps(float sigma, float* gk) {
    int ksz = max(3, (int)ceilf(__fmul_rn(6.f, sigma)));
    if ((ksz & 1) == 0) ++ksz;
    const int c = ksz / 2;
    if (ksz <= 512) {
        float tot = 0.f;
        for (int j = 0; j < ksz; ++j) {
            float x = (float)(j - c);
            float arg = __fdiv_rn(-__fmul_rn(x, x), __fmul_rn(__fmul_rn(2.f, sigma), sigma));
            float kv = (float)exp((double)arg);
            gk[j] = kv;
            tot = __fadd_rn(tot, kv);
        }
        for (int j = 0; j < ksz; ++j) gk[j] = __fdiv_rn(gk[j], tot);
    }
    return ksz;
}

// taps_ready: gk already holds gaussian_taps(info.gaussian_sigma) (clip mode computes them once per block)
//             sm_pre: clip mode, Kalman: this frame's filter output, computed by k_kalman_pass
static __device__ void smooth_and_setup(const LaneDev& L, WarpParams* wp_out, const StepInfo& info, float* gk, Traj path, Traj trf, Traj aux,
                                        bool taps_ready = false, const float* sm_pre = nullptr) {
    const int i = info.pop_index, n = info.path_len_at_pop;
    vs_output_record rec;
    rec.index = i; rec.passthrough = 0; rec.path_len = n; rec.radius = 0; rec.intent = 0;
    for (int k = 0; k < 3; ++k) rec.smoothed[k] = 0.f;
    for (int k = 0; k < 6; ++k) rec.T[k] = 0.f;
    WarpParams wp;
    if (i >= n) {                                           // Stabilizer.cpp:774-780
        rec.passthrough = 1;
        wp.passthrough = 1; wp.da = 0.f;
        for (int k = 0; k < 6; ++k) { wp.m[k] = (k == 0 || k == 4) ? 1. : 0.; wp.T[k] = (k == 0 || k == 4) ? 1.f : 0.f; }
        *wp_out = wp;
        if (info.n_out < L.record_capacity) L.orec[info.n_out] = rec;
        return;
    }
    float sm[3];
    bool have = false;
    if (info.method == 1) {                                 // gaussianFilterConvolve :1364-1413
        float sigma = info.gaussian_sigma;
        int ksz = max(3, (int)ceilf(__fmul_rn(6.f, sigma)));
        if ((ksz & 1) == 0) ++ksz;
        int c = ksz / 2;
        if (i - c < path.base) { path.p = L.path; path.base = 0; }      // window reaches below the staged tail
        if (n > c && ksz <= 512) {
            if (!taps_ready) gaussian_taps(sigma, gk);
            for (int comp = 0; comp < 3; ++comp) {
                float s = 0.f;
                for (int j = 0; j < ksz; ++j) {
                    int q = i + j;                          // index into the padded array
                    float v;
                    if (q < c) v = (c - q >= path.base) ? path.at(c - q, comp) : L.path[3 * (c - q) + comp];
                    else if (q < c + n) v = path.at(q - c, comp);
                    else v = path.at(n - 1 - (q - c - n), comp);
                    s = __fadd_rn(s, __fmul_rn(v, gk[j]));
                }
                sm[comp] = s;
            }
            have = true;
        }
    } else if (info.method == 2 && sm_pre) {
        sm[0] = sm_pre[0]; sm[1] = sm_pre[1]; sm[2] = sm_pre[2];
        have = true;
    } else if (info.method == 2) {                          // kalmanFilterSmooth :1416-1458, incremental
        for (int comp = 0; comp < 3; ++comp) {
            float* ks = L.kalman + 6 * comp;                // x0 x1 P00 P01 P10 P11
            float z = path.at(i, comp);
            if (i == 0) {
                ks[0] = z; ks[1] = 0.f; ks[2] = ks[3] = ks[4] = ks[5] = 0.f;
                sm[comp] = z;
            } else {
                // predict: x' = A x ; P' = A P A^T + Q, A = [1 1; 0 1], Q = 0.01 I
                float x0 = __fadd_rn(ks[0], ks[1]), x1 = ks[1];
                float t00 = __fadd_rn(ks[2], ks[4]), t01 = __fadd_rn(ks[3], ks[5]), t10 = ks[4], t11 = ks[5];
                float p00 = __fadd_rn(__fadd_rn(t00, t01), 0.01f), p01 = t01;
                float p10 = __fadd_rn(t10, t11), p11 = __fadd_rn(t11, 0.01f);
                // correct: H = [1 0], R = 0.1
                float s = __fadd_rn(p00, 0.1f);
                float k0 = __fdiv_rn(p00, s), k1 = __fdiv_rn(p01, s);
                float y = __fsub_rn(z, x0);
                ks[0] = __fadd_rn(x0, __fmul_rn(k0, y));
                ks[1] = __fadd_rn(x1, __fmul_rn(k1, y));
                ks[2] = __fsub_rn(p00, __fmul_rn(k0, p00)); ks[3] = __fsub_rn(p01, __fmul_rn(k0, p01));
                ks[4] = __fsub_rn(p10, __fmul_rn(k1, p00)); ks[5] = __fsub_rn(p11, __fmul_rn(k1, p01));
                sm[comp] = ks[0];
            }
        }
        have = true;
    }
    if (!have) {                                            // box with adaptive radius :808-823
        rec.radius = adaptive_radius(path, n, info.smoothing_radius);
        for (int comp = 0; comp < 3; ++comp) sm[comp] = box_at(path, comp, n, rec.radius, i, info.drone);
    }
    float raw[3], diff[3];
    for (int k = 0; k < 3; ++k) {
        raw[k] = trf.at(i, k);
        diff[k] = __fsub_rn(sm[k], path.at(i, k));
        rec.smoothed[k] = sm[k];
    }
    if (i > 0) {                                            // :854-888
        rec.intent = motion_intent(aux, n, raw, i);
        float sc = rec.intent == 1 ? 0.5f : rec.intent == 2 ? 1.0f : rec.intent == 3 ? 0.8f : 0.7f;
        for (int k = 0; k < 3; ++k) diff[k] = __fmul_rn(diff[k], sc);
    }
    float dx = __fadd_rn(raw[0], diff[0]), dy = __fadd_rn(raw[1], diff[1]), da = __fadd_rn(raw[2], diff[2]);
    if (info.horizon_lock) da = 0.f;
    float cs = f_cos(da), sn = f_sin(da);
    float T[6] = {cs, -sn, dx, sn, cs, dy};
    invert_affine(T, wp.m);
    for (int k = 0; k < 6; ++k) { wp.T[k] = T[k]; rec.T[k] = T[k]; }
    wp.passthrough = 0; wp.da = da;
    *wp_out = wp;
    if (info.n_out < L.record_capacity) L.orec[info.n_out] = rec;
}

// The same, executed by all 32 lanes of one warp (box smoothing; the gaussian / Kalman variants stay on lane 0).
static __device__ void smooth_and_setup_warp(const LaneDev& L, WarpParams* wp_out, const StepInfo& info, float* gk, Traj path, Traj trf,
                                             Traj aux, int lane) {
    const int i = info.pop_index, n = info.path_len_at_pop;
    if (i >= n || info.method != 0) {                       // pass-through frame, or a smoother with shared scratch
        if (lane == 0) smooth_and_setup(L, wp_out, info, gk, path, trf, aux);
        return;
    }
    vs_output_record rec;
    rec.index = i; rec.passthrough = 0; rec.path_len = n;
    float sm[3];
    rec.radius = adaptive_radius_warp(path, n, info.smoothing_radius, lane);       // :808-823
    box_at_warp(path, n, rec.radius, i, info.drone, lane, sm);
    float raw[3], diff[3];
    for (int k = 0; k < 3; ++k) {
        raw[k] = trf.at(i, k);
        diff[k] = __fsub_rn(sm[k], path.at(i, k));
        rec.smoothed[k] = sm[k];
    }
    rec.intent = 0;
    if (i > 0) {                                            // :854-888
        rec.intent = motion_intent_warp(aux, n, raw[2], i, lane);
        const float sc = rec.intent == 1 ? 0.5f : rec.intent == 2 ? 1.0f : rec.intent == 3 ? 0.8f : 0.7f;
        for (int k = 0; k < 3; ++k) diff[k] = __fmul_rn(diff[k], sc);
    }
    if (lane != 0) return;
    const float dx = __fadd_rn(raw[0], diff[0]), dy = __fadd_rn(raw[1], diff[1]);
    float da = __fadd_rn(raw[2], diff[2]);
    if (info.horizon_lock) da = 0.f;
    const float cs = f_cos(da), sn = f_sin(da);
    const float T[6] = {cs, -sn, dx, sn, cs, dy};
    WarpParams wp;
    invert_affine(T, wp.m);
    for (int k = 0; k < 6; ++k) { wp.T[k] = T[k]; rec.T[k] = T[k]; }
    wp.passthrough = 0; wp.da = da;
    *wp_out = wp;
    if (info.n_out < L.record_capacity) L.orec[info.n_out] = rec;
}

__global__ void __launch_bounds__(32) k_smooth_only(const LaneDev* __restrict__ lanes, StepInfo info) {
    __shared__ float gk[512];
    const LaneDev& L = lanes[blockIdx.z];
    smooth_and_setup_warp(L, L.wpb[info.wp_slot], info, gk, Traj{L.path, 0}, Traj{L.transforms, 0}, Traj{L.aux, 0}, threadIdx.x);
}

// phase 0: the whole step in one launch (single-stream engine, batch of one kernel).  The multi-stream engine splits it:
//   phase 1 (tracking stream, right behind LK, independent from frame to frame): status filter, RANSAC, refit -> L.fit[n]
//   phase 2 (motion stream, the sequential chain): decomposition, trajectory, records, smoothing, warp set-up
// so that only the short second half sits on the chain every frame has to wait for.
__global__ void __launch_bounds__(MO_THREADS) k_motion(const LaneDev* __restrict__ lanes, StepInfo info, int phase) {
    extern __shared__ unsigned char mo_raw[];
    MoSmem& S = *reinterpret_cast<MoSmem*>(mo_raw);
    const LaneDev& L = lanes[blockIdx.z];
    WarpParams* const wp_out = L.wpb[info.wp_slot];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    MotionFit* const fit_slot = L.fit + (info.frame_no & (VS_EV_RING - 1));
    int n_prev = phase == 2 ? fit_slot->n_prev : min(min(*L.kpc[info.kp_slot], L.kp_capacity), MO_MAXP);
    const int fidx = info.frame_no - 1;
    float2* lprev = nullptr; float2* lnext = nullptr; uint8_t* lstat = nullptr; uint8_t* lmask = nullptr;
    if (L.log_depth > 0) {
        size_t o = (size_t)(fidx % L.log_depth) * L.kp_capacity;
        lprev = L.log_prev + o; lnext = L.log_next + o; lstat = L.log_status + o; lmask = L.log_mask + o;
    }

    // ---- stage the tail of the trajectory (entries written by earlier steps) for the sequential part
    const int tail_base = max(0, fidx + 1 - MO_TAIL);
    if (phase != 1) {
        for (int i = tid; i < 3 * (fidx - tail_base); i += MO_THREADS) {
            S.tail_path[i] = L.path[3 * tail_base + i];
            S.tail_trf[i] = L.transforms[3 * tail_base + i];
        }
        for (int i = tid; i < 2 * (fidx - tail_base); i += MO_THREADS) S.tail_aux[i] = L.aux[2 * tail_base + i];
    }
    bool found = false;
    double A = 1., B = 0., TX = 0., TY = 0.;
    int n_inl = -1, n = 0, iters = 0;
    if (phase == 2) {
        const MotionFit f = *fit_slot;
        found = f.found != 0; A = f.A; B = f.B; TX = f.TX; TY = f.TY; n_inl = f.n_inl; n = f.n; iters = f.iters;
    } else {
    // ---- order-preserving compaction of the tracked pairs (status != 0)
    int running = 0;
    for (int base = 0; base < n_prev; base += MO_THREADS) {
        int i = base + tid;
        bool ok = false;
        float2 a = make_float2(0.f, 0.f), b = a;
        if (i < n_prev) {
            a = L.kpb[info.kp_slot][i]; b = L.lkn[info.lk_slot][i];
            uint8_t s = L.lks[info.lk_slot][i];
            ok = s != 0;
            if (lprev) { lprev[i] = a; lnext[i] = b; lstat[i] = s; }
        }
        unsigned m = __ballot_sync(FULL, ok);
        if (lane == 0) S.warp_tot[warp] = __popc(m);
        __syncthreads();
        int off = running;
        for (int w = 0; w < warp; ++w) off += S.warp_tot[w];
        if (ok) {
            int pos = off + __popc(m & ((1u << lane) - 1u));
            S.from[pos] = a; S.to[pos] = b;
        }
        int tot = 0;
        for (int w = 0; w < MO_THREADS / 32; ++w) tot += S.warp_tot[w];
        running += tot;
        __syncthreads();
    }
    n = running;

    if (tid == 0) {
        S.n = n; S.niters = VS_RANSAC_MAX_ITERS; S.iter = 0; S.max_good = 0; S.best_found = 0;
        S.rng = 0xFFFFFFFFFFFFFFFFull; S.cont = (n_prev > 0 && n >= 4) ? 1 : 0;
        S.fastmod_M = n > 0 ? 0xFFFFFFFFFFFFFFFFull / (unsigned)n + 1ull : 0ull;      // Lemire: r % n without a divide
    }
    __syncthreads();
    int batch = 8;               // the sequential loop usually stops after a handful of hypotheses

    // ---- RANSAC: rounds of 32 hypotheses, warp per hypothesis, sequential replay of the stop rule
    while (S.cont) {
        if (tid == 0) {
            unsigned long long r = S.rng;
            const unsigned long long M = S.fastmod_M;
            for (int h = 0; h < batch; ++h) {
                int i0 = (int)__umul64hi(M * rng_next(r), (unsigned long long)n), i1;
                do { i1 = (int)__umul64hi(M * rng_next(r), (unsigned long long)n); } while (i1 == i0);
                S.idx[h][0] = i0; S.idx[h][1] = i1;
            }
            S.rng = r;
        }
        __syncthreads();
        for (int hyp = warp; hyp < batch; hyp += MO_THREADS / 32) {
            double M[6];
            float F[6];
            int i0 = S.idx[hyp][0], i1 = S.idx[hyp][1];
            model_from_pair(S.from[i0], S.from[i1], S.to[i0], S.to[i1], M);
#pragma unroll
            for (int k = 0; k < 6; ++k) F[k] = (float)M[k];
            int cnt = 0;
            for (int i = lane; i < n + (32 - (n & 31)) % 32; i += 32) {
                bool in = i < n && is_inlier(F, S.from[i], S.to[i]);
                cnt += __popc(__ballot_sync(FULL, in));
            }
            if (lane == 0) S.good[hyp] = cnt;
        }
        __syncthreads();
        if (tid == 0) {
            int iter = S.iter, niters = S.niters, max_good = S.max_good, adv = 0, best = -1;
            for (int h = 0; h < batch && iter < niters; ++h, ++iter, ++adv) {
                int g = S.good[h];
                if (g > max(max_good, 1)) {
                    max_good = g; best = h;
                    niters = ransac_update_iters(0.99, (double)(n - g) / n, niters);
                }
            }
            // The RNG is consumed only by iterations that ran; draws for skipped hypotheses are
            // irrelevant because the loop ends (fresh RNG per call in the reference).
            S.iter = iter; S.niters = niters; S.max_good = max_good;
            if (best >= 0) {
                S.best_found = 1;
                double M[6];
                int i0 = S.idx[best][0], i1 = S.idx[best][1];
                model_from_pair(S.from[i0], S.from[i1], S.to[i0], S.to[i1], M);
                for (int k = 0; k < 6; ++k) { S.model[k] = M[k]; S.bestF[k] = (float)M[k]; }
            }
            S.cont = (iter < niters) ? 1 : 0;
        }
        __syncthreads();
        batch = MO_BATCH;
    }

    // ---- inlier mask of the winning hypothesis + least-squares similarity refit (double)
    found = S.best_found != 0;
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    if (found) {
        float F[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) F[k] = S.bestF[k];
        // pass 1: means over inliers
        double sx = 0, sy = 0, su = 0, sv = 0, cnt = 0;
        for (int i = tid; i < n; i += MO_THREADS) {
            bool in = is_inlier(F, S.from[i], S.to[i]);
            L.inlier_mask[i] = in;
            if (lmask) lmask[i] = in;
            if (in) { sx += S.from[i].x; sy += S.from[i].y; su += S.to[i].x; sv += S.to[i].y; cnt += 1.; }
        }
        acc[0] = sx; acc[1] = sy; acc[2] = su; acc[3] = sv; acc[4] = cnt;
    }
    auto block_reduce = [&](int nvals) {
        for (int k = 0; k < nvals; ++k) {
            double v = acc[k];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
            if (lane == 0) S.red[k][warp] = v;
        }
        __syncthreads();
        if (warp == 0) {
            for (int k = 0; k < nvals; ++k) {
                double v = lane < MO_THREADS / 32 ? S.red[k][lane] : 0.;
                for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
                if (lane == 0) S.red[k][0] = v;
            }
        }
        __syncthreads();
    };
    if (found) {                                            // uniform across the CTA
        block_reduce(5);
        double cnt = S.red[4][0];
        double mx = S.red[0][0] / cnt, my = S.red[1][0] / cnt, mu = S.red[2][0] / cnt, mv = S.red[3][0] / cnt;
        n_inl = (int)cnt;
        __syncthreads();
        for (int k = 0; k < 7; ++k) acc[k] = 0;
        float F2[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) F2[k] = S.bestF[k];
        for (int i = tid; i < n; i += MO_THREADS) {
            if (is_inlier(F2, S.from[i], S.to[i])) {    // recomputed: the global mask is shared scratch, two fits may be in flight
                double xc = S.from[i].x - mx, yc = S.from[i].y - my, uc = S.to[i].x - mu, vc = S.to[i].y - mv;
                acc[0] += xc * xc + yc * yc;
                acc[1] += xc * uc + yc * vc;
                acc[2] += xc * vc - yc * uc;
            }
        }
        block_reduce(3);
        double den = S.red[0][0];
        A = S.red[1][0] / den; B = S.red[2][0] / den;
        TX = mu - (A * mx - B * my);
        TY = mv - (B * mx + A * my);
    }
    iters = S.iter;
    if (phase == 1) {                                       // hand the fit to the motion stream and stop here
        if (tid == 0) {
            MotionFit f;
            f.A = A; f.B = B; f.TX = TX; f.TY = TY; f.found = found ? 1 : 0; f.n_inl = n_inl; f.n = n; f.n_prev = n_prev; f.iters = iters; f.pad = 0;
            *fit_slot = f;
        }
        return;
    }
    }   // phase != 2
    __syncthreads();

    // ---- decomposition, trajectory, record, smoothing (one thread; sequential float32 semantics)
    if (tid == 0) {
        float t[3] = {0.f, 0.f, 0.f};
        if (found) {
            float T00 = (float)A, T10 = (float)B;
            t[0] = (float)TX; t[1] = (float)TY;
            t[2] = f_atan2(T10, T00);
        }
        if (info.drone && n_prev > 0) hf_filters(L.hf, info, t);
        float* tr = L.transforms + 3 * fidx;
        float* pa = L.path + 3 * fidx;
        for (int k = 0; k < 3; ++k) {
            tr[k] = t[k];
            pa[k] = fidx == 0 ? t[k] : __fadd_rn(S.tail_path[3 * (fidx - 1 - tail_base) + k], t[k]);
            S.tail_trf[3 * (fidx - tail_base) + k] = t[k];
            S.tail_path[3 * (fidx - tail_base) + k] = pa[k];
        }
        {
            const float mag = f_sqrt(__fadd_rn(__fmul_rn(t[0], t[0]), __fmul_rn(t[1], t[1])));
            const float dir = f_atan2(t[1], t[0]);
            L.aux[2 * fidx] = mag; L.aux[2 * fidx + 1] = dir;
            S.tail_aux[2 * (fidx - tail_base)] = mag; S.tail_aux[2 * (fidx - tail_base) + 1] = dir;
        }
        if (fidx < L.record_capacity) {
            vs_frame_record& r = L.frec[fidx];
            r.frame_index = info.frame_no; r.n_prev_pts = n_prev; r.n_tracked = n;
            r.n_inliers = found ? n_inl : -1; r.ransac_iters = iters;
            if (!info.will_detect) r.n_detected = -1;          // else the detector (another stream) writes it
            for (int k = 0; k < 3; ++k) { r.transform[k] = t[k]; r.path[k] = pa[k]; }
            r.affine[0] = found ? A : 1.; r.affine[1] = found ? -B : 0.; r.affine[2] = found ? TX : 0.;
            r.affine[3] = found ? B : 0.; r.affine[4] = found ? A : 1.; r.affine[5] = found ? TY : 0.;
        }
        if (info.adaptive && info.frame_no >= 3) {          // adaptSmoothingRadius :1461-1492, :1562-1574
            float mag = f_sqrt(__fadd_rn(__fmul_rn(t[0], t[0]), __fmul_rn(t[1], t[1])));
            float sc = fmaxf(0.0f, fminf(1.0f, __fdiv_rn(mag, 50.0f)));
            sc = __fsub_rn(1.0f, sc);
            int nr = info.min_radius + (int)__fmul_rn(sc, (float)(info.max_radius - info.min_radius));
            L.kalman[VS_KAL_RADIUS_SLOT] = __int_as_float(nr);   // adaptive radius hand-off to the host
        }
        __threadfence();                                    // trajectory entries of this frame, for the global-memory view below
    }
    // ---- smoothing, intent and warp set-up of the frame being emitted: warp 0, ordered float32 sums fed by shuffles
    if (warp == 0 && info.pop_index >= 0) {
        __syncwarp();                                       // lane 0's tail entries are visible to the warp
        const bool in_tail = info.pop_index - (info.drone ? 50 : 20) >= tail_base || tail_base == 0;
        if (in_tail) smooth_and_setup_warp(L, wp_out, info, S.gk, Traj{S.tail_path, tail_base}, Traj{S.tail_trf, tail_base}, Traj{S.tail_aux, tail_base}, lane);
        else smooth_and_setup_warp(L, wp_out, info, S.gk, Traj{L.path, 0}, Traj{L.transforms, 0}, Traj{L.aux, 0}, lane);
    }
}

void launch_motion(const LaneDev* lanes, int n_lanes, StepInfo info, int phase, cudaStream_t st) {
    // per DEVICE, not per process (see launch_good_features)
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaFuncSetAttribute(k_motion, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MoSmem));
        attr_set[dev] = true;
    }
    k_motion<<<dim3(1, 1, n_lanes), MO_THREADS, sizeof(MoSmem), st>>>(lanes, info, phase);
}

void launch_smooth_only(const LaneDev* lanes, int n_lanes, StepInfo info, cudaStream_t st) {
    k_smooth_only<<<dim3(1, 1, n_lanes), 32, 0, st>>>(lanes, info);
}

// ------------------------------------------------------------------------------------------------
// Offline clip mode (temporal chunking): the per-frame transforms of the WHOLE clip
// are known up front (each rank analysed its chunk, one all-gather stitched them), so trajectory and
// smoothing become batch kernels.
//   k_traj_build    path_ = running sum of transforms_ in float32, STRICTLY in the reference's sequential
//                   order (Stabilizer.cpp:681-687) so the result is bit-identical to a streamed run; three
//                   lanes (x, y, angle) walk the clip tile by tile through shared memory, the |t| / atan2 terms
//                   are filled in parallel.
//   k_smooth_batch  one thread per output frame replays applyNextSmoothTransform's smoothing for the path
//                   length the streamed reference would have seen when that frame left the queue.
#define TRAJ_TILE 2048
__global__ void __launch_bounds__(256) k_traj_build(const LaneDev* __restrict__ lanes, int n_tr) {
    // The running sum is a chain of dependent float32 additions (one rounding per frame: no scan reproduces it), so the
    // chain itself stays on three threads; what a long clip needs is that each link costs an FADD, not a trip to HBM:
    // the block stages TRAJ_TILE frames in shared memory, the three chains run over the tile, the block writes it back.
    __shared__ float st[3 * TRAJ_TILE];
    const LaneDev& L = lanes[blockIdx.z];
    const int tid = threadIdx.x;
    for (int i = tid; i < n_tr; i += blockDim.x) {
        const float tx = L.transforms[3 * i], ty = L.transforms[3 * i + 1];
        L.aux[2 * i] = f_sqrt(__fadd_rn(__fmul_rn(tx, tx), __fmul_rn(ty, ty)));
        L.aux[2 * i + 1] = f_atan2(ty, tx);
    }
    float acc = 0.f;
    for (int base = 0; base < n_tr; base += TRAJ_TILE) {
        const int m = min(TRAJ_TILE, n_tr - base);
        for (int i = tid; i < 3 * m; i += blockDim.x) st[i] = L.transforms[3 * (size_t)base + i];
        __syncthreads();
        if (tid < 3) {
            for (int i = 0; i < m; ++i) {
                const float t = st[3 * i + tid];
                acc = (base + i == 0) ? t : __fadd_rn(acc, t);
                st[3 * i + tid] = acc;
            }
        }
        __syncthreads();
        for (int i = tid; i < 3 * m; i += blockDim.x) L.path[3 * (size_t)base + i] = st[i];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(128) k_smooth_batch(const LaneDev* __restrict__ lanes, StepInfo base, int first, int count,
                                                       int n_total, int gate, WarpParams* __restrict__ wps, int kal_from) {
    __shared__ float gk[512];
    const LaneDev& L = lanes[blockIdx.z];
    const int n_tr = n_total - 1;
    // box, gaussian and (the filter outputs being in wps[k].T already, k_kalman_pass) Kalman smoothing: one thread per output
    // frame replays its own window in the reference's order; the gaussian taps depend on sigma alone, once per block
    if (base.method == 1) {
        if (threadIdx.x == 0) gaussian_taps(base.gaussian_sigma, gk);
        __syncthreads();
    }
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int i = first + k;
    StepInfo info = base;
    info.pop_index = i;
    info.path_len_at_pop = min(i + gate - 1, n_tr);
    info.n_out = k;
    float pre[3] = {0.f, 0.f, 0.f};
    if (base.method == 2) { pre[0] = wps[k].T[0]; pre[1] = wps[k].T[1]; pre[2] = wps[k].T[2]; }
    smooth_and_setup(L, wps + k, info, gk, Traj{L.path, 0}, Traj{L.transforms, 0}, Traj{L.aux, 0}, true, base.method == 2 ? pre : nullptr);
}

// Clip mode, Kalman smoothing (kalmanFilterSmooth, Stabilizer.cpp:1416-1458): the filter is a recursion over the frames, one
// independent chain per component.  Three threads walk frames kal_from .. first + count - 1 with the state in registers (it lives
// in L.kalman between calls) and leave the outputs of frames >= first in wps[i - first].T[component] for k_smooth_batch.
__global__ void __launch_bounds__(32) k_kalman_pass(const LaneDev* __restrict__ lanes, int kal_from, int first, int count,
                                                    WarpParams* __restrict__ wps) {
    const LaneDev& L = lanes[blockIdx.z];
    const int comp = threadIdx.x;
    if (comp >= 3) return;
    float* ks = L.kalman + 6 * comp;                // x0 x1 P00 P01 P10 P11
    float k0s = ks[0], k1s = ks[1], k2s = ks[2], k3s = ks[3], k4s = ks[4], k5s = ks[5];
    for (int i = kal_from; i < first + count; ++i) {
        const float z = L.path[3 * (size_t)i + comp];
        float out;
        if (i == 0) {
            k0s = z; k1s = 0.f; k2s = k3s = k4s = k5s = 0.f;
            out = z;
        } else {
            // predict: x' = A x ; P' = A P A^T + Q, A = [1 1; 0 1], Q = 0.01 I
            float x0 = __fadd_rn(k0s, k1s), x1 = k1s;
            float t00 = __fadd_rn(k2s, k4s), t01 = __fadd_rn(k3s, k5s), t10 = k4s, t11 = k5s;
            float p00 = __fadd_rn(__fadd_rn(t00, t01), 0.01f), p01 = t01;
            float p10 = __fadd_rn(t10, t11), p11 = __fadd_rn(t11, 0.01f);
            // correct: H = [1 0], R = 0.1
            float sden = __fadd_rn(p00, 0.1f);
            float g0 = __fdiv_rn(p00, sden), g1 = __fdiv_rn(p01, sden);
            float y = __fsub_rn(z, x0);
            k0s = __fadd_rn(x0, __fmul_rn(g0, y));
            k1s = __fadd_rn(x1, __fmul_rn(g1, y));
            k2s = __fsub_rn(p00, __fmul_rn(g0, p00)); k3s = __fsub_rn(p01, __fmul_rn(g0, p01));
            k4s = __fsub_rn(p10, __fmul_rn(g1, p00)); k5s = __fsub_rn(p11, __fmul_rn(g1, p01));
            out = k0s;
        }
        if (i >= first) wps[i - first].T[comp] = out;
    }
    ks[0] = k0s; ks[1] = k1s; ks[2] = k2s; ks[3] = k3s; ks[4] = k4s; ks[5] = k5s;
}

void launch_traj_build(const LaneDev* lanes, int n_lanes, int n_tr, cudaStream_t st) {
    k_traj_build<<<dim3(1, 1, n_lanes), 256, 0, st>>>(lanes, n_tr);
}
void launch_smooth_batch(const LaneDev* lanes, int n_lanes, StepInfo base, int first, int count, int n_total, int gate,
                         WarpParams* wps, int kal_from, cudaStream_t st) {
    if (base.method == 2) k_kalman_pass<<<dim3(1, 1, n_lanes), 32, 0, st>>>(lanes, kal_from, first, count, wps);
    k_smooth_batch<<<dim3((count + 127) / 128, 1, n_lanes), 128, 0, st>>>(lanes, base, first, count, n_total, gate, wps, kal_from);
}
