// common.cuh — device-side data layout shared by every kernel of the stabilization path.
//
// Vocabulary follows the reference (src/Stabilizer.cpp): a *lane* is one independent video
// stream (one vs::Stabilizer instance); lanes of a batch advance in lock-step so every stage
// is ONE kernel launch over all lanes (blockIdx.z = lane).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/vstab_b200.h"

#define VS_PAD 16            // reflect-101 border kept around every gray pyramid level
#define VS_WIN 15            // LK window (Stabilizer.cpp:616)
#define VS_PYR_SLOTS 12      // pyramids kept per lane (frames n-1, n for LK + the run-ahead of the pyramid stream)
#define VS_GUARD_GROUP 4     // the slot guard is taken once per this many frames
#define VS_EV_RING 16        // per-frame event rings (must exceed VS_PYR_SLOTS - 1, the distance of the slot guard)
#define VS_LEVELS 3          // maxLevel 2 (Stabilizer.cpp:617)
#define VS_WP_SLOTS 8        // warp set-up buffers (LaneDev::wpb), by output index % VS_WP_SLOTS
#define VS_KP_SLOTS 8        // key-point buffers (LaneDev::kpb / kpc), by (detection frame / 2) % VS_KP_SLOTS
#define VS_LK_SLOTS 16       // tracker output buffers (LaneDev::lkn / lks), by frame % VS_LK_SLOTS
#define VS_AW 960            // analysis size (Stabilizer.cpp:410)
#define VS_AH 540
#define VS_FW 480            // first-frame analysis size (Stabilizer.cpp:277)
#define VS_FH 270
#define VS_RANSAC_MAX_ITERS 500
#define VS_KAL_FLOATS 32      // per-lane scalar block: 3 x 6 Kalman state floats + hand-off slots
#define VS_KAL_RADIUS_SLOT 18
#define VS_GRID_SLOTS 4      // max accepted corners per min-distance cell (geometric bound)

// One gray level in HBM: `base` addresses pixel (0,0); rows are `pitch` bytes apart; a VS_PAD-wide
// BORDER_REFLECT_101 frame around the image is materialised so window/tap reads never branch.
struct GrayLevel {
    uint8_t* base;
    int w, h, pitch;
};

struct Pyramid {
    GrayLevel lv[VS_LEVELS];
};

// Inverse-mapped warp set-up, produced on device by the motion kernel, consumed by the warp kernel.
struct WarpParams {
    double m[6];        // inverted matrix (cv::warpAffine's M after invertAffineTransform)
    float  T[6];        // forward float32 matrix (for the record)
    int    passthrough; // 1: copy the frame unchanged (Stabilizer.cpp:774-780)
    float  da;          // the correction angle behind T (the virtual canvas stage needs it unrounded by cos / sin)
};

// Per-lane device state.  An array of these lives in HBM; kernels index it with blockIdx.z.
// result of the frame-independent half of the motion step (k_motion phase 1), consumed by the sequential half (phase 2)
struct MotionFit {
    double A, B, TX, TY;
    int found, n_inl, n, n_prev, iters, pad;
};

struct LaneDev {
    Pyramid pyr[VS_PYR_SLOTS];      // analysis pyramids, slot = frame % VS_PYR_SLOTS: the pyramid stream can run ahead of tracking
    GrayLevel small0;               // 480x270 gray of the very first frame
    unsigned int* eig_max;          // max(eig) as float bits (non-negative => orderable)
    unsigned long long* cand;       // corner candidates: (float bits << 32) | linear address
    int* cand_count;
    unsigned int* grid;             // min-distance grid: per cell [count, slot0..slot3]
    // second set of detection scratch (generation 1): two detections can be in flight on two streams
    unsigned int* eig_max2;
    unsigned long long* cand2;
    int* cand_count2;
    unsigned int* grid2;
    float2* kp;                     // "prevKeypointsCPU_"  (slot 0; the single-kernel entry points use these four)
    int* kp_count;
    float2* lk_next;
    uint8_t* lk_status;
    // double-buffered copies so that detection / tracking / motion of neighbouring frames can overlap on
    // different CUDA streams (engine.cu): key points by detection generation, tracker output by frame parity
    float2* kpb[VS_KP_SLOTS];       // key points by detection index % VS_KP_SLOTS: a detection may finish while the motion
    int* kpc[VS_KP_SLOTS];          // kernel two detections back still reads its slot
    float2* lkn[VS_LK_SLOTS];       // tracker output by frame % VS_LK_SLOTS: LK(n) may run while motion(n-3) still reads its slot
    uint8_t* lks[VS_LK_SLOTS];
    uint8_t* inlier_mask;
    float* transforms;              // 3 floats per frame  (transforms_)
    float* path;                    // 3 floats per frame  (path_)
    float* aux;                     // 2 floats per frame: |t| and atan2(ty,tx) of each transform (motion-intent terms)
    float* kalman;                  // 3 x {x0,x1,P00,P01,P10,P11} incremental Kalman state
    MotionFit* fit;                 // VS_EV_RING fit records, by frame % VS_EV_RING
    float* hf;                      // drone high-frequency state (VS_HF_* slots), set at creation only
    vs_frame_record* frec;
    vs_output_record* orec;
    float2* log_prev;               // ring of per-frame point logs (tests)
    float2* log_next;
    uint8_t* log_status;
    uint8_t* log_mask;
    float2* log_detected;
    float2* first_corners;
    int* first_count;
    WarpParams* wp;                 // == wpb[0] (single-kernel entry points, clip mode scratch)
    WarpParams* wpb[VS_WP_SLOTS];   // warp set-ups by output % VS_WP_SLOTS: the motion kernel runs ahead of the warps
    int kp_capacity;
    int log_depth;
    int record_capacity;
    int pad0;
    const void* lk_maps;            // device array of CUtensorMap, [pyramid slot][level][0: 48x18 template box, 1: 48x32 search box] over
                                    // the PADDED gray plane of that level (k_lk.cu); null = the tracker stages with plain loads
};

// Host-known scalars of one lock-step frame step, passed by value to the motion kernel.
// LaneDev::hf layout: translation history (x,y) x 10 oldest first, then the scalars of Stabilizer.cpp:143-153
#define VS_HF_HIST 0
#define VS_HF_COUNT 20
#define VS_HF_MEDIAN 21
#define VS_HF_ROT_LP 23
#define VS_HF_IN_DEAD_ZONE 24
#define VS_HF_FREEZE_COUNTER 25
#define VS_HF_ACCUMULATOR 26
#define VS_HF_FLOATS 32

// One wait guards every slot ring: before gray(n) overwrites pyramid slot n % VS_PYR_SLOTS, motion(n - VS_PYR_SLOTS + 1) must
// be complete.  The pyramid stream takes the wait once per VS_GUARD_GROUP frames, on motion(n - VS_PYR_SLOTS +
// VS_GUARD_GROUP) at the first frame n of a group, which covers the whole group (the motion stream is in order).  The other
// rings must not be reused any sooner than the pyramid ring:
static_assert(VS_LK_SLOTS >= VS_PYR_SLOTS - 1, "tracker-output slot n % VS_LK_SLOTS is last read by motion(n - VS_LK_SLOTS)");
static_assert(2 * VS_KP_SLOTS - 2 >= VS_PYR_SLOTS - 1, "key-point slot of detect(n) is last read by motion(n - 2 * VS_KP_SLOTS + 2)");
static_assert(VS_EV_RING > VS_PYR_SLOTS - 1 && (VS_EV_RING & (VS_EV_RING - 1)) == 0, "event ring");
static_assert(VS_GUARD_GROUP >= 1 && VS_GUARD_GROUP < VS_PYR_SLOTS - 2, "guard group");

struct StepInfo {
    int frame_no;          // n >= 1: this is the n-th generateTransform() call (frame n)
    int cur;               // pyramid slot holding the current frame
    int pop_index;         // index of the frame to emit after this step, or -1
    int path_len_at_pop;   // path_.size() when that frame is popped
    int smoothing_radius;  // params_.smoothingRadius as of this step
    int method;            // 0 box, 1 gaussian, 2 kalman
    float gaussian_sigma;
    int horizon_lock;
    int n_out;             // output-record slot
    int adaptive;          // adaptiveSmoothing
    int min_radius, max_radius;
    int kp_slot;           // key-point buffer this frame tracks from (LaneDev::kpb)
    int lk_slot;           // tracker output buffer of this frame (LaneDev::lkn / lks)
    int will_detect;       // corners are re-detected on this frame (the detector writes n_detected itself)
    int wp_slot;           // warp set-up buffer of the output produced by this step (LaneDev::wpb)
    int drone;             // droneHighFreqMode (Stabilizer.cpp:666-671, :1144-1146)
    float hf_shake_px, hf_rot_lp_alpha, hf_dead_zone_threshold, hf_accumulator_decay;
    int hf_freeze_duration;
};

// Detection scratch / key-point buffers of generation `gen` (0 / 1).  Plain selects on the lane record in global
// memory: a by-value copy of LaneDev with run-time indexed members would live on the local-memory stack.
struct DetView {
    unsigned int* eig_max;
    unsigned long long* cand;
    int* cand_count;
    unsigned int* grid;
    float2* kp;
    int* kp_count;
};
static __device__ __forceinline__ DetView det_view(const LaneDev& L, int gen, int kp_slot) {
    DetView v;
    v.eig_max = gen ? L.eig_max2 : L.eig_max;
    v.cand = gen ? L.cand2 : L.cand;
    v.cand_count = gen ? L.cand_count2 : L.cand_count;
    v.grid = gen ? L.grid2 : L.grid;
    v.kp = L.kpb[kp_slot];
    v.kp_count = L.kpc[kp_slot];
    return v;
}

static __device__ __forceinline__ int reflect101(int p, int len) {
    // cv::borderInterpolate(BORDER_REFLECT_101) for |overshoot| < len
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * len - 2 - p;
    }
    return p;
}

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

#define CUDA_TRY(x)                                                                   \
    do {                                                                              \
        cudaError_t e__ = (x);                                                        \
        if (e__ != cudaSuccess) return vs_set_cuda_error(e__, #x, __FILE__, __LINE__); \
    } while (0)

vs_status vs_set_cuda_error(cudaError_t e, const char* what, const char* file, int line);
vs_status vs_set_error(vs_status st, const char* msg);
