// ref_shim.cpp — C entry points around the reference's OWN, UNMODIFIED translation units.  TEST INFRASTRUCTURE ONLY.
//
// oracle/build_ref.py compiles this file together with /root/reference/src/Stabilizer.cpp (read where it lies,
// never copied) against oracle/mini_cv (a stand-in for the OpenCV headers whose image operations call back
// into the real OpenCV of the cv2 wheel) into oracle/_ref/libvideostab_ref.so.  What runs through this library
// is therefore the reference's host logic itself: stabilize / generateTransform / applyNextSmoothTransform /
// boxFilterConvolve / gaussianFilterConvolve / kalmanFilterSmooth / calculateAdaptiveRadius / analyzeMotionIntent /
// the drone-HF chain / fade border — `Stabilizer.cpp` as written.
//
// Private members are reached by compiling THIS file with the access specifiers opened (the class layout does
// not depend on them); Stabilizer.cpp itself is compiled untouched.
#include <opencv2/opencv.hpp>

#define private public
#define protected public
#include "video/Stabilizer.h"
#undef private
#undef protected

#include <cstdio>
#include <cstring>
#include <string>

static const mini_cv_ops *g_ops = nullptr;
extern "C" void mini_cv_set_ops(const mini_cv_ops *ops) { g_ops = ops; }
extern "C" const mini_cv_ops *mini_cv_get_ops(void) { return g_ops; }

using vs::Stabilizer;
typedef Stabilizer::Parameters P;

#define NUM_FIELDS(X)                                                                                                  \
    X(useCuda) X(logging) X(smoothingRadius) X(maxCorners) X(qualityLevel) X(minDistance) X(blockSize) X(borderSize)   \
    X(cropNZoom) X(gaussianSigma) X(motionPrediction) X(horizonLock) X(orbFeatures) X(fastThreshold) X(useROI)         \
    X(adaptiveSmoothing) X(minSmoothingRadius) X(maxSmoothingRadius) X(outlierThreshold) X(intentionalMotionThreshold) \
    X(stageOneRadius) X(stageTwoRadius) X(useTemporalFiltering) X(temporalWindowSize) X(fadeAlpha) X(fadeDuration)     \
    X(motionThresholdLow) X(motionThresholdHigh) X(borderScaleFactor) X(rollCompensation) X(rollCompensationFactor)    \
    X(deepStabilization) X(separateTranslationRotation) X(useImuData) X(enableVirtualCanvas) X(canvasScaleFactor)      \
    X(temporalBufferSize) X(canvasBlendWeight) X(adaptiveCanvasSize) X(maxCanvasScale) X(minCanvasScale)               \
    X(preserveEdgeQuality) X(edgeBlendRadius) X(droneHighFreqMode) X(hfShakePx) X(hfAnalysisMaxWidth) X(hfRotLPAlpha)  \
    X(enableConditionalCLAHE) X(hfDeadZoneThreshold) X(hfFreezeDuration) X(hfMotionAccumulatorDecay)

template <typename T> static void assign_num(T &dst, double v) { dst = static_cast<T>(v); }

extern "C" {

void *vsref_params_new(void) { return new P(); }
void vsref_params_delete(void *p) { delete static_cast<P *>(p); }

// returns 0 when the field exists
int vsref_params_set_num(void *vp, const char *name, double v) {
    P &p = *static_cast<P *>(vp);
#define X(f) if (std::strcmp(name, #f) == 0) { assign_num(p.f, v); return 0; }
    NUM_FIELDS(X)
#undef X
    if (std::strcmp(name, "featureDetector") == 0) { p.featureDetector = static_cast<P::FeatureDetector>((int)v); return 0; }
    if (std::strcmp(name, "jitterFrequency") == 0) { p.jitterFrequency = static_cast<P::JitterFrequency>((int)v); return 0; }
    return 1;
}
int vsref_params_get_num(void *vp, const char *name, double *v) {
    P &p = *static_cast<P *>(vp);
#define X(f) if (std::strcmp(name, #f) == 0) { *v = static_cast<double>(p.f); return 0; }
    NUM_FIELDS(X)
#undef X
    return 1;
}
int vsref_params_set_str(void *vp, const char *name, const char *v) {
    P &p = *static_cast<P *>(vp);
    if (std::strcmp(name, "borderType") == 0) { p.borderType = v; return 0; }
    if (std::strcmp(name, "smoothingMethod") == 0) { p.smoothingMethod = v; return 0; }
    if (std::strcmp(name, "modelPath") == 0) { p.modelPath = v; return 0; }
    return 1;
}

void *vsref_new(void *params) {
    try {
        return new Stabilizer(*static_cast<P *>(params));
    } catch (const std::exception &e) {
        std::fprintf(stderr, "vsref_new: %s\n", e.what());
        return nullptr;
    }
}
void vsref_delete(void *s) { delete static_cast<Stabilizer *>(s); }
void vsref_clean(void *s) { static_cast<Stabilizer *>(s)->clean(); }

static int emit(const cv::Mat &out, unsigned char *dst, size_t cap, int *ow, int *oh) {
    if (out.empty()) return 0;
    if (out.type() != CV_8UC3) return -2;
    size_t need = (size_t)out.rows * out.cols * 3;
    *ow = out.cols;
    *oh = out.rows;
    if (need > cap) return -3;
    for (int y = 0; y < out.rows; y++) std::memcpy(dst + (size_t)y * out.cols * 3, out.ptr(y), (size_t)out.cols * 3);
    return 1;
}

// 1 = a frame was produced, 0 = empty Mat (not ready), < 0 = error.  The input is copied into a fresh Mat
// (the reference queues frames without cloning, Stabilizer.cpp:376; its callers hand it fresh clones).
int vsref_stabilize(void *s, const unsigned char *bgr, int w, int h, size_t stride, unsigned char *dst, size_t cap, int *ow, int *oh) {
    try {
        cv::Mat frame;
        if (bgr && w > 0 && h > 0) {
            frame.create(h, w, CV_8UC3);
            for (int y = 0; y < h; y++) std::memcpy(frame.ptr(y), bgr + (size_t)y * stride, (size_t)w * 3);
        }
        return emit(static_cast<Stabilizer *>(s)->stabilize(frame), dst, cap, ow, oh);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "vsref_stabilize: %s\n", e.what());
        return -1;
    }
}
// Timing variant (bench.py's CPU arm): the frame is a cv::Mat HEADER over the caller's buffer (no input copy: the reference
// queues the Mat it is given, Stabilizer.cpp:376, so the caller keeps the buffer alive) and the stabilized cv::Mat is only
// inspected, not copied out.  1 = produced, 0 = empty, < 0 = error.
int vsref_stabilize_nocopy(void *s, unsigned char *bgr, int w, int h, size_t stride, int *ow, int *oh) {
    try {
        cv::Mat frame(h, w, CV_8UC3, bgr, stride);
        cv::Mat out = static_cast<Stabilizer *>(s)->stabilize(frame);
        if (out.empty()) return 0;
        *ow = out.cols;
        *oh = out.rows;
        return 1;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "vsref_stabilize_nocopy: %s\n", e.what());
        return -1;
    }
}
int vsref_flush(void *s, unsigned char *dst, size_t cap, int *ow, int *oh) {
    try {
        return emit(static_cast<Stabilizer *>(s)->flush(), dst, cap, ow, oh);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "vsref_flush: %s\n", e.what());
        return -1;
    }
}

// ---- state read-back (private members)
int vsref_n_transforms(void *s) { return (int)static_cast<Stabilizer *>(s)->transforms_.size(); }
int vsref_queue_size(void *s) { return (int)static_cast<Stabilizer *>(s)->frameQueue_.size(); }
int vsref_smoothing_radius(void *s) { return static_cast<Stabilizer *>(s)->params_.smoothingRadius; }
static void copy3(const std::vector<cv::Vec3f> &v, float *dst, int cap) {
    int n = std::min((int)v.size(), cap);
    for (int i = 0; i < n; i++) { dst[3 * i] = v[i][0]; dst[3 * i + 1] = v[i][1]; dst[3 * i + 2] = v[i][2]; }
}
void vsref_get_transforms(void *s, float *dst, int cap) { copy3(static_cast<Stabilizer *>(s)->transforms_, dst, cap); }
void vsref_get_path(void *s, float *dst, int cap) { copy3(static_cast<Stabilizer *>(s)->path_, dst, cap); }
int vsref_n_smoothed(void *s) { return (int)static_cast<Stabilizer *>(s)->smoothedPath_.size(); }
void vsref_get_smoothed(void *s, float *dst, int cap) { copy3(static_cast<Stabilizer *>(s)->smoothedPath_, dst, cap); }
// virtual canvas stage called on its own, exactly as applyNextSmoothTransform does (Stabilizer.cpp:1130-1134): the private
// members updateTemporalFrameBuffer + applyVirtualCanvasStabilization on one frame and its correction.  transforms_ (read when the
// canvas is first sized) can be seeded with vsref_set_transforms.
int vsref_vc_apply(void *s, const unsigned char *bgr, int w, int h, size_t stride, const float *t3, unsigned char *dst, size_t cap, int *ow, int *oh) {
    try {
        auto *st = static_cast<Stabilizer *>(s);
        cv::Mat frame(h, w, CV_8UC3);
        for (int y = 0; y < h; y++) std::memcpy(frame.ptr(y), bgr + (size_t)y * stride, (size_t)w * 3);
        cv::Vec3f t(t3[0], t3[1], t3[2]);
        st->updateTemporalFrameBuffer(frame, t);
        return emit(st->applyVirtualCanvasStabilization(frame, t), dst, cap, ow, oh);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "vsref_vc_apply: %s\n", e.what());
        return -1;
    }
}
// virtual canvas (Stabilizer.cpp:2066-2167): the correction of the frame emitted last and the canvas scale in use
int vsref_vc_last(void *s, float *t3, float *scale) {
    auto *st = static_cast<Stabilizer *>(s);
    *scale = st->currentCanvasScale_;
    if (st->temporalTransformBuffer_.empty()) return 0;
    const cv::Vec3f &v = st->temporalTransformBuffer_.back();
    t3[0] = v[0]; t3[1] = v[1]; t3[2] = v[2];
    return (int)st->temporalTransformBuffer_.size();
}
int vsref_n_keypoints(void *s) { return (int)static_cast<Stabilizer *>(s)->prevKeypointsCPU_.size(); }
void vsref_get_keypoints(void *s, float *dst, int cap) {
    auto &v = static_cast<Stabilizer *>(s)->prevKeypointsCPU_;
    int n = std::min((int)v.size(), cap);
    for (int i = 0; i < n; i++) { dst[2 * i] = v[i].x; dst[2 * i + 1] = v[i].y; }
}

// ---- the pure-host functions, callable on their own (Stabilizer.cpp:1139-1172, 1364-1458, 1637-1780)
// `s` supplies params_ and — for analyzeMotionIntent — transforms_, which can be seeded with vsref_set_transforms.
void vsref_set_transforms(void *s, const float *t, int n) {
    auto &v = static_cast<Stabilizer *>(s)->transforms_;
    v.resize(n);
    for (int i = 0; i < n; i++) v[i] = cv::Vec3f(t[3 * i], t[3 * i + 1], t[3 * i + 2]);
}
static int put(const std::vector<float> &r, float *dst, int cap) {
    int n = std::min((int)r.size(), cap);
    std::memcpy(dst, r.data(), sizeof(float) * n);
    return (int)r.size();
}
int vsref_box_filter(void *s, const float *path, int n, float *dst) {
    return put(static_cast<Stabilizer *>(s)->boxFilterConvolve(std::vector<float>(path, path + n)), dst, n);
}
int vsref_gaussian_filter(void *s, const float *path, int n, float sigma, float *dst) {
    return put(static_cast<Stabilizer *>(s)->gaussianFilterConvolve(std::vector<float>(path, path + n), sigma), dst, n);
}
int vsref_kalman_filter(void *s, const float *path, int n, float *dst) {
    try {
        return put(static_cast<Stabilizer *>(s)->kalmanFilterSmooth(std::vector<float>(path, path + n)), dst, n);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "vsref_kalman_filter: %s\n", e.what());
        return -1;
    }
}
int vsref_adaptive_radius(void *s, const float *px, const float *py, const float *pa, int n) {
    return static_cast<Stabilizer *>(s)->calculateAdaptiveRadius(std::vector<float>(px, px + n), std::vector<float>(py, py + n),
                                                                 std::vector<float>(pa, pa + n));
}
int vsref_motion_intent(void *s, const float *motion3, int frame_index) {
    return (int)static_cast<Stabilizer *>(s)->analyzeMotionIntent(cv::Vec3f(motion3[0], motion3[1], motion3[2]), frame_index);
}
float vsref_stabilization_strength(void *s, int intent, const float *motion3) {
    return static_cast<Stabilizer *>(s)->calculateAdaptiveStabilizationStrength((vs::MotionIntent)intent, cv::Vec3f(motion3[0], motion3[1], motion3[2]));
}
float vsref_variance(void *s, const float *v, int n) { return static_cast<Stabilizer *>(s)->calculateVariance(std::vector<float>(v, v + n)); }
float vsref_consistency(void *s, const float *v, int n) { return static_cast<Stabilizer *>(s)->calculateConsistency(std::vector<float>(v, v + n)); }
void vsref_adapt_smoothing_radius(void *s, const float *motion3) {
    static_cast<Stabilizer *>(s)->adaptSmoothingRadius(cv::Vec3f(motion3[0], motion3[1], motion3[2]));
}
// drone-HF chain, one frame: dead zone -> micro shake -> rotation low pass -> history (Stabilizer.cpp:666-671 order)
void vsref_drone_chain(void *s, const float *in3, float *out3) {
    Stabilizer *st = static_cast<Stabilizer *>(s);
    vs::Transform t(in3[0], in3[1], in3[2]);
    t = st->applyDeadZoneFreeze(t);
    t = st->applyMicroShakeSuppression(t);
    t = st->applyRotationLowPass(t);
    st->updateTranslationHistory(cv::Vec2f(t.dx, t.dy));
    out3[0] = t.dx; out3[1] = t.dy; out3[2] = t.da;
}
void vsref_drone_analysis_size(void *s, int w, int h, int *aw, int *ah) {
    cv::Mat frame(h, w, CV_8UC3);
    cv::Size sz = static_cast<Stabilizer *>(s)->calculateDroneAnalysisSize(frame);
    *aw = sz.width;
    *ah = sz.height;
}

}  // extern "C"
