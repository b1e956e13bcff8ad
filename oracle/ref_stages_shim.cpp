// ref_stages_shim.cpp — C entry points around the reference's OWN RollCorrection.cpp and AutoZoomCrop.cpp (SURVEY.md section 8f,
// ranks 1 and 2).  TEST INFRASTRUCTURE ONLY.  oracle/build_ref.py compiles this one translation unit into
// oracle/_ref/libstages_ref.so; it #includes the two reference sources where they lie under /root/reference/src (nothing is
// copied), so their file-scope statics (RollCorrection.cpp:13-14 sFirstFrame / sSmoothedAngle) can be read back.  The cv::cuda::
// calls are served by oracle/mini_cv/opencv2/mini_cv_cuda.hpp (CPU functions of the same OpenCV; residuals stated there).
#include <opencv2/opencv.hpp>
#include <opencv2/cudaarithm.hpp>

#include "RollCorrection.cpp"     // -I /root/reference/src
#include "AutoZoomCrop.cpp"

#include <cstdio>
#include <cstring>

static const mini_cv_ops *g_ops = nullptr;
extern "C" void mini_cv_set_ops(const mini_cv_ops *ops) { g_ops = ops; }
extern "C" const mini_cv_ops *mini_cv_get_ops(void) { return g_ops; }

typedef vs::RollCorrection::Parameters RP;

extern "C" {

void *vsroll_params_new(void) { return new RP(); }
void vsroll_params_delete(void *p) { delete static_cast<RP *>(p); }
int vsroll_params_set(void *vp, const char *name, double v) {
    RP &p = *static_cast<RP *>(vp);
#define F(f, T) if (std::strcmp(name, #f) == 0) { p.f = (T)v; return 0; }
    F(scaleFactor, double) F(cannyThresholdLow, double) F(cannyThresholdHigh, double) F(cannyAperture, int) F(houghRho, float)
    F(houghTheta, float) F(houghThreshold, int) F(angleFilterMin, double) F(angleFilterMax, double) F(angleSmoothingAlpha, double)
    F(angleDecay, double) F(maxAngleChangeDeg, double)
#undef F
    return 1;
}
int vsroll_params_get(void *vp, const char *name, double *v) {
    RP &p = *static_cast<RP *>(vp);
#define F(f) if (std::strcmp(name, #f) == 0) { *v = (double)p.f; return 0; }
    F(scaleFactor) F(cannyThresholdLow) F(cannyThresholdHigh) F(cannyAperture) F(houghRho) F(houghTheta) F(houghThreshold)
    F(angleFilterMin) F(angleFilterMax) F(angleSmoothingAlpha) F(angleDecay) F(maxAngleChangeDeg)
#undef F
    return 1;
}

static int emit(const cv::Mat &out, unsigned char *dst, size_t cap, int *ow, int *oh) {
    if (out.empty()) return 0;
    if (out.type() != CV_8UC3) return -2;
    *ow = out.cols;
    *oh = out.rows;
    if ((size_t)out.rows * out.cols * 3 > cap) return -3;
    for (int y = 0; y < out.rows; y++) std::memcpy(dst + (size_t)y * out.cols * 3, out.ptr(y), (size_t)out.cols * 3);
    return 1;
}

// vs::RollCorrection::autoCorrectRoll (RollCorrection.cpp:16-155).  1 = frame produced, 0 = empty, < 0 = error.
int vsroll_correct(void *params, const unsigned char *bgr, int w, int h, size_t stride, unsigned char *dst, size_t cap, int *ow, int *oh) {
    try {
        cv::Mat frame;
        if (bgr && w > 0 && h > 0) {
            frame.create(h, w, CV_8UC3);
            for (int y = 0; y < h; y++) std::memcpy(frame.ptr(y), bgr + (size_t)y * stride, (size_t)w * 3);
        }
        return emit(vs::RollCorrection::autoCorrectRoll(frame, *static_cast<RP *>(params)), dst, cap, ow, oh);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "vsroll_correct: %s\n", e.what());
        return -1;
    }
}
double vsroll_smoothed_angle(void) { return vs::sSmoothedAngle; }
int vsroll_first_frame(void) { return vs::sFirstFrame ? 1 : 0; }

// vs::AutoZoomCrop::autoZoomCrop (AutoZoomCrop.cpp:102-283)
int vszoom_crop(const unsigned char *bgr, int w, int h, size_t stride, double margin, unsigned char *dst, size_t cap, int *ow, int *oh) {
    try {
        cv::Mat frame;
        if (bgr && w > 0 && h > 0) {
            frame.create(h, w, CV_8UC3);
            for (int y = 0; y < h; y++) std::memcpy(frame.ptr(y), bgr + (size_t)y * stride, (size_t)w * 3);
        }
        return emit(vs::AutoZoomCrop::autoZoomCrop(frame, margin), dst, cap, ow, oh);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "vszoom_crop: %s\n", e.what());
        return -1;
    }
}

}  // extern "C"
