// include/video/Stabilizer.h — drop-in replacement for the reference header of the same path
// (OmerMersin/video-stab include/video/Stabilizer.h:70-198): same namespace, class, nested
// `Parameters` (field names and defaults of Stabilizer.h:76-175) and public methods
//     explicit Stabilizer(const Parameters&);  ~Stabilizer();
//     cv::Mat stabilize(const cv::Mat& frame);  cv::Mat flush();  void clean();
// Header-only shim over the C-ABI in vstab_b200.h; link with -lvstab_b200.  All arithmetic runs in
// the CUDA library; there is no CPU fallback (the constructor throws std::runtime_error when no
// sm_100 device is available).
#ifndef VIDEO_STABILIZER_HPP
#define VIDEO_STABILIZER_HPP

#include <opencv2/core.hpp>

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../vstab_b200.h"

namespace vs {

struct Transform {
    float dx = 0.0f, dy = 0.0f, da = 0.0f;
    Transform() = default;
    Transform(float x, float y, float a) : dx(x), dy(y), da(a) {}
};

class Stabilizer {
public:
    struct Parameters {
        bool useCuda = false;
        bool logging = false;
        int smoothingRadius = 30;
        int maxCorners = 200;
        double qualityLevel = 0.01;
        double minDistance = 30.0;
        int blockSize = 3;
        std::string borderType = "black";
        int borderSize = 0;
        bool cropNZoom = false;
        std::string smoothingMethod = "box";
        double gaussianSigma = 2.0;
        bool motionPrediction = true;
        bool horizonLock = false;
        enum FeatureDetector { GFTT, ORB, FAST, BRISK };
        FeatureDetector featureDetector = GFTT;
        int orbFeatures = 500;
        int fastThreshold = 10;
        bool useROI = false;
        cv::Rect roi = cv::Rect();
        bool adaptiveSmoothing = false;
        int minSmoothingRadius = 5;
        int maxSmoothingRadius = 50;
        double outlierThreshold = 3.0;
        double intentionalMotionThreshold = 20.0;
        int stageOneRadius = 10;
        int stageTwoRadius = 25;
        bool useTemporalFiltering = false;
        int temporalWindowSize = 5;
        float fadeAlpha = 0.1f;
        int fadeDuration = 30;
        float motionThresholdLow = 5.0f;
        float motionThresholdHigh = 20.0f;
        float borderScaleFactor = 2.0f;
        bool rollCompensation = true;
        double rollCompensationFactor = 0.75;
        bool deepStabilization = false;
        std::string modelPath = "";
        enum JitterFrequency { LOW, MEDIUM, HIGH, ADAPTIVE };
        JitterFrequency jitterFrequency = ADAPTIVE;
        bool separateTranslationRotation = true;
        bool useImuData = false;
        bool enableVirtualCanvas = false;
        float canvasScaleFactor = 1.5f;
        int temporalBufferSize = 30;
        float canvasBlendWeight = 0.7f;
        bool adaptiveCanvasSize = true;
        float maxCanvasScale = 2.0f;
        float minCanvasScale = 1.2f;
        bool preserveEdgeQuality = true;
        int edgeBlendRadius = 20;
        bool droneHighFreqMode = false;
        float hfShakePx = 1.5f;
        int hfAnalysisMaxWidth = 960;
        float hfRotLPAlpha = 0.2f;
        bool enableConditionalCLAHE = true;
        float hfDeadZoneThreshold = 2.0f;
        int hfFreezeDuration = 10;
        float hfMotionAccumulatorDecay = 0.9f;
    };

    explicit Stabilizer(const Parameters& params) : params_(params) {
        vs_params p;
        to_c(params, &p);
        int device = 0;
        vs_status st = vs_stabilizer_create(&p, device, &h_);
        if (st != VS_OK) throw std::runtime_error(std::string("vs::Stabilizer: ") + vs_last_error());
    }
    ~Stabilizer() { vs_stabilizer_destroy(h_); }
    Stabilizer(const Stabilizer&) = delete;
    Stabilizer& operator=(const Stabilizer&) = delete;

    // Returns the stabilized frame, or an empty Mat while the reference would (first frame, latency gate).
    // Like the reference it never throws on the hot path: on an internal error the input frame is returned.
    // A frame that is not CV_8UC3 is refused (empty Mat, lastStatus() == VS_ERR_INVALID_ARG): the reference would raise a
    // cv::Exception inside cvtColor for anything but 3-channel input, and reading e.g. a BGRA or 16-bit Mat as packed BGR would
    // run past its rows.
    cv::Mat stabilize(const cv::Mat& frame) {
        if (frame.empty()) return cv::Mat();
        if (frame.type() != CV_8UC3) {
            status_ = VS_ERR_INVALID_ARG;
            error_ = "vs::Stabilizer::stabilize: frame must be CV_8UC3 (BGR)";
            return cv::Mat();
        }
        const int b = (params_.borderSize > 0 && !params_.cropNZoom) ? params_.borderSize : 0;
        cols_hint_ = frame.cols + 2 * b;
        rows_hint_ = frame.rows + 2 * b;
        cv::Mat out(rows_hint_, cols_hint_, CV_8UC3);
        int ow = 0, oh = 0, produced = 0;
        vs_status st = vs_stabilizer_push(h_, frame.data, frame.cols, frame.rows, (size_t)frame.step, out.data,
                                          (size_t)out.step, (size_t)out.step * out.rows, &ow, &oh, &produced);
        note(st);
        if (st != VS_OK) return frame;
        if (!produced) return cv::Mat();
        if (ow != out.cols || oh != out.rows) return out(cv::Rect(0, 0, ow, oh));   // pass-through of the last frame
        return out;
    }
    cv::Mat flush() {
        if (rows_hint_ == 0) return cv::Mat();          // nothing was ever pushed
        cv::Mat out(rows_hint_, cols_hint_, CV_8UC3);
        int ow = 0, oh = 0, produced = 0;
        vs_status st = vs_stabilizer_flush(h_, out.data, (size_t)out.step, (size_t)out.step * out.rows, &ow, &oh, &produced);
        note(st);
        if (st != VS_OK || !produced) return cv::Mat();
        if (ow != out.cols || oh != out.rows) return out(cv::Rect(0, 0, ow, oh));
        return out;
    }
    void clean() { note(vs_stabilizer_clean(h_)); }

    // Not part of the reference's interface: the reference swallows every failure (it returns the input frame or an empty
    // Mat), which hides real errors; these two let a caller see the status of the last call.
    vs_status lastStatus() const { return status_; }
    const std::string& lastError() const { return error_; }

    static void to_c(const Parameters& s, vs_params* p) {
        vs_params_default(p);
        p->use_cuda = s.useCuda; p->logging = s.logging; p->smoothing_radius = s.smoothingRadius;
        p->max_corners = s.maxCorners; p->quality_level = s.qualityLevel; p->min_distance = s.minDistance;
        p->block_size = s.blockSize;
        std::strncpy(p->border_type, s.borderType.c_str(), sizeof(p->border_type) - 1);
        p->border_size = s.borderSize; p->crop_n_zoom = s.cropNZoom;
        std::strncpy(p->smoothing_method, s.smoothingMethod.c_str(), sizeof(p->smoothing_method) - 1);
        p->gaussian_sigma = s.gaussianSigma; p->motion_prediction = s.motionPrediction; p->horizon_lock = s.horizonLock;
        p->feature_detector = (int)s.featureDetector; p->orb_features = s.orbFeatures; p->fast_threshold = s.fastThreshold;
        p->use_roi = s.useROI; p->roi_x = s.roi.x; p->roi_y = s.roi.y; p->roi_width = s.roi.width; p->roi_height = s.roi.height;
        p->adaptive_smoothing = s.adaptiveSmoothing; p->min_smoothing_radius = s.minSmoothingRadius;
        p->max_smoothing_radius = s.maxSmoothingRadius; p->outlier_threshold = s.outlierThreshold;
        p->intentional_motion_threshold = s.intentionalMotionThreshold; p->stage_one_radius = s.stageOneRadius;
        p->stage_two_radius = s.stageTwoRadius; p->use_temporal_filtering = s.useTemporalFiltering;
        p->temporal_window_size = s.temporalWindowSize; p->fade_alpha = s.fadeAlpha; p->fade_duration = s.fadeDuration;
        p->motion_threshold_low = s.motionThresholdLow; p->motion_threshold_high = s.motionThresholdHigh;
        p->border_scale_factor = s.borderScaleFactor; p->roll_compensation = s.rollCompensation;
        p->roll_compensation_factor = s.rollCompensationFactor; p->deep_stabilization = s.deepStabilization;
        std::strncpy(p->model_path, s.modelPath.c_str(), sizeof(p->model_path) - 1);
        p->jitter_frequency = (int)s.jitterFrequency; p->separate_translation_rotation = s.separateTranslationRotation;
        p->use_imu_data = s.useImuData; p->enable_virtual_canvas = s.enableVirtualCanvas;
        p->canvas_scale_factor = s.canvasScaleFactor; p->temporal_buffer_size = s.temporalBufferSize;
        p->canvas_blend_weight = s.canvasBlendWeight; p->adaptive_canvas_size = s.adaptiveCanvasSize;
        p->max_canvas_scale = s.maxCanvasScale; p->min_canvas_scale = s.minCanvasScale;
        p->preserve_edge_quality = s.preserveEdgeQuality; p->edge_blend_radius = s.edgeBlendRadius;
        p->drone_high_freq_mode = s.droneHighFreqMode; p->hf_shake_px = s.hfShakePx;
        p->hf_analysis_max_width = s.hfAnalysisMaxWidth; p->hf_rot_lp_alpha = s.hfRotLPAlpha;
        p->enable_conditional_clahe = s.enableConditionalCLAHE; p->hf_dead_zone_threshold = s.hfDeadZoneThreshold;
        p->hf_freeze_duration = s.hfFreezeDuration; p->hf_motion_accumulator_decay = s.hfMotionAccumulatorDecay;
    }

private:
    void note(vs_status st) {
        status_ = st;
        if (st != VS_OK) error_ = vs_last_error();
    }
    Parameters params_;
    vs_status status_ = VS_OK;
    std::string error_;
    vs_stabilizer* h_ = nullptr;
    int cols_hint_ = 0, rows_hint_ = 0;      // output geometry remembered for flush()
};

}  // namespace vs

#endif  // VIDEO_STABILIZER_HPP
