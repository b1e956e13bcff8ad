// video/RollCorrection.h — drop-in for the reference's include/video/RollCorrection.h (vs::RollCorrection), header-only, over the
// C-ABI of libvstab_b200 (include/vstab_b200.h, vs_roll_*).  Same nested Parameters (names and defaults of RollCorrection.h:16-38) and
// the same static entry point `cv::Mat autoCorrectRoll(const cv::Mat&, const Parameters&)` (RollCorrection.h:47).
//
// The reference keeps its smoothed angle in two process-global statics (src/RollCorrection.cpp:13-14); so does this shim: one device
// handle per process, created on first use and re-created when the parameters change (the reference would simply read the new values).
#ifndef ROLL_CORRECTION_H
#define ROLL_CORRECTION_H

#include <opencv2/core.hpp>

#include <cstring>
#include <stdexcept>
#include <string>

#include "../vstab_b200.h"

namespace vs {

class RollCorrection {
public:
    struct Parameters {
        double scaleFactor = 0.25;
        double cannyThresholdLow = 50.0;
        double cannyThresholdHigh = 150.0;
        int cannyAperture = 3;
        float houghRho = 1.0f;
        float houghTheta = static_cast<float>(3.1415926535897932384626433832795 / 180.0f);
        int houghThreshold = 100;
        double angleFilterMin = -10.0;
        double angleFilterMax = 10.0;
        double angleSmoothingAlpha = 0.1;
        double angleDecay = 0.995;
        double maxAngleChangeDeg = 0.5;
    };

    // Returns the roll-corrected frame (same size), an empty Mat for an empty or non-CV_8UC3 input, and — like the reference's
    // callers expect (vsg.cpp:1272-1274) — never throws on the hot path: on a device error the input frame comes back.
    static cv::Mat autoCorrectRoll(const cv::Mat& input) {          // examples/roll-correction-file.cpp:61 calls it like this
        const Parameters defaults;
        return autoCorrectRoll(input, defaults);
    }
    static cv::Mat autoCorrectRoll(const cv::Mat& input, const Parameters& params) {
        if (input.empty() || input.type() != CV_8UC3) return cv::Mat();
        State& s = state();
        vs_roll_params p;
        to_c(params, &p);
        if (!s.h || std::memcmp(&p, &s.p, sizeof(p)) != 0) {
            // carry the smoothed angle over a parameter change?  The reference's statics survive it; a new handle starts at the
            // decay-free first-frame state, which is what a fresh process would do.  Parameter changes are rare (config reload).
            if (s.h) vs_roll_destroy(s.h);
            s.h = nullptr;
            if (vs_roll_create(&p, 0, &s.h) != VS_OK) throw std::runtime_error(std::string("vs::RollCorrection: ") + vs_last_error());
            s.p = p;
        }
        cv::Mat out(input.rows, input.cols, CV_8UC3);
        const vs_status st = vs_roll_correct(s.h, input.data, input.cols, input.rows, (size_t)input.step, out.data, (size_t)out.step);
        return st == VS_OK ? out : input;
    }

    static void to_c(const Parameters& s, vs_roll_params* p) {
        std::memset(p, 0, sizeof(*p));
        p->scale_factor = s.scaleFactor; p->canny_threshold_low = s.cannyThresholdLow; p->canny_threshold_high = s.cannyThresholdHigh;
        p->canny_aperture = s.cannyAperture; p->hough_rho = s.houghRho; p->hough_theta = s.houghTheta; p->hough_threshold = s.houghThreshold;
        p->angle_filter_min = s.angleFilterMin; p->angle_filter_max = s.angleFilterMax; p->angle_smoothing_alpha = s.angleSmoothingAlpha;
        p->angle_decay = s.angleDecay; p->max_angle_change_deg = s.maxAngleChangeDeg;
    }

private:
    struct State {
        vs_roll* h = nullptr;
        vs_roll_params p{};
        ~State() { if (h) vs_roll_destroy(h); }
    };
    static State& state() { static State s; return s; }
};

}  // namespace vs

#endif  // ROLL_CORRECTION_H
