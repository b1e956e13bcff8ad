// roll.h — vs::RollCorrection (reference src/RollCorrection.cpp, include/video/RollCorrection.h) on the device: see k_roll.cu
#pragma once
#include "common.cuh"

struct RollState {            // device-resident; the reference keeps this in two file-scope statics (RollCorrection.cpp:13-14)
    double smoothed_angle;    // sSmoothedAngle, degrees
    int first;                // sFirstFrame
    int n_lines;              // lines found on the last frame
    float coef[6];            // inverse rotation, float (cv::cuda::buildWarpAffineMaps)
};
struct RollParamsDev {
    double angle_filter_min, angle_filter_max, angle_smoothing_alpha, angle_decay, max_angle_change_deg;
};

#define VS_TRY_ROLL(x) do { vs_status s__ = (x); if (s__ != VS_OK) return s__; } while (0)

class RollCorrector {
public:
    static vs_status create(const vs_roll_params& p, int device, RollCorrector** out);
    ~RollCorrector();
    // asynchronous on `st`; src and dst must not overlap
    vs_status correct_device(const uint8_t* d_src, int w, int h, size_t stride, uint8_t* d_dst, size_t dstride, cudaStream_t st);
    vs_status correct_host(const uint8_t* src, int w, int h, size_t stride, uint8_t* dst, size_t dstride);
    vs_status reset();
    vs_status state(RollState* out, int* n_edges);
    vs_status debug(uint8_t* gray, uint8_t* edges, float* lines, int cap_lines);
    int small_w() const { return sw_; }
    int small_h() const { return sh_; }
    uint64_t launches() const { return launches_; }

private:
    RollCorrector() = default;
    vs_status ensure(int w, int h);
    vs_roll_params p_{};
    int device_ = 0, w_ = 0, h_ = 0, sw_ = 0, sh_ = 0, numangle_ = 0, numrho_ = 0;
    float* d_tab_ = nullptr;          // tabSin[numangle] then tabCos[numangle]
    RollState* d_state_ = nullptr;
    float* d_lines_ = nullptr;        // (rho, theta, votes) of the last frame, in cv::HoughLines order
    unsigned long long* d_cand_ = nullptr;   // local maxima of the accumulator, unsorted keys
    uint8_t *d_gray_ = nullptr, *d_edges_ = nullptr, *d_in_ = nullptr, *d_out_ = nullptr;
    unsigned int *d_map_ = nullptr, *d_list_ = nullptr;
    int *d_queue_ = nullptr, *d_counters_ = nullptr, *d_accum_ = nullptr;
    uint64_t launches_ = 0;
};
