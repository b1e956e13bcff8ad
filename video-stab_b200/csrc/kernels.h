// kernels.h — host-callable launchers of the stabilization kernels (one .cu per stage).
#pragma once
#include "common.cuh"

#define VS_MAX_GROUP 64     // lanes per launch group (frame pointers travel as kernel arguments)

struct PtrPack {
    const uint8_t* p[VS_MAX_GROUP];
};
struct MutPtrPack {
    uint8_t* p[VS_MAX_GROUP];
};

// ---- k_pyramid.cu : resize + gray + pyrDown (Stabilizer.cpp:304-305,449-450,602; pyramid of :611)
// full-res BGR -> padded gray level `dst_level` of every lane's pyramid slot `slot`
// (slot < 0: the lanes' `small0` first-frame level)
// aw x ah: the analysis size the lanes' pyramids were allocated with (960 x 540 unless drone_high_freq_mode chose another)
void launch_gray_resize(const LaneDev* lanes, int n_lanes, const PtrPack& src, int w, int h, size_t stride,
                        int slot, cudaStream_t st, int aw = VS_AW, int ah = VS_AH);
// gray `small0` (480x270) -> level 0 of pyramid slot `slot`, cv::resize INTER_LINEAR up-sampling (:602)
void launch_upsample_small(const LaneDev* lanes, int n_lanes, int slot, cudaStream_t st, int aw = VS_AW, int ah = VS_AH);
// levels 1 and 2 of pyramid slot `slot` from level 0 (cv::pyrDown x2)
void launch_pyrdown(const LaneDev* lanes, int n_lanes, int slot, cudaStream_t st, int aw = VS_AW, int ah = VS_AH);
// generic cv::resize INTER_LINEAR on tightly addressed 8UC1/8UC3 (tests, crop+zoom second pass)
void launch_resize_linear(const uint8_t* src, int sw, int sh, size_t sstride, int ch,
                          uint8_t* dst, int dw, int dh, size_t dstride, cudaStream_t st);
// copy a tightly packed gray image into a padded level / back (tests)
void launch_pack_level(const uint8_t* src, GrayLevel dst, cudaStream_t st);
void launch_unpack_level(GrayLevel src, uint8_t* dst, cudaStream_t st);

// ---- k_gftt.cu : cv::goodFeaturesToTrack (Stabilizer.cpp:355-357, 740-744)
// source = pyramid slot `slot` level 0 (slot >= 0) or small0 (slot < 0); result -> lanes[].kp / kp_count
// (and first_corners when slot < 0).  record_frame_no > 0: also log into the frame record ring.
void launch_good_features(const LaneDev* lanes, int n_lanes, int slot, int max_corners, double quality,
                          double min_dist, int record_frame_no, int gen, int kp_slot, cudaStream_t st, int block_size = 3,
                          float* eig_scratch = nullptr, int aw = VS_AW, int ah = VS_AH);   // block_size != 3: per-pixel eigenvalue map in eig_scratch (n_lanes * w * h floats)
size_t gftt_grid_words(int w, int h, double min_dist);

// ---- k_lk.cu : cv::calcOpticalFlowPyrLK (Stabilizer.cpp:611-619)
// tracks lanes[].kp from pyramid slot `prev` to slot `cur`; writes lk_next / lk_status
void launch_pyr_lk(const LaneDev* lanes, int n_lanes, int prev, int cur, int max_pts, int kp_slot, int lk_slot, cudaStream_t st, bool tma = false);
// tensor maps of one lane's pyramid planes for the TMA tracker: out_host = VS_PYR_SLOTS * VS_LEVELS * 2 CUtensorMap (128 bytes each)
bool lk_encode_maps(const LaneDev& host_lane, void* out_host);

// ---- k_motion.cu : status filter + estimateAffinePartial2D + decomposition + trajectory +
//                    smoothing + warp set-up (Stabilizer.cpp:629-688, 783-908, 1139-1172, 1364-1458, 1637-1780)
// phase 0: whole step; 1: status filter + RANSAC + refit -> LaneDev::fit; 2: trajectory, smoothing, set-up from LaneDev::fit
void launch_motion(const LaneDev* lanes, int n_lanes, StepInfo info, int phase, cudaStream_t st);
// flush-time variant: no new frame, only the smoothing + warp set-up for `info.pop_index`
void launch_smooth_only(const LaneDev* lanes, int n_lanes, StepInfo info, cudaStream_t st);

// offline clip mode: trajectory of the whole clip + batched smoothing (k_motion.cu)
void launch_traj_build(const LaneDev* lanes, int n_lanes, int n_tr, cudaStream_t st);
void launch_smooth_batch(const LaneDev* lanes, int n_lanes, StepInfo base, int first, int count, int n_total, int gate,
                         WarpParams* wps, int kal_from, cudaStream_t st);

// ---- k_warp.cu : copyMakeBorder + warpAffine + crop/zoom (Stabilizer.cpp:981-990, 1056-1060, 1108-1124)
struct WarpGeom {
    int src_w, src_h;       // frame as pushed
    size_t src_stride;
    int mode;               // 0 plain, 1 border (output grows by 2b), 2 crop+zoom
    int border, border_mode;
    int out_w, out_h;
    size_t out_stride;
    int wp_slot;            // LaneDev::wpb index holding the warp set-up of this output
    void* d_tmaps;          // device scratch for VS_MAX_GROUP tensor maps (batches of more than 8 lanes), or nullptr
};
// both output-stage launchers return the number of kernels they launched
int launch_warp(const LaneDev* lanes, int n_lanes, const PtrPack& src, const MutPtrPack& dst, WarpGeom g,
                uint8_t* const* scratch, cudaStream_t st);
// stand-alone batched warp with host-supplied matrices (tests + roofline bench)
void launch_warp_matrices(const uint8_t* src, int sw, int sh, size_t sstride, size_t sframe,
                          uint8_t* dst, int dw, int dh, size_t dstride, size_t dframe,
                          const WarpParams* d_wp, int n_frames, cudaStream_t st);
void warp_params_from_T(const float* T, WarpParams* wp);   // host: cv::warpAffine's matrix inversion
// batched output stage for contiguous frames with device-resident warp set-ups (offline clip mode)
int launch_warp_frames_mode(const uint8_t* src, int sw, int sh, size_t sstride, size_t sframe, uint8_t* dst,
                            size_t dstride, size_t dframe, const WarpParams* d_wp, int n_frames, int mode, int border,
                            int border_mode, uint8_t* scratch, cudaStream_t st);

// border_type "fade" (Stabilizer.cpp:914-978, 1070-1106): history blend before the warp, history update after it
void launch_fade_blend(const PtrPack& frames, int n_lanes, int w, int h, size_t stride, int b, uint8_t* hist, uint8_t* blend,
                       float alpha, float beta, bool init, cudaStream_t st);
void launch_fade_update(uint8_t* hist, const MutPtrPack& outs, size_t out_stride, int n_lanes, int w, int h, int b, cudaStream_t st);

// vs::AutoZoomCrop (k_autozoom.cu): content mask on the device; the whole call on a device frame
void launch_content_mask(const uint8_t* d_bgr, int w, int h, size_t stride, uint8_t* d_mask, uint8_t* d_scratch, cudaStream_t st);
vs_status auto_zoom_crop_device(const uint8_t* d_bgr, int w, int h, size_t stride, uint8_t* d_out, size_t out_stride, size_t out_capacity,
                                int* ow, int* oh, cudaStream_t st);

// ---- k_nv12.cu : NV12 <-> packed BGR on the device (decoder / encoder hand-off; OpenCV's 8-bit BT.601 fixed point)
void launch_nv12_to_bgr(const uint8_t* y, size_t y_stride, const uint8_t* uv, size_t uv_stride, int w, int h, uint8_t* bgr,
                        size_t bgr_stride, cudaStream_t st);
void launch_bgr_to_nv12(const uint8_t* bgr, size_t bgr_stride, int w, int h, uint8_t* y, size_t y_stride, uint8_t* uv,
                        size_t uv_stride, cudaStream_t st);
