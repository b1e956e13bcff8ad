/* vstab_b200.h — C-ABI of the B200-native stabilization hot path.
 *
 * Drop-in boundary for OmerMersin/video-stab's `vs::Stabilizer` (reference
 * include/video/Stabilizer.h:70-198, src/Stabilizer.cpp).  Every entry point below names
 * the reference interface it replaces.  Plain pointers and sizes only; no C++/torch types.
 * The header-only C++ shim `include/video/Stabilizer.h` re-creates `vs::Stabilizer` on top
 * of this ABI; `INTEGRATION.md` shows the binding a maintainer would add.
 *
 * Conventions
 *  - every function returns a vs_status (0 = OK); nothing throws or aborts across the ABI
 *    (the reference never throws on the hot path either: Stabilizer.cpp:620-626,653-658,1061-1066)
 *  - "not ready yet" (reference: empty cv::Mat, Stabilizer.cpp:367,384-387) is *produced = 0
 *  - frames are 8-bit BGR interleaved (CV_8UC3), arbitrary row stride in bytes
 *  - one handle = one CUDA stream; calls on one handle must be externally serialised
 *    (same contract as the reference: examples/vsg.cpp:185-228); handles are independent
 *  - there is NO CPU fallback: without a CUDA device vs_stabilizer_create fails with
 *    VS_ERR_NO_DEVICE
 */
#ifndef VSTAB_B200_H
#define VSTAB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSTAB_B200_ABI_VERSION 1

typedef enum vs_status {
    VS_OK = 0,
    VS_ERR_INVALID_ARG = 1,
    VS_ERR_NO_DEVICE = 2,      /* no CUDA device / not an sm_100 part */
    VS_ERR_CUDA = 3,           /* a CUDA call failed; vs_last_error() has the text */
    VS_ERR_OUT_OF_MEMORY = 4,
    VS_ERR_BUFFER_TOO_SMALL = 5,
    VS_ERR_IO = 6,             /* config file unreadable */
    VS_ERR_UNSUPPORTED = 7     /* flag accepted by the reference but not built here yet */
} vs_status;

/* ---- vs::Stabilizer::Parameters (Stabilizer.h:76-175), field for field -------------------
 * std::string fields are fixed-size char arrays.  "live" = changes stabilize() output on the
 * reference's CPU path today; inert fields are stored and ignored, exactly
 * like the reference. */
typedef struct vs_params {
    int32_t use_cuda;                 /* useCuda — accepted, ignored (always GPU)           */
    int32_t logging;                  /* logging                                             */
    int32_t smoothing_radius;         /* smoothingRadius (30) live                           */
    int32_t max_corners;              /* maxCorners (200) live, first frame; 1..2048 (else UNSUPPORTED) */
    double  quality_level;            /* qualityLevel (0.01) live, first frame               */
    double  min_distance;             /* minDistance (30.0) live, first frame                */
    int32_t block_size;               /* blockSize (3) live, first frame only; 1..23          */
    char    border_type[32];          /* borderType ("black") live                           */
    int32_t border_size;              /* borderSize (0) live                                 */
    int32_t crop_n_zoom;              /* cropNZoom (false) live                              */
    char    smoothing_method[32];     /* smoothingMethod ("box") live                        */
    double  gaussian_sigma;           /* gaussianSigma (2.0) live with "gaussian"            */
    int32_t motion_prediction;        /* motionPrediction — inert                            */
    int32_t horizon_lock;             /* horizonLock (false) live                            */
    int32_t feature_detector;         /* featureDetector — inert (detectFeatures is dead)    */
    int32_t orb_features;             /* inert */
    int32_t fast_threshold;           /* inert */
    int32_t use_roi;                  /* inert */
    int32_t roi_x, roi_y, roi_width, roi_height;   /* inert */
    int32_t adaptive_smoothing;       /* adaptiveSmoothing (false) live                      */
    int32_t min_smoothing_radius;     /* (5) live with adaptive_smoothing                    */
    int32_t max_smoothing_radius;     /* (50) live with adaptive_smoothing                   */
    double  outlier_threshold;        /* inert */
    double  intentional_motion_threshold; /* inert */
    int32_t stage_one_radius;         /* inert */
    int32_t stage_two_radius;         /* inert */
    int32_t use_temporal_filtering;   /* inert */
    int32_t temporal_window_size;     /* inert */
    float   fade_alpha;               /* fadeAlpha (0.1) live with border_type "fade"        */
    int32_t fade_duration;            /* fadeDuration (30) live with border_type "fade"      */
    float   motion_threshold_low;     /* inert */
    float   motion_threshold_high;    /* inert */
    float   border_scale_factor;      /* inert */
    int32_t roll_compensation;        /* inert */
    double  roll_compensation_factor; /* inert */
    int32_t deep_stabilization;       /* inert */
    char    model_path[256];          /* inert */
    int32_t jitter_frequency;         /* inert */
    int32_t separate_translation_rotation; /* inert */
    int32_t use_imu_data;             /* inert */
    int32_t enable_virtual_canvas;    /* enableVirtualCanvas — single-stream handles, scales >= 1 */
    float   canvas_scale_factor;
    int32_t temporal_buffer_size;
    float   canvas_blend_weight;
    int32_t adaptive_canvas_size;
    float   max_canvas_scale;
    float   min_canvas_scale;
    int32_t preserve_edge_quality;
    int32_t edge_blend_radius;
    int32_t drone_high_freq_mode;     /* droneHighFreqMode (analysis size >= 62x62)          */
    float   hf_shake_px;
    int32_t hf_analysis_max_width;
    float   hf_rot_lp_alpha;
    int32_t enable_conditional_clahe;
    float   hf_dead_zone_threshold;
    int32_t hf_freeze_duration;
    float   hf_motion_accumulator_decay;
} vs_params;

/* Defaults of Stabilizer.h:78-174. */
vs_status vs_params_default(vs_params* p);
/* Reads the `stabilizer:` section of a reference config.yaml (keys of examples/vsg.cpp:1003-1114,
 * OpenCV FileStorage "%YAML:1.0" subset).  Missing keys keep their current value, like
 * cv::FileNode >> does. */
vs_status vs_params_from_yaml(const char* path, vs_params* p);
vs_status vs_params_from_yaml_string(const char* text, vs_params* p);

/* ---- per-frame record, for tests and diagnostics --------------------------------------- */
typedef struct vs_frame_record {
    int32_t frame_index;      /* n = index of the generateTransform() call (frame n), 1-based  */
    int32_t n_prev_pts;       /* key points LK started from                                    */
    int32_t n_tracked;        /* status != 0                                                   */
    int32_t n_inliers;        /* RANSAC inliers (-1: estimate not run / failed)                */
    int32_t ransac_iters;     /* hypotheses the sequential reference loop would have evaluated */
    int32_t n_detected;       /* corners re-detected on this frame (-1: no detection)          */
    float   transform[3];     /* dx, dy, da   (Stabilizer.cpp:660-662)                         */
    float   path[3];          /* cumulative   (Stabilizer.cpp:681-687)                         */
    double  affine[6];        /* refined 2x3 from the partial-affine fit                       */
} vs_frame_record;

typedef struct vs_output_record {
    int32_t index;            /* frame index this output belongs to                            */
    int32_t passthrough;      /* 1: returned un-warped (Stabilizer.cpp:774-780)                */
    int32_t path_len;
    int32_t radius;           /* adaptive box radius before the [2,8] clamp                    */
    int32_t intent;           /* 0 NORMAL 1 DELIBERATE_PAN 2 SHAKE_REMOVAL 3 FOLLOW_ACTION     */
    float   smoothed[3];
    float   T[6];             /* 2x3 float32 handed to the warp (Stabilizer.cpp:902-908)       */
} vs_output_record;

/* ---- the stabilizer handle -------------------------------------------------------------- */
typedef struct vs_stabilizer vs_stabilizer;

/* Stabilizer::Stabilizer(const Parameters&)  — Stabilizer.h:177, Stabilizer.cpp:50-164 */
vs_status vs_stabilizer_create(const vs_params* params, int device, vs_stabilizer** out);
/* Stabilizer::~Stabilizer()                  — Stabilizer.cpp:216-219 */
void      vs_stabilizer_destroy(vs_stabilizer* s);

/* cv::Mat Stabilizer::stabilize(const cv::Mat&) — Stabilizer.h:187, Stabilizer.cpp:258-392.
 * Host frame in, host frame out (synchronous).  out must hold out_capacity bytes;
 * worst case (width+2*border_size)*3 per row * (height+2*border_size) rows.
 * *produced = 0 while the reference would return an empty Mat. */
vs_status vs_stabilizer_push(vs_stabilizer* s, const uint8_t* bgr, int width, int height, size_t stride,
                             uint8_t* out, size_t out_stride, size_t out_capacity,
                             int* out_width, int* out_height, int* produced);
/* cv::Mat Stabilizer::flush()                — Stabilizer.h:193, Stabilizer.cpp:394-400 */
vs_status vs_stabilizer_flush(vs_stabilizer* s, uint8_t* out, size_t out_stride, size_t out_capacity,
                              int* out_width, int* out_height, int* produced);
/* n consecutive stabilize() calls in one: the loop of the reference's file-processing apps
 * (examples/file-capture.cpp:55-75: read frame, stabilize, keep if non-empty).  Frame k is at bgr + k*frame_step,
 * produced frame j is written at out + j*out_frame_capacity.  Results are identical to n vs_stabilizer_push() calls;
 * the difference is that frame k+1's host->device copy, frame k's kernels and frame k-1's device->host copy overlap
 * (copy-in / compute / copy-out streams), so throughput is bounded by PCIe rather than by the sum.  Returns when
 * all *n_produced frames are in `out`.  Host buffers should be page-locked for the copies to overlap. */
vs_status vs_stabilizer_push_many(vs_stabilizer* s, const uint8_t* bgr, size_t frame_step, int n_frames, int width, int height,
                                  size_t stride, uint8_t* out, size_t out_stride, size_t out_frame_capacity,
                                  int* out_width, int* out_height, int* n_produced);
/* the `while (!(f = flush()).empty())` drain loop, up to max_frames frames */
vs_status vs_stabilizer_flush_many(vs_stabilizer* s, uint8_t* out, size_t out_stride, size_t out_frame_capacity, int max_frames,
                                   int* out_width, int* out_height, int* n_produced);
/* void Stabilizer::clean()                   — Stabilizer.h:198, Stabilizer.cpp:221-256 */
vs_status vs_stabilizer_clean(vs_stabilizer* s);

/* Device-resident variants of stabilize()/flush(): `d_bgr` and `d_out` are device pointers on the
 * handle's device.  Asynchronous on the handle's stream; call vs_stabilizer_sync() before reading
 * d_out.  flags: VS_PUSH_BORROW — the caller keeps d_bgr alive and unmodified until the output of
 * that frame has been produced (the reference itself queues frames without cloning them,
 * Stabilizer.cpp:376); without it the frame is copied into an internal ring.
 * The handle's streams are non-blocking: they do NOT order against the caller's streams.  d_bgr must be complete
 * when the call is made — either synchronise its producer, or record a cudaEvent_t after the producer and hand it
 * to vs_stabilizer_wait_event() first (stream-ordered hand-off from a decoder, no host synchronisation). */
#define VS_PUSH_BORROW 1u
vs_status vs_stabilizer_push_device(vs_stabilizer* s, const uint8_t* d_bgr, int width, int height, size_t stride,
                                    uint8_t* d_out, size_t out_stride, size_t out_capacity, unsigned flags,
                                    int* out_width, int* out_height, int* produced);
/* The stabilize() loop over n_frames device-resident frames (frame k at d_bgr + k * frame_step) as one call; outputs are
 * written consecutively from d_out, out_frame_capacity bytes apart.  Asynchronous like vs_stabilizer_push_device. */
vs_status vs_stabilizer_push_many_device(vs_stabilizer* s, const uint8_t* d_bgr, size_t frame_step, int n_frames, int width, int height,
                                         size_t stride, uint8_t* d_out, size_t out_stride, size_t out_frame_capacity, unsigned flags,
                                         int* out_width, int* out_height, int* n_produced);
vs_status vs_stabilizer_flush_device(vs_stabilizer* s, uint8_t* d_out, size_t out_stride, size_t out_capacity,
                                     int* out_width, int* out_height, int* produced);
vs_status vs_stabilizer_sync(vs_stabilizer* s);
/* Everything pushed after this call waits (on the device) for `cuda_event` (a cudaEvent_t as void*). */
vs_status vs_stabilizer_wait_event(vs_stabilizer* s, void* cuda_event);
/* cudaStream_t of the handle (as void*), so callers can time/order work on it.  A handle runs its analysis and
 * corner-detection kernels on two further internal streams; every OUTPUT frame is produced on this public
 * stream.  vs_stabilizer_join() makes the public stream wait for everything enqueued so far on the internal
 * ones (so an event recorded on it afterwards covers all the handle's work) without blocking the host. */
void*     vs_stabilizer_stream(vs_stabilizer* s);
vs_status vs_stabilizer_join(vs_stabilizer* s);

/* Diagnostics: number of frames analysed so far / outputs produced so far, and their records
 * (synchronises the stream).  Used by the parity tests; not on the hot path. */
vs_status vs_stabilizer_counts(vs_stabilizer* s, int* n_frame_records, int* n_output_records);
vs_status vs_stabilizer_frame_record(vs_stabilizer* s, int i, vs_frame_record* rec);
vs_status vs_stabilizer_output_record(vs_stabilizer* s, int i, vs_output_record* rec);
/* Copies the points of frame record i: prev/next are n_prev_pts*2 floats, status n_prev_pts bytes,
 * inlier_mask n_tracked bytes, detected n_detected*2 floats (any pointer may be NULL). */
vs_status vs_stabilizer_frame_points(vs_stabilizer* s, int i, float* prev_xy, float* next_xy, uint8_t* status,
                                     uint8_t* inlier_mask, float* detected_xy);
/* corners found on the very first frame (Stabilizer.cpp:355-357); returns count in *n. */
vs_status vs_stabilizer_first_corners(vs_stabilizer* s, float* xy, int capacity, int* n);
/* kernels this handle has launched since creation (the bench's gpu_launches claim). */
vs_status vs_stabilizer_launch_count(vs_stabilizer* s, uint64_t* n);

/* Per-stage device timing with CUDA events on the handle's stream (off by default; bench/profiles).
 * stage: 0 resize+gray, 1 pyrDown, 2 PyrLK, 3 RANSAC+trajectory+smoothing, 4 GFTT, 5 warp, 6 / 7 copy-in / copy-out of the host path.
 * vs_*_stage_time synchronises and returns the summed duration and launch-group count since enabling. */
vs_status vs_stabilizer_set_timing(vs_stabilizer* s, int enable);
vs_status vs_stabilizer_stage_time(vs_stabilizer* s, int stage, double* total_ms, long long* count);
/* Diagnostics: the timeline of the stage launches timed since the previous query, as (stage, start us, end us) float
 * triples relative to the first of them; *n = number of triples available (at most `capacity` are written). */
vs_status vs_stabilizer_trace(vs_stabilizer* s, float* out, int capacity, int* n);

/* ---- offline clip mode: one temporal chunk of a long clip per handle / GPU (BASELINE config 5) -----------
 * The reference has no such mode (it is single-stream, single-process); results are defined as "identical to
 * pushing the whole clip through stabilize()+flush() on one handle".  Motion estimation is pairwise-local, so
 * a chunk [first, first+count) needs only vs_clip_halo(first) <= 2 leading frames of pixels; the trajectory is
 * then stitched from ALL chunks' transforms (a few hundred KB: one all-gather) and accumulated in the
 * reference's sequential float32 order, so every rank's path_ is bit-identical to a single-GPU run.
 *   vs_clip_analyze: d_frames = frames [first - vs_clip_halo(first), first + count), tight rows, device memory.
 *                    Writes transforms_[n-1] for n = max(first,1) .. first+count-1 (3 floats each) to host.
 *   vs_clip_render : all_transforms_host = the n_total-1 transforms of the whole clip; d_frames = frames
 *                    [first, first+count); d_out = count output frames, tight rows of out_width*3 bytes.
 * Not available with adaptive_smoothing (the latency gate becomes data dependent). */
int       vs_clip_halo(int first);
vs_status vs_clip_analyze(vs_stabilizer* s, const uint8_t* d_frames, int width, int height, int first, int count,
                          float* transforms_out_host, int* n_out);
vs_status vs_clip_render(vs_stabilizer* s, const float* all_transforms_host, int n_total, const uint8_t* d_frames,
                         int width, int height, int first, int count, uint8_t* d_out, int* out_width, int* out_height);
/* Device-resident variants: the transforms stay in device memory, so the exchange step (an NCCL all-gather of
 * 12 bytes per frame) reads and writes them in place and nothing on the path blocks the host.  Both calls are
 * asynchronous on vs_stabilizer_stream(): order the collective after vs_clip_analyze_device and
 * vs_clip_render_device after the collective with events on that stream (or vs_stabilizer_wait_event). */
vs_status vs_clip_analyze_device(vs_stabilizer* s, const uint8_t* d_frames, int width, int height, int first, int count,
                                 float* d_transforms_out, int* n_out);
vs_status vs_clip_render_device(vs_stabilizer* s, const float* d_all_transforms, int n_total, const uint8_t* d_frames,
                                int width, int height, int first, int count, uint8_t* d_out, int* out_width, int* out_height);
/* vs_clip_render_device in two steps, for a rank that renders several chunks of the same clip: the transform list is
 * installed (and the float32 trajectory rebuilt in the reference's order) ONCE, then each chunk is smoothed and warped. */
vs_status vs_clip_set_transforms_device(vs_stabilizer* s, const float* d_all_transforms, int n_total, int width, int height);
vs_status vs_clip_render_prepared_device(vs_stabilizer* s, const uint8_t* d_frames, int width, int height, int first, int count,
                                         uint8_t* d_out, int* out_width, int* out_height);

/* ---- decoder / encoder hand-off ---------------------------------------------------
 * The reference ingests BGR from a GStreamer appsink and hands BGR to an appsrc; the colour conversion from and to the codecs'
 * NV12 happens on the CPU inside its pipelines (`videoconvert`, examples/vsg.cpp:91-134, 230-311).  With a hardware decoder the
 * frames are NV12 surfaces in device memory: these two calls convert on the device, so decode -> vs_stabilizer_push_device
 * (VS_PUSH_BORROW) -> encode never touches the host.  Asynchronous on `stream` (a cudaStream_t as void*, NULL = the default
 * stream); order the stabilizer behind the conversion with vs_stabilizer_wait_event.  d_y: height rows of width bytes;
 * d_uv: height / 2 rows of width bytes (U, V interleaved); width and height even.  Arithmetic: OpenCV's 8-bit BT.601
 * limited-range fixed point, bit-exact with cv::cvtColor(COLOR_YUV2BGR_NV12) and cv::cvtColor(COLOR_BGR2YUV_I420). */
vs_status vs_nv12_to_bgr_device(const uint8_t* d_y, size_t y_stride, const uint8_t* d_uv, size_t uv_stride, int width, int height,
                                uint8_t* d_bgr, size_t bgr_stride, void* stream);
vs_status vs_bgr_to_nv12_device(const uint8_t* d_bgr, size_t bgr_stride, int width, int height, uint8_t* d_y, size_t y_stride,
                                uint8_t* d_uv, size_t uv_stride, void* stream);

/* ---- roll correction ------------------------------------------------------------
 * Replaces vs::RollCorrection::autoCorrectRoll(input, params) (reference include/video/RollCorrection.h:16-51,
 * src/RollCorrection.cpp:16-155), the stage that runs immediately before stabilize() in the reference's pipeline
 * (examples/vsg.cpp:1272-1285): downscale -> gray -> Canny -> Hough lines -> mean angle in the filter band ->
 * exponential smoothing with a per-frame clamp (decay when no line qualifies) -> rotate the full frame about its centre
 * (INTER_LINEAR, BORDER_REPLICATE).  The reference keeps the smoothed angle in two process-global statics; here it is
 * per handle and lives on the device.  Output size == input size.  Fields mirror RollCorrection::Parameters; YAML keys are
 * those of the `roll_correction:` section (examples/vsg.cpp:988-1000). */
typedef struct vs_roll_params {
    double  scale_factor;            /* scaleFactor (0.25)          */
    double  canny_threshold_low;     /* cannyThresholdLow (50)      */
    double  canny_threshold_high;    /* cannyThresholdHigh (150)    */
    int32_t canny_aperture;          /* cannyAperture (3); only 3   */
    float   hough_rho;               /* houghRho (1)                */
    float   hough_theta;             /* houghTheta (pi/180)         */
    int32_t hough_threshold;         /* houghThreshold (100)        */
    double  angle_filter_min;        /* angleFilterMin (-10)        */
    double  angle_filter_max;        /* angleFilterMax (+10)        */
    double  angle_smoothing_alpha;   /* angleSmoothingAlpha (0.1)   */
    double  angle_decay;             /* angleDecay (0.995)          */
    double  max_angle_change_deg;    /* maxAngleChangeDeg (0.5)     */
} vs_roll_params;
typedef struct vs_roll vs_roll;
vs_status vs_roll_params_default(vs_roll_params* p);
vs_status vs_roll_params_from_yaml(const char* path, vs_roll_params* p);
vs_status vs_roll_params_from_yaml_string(const char* text, vs_roll_params* p);
vs_status vs_roll_create(const vs_roll_params* params, int device, vs_roll** out);
void      vs_roll_destroy(vs_roll* r);
/* host frame in, host frame out (synchronous), like the reference's cv::Mat -> cv::Mat call */
vs_status vs_roll_correct(vs_roll* r, const uint8_t* bgr, int width, int height, size_t stride, uint8_t* out, size_t out_stride);
/* device frame in, device frame out, asynchronous on `stream` (a cudaStream_t; NULL = the default stream) */
vs_status vs_roll_correct_device(vs_roll* r, const uint8_t* d_bgr, int width, int height, size_t stride, uint8_t* d_out,
                                 size_t out_stride, void* stream);
vs_status vs_roll_reset(vs_roll* r);                                   /* sFirstFrame = true */
/* diagnostics (synchronise): smoothed angle in degrees, lines / edge pixels found on the last frame, kernels launched */
vs_status vs_roll_state(vs_roll* r, double* smoothed_angle_deg, int* n_lines, int* n_edges, uint64_t* launches);
/* analysis image of the last frame: size, gray and edge planes (small_w * small_h bytes each), lines as (rho, theta, votes) */
vs_status vs_roll_debug(vs_roll* r, int* small_w, int* small_h, uint8_t* gray_out, uint8_t* edges_out, float* lines_out, int lines_capacity);

/* ---- auto zoom-crop ------------------------------------------------------------
 * Replaces vs::AutoZoomCrop::autoZoomCrop(corrected, marginPercent) (reference include/video/AutoZoomCrop.h:8-16,
 * src/AutoZoomCrop.cpp:102-283): finds the largest axis-aligned rectangle of the frame's aspect ratio inside the non-black
 * content (the black corners a rotation leaves), crops it and scales it to the reference's hard-coded 640 x 360
 * (AutoZoomCrop.cpp:246-261; marginPercent is ignored there and here).  Device: gray, threshold, 5x5 elliptic close, crop +
 * warpAffine scale.  Host (as in the reference, which downloads the mask and calls cv::findContours): border following of the
 * content region and the greedy rectangle-shrinking loop.  Output is out_width x out_height (640 x 360, or the input size when
 * the reference would hand the frame back unchanged). */
vs_status vs_auto_zoom_crop(const uint8_t* bgr, int width, int height, size_t stride, double margin_percent, int device,
                            uint8_t* out, size_t out_stride, size_t out_capacity, int* out_width, int* out_height);
vs_status vs_auto_zoom_crop_device(const uint8_t* d_bgr, int width, int height, size_t stride, double margin_percent,
                                   uint8_t* d_out, size_t out_stride, size_t out_capacity, int* out_width, int* out_height,
                                   void* stream);
/* ---- Stabilizer::applyVirtualCanvasStabilization on its own (Stabilizer.cpp:2066-2443) ----------------------------------
 * The output stage `enable_virtual_canvas` switches on inside vs_stabilizer_push*, as a handle of its own: it keeps the temporal
 * frame buffer (device) and the canvas size.  One call = updateTemporalFrameBuffer (:2153-2167) + applyVirtualCanvasStabilization
 * for one frame with its correction (dx, dy, da).  `recent_transforms` (3 floats each, oldest first; the tail of the reference's
 * transforms_) is read by the first call only, where adaptive_canvas_size sizes the canvas (:2280-2314).  Device frames;
 * the call synchronises `stream` once (twice on frames that contain pixels of gray <= 1, whose mask goes to the host for
 * cv::findContours as in the reference) and returns with the output enqueued.  Only the canvas fields of vs_params are read. */
typedef struct vs_canvas vs_canvas;
vs_status vs_canvas_create(const vs_params* params, int device, vs_canvas** out);
void      vs_canvas_destroy(vs_canvas* c);
vs_status vs_canvas_apply_device(vs_canvas* c, const uint8_t* d_bgr, int width, int height, size_t stride, const float* transform3,
                                 const float* recent_transforms, int n_recent, uint8_t* d_out, size_t out_stride, void* stream);
/* canvas scale chosen by the first call (0 before it) and the number of regions the last call filled */
vs_status vs_canvas_info(vs_canvas* c, float* scale, int* regions_filled);

/* the host half on its own: crop rectangle from a (closed) content mask in host memory.  found = 0: no contour (frame unchanged) */
vs_status vs_auto_zoom_rect_from_mask(const uint8_t* mask, int width, int height, size_t stride, int* x, int* y, int* w, int* h,
                                      int* found);
/* cv::findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) of a host mask: points_xy (2 ints per point, contours back to back) and the
 * length of every contour, in cv::findContours order (tests) */
vs_status vs_k_find_external_contours(const uint8_t* mask, int width, int height, size_t stride, int* points_xy, int points_capacity,
                                      int* lengths, int lengths_capacity, int* n_contours);
/* the content mask of a device frame (gray > 1, then the 5x5 elliptic close), device -> device; `stream` as above */
vs_status vs_k_content_mask(const uint8_t* d_bgr, int width, int height, size_t stride, uint8_t* d_mask, uint8_t* d_scratch, void* stream);

/* ---- multi-stream batch: N independent streams advanced in lock-step, one kernel launch per
 * stage for the whole batch (BASELINE config 4).  Semantically N vs_stabilizers. -------------- */
typedef struct vs_batch vs_batch;
vs_status vs_batch_create(const vs_params* params, int device, int n_streams, vs_batch** out);
void      vs_batch_destroy(vs_batch* b);
/* d_frames / d_outs: host arrays of n_streams device pointers (same geometry for all streams). */
vs_status vs_batch_push_device(vs_batch* b, const uint8_t* const* d_frames, int width, int height, size_t stride,
                               uint8_t* const* d_outs, size_t out_stride, size_t out_capacity, unsigned flags,
                               int* out_width, int* out_height, int* produced);
vs_status vs_batch_flush_device(vs_batch* b, uint8_t* const* d_outs, size_t out_stride, size_t out_capacity,
                                int* out_width, int* out_height, int* produced);
/* Offline clip mode, lock-step: the batch's n_streams lanes analyse n_streams temporal chunks of a long clip at once (one
 * launch per stage for all of them instead of one per frame and chunk).  Lane l analyses `count` frames starting at an EVEN
 * clip frame number first_l >= 4; d_frames[l] points at frame first_l - 2 (the two-frame halo of vs_clip_halo), tight rows,
 * count + 2 frames.  d_transforms_out[l] (device memory) receives the `count` transforms of generateTransform calls
 * first_l .. first_l + count - 1, bit-identical to vs_clip_analyze_device on the same chunk (motion estimation is
 * pairwise-local, Stabilizer.cpp:594-678: only the parity of the chunk start matters).  Asynchronous on the batch's stream. */
vs_status vs_batch_clip_analyze_device(vs_batch* b, const uint8_t* const* d_frames, int width, int height, int count,
                                       float* const* d_transforms_out);
/* as vs_stabilizer_wait_event: everything pushed after this call waits (on the device) for `cuda_event` */
vs_status vs_batch_wait_event(vs_batch* b, void* cuda_event);
vs_status vs_batch_sync(vs_batch* b);
vs_status vs_batch_join(vs_batch* b);
void*     vs_batch_stream(vs_batch* b);
vs_status vs_batch_launch_count(vs_batch* b, uint64_t* n);
vs_status vs_batch_set_timing(vs_batch* b, int enable);
vs_status vs_batch_stage_time(vs_batch* b, int stage, double* total_ms, long long* count);
/* cv::resize + cvtColor + both pyrDown levels (Stabilizer.cpp:449-450, the pyramid of :611) for every stream of the
 * batch, nothing else: the analysis-image build timed alone (pyramid roofline in bench.py). */
vs_status vs_batch_build_pyramids(vs_batch* b, const uint8_t* const* d_frames, int width, int height, size_t stride);
/* The two halves separately: parts & 1 = level 0 (resize + gray, the HBM-bound kernel), parts & 2 = both pyrDown levels. */
vs_status vs_batch_build_levels(vs_batch* b, const uint8_t* const* d_frames, int width, int height, size_t stride, int parts);
vs_status vs_batch_stream_counts(vs_batch* b, int stream, int* n_frame_records, int* n_output_records);
vs_status vs_batch_frame_record(vs_batch* b, int stream, int i, vs_frame_record* rec);
vs_status vs_batch_output_record(vs_batch* b, int stream, int i, vs_output_record* rec);

/* ---- single-kernel entry points (device pointers; stream = cudaStream_t as void*, NULL = default).
 * One per OpenCV call the reference makes on the path; used by the per-kernel parity tests and
 * by the bench's roofline measurement. ------------------------------------------------------ */
/* cv::warpAffine(src, dst, T, dsize, INTER_LINEAR, BORDER_CONSTANT) — Stabilizer.cpp:1056-1060.
 * n_frames frames in one launch: frame i at d_src + i*src_frame_bytes, matrix T + 6*i. */
vs_status vs_k_warp_affine_bgr8(const uint8_t* d_src, int src_w, int src_h, size_t src_stride, size_t src_frame_bytes,
                                uint8_t* d_dst, int dst_w, int dst_h, size_t dst_stride, size_t dst_frame_bytes,
                                const float* T_host, int n_frames, void* stream);
/* cv::resize(INTER_LINEAR) + cv::cvtColor(BGR2GRAY) + the 3-level cv::pyrDown pyramid PyrLK builds —
 * Stabilizer.cpp:449-450, 611.  Writes three tightly packed gray levels (aw x ah, then halves). */
vs_status vs_k_gray_pyramid(const uint8_t* d_bgr, int w, int h, size_t stride, int aw, int ah,
                            uint8_t* d_l0, uint8_t* d_l1, uint8_t* d_l2, void* stream);
/* cv::resize(INTER_LINEAR) on 8UC1 / 8UC3 — Stabilizer.cpp:304,602,1121 */
vs_status vs_k_resize_linear_u8(const uint8_t* d_src, int sw, int sh, size_t sstride, int channels,
                                uint8_t* d_dst, int dw, int dh, size_t dstride, void* stream);
/* cv::goodFeaturesToTrack(gray, maxCorners, q, minDist, noArray, 3) — Stabilizer.cpp:355-357,740-744.
 * d_gray tightly packed w x h.  Writes up to capacity (x,y) pairs to host xy_out. */
vs_status vs_k_good_features(const uint8_t* d_gray, int w, int h, int max_corners, double quality, double min_dist,
                             float* xy_out_host, int capacity, int* n_out, void* stream);
/* the same with cv::goodFeaturesToTrack's blockSize argument (1..23; the first-frame detection passes params.blockSize) */
vs_status vs_k_good_features_block(const uint8_t* d_gray, int w, int h, int max_corners, double quality, double min_dist,
                                   int block_size, float* xy_out_host, int capacity, int* n_out, void* stream);
/* cv::calcOpticalFlowPyrLK(prev, next, pts, 15x15, maxLevel 2, COUNT+EPS 20/0.03) — Stabilizer.cpp:611-619.
 * d_prev/d_next tightly packed w x h gray. */
vs_status vs_k_pyr_lk(const uint8_t* d_prev, const uint8_t* d_next, int w, int h, const float* pts_xy_host, int n,
                      float* next_xy_host, uint8_t* status_host, void* stream);
/* cv::estimateAffinePartial2D(from, to, noArray, RANSAC, 5.0, 500) — Stabilizer.cpp:647-649.
 * affine_out: 6 doubles; returns *ok = 0 when the reference would get an empty Mat. */
vs_status vs_k_estimate_affine_partial(const float* from_xy_host, const float* to_xy_host, int n,
                                       double* affine_out, uint8_t* inlier_mask_host, int* iters_out, int* ok,
                                       void* stream);
/* cv::copyMakeBorder + warpAffine (Stabilizer.cpp:982-987,1056-1060) or warpAffine + crop + resize
 * (Stabilizer.cpp:1108-1124), fused.  mode: 0 plain, 1 border (border_mode = cv border code), 2 crop+zoom. */
vs_status vs_k_warp_output(const uint8_t* d_src, int w, int h, size_t stride, const float* T_host,
                           int mode, int border_size, int border_mode,
                           uint8_t* d_dst, size_t dst_stride, int* out_w, int* out_h, void* stream);

const char* vs_last_error(void);
const char* vs_version(void);
int         vs_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* VSTAB_B200_H */
