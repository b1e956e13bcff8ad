// capi.cu — the extern "C" boundary declared in include/vstab_b200.h.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "engine.h"
#include "roll.h"
#include "autozoom_host.h"

struct vs_stabilizer { Engine* eng; };
struct vs_batch { Engine* eng; };

#define API_BEGIN try {
#define API_END } catch (const std::bad_alloc&) { return vs_set_error(VS_ERR_OUT_OF_MEMORY, "host allocation failed"); } \
                  catch (...) { return vs_set_error(VS_ERR_CUDA, "unexpected C++ exception"); }

extern "C" {

const char* vs_version(void) { return "vstab_b200 0.1 (sm_100a)"; }
int vs_abi_version(void) { return VSTAB_B200_ABI_VERSION; }

// ------------------------------------------------------------------------------------ parameters
vs_status vs_params_default(vs_params* p) {
    if (!p) return vs_set_error(VS_ERR_INVALID_ARG, "null params");
    memset(p, 0, sizeof(*p));
    // Stabilizer.h:78-174
    p->use_cuda = 0; p->logging = 0;
    p->smoothing_radius = 30; p->max_corners = 200; p->quality_level = 0.01; p->min_distance = 30.0; p->block_size = 3;
    strcpy(p->border_type, "black"); p->border_size = 0; p->crop_n_zoom = 0;
    strcpy(p->smoothing_method, "box"); p->gaussian_sigma = 2.0; p->motion_prediction = 1; p->horizon_lock = 0;
    p->feature_detector = 0; p->orb_features = 500; p->fast_threshold = 10;
    p->use_roi = 0;
    p->adaptive_smoothing = 0; p->min_smoothing_radius = 5; p->max_smoothing_radius = 50;
    p->outlier_threshold = 3.0; p->intentional_motion_threshold = 20.0;
    p->stage_one_radius = 10; p->stage_two_radius = 25; p->use_temporal_filtering = 0; p->temporal_window_size = 5;
    p->fade_alpha = 0.1f; p->fade_duration = 30;
    p->motion_threshold_low = 5.0f; p->motion_threshold_high = 20.0f; p->border_scale_factor = 2.0f;
    p->roll_compensation = 1; p->roll_compensation_factor = 0.75;
    p->deep_stabilization = 0; p->model_path[0] = 0;
    p->jitter_frequency = 3; p->separate_translation_rotation = 1; p->use_imu_data = 0;
    p->enable_virtual_canvas = 0; p->canvas_scale_factor = 1.5f; p->temporal_buffer_size = 30; p->canvas_blend_weight = 0.7f;
    p->adaptive_canvas_size = 1; p->max_canvas_scale = 2.0f; p->min_canvas_scale = 1.2f; p->preserve_edge_quality = 1;
    p->edge_blend_radius = 20;
    p->drone_high_freq_mode = 0; p->hf_shake_px = 1.5f; p->hf_analysis_max_width = 960; p->hf_rot_lp_alpha = 0.2f;
    p->enable_conditional_clahe = 1; p->hf_dead_zone_threshold = 2.0f; p->hf_freeze_duration = 10;
    p->hf_motion_accumulator_decay = 0.9f;
    return VS_OK;
}

static std::string trim(const std::string& s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
static bool to_bool(const std::string& v) {
    return v == "true" || v == "True" || v == "TRUE" || v == "1" || v == "yes" || v == "on";
}
static void set_str(char* dst, size_t cap, const std::string& v) {
    std::string s = v;
    if (s.size() >= 2 && (s.front() == '"' || s.front() == '\'') && s.back() == s.front()) s = s.substr(1, s.size() - 2);
    snprintf(dst, cap, "%s", s.c_str());
}

// The `stabilizer:` keys read by the reference's canonical reader, examples/vsg.cpp:1003-1114.
static void apply_key(vs_params* p, const std::string& k, const std::string& v) {
#define KI(name, field) if (k == name) { p->field = (int32_t)strtol(v.c_str(), nullptr, 10); return; }
#define KB(name, field) if (k == name) { p->field = to_bool(v) ? 1 : 0; return; }
#define KD(name, field) if (k == name) { p->field = strtod(v.c_str(), nullptr); return; }
#define KF(name, field) if (k == name) { p->field = (float)strtod(v.c_str(), nullptr); return; }
    KI("smoothing_radius", smoothing_radius)
    if (k == "border_type") { set_str(p->border_type, sizeof(p->border_type), v); return; }
    KI("border_size", border_size) KB("crop_n_zoom", crop_n_zoom) KB("logging", logging) KB("use_cuda", use_cuda)
    if (k == "smoothing_method") { set_str(p->smoothing_method, sizeof(p->smoothing_method), v); return; }
    KD("gaussian_sigma", gaussian_sigma)
    KI("stage_one_radius", stage_one_radius) KI("stage_two_radius", stage_two_radius)
    KB("use_temporal_filtering", use_temporal_filtering) KI("temporal_window_size", temporal_window_size)
    KB("adaptive_smoothing", adaptive_smoothing) KI("min_smoothing_radius", min_smoothing_radius)
    KI("max_smoothing_radius", max_smoothing_radius)
    KI("max_corners", max_corners) KD("quality_level", quality_level) KD("min_distance", min_distance) KI("block_size", block_size)
    KD("outlier_threshold", outlier_threshold) KB("motion_prediction", motion_prediction)
    KD("intentional_motion_threshold", intentional_motion_threshold)
    KI("jitter_frequency", jitter_frequency) KB("separate_translation_rotation", separate_translation_rotation)
    KB("deep_stabilization", deep_stabilization)
    if (k == "model_path") { set_str(p->model_path, sizeof(p->model_path), v); return; }
    KB("roll_compensation", roll_compensation) KD("roll_compensation_factor", roll_compensation_factor)
    KB("use_roi", use_roi) KI("roi_x", roi_x) KI("roi_y", roi_y) KI("roi_width", roi_width) KI("roi_height", roi_height)
    KB("horizon_lock", horizon_lock)
    KI("feature_detector_type", feature_detector) KI("fast_threshold", fast_threshold) KI("orb_features", orb_features)
    KF("border_scale_factor", border_scale_factor) KF("motion_threshold_low", motion_threshold_low)
    KF("motion_threshold_high", motion_threshold_high)
    KI("fadeDuration", fade_duration) KF("fadeAlpha", fade_alpha)
    KB("use_imu_data", use_imu_data)
    KB("enable_virtual_canvas", enable_virtual_canvas) KF("canvas_scale_factor", canvas_scale_factor)
    KI("temporal_buffer_size", temporal_buffer_size) KF("canvas_blend_weight", canvas_blend_weight)
    KB("adaptive_canvas_size", adaptive_canvas_size) KF("max_canvas_scale", max_canvas_scale)
    KF("min_canvas_scale", min_canvas_scale) KB("preserve_edge_quality", preserve_edge_quality)
    KI("edge_blend_radius", edge_blend_radius)
    KB("drone_high_freq_mode", drone_high_freq_mode) KF("hf_shake_px", hf_shake_px)
    KI("hf_analysis_max_width", hf_analysis_max_width) KF("hf_rot_lp_alpha", hf_rot_lp_alpha)
    KB("enable_conditional_clahe", enable_conditional_clahe)
    KF("hf_dead_zone_threshold", hf_dead_zone_threshold) KI("hf_freeze_duration", hf_freeze_duration)
    KF("hf_motion_accumulator_decay", hf_motion_accumulator_decay)
    // shake_level_threshold / walking_detection_threshold / vehicle_detection_threshold / outlier_rejection:
    // read into locals and dropped by the reference (vsg.cpp:1109-1113) -> ignored
#undef KI
#undef KB
#undef KD
#undef KF
}

vs_status vs_params_from_yaml_string(const char* text, vs_params* p) {
    if (!text || !p) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    API_BEGIN
    std::istringstream in(text);
    std::string line;
    bool in_section = false;
    // roi_* are only consumed when use_roi is set (vsg.cpp:1054); harmless either way (inert fields)
    while (std::getline(in, line)) {
        size_t hash = std::string::npos;
        bool q = false;
        for (size_t i = 0; i < line.size(); ++i) {          // strip comments outside quotes
            if (line[i] == '"') q = !q;
            if (line[i] == '#' && !q) { hash = i; break; }
        }
        if (hash != std::string::npos) line = line.substr(0, hash);
        if (trim(line).empty() || line[0] == '%' || trim(line) == "---") continue;
        bool indented = line[0] == ' ' || line[0] == '\t';
        size_t colon = line.find(':');
        if (colon == std::string::npos) continue;
        std::string key = trim(line.substr(0, colon)), val = trim(line.substr(colon + 1));
        if (!indented) { in_section = (key == "stabilizer"); continue; }
        if (in_section && !val.empty()) apply_key(p, key, val);
    }
    return VS_OK;
    API_END
}

vs_status vs_params_from_yaml(const char* path, vs_params* p) {
    if (!path || !p) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    API_BEGIN
    std::ifstream f(path);
    if (!f) return vs_set_error(VS_ERR_IO, "cannot open config file");
    std::stringstream ss;
    ss << f.rdbuf();
    return vs_params_from_yaml_string(ss.str().c_str(), p);
    API_END
}

// ------------------------------------------------------------------------------------ stabilizer
vs_status vs_stabilizer_create(const vs_params* params, int device, vs_stabilizer** out) {
    if (!params || !out) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    API_BEGIN
    Engine* e = nullptr;
    vs_status st = Engine::create(*params, device, 1, &e);
    if (st != VS_OK) { *out = nullptr; return st; }
    *out = new vs_stabilizer{e};
    return VS_OK;
    API_END
}
void vs_stabilizer_destroy(vs_stabilizer* s) {
    if (!s) return;
    delete s->eng;
    delete s;
}
vs_status vs_stabilizer_push(vs_stabilizer* s, const uint8_t* bgr, int width, int height, size_t stride, uint8_t* out,
                             size_t out_stride, size_t out_capacity, int* out_width, int* out_height, int* produced) {
    if (!s || !produced || !out_width || !out_height) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    API_BEGIN
    const uint8_t* f[1] = {bgr};
    uint8_t* o[1] = {out};
    vs_status rc = s->eng->push(bgr ? f : nullptr, width, height, stride, o, out_stride, out_capacity, 0, true, out_width, out_height, produced);
    if ((rc == VS_ERR_CUDA || rc == VS_ERR_OUT_OF_MEMORY) && bgr && out && width > 0 && height > 0) {
        // The reference never throws on this path: every cv::Exception degrades to the un-warped frame (Stabilizer.cpp:609-626,
        // :1049-1066).  There is no CPU fallback here, so a device failure is REPORTED (status + vs_last_error()) and the
        // input frame is handed back unchanged when the caller's buffer can take it.
        const size_t tight = (size_t)width * 3, is = stride ? stride : tight, os = out_stride ? out_stride : tight;
        if (is >= tight && os >= tight && os * (size_t)(height - 1) + tight <= out_capacity) {
            for (int y = 0; y < height; ++y) memcpy(out + (size_t)y * os, bgr + (size_t)y * is, tight);
            *out_width = width; *out_height = height; *produced = 1;
        }
    }
    return rc;
    API_END
}
vs_status vs_stabilizer_flush(vs_stabilizer* s, uint8_t* out, size_t out_stride, size_t out_capacity, int* out_width,
                              int* out_height, int* produced) {
    if (!s || !produced || !out_width || !out_height) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    API_BEGIN
    uint8_t* o[1] = {out};
    return s->eng->flush(o, out_stride, out_capacity, true, out_width, out_height, produced);
    API_END
}
vs_status vs_stabilizer_push_many(vs_stabilizer* s, const uint8_t* bgr, size_t frame_step, int n_frames, int width, int height,
                                  size_t stride, uint8_t* out, size_t out_stride, size_t out_frame_capacity,
                                  int* out_width, int* out_height, int* n_produced) {
    if (!s || !n_produced || !out_width || !out_height || !bgr || !out || n_frames < 0)
        return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    API_BEGIN
    return s->eng->push_many(bgr, frame_step, n_frames, width, height, stride, out, out_stride, out_frame_capacity,
                             out_width, out_height, n_produced);
    API_END
}
vs_status vs_stabilizer_push_many_device(vs_stabilizer* s, const uint8_t* d_bgr, size_t frame_step, int n_frames, int width, int height,
                                         size_t stride, uint8_t* d_out, size_t out_stride, size_t out_frame_capacity, unsigned flags,
                                         int* out_width, int* out_height, int* n_produced) {
    if (!s || !n_produced || !out_width || !out_height || !d_bgr || !d_out || n_frames < 0)
        return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    API_BEGIN
    return s->eng->push_many_device(d_bgr, frame_step, n_frames, width, height, stride, d_out, out_stride, out_frame_capacity, flags,
                                    out_width, out_height, n_produced);
    API_END
}
vs_status vs_stabilizer_flush_many(vs_stabilizer* s, uint8_t* out, size_t out_stride, size_t out_frame_capacity, int max_frames,
                                   int* out_width, int* out_height, int* n_produced) {
    if (!s || !n_produced || !out_width || !out_height || !out) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    API_BEGIN
    return s->eng->flush_many(out, out_stride, out_frame_capacity, max_frames, out_width, out_height, n_produced);
    API_END
}
vs_status vs_stabilizer_clean(vs_stabilizer* s) {
    if (!s) return vs_set_error(VS_ERR_INVALID_ARG, "null handle");
    return s->eng->clean();
}
vs_status vs_stabilizer_push_device(vs_stabilizer* s, const uint8_t* d_bgr, int width, int height, size_t stride,
                                    uint8_t* d_out, size_t out_stride, size_t out_capacity, unsigned flags,
                                    int* out_width, int* out_height, int* produced) {
    if (!s || !produced || !out_width || !out_height) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    API_BEGIN
    const uint8_t* f[1] = {d_bgr};
    uint8_t* o[1] = {d_out};
    return s->eng->push(d_bgr ? f : nullptr, width, height, stride, o, out_stride, out_capacity, flags, false, out_width, out_height, produced);
    API_END
}
vs_status vs_stabilizer_flush_device(vs_stabilizer* s, uint8_t* d_out, size_t out_stride, size_t out_capacity,
                                     int* out_width, int* out_height, int* produced) {
    if (!s || !produced || !out_width || !out_height) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    API_BEGIN
    uint8_t* o[1] = {d_out};
    return s->eng->flush(o, out_stride, out_capacity, false, out_width, out_height, produced);
    API_END
}
vs_status vs_stabilizer_sync(vs_stabilizer* s) { return s ? s->eng->sync() : vs_set_error(VS_ERR_INVALID_ARG, "null handle"); }
vs_status vs_stabilizer_wait_event(vs_stabilizer* s, void* cuda_event) {
    if (!s || !cuda_event) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    return s->eng->wait_external((cudaEvent_t)cuda_event);
}
vs_status vs_stabilizer_join(vs_stabilizer* s) { return s ? s->eng->join() : vs_set_error(VS_ERR_INVALID_ARG, "null handle"); }
void* vs_stabilizer_stream(vs_stabilizer* s) { return s ? (void*)s->eng->stream() : nullptr; }
vs_status vs_stabilizer_counts(vs_stabilizer* s, int* nf, int* no) {
    if (!s) return vs_set_error(VS_ERR_INVALID_ARG, "null handle");
    if (nf) *nf = s->eng->n_frame_records();
    if (no) *no = s->eng->n_output_records();
    return VS_OK;
}
vs_status vs_stabilizer_frame_record(vs_stabilizer* s, int i, vs_frame_record* r) {
    return s ? s->eng->frame_record(0, i, r) : vs_set_error(VS_ERR_INVALID_ARG, "null handle");
}
vs_status vs_stabilizer_output_record(vs_stabilizer* s, int i, vs_output_record* r) {
    return s ? s->eng->output_record(0, i, r) : vs_set_error(VS_ERR_INVALID_ARG, "null handle");
}
vs_status vs_stabilizer_frame_points(vs_stabilizer* s, int i, float* prev_xy, float* next_xy, uint8_t* status,
                                     uint8_t* inlier_mask, float* detected_xy) {
    return s ? s->eng->frame_points(0, i, prev_xy, next_xy, status, inlier_mask, detected_xy)
             : vs_set_error(VS_ERR_INVALID_ARG, "null handle");
}
vs_status vs_stabilizer_first_corners(vs_stabilizer* s, float* xy, int capacity, int* n) {
    return s ? s->eng->first_corners(0, xy, capacity, n) : vs_set_error(VS_ERR_INVALID_ARG, "null handle");
}
vs_status vs_stabilizer_launch_count(vs_stabilizer* s, uint64_t* n) {
    if (!s || !n) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    *n = s->eng->launches();
    return VS_OK;
}

vs_status vs_stabilizer_set_timing(vs_stabilizer* s, int enable) {
    if (!s) return vs_set_error(VS_ERR_INVALID_ARG, "null handle");
    s->eng->set_timing(enable != 0);
    return VS_OK;
}
vs_status vs_stabilizer_stage_time(vs_stabilizer* s, int stage, double* total_ms, long long* count) {
    if (!s || !total_ms || !count) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    s->eng->stage_time(stage, total_ms, count);
    return VS_OK;
}
vs_status vs_stabilizer_trace(vs_stabilizer* s, float* out, int capacity, int* n) {
    if (!s || !n) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    const std::vector<float>& t = s->eng->trace();
    const int m = (int)(t.size() / 3);
    *n = m;
    if (out) memcpy(out, t.data(), sizeof(float) * 3 * (size_t)(m < capacity ? m : capacity));
    return VS_OK;
}
vs_status vs_batch_set_timing(vs_batch* b, int enable) {
    if (!b) return vs_set_error(VS_ERR_INVALID_ARG, "null handle");
    b->eng->set_timing(enable != 0);
    return VS_OK;
}
vs_status vs_batch_stage_time(vs_batch* b, int stage, double* total_ms, long long* count) {
    if (!b || !total_ms || !count) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    b->eng->stage_time(stage, total_ms, count);
    return VS_OK;
}

// ------------------------------------------------------------------------------------ batch
vs_status vs_batch_create(const vs_params* params, int device, int n_streams, vs_batch** out) {
    if (!params || !out) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    API_BEGIN
    Engine* e = nullptr;
    vs_status st = Engine::create(*params, device, n_streams, &e);
    if (st != VS_OK) { *out = nullptr; return st; }
    *out = new vs_batch{e};
    return VS_OK;
    API_END
}
void vs_batch_destroy(vs_batch* b) {
    if (!b) return;
    delete b->eng;
    delete b;
}
vs_status vs_batch_push_device(vs_batch* b, const uint8_t* const* d_frames, int width, int height, size_t stride,
                               uint8_t* const* d_outs, size_t out_stride, size_t out_capacity, unsigned flags,
                               int* out_width, int* out_height, int* produced) {
    if (!b || !produced || !out_width || !out_height) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    API_BEGIN
    return b->eng->push(d_frames, width, height, stride, d_outs, out_stride, out_capacity, flags, false, out_width, out_height, produced);
    API_END
}
vs_status vs_batch_flush_device(vs_batch* b, uint8_t* const* d_outs, size_t out_stride, size_t out_capacity,
                                int* out_width, int* out_height, int* produced) {
    if (!b || !produced || !out_width || !out_height) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    API_BEGIN
    return b->eng->flush(d_outs, out_stride, out_capacity, false, out_width, out_height, produced);
    API_END
}
vs_status vs_batch_clip_analyze_device(vs_batch* b, const uint8_t* const* d_frames, int width, int height, int count,
                                       float* const* d_transforms_out) {
    if (!b || !d_frames || !d_transforms_out) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    API_BEGIN
    return b->eng->analyze_chunks_lockstep(d_frames, width, height, count, d_transforms_out);
    API_END
}
vs_status vs_batch_wait_event(vs_batch* b, void* cuda_event) {
    if (!b || !cuda_event) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    return b->eng->wait_external((cudaEvent_t)cuda_event);
}
vs_status vs_batch_sync(vs_batch* b) { return b ? b->eng->sync() : vs_set_error(VS_ERR_INVALID_ARG, "null handle"); }
vs_status vs_batch_join(vs_batch* b) { return b ? b->eng->join() : vs_set_error(VS_ERR_INVALID_ARG, "null handle"); }
void* vs_batch_stream(vs_batch* b) { return b ? (void*)b->eng->stream() : nullptr; }
vs_status vs_batch_launch_count(vs_batch* b, uint64_t* n) {
    if (!b || !n) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    *n = b->eng->launches();
    return VS_OK;
}
// analysis-image build alone (cv::resize + cvtColor + both pyrDown levels) for every stream of the batch, into pyramid
// slot 0, on the batch's public stream: the launch bench.py times for the pyramid roofline
vs_status vs_batch_build_levels(vs_batch* b, const uint8_t* const* d_frames, int width, int height, size_t stride, int parts) {
    if (!b || !d_frames) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    API_BEGIN
    Engine* e = b->eng;
    PtrPack src;
    for (int l = 0; l < e->n_lanes(); ++l) src.p[l] = d_frames[l];
    if (parts & 1) launch_gray_resize(e->d_lanes(), e->n_lanes(), src, width, height, stride ? stride : (size_t)width * 3, 0, e->stream());
    if (parts & 2) launch_pyrdown(e->d_lanes(), e->n_lanes(), 0, e->stream());
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) return vs_set_cuda_error(ce, "vs_batch_build_levels", __FILE__, __LINE__);
    return VS_OK;
    API_END
}
vs_status vs_batch_build_pyramids(vs_batch* b, const uint8_t* const* d_frames, int width, int height, size_t stride) {
    return vs_batch_build_levels(b, d_frames, width, height, stride, 3);
}
vs_status vs_batch_stream_counts(vs_batch* b, int stream, int* nf, int* no) {
    if (!b || stream < 0 || stream >= b->eng->n_lanes()) return vs_set_error(VS_ERR_INVALID_ARG, "bad stream");
    if (nf) *nf = b->eng->n_frame_records();
    if (no) *no = b->eng->n_output_records();
    return VS_OK;
}
vs_status vs_batch_frame_record(vs_batch* b, int stream, int i, vs_frame_record* rec) {
    return b ? b->eng->frame_record(stream, i, rec) : vs_set_error(VS_ERR_INVALID_ARG, "null handle");
}
vs_status vs_batch_output_record(vs_batch* b, int stream, int i, vs_output_record* rec) {
    return b ? b->eng->output_record(stream, i, rec) : vs_set_error(VS_ERR_INVALID_ARG, "null handle");
}

// ------------------------------------------------------------------------------------ offline clip mode
int vs_clip_halo(int first) { return Engine::chunk_halo(first); }
vs_status vs_clip_analyze(vs_stabilizer* s, const uint8_t* d_frames, int width, int height, int first, int count,
                          float* transforms_out_host, int* n_out) {
    if (!s) return vs_set_error(VS_ERR_INVALID_ARG, "null handle");
    API_BEGIN
    return s->eng->analyze_chunk(d_frames, width, height, first, count, transforms_out_host, false, n_out);
    API_END
}
vs_status vs_clip_analyze_device(vs_stabilizer* s, const uint8_t* d_frames, int width, int height, int first, int count,
                                 float* d_transforms_out, int* n_out) {
    if (!s) return vs_set_error(VS_ERR_INVALID_ARG, "null handle");
    API_BEGIN
    return s->eng->analyze_chunk(d_frames, width, height, first, count, d_transforms_out, true, n_out);
    API_END
}
vs_status vs_clip_render(vs_stabilizer* s, const float* all_transforms_host, int n_total, const uint8_t* d_frames,
                         int width, int height, int first, int count, uint8_t* d_out, int* out_width, int* out_height) {
    if (!s || !out_width || !out_height) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    API_BEGIN
    return s->eng->render_chunk(all_transforms_host, false, n_total, d_frames, width, height, first, count, d_out, out_width, out_height);
    API_END
}
vs_status vs_clip_set_transforms_device(vs_stabilizer* s, const float* d_all_transforms, int n_total, int width, int height) {
    if (!s) return vs_set_error(VS_ERR_INVALID_ARG, "null handle");
    API_BEGIN
    return s->eng->set_clip_transforms(d_all_transforms, true, n_total, width, height);
    API_END
}
vs_status vs_clip_render_prepared_device(vs_stabilizer* s, const uint8_t* d_frames, int width, int height, int first, int count,
                                         uint8_t* d_out, int* out_width, int* out_height) {
    if (!s || !out_width || !out_height) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    API_BEGIN
    return s->eng->render_prepared(d_frames, width, height, first, count, d_out, out_width, out_height, false);
    API_END
}
vs_status vs_clip_render_device(vs_stabilizer* s, const float* d_all_transforms, int n_total, const uint8_t* d_frames,
                                int width, int height, int first, int count, uint8_t* d_out, int* out_width, int* out_height) {
    if (!s || !out_width || !out_height) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    API_BEGIN
    return s->eng->render_chunk(d_all_transforms, true, n_total, d_frames, width, height, first, count, d_out, out_width, out_height);
    API_END
}

// ------------------------------------------------------------------------------------ roll correction
struct vs_roll { RollCorrector* rc; };

vs_status vs_roll_params_default(vs_roll_params* p) {
    if (!p) return vs_set_error(VS_ERR_INVALID_ARG, "null params");
    // RollCorrection.h:16-38
    p->scale_factor = 0.25; p->canny_threshold_low = 50.0; p->canny_threshold_high = 150.0; p->canny_aperture = 3;
    p->hough_rho = 1.0f; p->hough_theta = (float)(3.1415926535897932384626433832795 / 180.0f); p->hough_threshold = 100;
    p->angle_filter_min = -10.0; p->angle_filter_max = 10.0; p->angle_smoothing_alpha = 0.1; p->angle_decay = 0.995;
    p->max_angle_change_deg = 0.5;
    return VS_OK;
}
vs_status vs_roll_params_from_yaml_string(const char* text, vs_roll_params* p) {
    if (!text || !p) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    API_BEGIN
    std::istringstream in(text);
    std::string line;
    bool in_section = false;
    while (std::getline(in, line)) {                         // the `roll_correction:` section, keys of examples/vsg.cpp:988-1000
        size_t hash = line.find('#');
        if (hash != std::string::npos) line = line.substr(0, hash);
        if (trim(line).empty() || line[0] == '%' || trim(line) == "---") continue;
        const bool indented = line[0] == ' ' || line[0] == '\t';
        size_t colon = line.find(':');
        if (colon == std::string::npos) continue;
        std::string key = trim(line.substr(0, colon)), val = trim(line.substr(colon + 1));
        if (!indented) { in_section = (key == "roll_correction"); continue; }
        if (!in_section || val.empty()) continue;
        const double v = atof(val.c_str());
        if (key == "scale_factor") p->scale_factor = v;
        else if (key == "canny_threshold_low") p->canny_threshold_low = v;
        else if (key == "canny_threshold_high") p->canny_threshold_high = v;
        else if (key == "canny_aperture") p->canny_aperture = (int)v;
        else if (key == "hough_rho") p->hough_rho = (float)v;
        else if (key == "hough_theta") p->hough_theta = (float)v;
        else if (key == "hough_threshold") p->hough_threshold = (int)v;
        else if (key == "angle_smoothing_alpha") p->angle_smoothing_alpha = v;
        else if (key == "angle_decay") p->angle_decay = v;
        else if (key == "angle_filter_min") p->angle_filter_min = v;
        else if (key == "angle_filter_max") p->angle_filter_max = v;
        else if (key == "max_angle_change_deg") p->max_angle_change_deg = v;   // not read by vsg.cpp; accepted
    }
    return VS_OK;
    API_END
}
vs_status vs_roll_params_from_yaml(const char* path, vs_roll_params* p) {
    if (!path || !p) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    API_BEGIN
    std::ifstream f(path);
    if (!f) return vs_set_error(VS_ERR_IO, "cannot open config file");
    std::stringstream ss;
    ss << f.rdbuf();
    return vs_roll_params_from_yaml_string(ss.str().c_str(), p);
    API_END
}
// ---- virtual canvas stage on its own (canvas.h)
struct vs_canvas { VirtualCanvas vc; int device; };
static vs_status canvas_params_ok(const vs_params& p) {
    if (p.temporal_buffer_size < 0 || p.temporal_buffer_size > 240)
        return vs_set_error(VS_ERR_UNSUPPORTED, "temporal_buffer_size must be 0..240 (frames kept on the device)");
    const float lo = p.adaptive_canvas_size ? (p.min_canvas_scale < p.canvas_scale_factor ? p.min_canvas_scale : p.canvas_scale_factor) : p.canvas_scale_factor;
    if (!(lo >= 1.0f) || !(p.max_canvas_scale <= 8.0f) || !(p.canvas_scale_factor <= 8.0f))
        return vs_set_error(VS_ERR_UNSUPPORTED, "virtual canvas scales must lie in 1..8 (a canvas smaller than the frame is not built)");
    return VS_OK;
}
vs_status vs_canvas_create(const vs_params* params, int device, vs_canvas** out) {
    if (!params || !out) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    API_BEGIN
    *out = nullptr;
    vs_status st = canvas_params_ok(*params);
    if (st != VS_OK) return st;
    vs_canvas* c = new vs_canvas;
    c->device = device;
    c->vc.configure(*params);
    *out = c;
    return VS_OK;
    API_END
}
void vs_canvas_destroy(vs_canvas* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    delete c;
}
vs_status vs_canvas_apply_device(vs_canvas* c, const uint8_t* d_bgr, int width, int height, size_t stride, const float* transform3,
                                 const float* recent_transforms, int n_recent, uint8_t* d_out, size_t out_stride, void* stream) {
    if (!c || !d_bgr || !d_out || !transform3 || width < 4 || height < 4 || (n_recent > 0 && !recent_transforms))
        return vs_set_error(VS_ERR_INVALID_ARG, "virtual canvas: bad argument");
    API_BEGIN
    CUDA_TRY(cudaSetDevice(c->device));
    if (stride == 0) stride = (size_t)width * 3;
    if (out_stride == 0) out_stride = (size_t)width * 3;
    if (stride < (size_t)width * 3 || out_stride < (size_t)width * 3) return vs_set_error(VS_ERR_INVALID_ARG, "stride smaller than a row");
    if (n_recent > 30) { recent_transforms += 3 * (size_t)(n_recent - 30); n_recent = 30; }
    int launches = 0;
    return c->vc.apply(d_bgr, width, height, stride, transform3, recent_transforms, n_recent < 0 ? 0 : n_recent, d_out, out_stride,
                       (cudaStream_t)stream, &launches);
    API_END
}
vs_status vs_canvas_info(vs_canvas* c, float* scale, int* regions_filled) {
    if (!c) return vs_set_error(VS_ERR_INVALID_ARG, "null handle");
    if (scale) *scale = c->vc.scale();
    if (regions_filled) *regions_filled = c->vc.regions_last();
    return VS_OK;
}

vs_status vs_roll_create(const vs_roll_params* params, int device, vs_roll** out) {
    if (!params || !out) return vs_set_error(VS_ERR_INVALID_ARG, "null argument");
    API_BEGIN
    RollCorrector* rc = nullptr;
    vs_status st = RollCorrector::create(*params, device, &rc);
    if (st != VS_OK) { *out = nullptr; return st; }
    *out = new vs_roll{rc};
    return VS_OK;
    API_END
}
void vs_roll_destroy(vs_roll* r) {
    if (!r) return;
    delete r->rc;
    delete r;
}
vs_status vs_roll_correct(vs_roll* r, const uint8_t* bgr, int width, int height, size_t stride, uint8_t* out, size_t out_stride) {
    if (!r) return vs_set_error(VS_ERR_INVALID_ARG, "null handle");
    API_BEGIN
    return r->rc->correct_host(bgr, width, height, stride, out, out_stride);
    API_END
}
vs_status vs_roll_correct_device(vs_roll* r, const uint8_t* d_bgr, int width, int height, size_t stride, uint8_t* d_out,
                                 size_t out_stride, void* stream) {
    if (!r) return vs_set_error(VS_ERR_INVALID_ARG, "null handle");
    API_BEGIN
    return r->rc->correct_device(d_bgr, width, height, stride, d_out, out_stride, (cudaStream_t)stream);
    API_END
}
vs_status vs_roll_reset(vs_roll* r) { return r ? r->rc->reset() : vs_set_error(VS_ERR_INVALID_ARG, "null handle"); }
vs_status vs_roll_state(vs_roll* r, double* smoothed_angle_deg, int* n_lines, int* n_edges, uint64_t* launches) {
    if (!r) return vs_set_error(VS_ERR_INVALID_ARG, "null handle");
    RollState s{};
    vs_status st = r->rc->state(&s, n_edges);
    if (st != VS_OK) return st;
    if (smoothed_angle_deg) *smoothed_angle_deg = s.smoothed_angle;
    if (n_lines) *n_lines = s.n_lines;
    if (launches) *launches = r->rc->launches();
    return VS_OK;
}
vs_status vs_roll_debug(vs_roll* r, int* small_w, int* small_h, uint8_t* gray_out, uint8_t* edges_out, float* lines_out, int lines_capacity) {
    if (!r) return vs_set_error(VS_ERR_INVALID_ARG, "null handle");
    if (small_w) *small_w = r->rc->small_w();
    if (small_h) *small_h = r->rc->small_h();
    if (!gray_out && !edges_out && !lines_out) return VS_OK;
    return r->rc->debug(gray_out, edges_out, lines_out, lines_capacity);
}

// ------------------------------------------------------------------------------------ auto zoom-crop
vs_status vs_auto_zoom_crop_device(const uint8_t* d_bgr, int width, int height, size_t stride, double /*margin_percent*/,
                                   uint8_t* d_out, size_t out_stride, size_t out_capacity, int* out_width, int* out_height, void* stream) {
    API_BEGIN
    return auto_zoom_crop_device(d_bgr, width, height, stride, d_out, out_stride, out_capacity, out_width, out_height, (cudaStream_t)stream);
    API_END
}
vs_status vs_auto_zoom_crop(const uint8_t* bgr, int width, int height, size_t stride, double /*margin_percent*/, int device,
                            uint8_t* out, size_t out_stride, size_t out_capacity, int* out_width, int* out_height) {
    if (!bgr || !out || !out_width || !out_height || width < 4 || height < 4) return vs_set_error(VS_ERR_INVALID_ARG, "auto zoom-crop: bad argument");
    API_BEGIN
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return vs_set_error(VS_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
    }
    if (device < 0 || device >= count) return vs_set_error(VS_ERR_INVALID_ARG, "bad device ordinal");
    CUDA_TRY(cudaSetDevice(device));
    if (stride == 0) stride = (size_t)width * 3;
    const size_t tight = (size_t)width * 3;
    const size_t ocap = std::max(tight * height, (size_t)640 * 360 * 3);
    // device staging frames kept per host thread and device (a per-frame call must not pay for allocations)
    struct Stage { int dev = -1; size_t in_cap = 0, out_cap = 0; uint8_t *d_in = nullptr, *d_o = nullptr; };
    static thread_local Stage sg;
    if (sg.dev != device || tight * height > sg.in_cap || ocap > sg.out_cap) {
        if (sg.d_in) cudaFree(sg.d_in);
        if (sg.d_o) cudaFree(sg.d_o);
        sg = Stage();
        CUDA_TRY(cudaMalloc((void**)&sg.d_in, tight * height));
        CUDA_TRY(cudaMalloc((void**)&sg.d_o, ocap));
        sg.dev = device; sg.in_cap = tight * height; sg.out_cap = ocap;
    }
    uint8_t *d_in = sg.d_in, *d_o = sg.d_o;
    cudaError_t e = cudaSuccess;
    vs_status rc = VS_OK;
    if (e == cudaSuccess) e = cudaMemcpy2D(d_in, tight, bgr, stride, tight, height, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        rc = auto_zoom_crop_device(d_in, width, height, tight, d_o, 0, ocap, out_width, out_height, 0);
        if (rc == VS_OK) {
            const size_t ot = (size_t)*out_width * 3;
            if (out_stride == 0) out_stride = ot;
            if (out_stride < ot || out_stride * (size_t)(*out_height - 1) + ot > out_capacity)
                rc = vs_set_error(VS_ERR_BUFFER_TOO_SMALL, "auto zoom-crop: output buffer too small");
            else e = cudaMemcpy2D(out, out_stride, d_o, ot, ot, *out_height, cudaMemcpyDeviceToHost);
        }
    }
    CUDA_TRY(e);
    return rc;
    API_END
}
vs_status vs_auto_zoom_rect_from_mask(const uint8_t* mask, int width, int height, size_t stride, int* x, int* y, int* w, int* h, int* found) {
    if (!mask || !x || !y || !w || !h || !found || width < 1 || height < 1) return vs_set_error(VS_ERR_INVALID_ARG, "bad argument");
    API_BEGIN
    azc::Rect r;
    *found = azc::crop_rect_from_mask(mask, width, height, stride ? stride : (size_t)width, &r) ? 1 : 0;
    *x = r.x; *y = r.y; *w = r.width; *h = r.height;
    return VS_OK;
    API_END
}
vs_status vs_k_find_external_contours(const uint8_t* mask, int width, int height, size_t stride, int* points_xy, int points_capacity,
                                      int* lengths, int lengths_capacity, int* n_contours) {
    if (!mask || !points_xy || !lengths || !n_contours || width < 1 || height < 1) return vs_set_error(VS_ERR_INVALID_ARG, "bad argument");
    API_BEGIN
    std::vector<std::vector<azc::Pt>> cs;
    std::vector<signed char> work;
    azc::find_external_contours(mask, width, height, stride ? stride : (size_t)width, cs, work);
    *n_contours = (int)cs.size();
    if ((int)cs.size() > lengths_capacity) return vs_set_error(VS_ERR_BUFFER_TOO_SMALL, "too many contours");
    int k = 0;
    for (size_t i = 0; i < cs.size(); ++i) {
        lengths[i] = (int)cs[i].size();
        if (k + (int)cs[i].size() > points_capacity) return vs_set_error(VS_ERR_BUFFER_TOO_SMALL, "too many contour points");
        for (auto& p : cs[i]) { points_xy[2 * k] = p.x; points_xy[2 * k + 1] = p.y; ++k; }
    }
    return VS_OK;
    API_END
}
vs_status vs_k_content_mask(const uint8_t* d_bgr, int width, int height, size_t stride, uint8_t* d_mask, uint8_t* d_scratch, void* stream) {
    if (!d_bgr || !d_mask || !d_scratch || width < 1 || height < 1) return vs_set_error(VS_ERR_INVALID_ARG, "bad argument");
    launch_content_mask(d_bgr, width, height, stride ? stride : (size_t)width * 3, d_mask, d_scratch, (cudaStream_t)stream);
    CUDA_TRY(cudaGetLastError());
    return VS_OK;
}

// ------------------------------------------------------------------------------------ single kernels
static vs_status scratch_engine(Engine** e, int max_corners = 200, double min_dist = 1.0) {
    vs_params p;
    vs_params_default(&p);
    p.max_corners = (max_corners <= 0 || max_corners > 2048) ? 2048 : max_corners;   // sizes the key-point buffers only
    p.min_distance = min_dist;
    int dev = 0;
    cudaGetDevice(&dev);
    return Engine::create(p, dev, 1, e);
}

vs_status vs_k_warp_affine_bgr8(const uint8_t* d_src, int src_w, int src_h, size_t src_stride, size_t src_frame_bytes,
                                uint8_t* d_dst, int dst_w, int dst_h, size_t dst_stride, size_t dst_frame_bytes,
                                const float* T_host, int n_frames, void* stream) {
    if (!d_src || !d_dst || !T_host || n_frames < 1) return vs_set_error(VS_ERR_INVALID_ARG, "bad argument");
    API_BEGIN
    cudaStream_t st = (cudaStream_t)stream;
    // One set-up block per (host thread, device, stream): calls on the same stream are ordered by the stream itself (the copy
    // below queues behind the previous launch), calls on different streams or devices never share a block.
    struct WpBlock { int dev; cudaStream_t st; WarpParams* p; int cap; };
    static thread_local std::vector<WpBlock> blocks;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    WpBlock* blk = nullptr;
    for (auto& b : blocks) if (b.dev == dev && b.st == st) blk = &b;
    if (!blk) { blocks.push_back({dev, st, nullptr, 0}); blk = &blocks.back(); }
    if (n_frames > blk->cap) {
        if (blk->p) { CUDA_TRY(cudaStreamSynchronize(st)); cudaFree(blk->p); blk->p = nullptr; blk->cap = 0; }
        CUDA_TRY(cudaMalloc((void**)&blk->p, sizeof(WarpParams) * n_frames));
        blk->cap = n_frames;
    }
    std::vector<WarpParams> h(n_frames);
    for (int i = 0; i < n_frames; ++i) warp_params_from_T(T_host + 6 * i, &h[i]);
    CUDA_TRY(cudaMemcpyAsync(blk->p, h.data(), sizeof(WarpParams) * n_frames, cudaMemcpyHostToDevice, st));   // pageable source: staged before return
    launch_warp_matrices(d_src, src_w, src_h, src_stride, src_frame_bytes, d_dst, dst_w, dst_h, dst_stride,
                         dst_frame_bytes, blk->p, n_frames, st);
    CUDA_TRY(cudaGetLastError());
    return VS_OK;
    API_END
}

vs_status vs_nv12_to_bgr_device(const uint8_t* d_y, size_t y_stride, const uint8_t* d_uv, size_t uv_stride, int width, int height,
                                uint8_t* d_bgr, size_t bgr_stride, void* stream) {
    if (!d_y || !d_uv || !d_bgr || width < 2 || height < 2 || (width & 1) || (height & 1))
        return vs_set_error(VS_ERR_INVALID_ARG, "NV12 needs non-null planes and even dimensions");
    if (y_stride < (size_t)width || uv_stride < (size_t)width || bgr_stride < (size_t)width * 3)
        return vs_set_error(VS_ERR_INVALID_ARG, "row stride smaller than a row");
    launch_nv12_to_bgr(d_y, y_stride, d_uv, uv_stride, width, height, d_bgr, bgr_stride, (cudaStream_t)stream);
    CUDA_TRY(cudaGetLastError());
    return VS_OK;
}
vs_status vs_bgr_to_nv12_device(const uint8_t* d_bgr, size_t bgr_stride, int width, int height, uint8_t* d_y, size_t y_stride,
                                uint8_t* d_uv, size_t uv_stride, void* stream) {
    if (!d_y || !d_uv || !d_bgr || width < 2 || height < 2 || (width & 1) || (height & 1))
        return vs_set_error(VS_ERR_INVALID_ARG, "NV12 needs non-null planes and even dimensions");
    if (y_stride < (size_t)width || uv_stride < (size_t)width || bgr_stride < (size_t)width * 3)
        return vs_set_error(VS_ERR_INVALID_ARG, "row stride smaller than a row");
    launch_bgr_to_nv12(d_bgr, bgr_stride, width, height, d_y, y_stride, d_uv, uv_stride, (cudaStream_t)stream);
    CUDA_TRY(cudaGetLastError());
    return VS_OK;
}

vs_status vs_k_resize_linear_u8(const uint8_t* d_src, int sw, int sh, size_t sstride, int channels, uint8_t* d_dst,
                                int dw, int dh, size_t dstride, void* stream) {
    if (!d_src || !d_dst || (channels != 1 && channels != 3)) return vs_set_error(VS_ERR_INVALID_ARG, "bad argument");
    launch_resize_linear(d_src, sw, sh, sstride, channels, d_dst, dw, dh, dstride, (cudaStream_t)stream);
    CUDA_TRY(cudaGetLastError());
    return VS_OK;
}

vs_status vs_k_gray_pyramid(const uint8_t* d_bgr, int w, int h, size_t stride, int aw, int ah, uint8_t* d_l0,
                            uint8_t* d_l1, uint8_t* d_l2, void* stream) {
    if (!d_bgr || !d_l0) return vs_set_error(VS_ERR_INVALID_ARG, "bad argument");
    if (!((aw == VS_AW && ah == VS_AH) || (aw == VS_FW && ah == VS_FH)))
        return vs_set_error(VS_ERR_INVALID_ARG, "analysis size must be 960x540 or 480x270");
    API_BEGIN
    Engine* e = nullptr;
    vs_status s = scratch_engine(&e);
    if (s != VS_OK) return s;
    cudaStream_t st = e->stream();
    PtrPack src;
    src.p[0] = d_bgr;
    const LaneDev& L = e->h_lane(0);
    if (aw == VS_FW) {
        launch_gray_resize(e->d_lanes(), 1, src, w, h, stride, -1, st);
        launch_unpack_level(L.small0, d_l0, st);
    } else {
        launch_gray_resize(e->d_lanes(), 1, src, w, h, stride, 0, st);
        launch_pyrdown(e->d_lanes(), 1, 0, st);
        launch_unpack_level(L.pyr[0].lv[0], d_l0, st);
        if (d_l1) launch_unpack_level(L.pyr[0].lv[1], d_l1, st);
        if (d_l2) launch_unpack_level(L.pyr[0].lv[2], d_l2, st);
    }
    cudaError_t ce = cudaStreamSynchronize(st);
    delete e;
    if (ce != cudaSuccess) return vs_set_cuda_error(ce, "vs_k_gray_pyramid", __FILE__, __LINE__);
    (void)stream;
    return VS_OK;
    API_END
}

vs_status vs_k_good_features_block(const uint8_t* d_gray, int w, int h, int max_corners, double quality, double min_dist,
                                   int block_size, float* xy_out_host, int capacity, int* n_out, void* stream) {
    if (!d_gray || !n_out || block_size < 1 || block_size > 23) return vs_set_error(VS_ERR_INVALID_ARG, "bad argument");
    if (!((w == VS_AW && h == VS_AH) || (w == VS_FW && h == VS_FH)))
        return vs_set_error(VS_ERR_INVALID_ARG, "gray size must be 960x540 or 480x270");
    API_BEGIN
    Engine* e = nullptr;
    vs_status s = scratch_engine(&e, max_corners, min_dist < 1.0 ? 1.0 : min_dist);
    if (s != VS_OK) return s;
    cudaStream_t st = e->stream();
    const LaneDev& L = e->h_lane(0);
    const int slot = (w == VS_FW) ? -1 : 0;
    launch_pack_level(d_gray, slot < 0 ? L.small0 : L.pyr[0].lv[0], st);
    e->reset_detect_counters();
    float* eig = nullptr;
    if (block_size != 3 && cudaMalloc((void**)&eig, sizeof(float) * (size_t)w * h) != cudaSuccess) { delete e; return vs_set_cuda_error(cudaGetLastError(), "cudaMalloc", __FILE__, __LINE__); }
    launch_good_features(e->d_lanes(), 1, slot, max_corners, quality, min_dist, 0, 0, 0, st, block_size, eig);
    int n = 0;
    cudaError_t ce = cudaMemcpyAsync(&n, L.kp_count, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce == cudaSuccess && xy_out_host && n > 0)
        ce = cudaMemcpy(xy_out_host, L.kp, sizeof(float2) * (n < capacity ? n : capacity), cudaMemcpyDeviceToHost);
    delete e;
    if (eig) cudaFree(eig);
    if (ce != cudaSuccess) return vs_set_cuda_error(ce, "vs_k_good_features", __FILE__, __LINE__);
    *n_out = n;
    (void)stream;
    return VS_OK;
    API_END
}
vs_status vs_k_good_features(const uint8_t* d_gray, int w, int h, int max_corners, double quality, double min_dist,
                             float* xy_out_host, int capacity, int* n_out, void* stream) {
    return vs_k_good_features_block(d_gray, w, h, max_corners, quality, min_dist, 3, xy_out_host, capacity, n_out, stream);
}

vs_status vs_k_pyr_lk(const uint8_t* d_prev, const uint8_t* d_next, int w, int h, const float* pts_xy_host, int n,
                      float* next_xy_host, uint8_t* status_host, void* stream) {
    if (!d_prev || !d_next || !pts_xy_host || n < 0) return vs_set_error(VS_ERR_INVALID_ARG, "bad argument");
    if (w != VS_AW || h != VS_AH) return vs_set_error(VS_ERR_INVALID_ARG, "gray size must be 960x540");
    if (n > 2048) return vs_set_error(VS_ERR_INVALID_ARG, "at most 2048 points");
    API_BEGIN
    Engine* e = nullptr;
    vs_status s = scratch_engine(&e, n > 200 ? n : 200);
    if (s != VS_OK) return s;
    cudaStream_t st = e->stream();
    const LaneDev& L = e->h_lane(0);
    launch_pack_level(d_prev, L.pyr[0].lv[0], st);
    launch_pack_level(d_next, L.pyr[1].lv[0], st);
    launch_pyrdown(e->d_lanes(), 1, 0, st);
    launch_pyrdown(e->d_lanes(), 1, 1, st);
    cudaError_t ce = cudaMemcpyAsync(L.kp, pts_xy_host, sizeof(float2) * n, cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(L.kp_count, &n, sizeof(int), cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) {
        launch_pyr_lk(e->d_lanes(), 1, 0, 1, n, 0, 0, st, e->tracker_uses_tma());
        ce = cudaStreamSynchronize(st);
    }
    if (ce == cudaSuccess && n > 0) {
        if (next_xy_host) ce = cudaMemcpy(next_xy_host, L.lk_next, sizeof(float2) * n, cudaMemcpyDeviceToHost);
        if (ce == cudaSuccess && status_host) ce = cudaMemcpy(status_host, L.lk_status, n, cudaMemcpyDeviceToHost);
    }
    delete e;
    if (ce != cudaSuccess) return vs_set_cuda_error(ce, "vs_k_pyr_lk", __FILE__, __LINE__);
    (void)stream;
    return VS_OK;
    API_END
}

vs_status vs_k_estimate_affine_partial(const float* from_xy_host, const float* to_xy_host, int n, double* affine_out,
                                       uint8_t* inlier_mask_host, int* iters_out, int* ok, void* stream) {
    if (!from_xy_host || !to_xy_host || n < 0 || n > 2048 || !ok) return vs_set_error(VS_ERR_INVALID_ARG, "bad argument");
    API_BEGIN
    Engine* e = nullptr;
    vs_status s = scratch_engine(&e, n > 200 ? n : 200);
    if (s != VS_OK) return s;
    cudaStream_t st = e->stream();
    const LaneDev& L = e->h_lane(0);
    std::vector<uint8_t> ones(n > 0 ? n : 1, 1);
    cudaError_t ce = cudaMemcpyAsync(L.kp, from_xy_host, sizeof(float2) * n, cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(L.lk_next, to_xy_host, sizeof(float2) * n, cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(L.lk_status, ones.data(), n, cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(L.kp_count, &n, sizeof(int), cudaMemcpyHostToDevice, st);
    vs_frame_record rec{};
    if (ce == cudaSuccess) {
        StepInfo info{};
        info.frame_no = 1; info.cur = 1; info.pop_index = -1; info.smoothing_radius = 30;
        launch_motion(e->d_lanes(), 1, info, 0, st);
        ce = cudaStreamSynchronize(st);
    }
    if (ce == cudaSuccess) ce = cudaMemcpy(&rec, L.frec, sizeof(rec), cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess && inlier_mask_host && rec.n_inliers >= 0 && n > 0)
        ce = cudaMemcpy(inlier_mask_host, L.inlier_mask, n, cudaMemcpyDeviceToHost);
    delete e;
    if (ce != cudaSuccess) return vs_set_cuda_error(ce, "vs_k_estimate_affine_partial", __FILE__, __LINE__);
    *ok = rec.n_inliers >= 0 ? 1 : 0;
    if (iters_out) *iters_out = rec.ransac_iters;
    if (affine_out) for (int i = 0; i < 6; ++i) affine_out[i] = rec.affine[i];
    (void)stream;
    return VS_OK;
    API_END
}

vs_status vs_k_warp_output(const uint8_t* d_src, int w, int h, size_t stride, const float* T_host, int mode,
                           int border_size, int border_mode, uint8_t* d_dst, size_t dst_stride, int* out_w, int* out_h,
                           void* stream) {
    if (!d_src || !d_dst || !T_host || !out_w || !out_h) return vs_set_error(VS_ERR_INVALID_ARG, "bad argument");
    API_BEGIN
    Engine* e = nullptr;
    vs_status s = scratch_engine(&e);
    if (s != VS_OK) return s;
    cudaStream_t st = e->stream();
    const LaneDev& L = e->h_lane(0);
    WarpParams wp;
    warp_params_from_T(T_host, &wp);
    uint8_t* scratch = nullptr;
    cudaError_t ce = cudaMemcpyAsync(L.wp, &wp, sizeof(wp), cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess && mode == 2) ce = cudaMalloc((void**)&scratch, (size_t)w * 3 * h);
    if (ce == cudaSuccess) {
        WarpGeom g{};
        g.src_w = w; g.src_h = h; g.src_stride = stride; g.mode = mode; g.border = border_size; g.border_mode = border_mode;
        g.out_w = mode == 1 ? w + 2 * border_size : w;
        g.out_h = mode == 1 ? h + 2 * border_size : h;
        g.out_stride = dst_stride ? dst_stride : (size_t)g.out_w * 3;
        *out_w = g.out_w; *out_h = g.out_h;
        PtrPack src; src.p[0] = d_src;
        MutPtrPack dst; dst.p[0] = d_dst;
        uint8_t* sc[1] = {scratch};
        launch_warp(e->d_lanes(), 1, src, dst, g, sc, st);
        ce = cudaStreamSynchronize(st);
    }
    if (scratch) cudaFree(scratch);
    delete e;
    if (ce != cudaSuccess) return vs_set_cuda_error(ce, "vs_k_warp_output", __FILE__, __LINE__);
    (void)stream;
    return VS_OK;
    API_END
}

}  // extern "C"
