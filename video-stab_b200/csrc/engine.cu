// engine.cu — host orchestration.  Mirrors the control flow of Stabilizer::stabilize / flush / clean /
// generateTransform / applyNextSmoothTransform (Stabilizer.cpp:221-400, 402-761, 763-1137): every
// decision the reference takes on the host from frame COUNTS stays on the host; every decision it
// takes from DATA (tracked-pair count, RANSAC outcome, adaptive radius, motion intent) is taken on
// the device, so a step is a pure launch sequence with no host<->device round trip.
#include "engine.h"

#include <nvtx3/nvToolsExt.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

static thread_local char g_err[512] = "";

vs_status vs_set_cuda_error(cudaError_t e, const char* what, const char* file, int line) {
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    if (e == cudaErrorMemoryAllocation) return VS_ERR_OUT_OF_MEMORY;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice) return VS_ERR_NO_DEVICE;
    return VS_ERR_CUDA;
}
vs_status vs_set_error(vs_status st, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return st;
}
extern "C" const char* vs_last_error(void) { return g_err; }

// ---- optional per-stage CUDA-event timing (bench / profiles) and NVTX ranges (VS_NVTX=1: the enqueue of every stage shows
//      up as a named range in ncu / Nsight timelines and can be filtered on with `ncu --nvtx --nvtx-include`); both off by default
static bool nvtx_on() {
    static const bool v = [] { const char* e = getenv("VS_NVTX"); return e && *e == '1'; }();
    return v;
}
static const char* const k_stage_names[] = {"vs.gray_resize", "vs.pyrdown", "vs.pyr_lk", "vs.motion", "vs.gftt", "vs.warp", "vs.copy_in", "vs.copy_out"};
struct StageScope {
    Engine* e; int stage; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr; bool range = false;
    StageScope(Engine* e_, int s, cudaStream_t st_) : e(e_), stage(s), st(st_) {
        if (nvtx_on() && s >= 0 && s < (int)(sizeof(k_stage_names) / sizeof(k_stage_names[0]))) { nvtxRangePushA(k_stage_names[s]); range = true; }
        if (e->timing_on() && (only_stage() < 0 || only_stage() == s)) { a = e->take_event(); b = e->take_event(); cudaEventRecord(a, st); }
    }
    // VS_TRACE_STAGE=<stage>: time only that stage (two events per frame instead of twelve: an almost unperturbed pipeline)
    static int only_stage() {
        static int v = [] { const char* e = getenv("VS_TRACE_STAGE"); return e ? atoi(e) : -1; }();
        return v;
    }
    ~StageScope() {
        if (a) { cudaEventRecord(b, st); e->add_pending(stage, a, b); }
        if (range) nvtxRangePop();
    }
};

cudaEvent_t Engine::take_event() {
    if (!event_pool_.empty()) { cudaEvent_t ev = event_pool_.back(); event_pool_.pop_back(); return ev; }
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    return ev;
}
void Engine::add_pending(int stage, cudaEvent_t a, cudaEvent_t b) {
    pending_.push_back({stage, a, b});
    if (pending_.size() >= 8192) collect_timing();
}
void Engine::collect_timing() {
    if (pending_.empty()) return;
    sync();
    trace_.clear();
    for (auto& p : pending_) {
        float ms = 0.f, t0 = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { stage_ms_[p.stage] += ms; stage_n_[p.stage] += 1; }
        // timeline of the last collected batch, relative to its first event (diagnostics: vs_stabilizer_trace)
        if (cudaEventElapsedTime(&t0, pending_.front().a, p.a) == cudaSuccess) {
            trace_.push_back((float)p.stage); trace_.push_back(t0 * 1e3f); trace_.push_back((t0 + ms) * 1e3f);
        }
    }
    for (auto& p : pending_) { event_pool_.push_back(p.a); event_pool_.push_back(p.b); }
    pending_.clear();
}
void Engine::set_timing(bool on) {
    collect_timing();
    timing_ = on;
    for (int i = 0; i < VS_N_STAGES; ++i) { stage_ms_[i] = 0.; stage_n_[i] = 0; }
}
void Engine::stage_time(int stage, double* ms, long long* n) {
    collect_timing();
    *ms = (stage >= 0 && stage < VS_N_STAGES) ? stage_ms_[stage] : 0.;
    *n = (stage >= 0 && stage < VS_N_STAGES) ? stage_n_[stage] : 0;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

#define MO_MAXP_HOST 2048

vs_status Engine::create(const vs_params& p, int device, int n_lanes, Engine** out) {
    *out = nullptr;
    if (n_lanes < 1 || n_lanes > VS_MAX_GROUP) return vs_set_error(VS_ERR_INVALID_ARG, "n_streams must be 1..64");
    if (p.enable_virtual_canvas && !p.crop_n_zoom) {                                // the stage is skipped with crop_n_zoom (Stabilizer.cpp:1110-1127)
        if (n_lanes > 1) return vs_set_error(VS_ERR_UNSUPPORTED, "enable_virtual_canvas reads one transform back per frame; single-stream handles only");
        if (p.temporal_buffer_size < 0 || p.temporal_buffer_size > 240)
            return vs_set_error(VS_ERR_UNSUPPORTED, "temporal_buffer_size must be 0..240 (frames kept on the device)");
        const float lo = p.adaptive_canvas_size ? (p.min_canvas_scale < p.canvas_scale_factor ? p.min_canvas_scale : p.canvas_scale_factor) : p.canvas_scale_factor;
        if (!(lo >= 1.0f) || !(p.max_canvas_scale <= 8.0f) || !(p.canvas_scale_factor <= 8.0f))
            return vs_set_error(VS_ERR_UNSUPPORTED, "virtual canvas scales must lie in 1..8 (a canvas smaller than the frame is not built)");
    }
    // block_size applies to the first-frame detection only (Stabilizer.cpp:355-357); the window must stay inside the 16-pixel
    // reflect frame kept around every gray level
    if (p.block_size < 1 || p.block_size > 23) return vs_set_error(VS_ERR_UNSUPPORTED, "block_size must be 1..23");
    if (p.max_corners <= 0 || p.max_corners > MO_MAXP_HOST)
        return vs_set_error(VS_ERR_UNSUPPORTED, "max_corners must be 1..2048 (<= 0 means 'unlimited' to cv::goodFeaturesToTrack; the corner buffers are fixed-size)");
    if (p.adaptive_smoothing && n_lanes > 1)
        return vs_set_error(VS_ERR_UNSUPPORTED, "adaptive_smoothing makes the latency gate data dependent; single-stream handles only");
    // A handle uses up to nine streams (seven + two copy streams), one more than the default number of hardware queues
    // (CUDA_DEVICE_MAX_CONNECTIONS = 8).  The library does not touch that setting: on the B200 boxes it was measured on, raising
    // it helps a process that drives several handles at once (+5 % for two lock-step groups) and costs the page-locked copies of
    // the host-buffer path 8 - 24 % (profiles/r02_summary.md); the application knows which of the two it is.
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return vs_set_error(VS_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
    }
    if (device < 0 || device >= count) return vs_set_error(VS_ERR_INVALID_ARG, "bad device ordinal");
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return vs_set_error(VS_ERR_NO_DEVICE, "kernels are built for sm_100a (B200) only");
    Engine* eng = new Engine();
    vs_status st = eng->init(p, device, n_lanes);
    if (st != VS_OK) { delete eng; return st; }
    *out = eng;
    return VS_OK;
}

template <typename T>
static vs_status dalloc(std::vector<void*>& allocs, T** out, size_t n) {
    void* ptr = nullptr;
    CUDA_TRY(cudaMalloc(&ptr, n * sizeof(T)));
    allocs.push_back(ptr);
    CUDA_TRY(cudaMemset(ptr, 0, n * sizeof(T)));
    *out = (T*)ptr;
    return VS_OK;
}
#define VS_TRY(x) do { vs_status s__ = (x); if (s__ != VS_OK) return s__; } while (0)

static vs_status alloc_level(std::vector<void*>& allocs, int w, int h, GrayLevel* lv) {
    int pitch = (int)align_up((size_t)w + 2 * VS_PAD, 16);
    uint8_t* mem = nullptr;
    VS_TRY(dalloc(allocs, &mem, (size_t)pitch * (h + 2 * VS_PAD) + 64));   // +64: aligned 32-bit patch loads may over-read a few bytes
    lv->base = mem + (size_t)VS_PAD * pitch + VS_PAD;
    lv->w = w; lv->h = h; lv->pitch = pitch;
    return VS_OK;
}

vs_status Engine::init(const vs_params& p, int device, int n_lanes) {
    p_ = p;
    device_ = device;
    n_lanes_ = n_lanes;
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    // Seven streams per handle (engine.h): pyramid, two tracking, motion, two detection, and the public stream for the
    // warp.  adaptive_smoothing reads one int back per frame, so that mode stays on one stream.
    // VS_SINGLE_STREAM=1 runs everything on the public stream (verification: the multi-stream engine must reproduce it)
    multi_ = !p.adaptive_smoothing && !getenv("VS_SINGLE_STREAM");
    // The frame-independent half of the motion step (status filter, RANSAC, refit) runs behind LK on the tracking
    // streams, which shortens the sequential chain on the motion stream (27.2 -> 24.5 us device-side) for one more
    // launch per frame.  Measured on config 2 (two boxes, alternating runs): +12 % frames/s with the default 8 CUDA
    // connections (42.6k vs 37.8k), -3 % with 32 connections; batches and the host-buffer path are unaffected or
    // better.  VS_SPLIT_MOTION=0 keeps the step in one kernel.
    { const char* e = getenv("VS_SPLIT_MOTION"); split_motion_ = !(e && e[0] == '0'); }
    if (const char* e = getenv("VS_TRACK_N")) { const int v = atoi(e); if (v >= 1 && v <= VS_TRACK_STREAMS) track_n_ = v; }
    if (multi_) {
        // The analysis kernels are small and latency-critical (a single CTA for k_motion / k_select), the warp is one
        // machine-filling grid: the analysis streams get the higher priority so their CTAs are placed first whenever
        // a warp CTA retires, instead of queueing behind the rest of the warp grid.
        int prio_lo = 0, prio_hi = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        for (auto& st : sA_) CUDA_TRY(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio_hi));
        for (auto& st : sC_) CUDA_TRY(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio_hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&sP_, cudaStreamNonBlocking, prio_hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&sM_, cudaStreamNonBlocking, prio_hi));
        for (auto& ev : evS_) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        for (auto& ev : evW_) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        for (auto& ev : evP_) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        for (auto& ev : evA_) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        for (auto& ev : evB_) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        for (auto& ev : evJ_) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&evG_, cudaEventDisableTiming));
        for (auto& ev : evC_) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        // pipelined host I/O (push_many / flush_many): the copy-in and copy-out streams are created on first use
        // (a handle that never calls push_many stays at six streams: the default number of hardware queues is 8,
        // CUDA_DEVICE_MAX_CONNECTIONS, and streams beyond it pick up false dependencies)
        for (auto& ev : evH_) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        for (auto& ev : evRing_) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        for (auto& ev : evOutReady_) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        for (auto& ev : evOutFree_) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    }
    // mapBorderMode, Stabilizer.cpp:31-38 and the cropNZoom override :67-71
    border_mode_ = !strcmp(p.border_type, "reflect") ? 2 : !strcmp(p.border_type, "reflect_101") ? 4
                 : !strcmp(p.border_type, "replicate") ? 1 : !strcmp(p.border_type, "wrap") ? 3 : 0;
    if (p.crop_n_zoom) border_mode_ = 0;
    fade_ = !strcmp(p.border_type, "fade") && p.border_size > 0 && !p.crop_n_zoom;     // :914-916
    vc_on_ = p.enable_virtual_canvas && !p.crop_n_zoom;                                // :1110-1134
    canvas_.configure(p);
    method_ = !strcmp(p.smoothing_method, "gaussian") ? 1 : !strcmp(p.smoothing_method, "kalman") ? 2 : 0;
    smoothing_radius_ = p.smoothing_radius;
    cap_first_ = p.max_corners > 0 ? (p.max_corners < MO_MAXP_HOST ? p.max_corners : MO_MAXP_HOST) : MO_MAXP_HOST;
    int mc = p.max_corners < 200 ? p.max_corners : 200;                 // std::min(maxCorners, 200) :741
    cap_redetect_ = mc > 0 ? mc : MO_MAXP_HOST;
    kp_cap_ = cap_first_ > cap_redetect_ ? cap_first_ : cap_redetect_;
    log_depth_ = n_lanes == 1 ? 512 : 8;
    traj_cap_ = 1 << 15;
    // VS_TRAJ_CAP=<frames>: initial capacity of the growable trajectory / record arrays (tests use a small value to
    // exercise grow_trajectory() without a 20-minute clip)
    if (const char* tc = getenv("VS_TRAJ_CAP")) { int v = atoi(tc); if (v >= 64) traj_cap_ = v; }
    return alloc_fixed();
}

vs_status Engine::alloc_fixed() {
    h_lanes_.assign(n_lanes_, LaneDev{});
    VS_TRY(dalloc(allocs_, &d_lanes_, (size_t)n_lanes_));
    VS_TRY(dalloc(allocs_, &d_detect_counters_, (size_t)n_lanes_ * 4));      // [generation][lane][eig_max, cand_count]
    int* small_counters = nullptr;
    VS_TRY(dalloc(allocs_, &small_counters, (size_t)n_lanes_ * (1 + VS_KP_SLOTS)));
    traj_bufs_.clear();
    if (n_lanes_ > 8) VS_TRY(dalloc(allocs_, &d_tmaps_, (size_t)VS_MAX_GROUP * 128));
    for (int l = 0; l < n_lanes_; ++l) {
        LaneDev& L = h_lanes_[l];
        VS_TRY(alloc_level(allocs_, VS_FW, VS_FH, &L.small0));
        L.eig_max = d_detect_counters_ + 2 * l;
        L.cand_count = (int*)(d_detect_counters_ + 2 * l + 1);
        L.eig_max2 = d_detect_counters_ + 2 * n_lanes_ + 2 * l;
        L.cand_count2 = (int*)(d_detect_counters_ + 2 * n_lanes_ + 2 * l + 1);
        L.kp_count = small_counters + (1 + VS_KP_SLOTS) * l;
        L.first_count = small_counters + (1 + VS_KP_SLOTS) * l + 1;
        L.kpc[0] = L.kp_count;
        for (int k = 1; k < VS_KP_SLOTS; ++k) L.kpc[k] = small_counters + (1 + VS_KP_SLOTS) * l + 1 + k;
        VS_TRY(dalloc(allocs_, &L.kp, (size_t)kp_cap_));
        VS_TRY(dalloc(allocs_, &L.lk_next, (size_t)kp_cap_));
        VS_TRY(dalloc(allocs_, &L.lk_status, (size_t)kp_cap_));
        VS_TRY(dalloc(allocs_, &L.inlier_mask, (size_t)kp_cap_));
        L.kpb[0] = L.kp; L.lkn[0] = L.lk_next; L.lks[0] = L.lk_status;
        for (int k = 1; k < VS_KP_SLOTS; ++k) VS_TRY(dalloc(allocs_, &L.kpb[k], (size_t)kp_cap_));
        for (int k = 1; k < VS_LK_SLOTS; ++k) {
            VS_TRY(dalloc(allocs_, &L.lkn[k], (size_t)kp_cap_));
            VS_TRY(dalloc(allocs_, &L.lks[k], (size_t)kp_cap_));
        }
        VS_TRY(dalloc(allocs_, &L.first_corners, (size_t)kp_cap_));
        VS_TRY(dalloc(allocs_, &L.kalman, (size_t)VS_KAL_FLOATS));
        VS_TRY(dalloc(allocs_, &L.hf, (size_t)VS_HF_FLOATS));
        VS_TRY(dalloc(allocs_, &L.fit, (size_t)VS_EV_RING));
        VS_TRY(dalloc(allocs_, &L.wp, (size_t)VS_WP_SLOTS));
        for (int k = 0; k < VS_WP_SLOTS; ++k) L.wpb[k] = L.wp + k;
        size_t ln = (size_t)log_depth_ * kp_cap_;
        VS_TRY(dalloc(allocs_, &L.log_prev, ln));
        VS_TRY(dalloc(allocs_, &L.log_next, ln));
        VS_TRY(dalloc(allocs_, &L.log_status, ln));
        VS_TRY(dalloc(allocs_, &L.log_mask, ln));
        VS_TRY(dalloc(allocs_, &L.log_detected, ln));
        // trajectory + records: separately tracked so they can grow
        CUDA_TRY(cudaMalloc((void**)&L.transforms, (size_t)traj_cap_ * 3 * sizeof(float)));
        CUDA_TRY(cudaMalloc((void**)&L.path, (size_t)traj_cap_ * 3 * sizeof(float)));
        CUDA_TRY(cudaMalloc((void**)&L.aux, (size_t)traj_cap_ * 2 * sizeof(float)));
        CUDA_TRY(cudaMalloc((void**)&L.frec, (size_t)traj_cap_ * sizeof(vs_frame_record)));
        CUDA_TRY(cudaMalloc((void**)&L.orec, (size_t)traj_cap_ * sizeof(vs_output_record)));
        L.kp_capacity = kp_cap_;
        L.log_depth = log_depth_;
        L.record_capacity = traj_cap_;
    }
    return alloc_analysis(VS_AW, VS_AH);
}

// Everything whose size follows the analysis image: the pyramids, the detector's candidate list and min-distance grids, and
// the tracker's tensor maps.  960 x 540 (Stabilizer.cpp:410) at creation; drone_high_freq_mode picks its own size on the
// first frame (calculateDroneAnalysisSize, :2447-2466) and calls this again (the previous buffers stay allocated until the
// handle is destroyed).
vs_status Engine::alloc_analysis(int aw, int ah) {
    aw_ = aw; ah_ = ah;
    size_t gw = gftt_grid_words(VS_FW, VS_FH, p_.min_distance);
    size_t gw2 = gftt_grid_words(aw, ah, 15.0);
    if (gw2 > gw) gw = gw2;
    gw2 = gftt_grid_words(aw, ah, p_.min_distance);           // single-kernel entry point on an analysis-size image
    if (gw2 > gw) gw = gw2;
    for (int l = 0; l < n_lanes_; ++l) {
        LaneDev& L = h_lanes_[l];
        for (int s = 0; s < VS_PYR_SLOTS; ++s) {
            int w = aw, h = ah;
            for (int k = 0; k < VS_LEVELS; ++k) {
                VS_TRY(alloc_level(allocs_, w, h, &L.pyr[s].lv[k]));
                w = (w + 1) / 2; h = (h + 1) / 2;
            }
        }
        VS_TRY(dalloc(allocs_, &L.cand, (size_t)aw * ah));
        VS_TRY(dalloc(allocs_, &L.grid, gw));
        VS_TRY(dalloc(allocs_, &L.cand2, (size_t)aw * ah));
        VS_TRY(dalloc(allocs_, &L.grid2, gw));
        // tensor maps of this lane's pyramid planes (TMA-staged tracker, k_lk.cu); without them the tracker uses plain loads
        L.lk_maps = nullptr;
        if (l == 0 || lk_tma_) {
            std::vector<unsigned char> hm((size_t)VS_PYR_SLOTS * VS_LEVELS * 2 * 128 + 64);
            unsigned char* aligned = reinterpret_cast<unsigned char*>(((uintptr_t)hm.data() + 63) & ~(uintptr_t)63);
            if (lk_encode_maps(L, aligned)) {
                unsigned char* dm = nullptr;
                VS_TRY(dalloc(allocs_, &dm, (size_t)VS_PYR_SLOTS * VS_LEVELS * 2 * 128));
                CUDA_TRY(cudaMemcpy(dm, aligned, (size_t)VS_PYR_SLOTS * VS_LEVELS * 2 * 128, cudaMemcpyHostToDevice));
                L.lk_maps = dm;
                lk_tma_ = true;
            } else lk_tma_ = false;
        }
    }
    if (!lk_tma_) for (auto& L : h_lanes_) L.lk_maps = nullptr;
    CUDA_TRY(cudaMemcpy(d_lanes_, h_lanes_.data(), sizeof(LaneDev) * n_lanes_, cudaMemcpyHostToDevice));
    return VS_OK;
}

vs_status Engine::grow_trajectory() {
    VS_TRY(sync());
    int ncap = traj_cap_ * 2;
    for (int l = 0; l < n_lanes_; ++l) {
        LaneDev& L = h_lanes_[l];
        float *t = nullptr, *pa = nullptr, *ax = nullptr;
        vs_frame_record* fr = nullptr;
        vs_output_record* orr = nullptr;
        CUDA_TRY(cudaMalloc((void**)&t, (size_t)ncap * 3 * sizeof(float)));
        CUDA_TRY(cudaMalloc((void**)&pa, (size_t)ncap * 3 * sizeof(float)));
        CUDA_TRY(cudaMalloc((void**)&ax, (size_t)ncap * 2 * sizeof(float)));
        CUDA_TRY(cudaMalloc((void**)&fr, (size_t)ncap * sizeof(vs_frame_record)));
        CUDA_TRY(cudaMalloc((void**)&orr, (size_t)ncap * sizeof(vs_output_record)));
        CUDA_TRY(cudaMemcpy(t, L.transforms, (size_t)traj_cap_ * 3 * sizeof(float), cudaMemcpyDeviceToDevice));
        CUDA_TRY(cudaMemcpy(pa, L.path, (size_t)traj_cap_ * 3 * sizeof(float), cudaMemcpyDeviceToDevice));
        CUDA_TRY(cudaMemcpy(ax, L.aux, (size_t)traj_cap_ * 2 * sizeof(float), cudaMemcpyDeviceToDevice));
        CUDA_TRY(cudaMemcpy(fr, L.frec, (size_t)traj_cap_ * sizeof(vs_frame_record), cudaMemcpyDeviceToDevice));
        CUDA_TRY(cudaMemcpy(orr, L.orec, (size_t)traj_cap_ * sizeof(vs_output_record), cudaMemcpyDeviceToDevice));
        cudaFree(L.transforms); cudaFree(L.path); cudaFree(L.aux); cudaFree(L.frec); cudaFree(L.orec);
        L.transforms = t; L.path = pa; L.aux = ax; L.frec = fr; L.orec = orr;
        L.record_capacity = ncap;
    }
    traj_cap_ = ncap;
    CUDA_TRY(cudaMemcpy(d_lanes_, h_lanes_.data(), sizeof(LaneDev) * n_lanes_, cudaMemcpyHostToDevice));
    return VS_OK;
}

void Engine::free_all() {
    sync();
    collect_timing();
    for (auto& ev : evA_) if (ev) { cudaEventDestroy(ev); ev = nullptr; }
    for (auto& ev : evP_) if (ev) { cudaEventDestroy(ev); ev = nullptr; }
    if (sP_) { cudaStreamDestroy(sP_); sP_ = nullptr; }
    if (sM_) { cudaStreamDestroy(sM_); sM_ = nullptr; }
    for (auto& ev : evS_) if (ev) { cudaEventDestroy(ev); ev = nullptr; }
    for (auto& ev : evW_) if (ev) { cudaEventDestroy(ev); ev = nullptr; }
    for (auto& ev : evB_) if (ev) { cudaEventDestroy(ev); ev = nullptr; }
    for (auto& ev : evJ_) if (ev) { cudaEventDestroy(ev); ev = nullptr; }
    if (evG_) { cudaEventDestroy(evG_); evG_ = nullptr; }
    for (auto& ev : evC_) if (ev) { cudaEventDestroy(ev); ev = nullptr; }
    for (auto& ev : evH_) if (ev) { cudaEventDestroy(ev); ev = nullptr; }
    for (auto& ev : evRing_) if (ev) { cudaEventDestroy(ev); ev = nullptr; }
    for (auto& ev : evOutReady_) if (ev) { cudaEventDestroy(ev); ev = nullptr; }
    for (auto& ev : evOutFree_) if (ev) { cudaEventDestroy(ev); ev = nullptr; }
    for (auto& st : sA_) if (st) { cudaStreamDestroy(st); st = nullptr; }
    for (auto& st : sC_) if (st) { cudaStreamDestroy(st); st = nullptr; }
    if (sH_) { cudaStreamDestroy(sH_); sH_ = nullptr; }
    if (sO_) { cudaStreamDestroy(sO_); sO_ = nullptr; }
    for (cudaEvent_t ev : event_pool_) cudaEventDestroy(ev);
    event_pool_.clear();
    for (auto& L : h_lanes_) {
        cudaFree(L.transforms); cudaFree(L.path); cudaFree(L.aux); cudaFree(L.frec); cudaFree(L.orec);
    }
    h_lanes_.clear();
    for (void* p : allocs_) cudaFree(p);
    allocs_.clear();
    if (d_ring_) cudaFree(d_ring_);
    if (d_out_) cudaFree(d_out_);
    if (d_scratch_) cudaFree(d_scratch_);
    if (d_fade_) cudaFree(d_fade_);
    d_fade_ = nullptr;
    if (d_wp_batch_) cudaFree(d_wp_batch_);
    d_wp_batch_ = nullptr; wp_batch_cap_ = 0;
    d_ring_ = d_out_ = d_scratch_ = nullptr;
    if (stream_) cudaStreamDestroy(stream_);
    stream_ = nullptr;
}

Engine::~Engine() {
    cudaSetDevice(device_);
    canvas_.reset();
    if (h_vc_wp_) cudaFreeHost(h_vc_wp_);
    free_all();
}

vs_status Engine::sync() {
    if (sH_) CUDA_TRY(cudaStreamSynchronize(sH_));
    if (sP_) CUDA_TRY(cudaStreamSynchronize(sP_));
    if (sM_) CUDA_TRY(cudaStreamSynchronize(sM_));
    for (auto& st : sA_) if (st) CUDA_TRY(cudaStreamSynchronize(st));
    for (auto& st : sC_) if (st) CUDA_TRY(cudaStreamSynchronize(st));
    if (stream_) CUDA_TRY(cudaStreamSynchronize(stream_));
    if (sO_) CUDA_TRY(cudaStreamSynchronize(sO_));
    return VS_OK;
}

// Orders everything enqueued so far (on all three streams) before whatever is enqueued on the public stream next.
vs_status Engine::join() {
    if (!multi_) return VS_OK;
    for (int k = 0; k < VS_TRACK_STREAMS; ++k) CUDA_TRY(cudaEventRecord(evJ_[k], sA_[k]));
    CUDA_TRY(cudaEventRecord(evJ_[VS_TRACK_STREAMS], sC_[0]));
    CUDA_TRY(cudaEventRecord(evJ_[VS_TRACK_STREAMS + 1], sC_[1]));
    CUDA_TRY(cudaEventRecord(evJ_[VS_TRACK_STREAMS + 2], sP_));
    CUDA_TRY(cudaEventRecord(evJ_[VS_TRACK_STREAMS + 3], sM_));
    for (auto& ev : evJ_) CUDA_TRY(cudaStreamWaitEvent(stream_, ev, 0));
    return VS_OK;
}

vs_status Engine::clean() {
    // Stabilizer::clean, Stabilizer.cpp:221-256
    VS_TRY(sync());
    for (bool& b : evB_set_) b = false;
    for (bool& b : evA_set_) b = false;
    for (bool& b : evW_set_) b = false;
    last_detect_frame_ = -100;
    for (bool& b : ring_ev_set_) b = false;
    for (bool& b : out_free_set_) b = false;
    for (bool& b : c_pending_) b = false;
    queue_.clear();
    kalman_next_ = -1; clip_total_ = 0;
    first_ = true;
    next_index_ = 0;
    n_frames_ = 0;
    n_out_ = 0;
    detect_counter_ = 0;
    smoothing_radius_ = p_.smoothing_radius;
    W_ = H_ = 0;
    return VS_OK;
}

vs_status Engine::ensure_geometry(int w, int h, bool need_ring, bool need_out, bool need_scratch) {
    if (p_.drone_high_freq_mode) {
        // calculateDroneAnalysisSize (Stabilizer.cpp:2447-2466)
        const int mw = p_.hf_analysis_max_width < w ? p_.hf_analysis_max_width : w;
        const float aspect = (float)h / (float)w;
        const int ah0 = (int)((float)mw * aspect);
        const int aw = (mw / 2) * 2, ah = (ah0 / 2) * 2;
        if (aw != aw_ || ah != ah_) {
            // the tracker always runs three pyramid levels; OpenCV drops the levels that are no larger than the 15 x 15 window (:611-619)
            if (aw > 4096 || ((aw + 1) / 2 + 1) / 2 <= VS_WIN || ((ah + 1) / 2 + 1) / 2 <= VS_WIN)
                return vs_set_error(VS_ERR_UNSUPPORTED, "drone_high_freq_mode: the drone analysis size (min(hf_analysis_max_width, width) x matching "
                                                        "height, both made even) must be at least 62 x 62 and at most 4096 wide");
            VS_TRY(sync());
            VS_TRY(alloc_analysis(aw, ah));
        }
    }
    if (W_ == 0) {
        W_ = w; H_ = h;
        frame_bytes_ = (size_t)w * 3 * h;
        int b = (p_.border_size > 0 && !p_.crop_n_zoom) ? p_.border_size : 0;
        out_bytes_ = (size_t)(w + 2 * b) * 3 * (h + 2 * b);
        ring_slots_ = 36;           // latency gate is at most 35 frames (Stabilizer.cpp:383)
        if (d_ring_) { cudaFree(d_ring_); d_ring_ = nullptr; }
        if (d_out_) { cudaFree(d_out_); d_out_ = nullptr; }
        if (d_scratch_) { cudaFree(d_scratch_); d_scratch_ = nullptr; }
        // borderHistory_ / fadeFrameCount_ survive clean() in the reference (Stabilizer.cpp:221-256 does not touch them):
        // keep them unless the geometry changed
        if (d_fade_ && (w != fade_w_ || h != fade_h_)) { cudaFree(d_fade_); d_fade_ = nullptr; fade_hist_valid_ = false; fade_count_ = 0; }
        fade_w_ = w; fade_h_ = h;
    } else if (w != W_ || h != H_) {
        return vs_set_error(VS_ERR_INVALID_ARG, "frame size changed mid-stream (call clean() first)");
    }
    if (need_ring && !d_ring_) CUDA_TRY(cudaMalloc((void**)&d_ring_, frame_bytes_ * ring_slots_ * n_lanes_));
    if (need_out && !d_out_) CUDA_TRY(cudaMalloc((void**)&d_out_, out_bytes_ * n_lanes_ * VS_OUT_SLOTS));
    if (need_scratch && !d_scratch_) CUDA_TRY(cudaMalloc((void**)&d_scratch_, frame_bytes_ * n_lanes_));
    if (fade_ && !d_fade_) CUDA_TRY(cudaMalloc((void**)&d_fade_, out_bytes_ * n_lanes_ * 2));     // history + blended source
    return VS_OK;
}

StepInfo Engine::step_info(int pop_index) const {
    StepInfo s{};
    s.frame_no = n_frames_;
    s.cur = n_frames_ % VS_PYR_SLOTS;
    s.pop_index = pop_index;
    s.path_len_at_pop = n_frames_;
    s.smoothing_radius = smoothing_radius_;
    s.method = method_;
    s.gaussian_sigma = (float)p_.gaussian_sigma;
    s.horizon_lock = p_.horizon_lock;
    s.n_out = n_out_;
    s.adaptive = p_.adaptive_smoothing;
    s.min_radius = p_.min_smoothing_radius;
    s.max_radius = p_.max_smoothing_radius;
    s.kp_slot = ((n_frames_ - 1) / 2) % VS_KP_SLOTS;
    s.lk_slot = n_frames_ % VS_LK_SLOTS;
    s.will_detect = ((detect_counter_ + 1) % 2) == 0;            // (++featureDetectionCounter % 2) == 0, Stabilizer.cpp:696-697
    s.wp_slot = n_out_ % VS_WP_SLOTS;
    s.drone = p_.drone_high_freq_mode;
    s.hf_shake_px = p_.hf_shake_px;
    s.hf_rot_lp_alpha = p_.hf_rot_lp_alpha;
    s.hf_dead_zone_threshold = p_.hf_dead_zone_threshold;
    s.hf_accumulator_decay = p_.hf_motion_accumulator_decay;
    s.hf_freeze_duration = p_.hf_freeze_duration;
    return s;
}

// First-frame analysis (Stabilizer.cpp:271-368): 480x270 gray + GFTT with the user's parameters -> key-point slot 0
vs_status Engine::first_frame_detect(const PtrPack& src, int w, int h, size_t stride) {
    launch_gray_resize(d_lanes_, n_lanes_, src, w, h, stride, -1, sp());                  // :304-305
    if (multi_) { CUDA_TRY(cudaEventRecord(evG_, sp())); CUDA_TRY(cudaStreamWaitEvent(sc(0), evG_, 0)); }
    CUDA_TRY(cudaMemsetAsync(d_detect_counters_, 0, sizeof(unsigned int) * 2 * n_lanes_, sc(0)));
    if (p_.block_size != 3 && !d_eig_generic_) {
        CUDA_TRY(cudaMalloc((void**)&d_eig_generic_, sizeof(float) * VS_FW * VS_FH * n_lanes_));
        allocs_.push_back(d_eig_generic_);
    }
    launch_good_features(d_lanes_, n_lanes_, -1, p_.max_corners, p_.quality_level, p_.min_distance, 0, 0, 0, sc(0), p_.block_size, d_eig_generic_);  // :355-357
    if (multi_) { CUDA_TRY(cudaEventRecord(evC_[0], sc(0))); c_pending_[0] = true; }
    launches_ += 3;
    return VS_OK;
}

// Corner re-detection on analysis frame `frame_no` (pyramid slot `cur`) -> key-point slot (frame_no / 2) & 1, on the
// detection stream.  It reads level 0 of the frame (ready at evG_) and overwrites the key points last read by
// the motion kernel of frame_no - 2.
vs_status Engine::redetect(int cur, int frame_no, int record_frame_no, cudaEvent_t level0_ready) {
    const int gen = (frame_no / 2) & 1;               // detection generation: scratch set and stream
    const int ks = (frame_no / 2) % VS_KP_SLOTS;      // key-point slot written (LK / motion of frame_no + 1, + 2 read it)
    if (multi_) {
        CUDA_TRY(cudaStreamWaitEvent(sc(gen), level0_ready, 0));
        // slot ks was last read by the motion kernel of frame_no - 6: complete, see the guard in generate_transform
        // (with two slots the explicit wait on motion(frame_no - 2) closed a loop detect(n-2) -> LK(n-1) -> LK(n) ->
        // motion(n) -> detect(n+2) that set the pipeline period)
    }
    StageScope t(this, VS_STAGE_GFTT, sc(gen));
    int mc = p_.max_corners < 200 ? p_.max_corners : 200;       // (the counters were left zeroed by the previous k_select)
    launch_good_features(d_lanes_, n_lanes_, cur, mc, 0.02, 15.0, record_frame_no, gen, ks, sc(gen), 3, nullptr, aw_, ah_);   // :740-744
    if (multi_) { CUDA_TRY(cudaEventRecord(evC_[ks], sc(gen))); c_pending_[ks] = true; last_detect_frame_ = frame_no; }
    launches_ += 2;
    return VS_OK;
}

// generateTransform, Stabilizer.cpp:402-761 (CPU branch) as a launch sequence over the handle's streams:
//   P  (pyramid)    gray -> pyramid               needs: motion(n-8) complete, checked every 4th frame - the one slot guard (below)
//   A0/A1 (tracking) LK(n) on stream n & 1        needs: pyramid(n) (evP_), key points of the last detection (evC_)
//   M  (motion)     RANSAC .. smoothing, set-up   needs: LK of this frame (evA_), warp set-up slot (evW_, every 4th output)
//   O  (public)     warp, in emit()               needs: motion of this step (evB_)
//   C0/C1 (detection) eig+NMS -> select           needs: pyramid of this frame (evP_)
// Pyramids live in 12 slots, tracker output in 16, key points in 8, warp set-ups in 8, detection scratch in 2; every
// kernel of frame n-8 and older is complete once motion(n-8) is (the motion stream is in order and consumes LK, which
// consumes the detection), and LK(n) / detect(n) start after pyramid(n), so one wait per four frames guards all of them.
vs_status Engine::generate_transform(const QueueEntry& e, bool* will_pop) {
    if (n_frames_ + 1 >= traj_cap_) VS_TRY(grow_trajectory());
    const int frame_no = ++n_frames_;
    const int cur = frame_no % VS_PYR_SLOTS, prev = (frame_no - 1) % VS_PYR_SLOTS;
    const int kp_slot = ((frame_no - 1) / 2) % VS_KP_SLOTS, lk_slot = frame_no % VS_LK_SLOTS;
    const bool detect = ((detect_counter_ + 1) % 2) == 0;                             // :696-697
    PtrPack src;
    for (int l = 0; l < n_lanes_; ++l) src.p[l] = e.frames[l];
    if (frame_no == 1) {
        // prevGray is still the 480x270 first-frame image: cv::resize it up (Stabilizer.cpp:598-603)
        launch_upsample_small(d_lanes_, n_lanes_, prev, sp(), aw_, ah_);
        launch_pyrdown(d_lanes_, n_lanes_, prev, sp(), aw_, ah_);
        launches_ += 2;
    }
    if (multi_) {
        // Slot `cur` was last read by LK(frame_no - VS_PYR_SLOTS + 1) (as its previous frame), LK(frame_no - VS_PYR_SLOTS)
        // (as its current frame) and the detection of frame_no - VS_PYR_SLOTS.  ONE wait covers them and more:
        // motion(guard) complete means LK(guard) complete (it consumed its output), hence the detection whose key points
        // that LK used, and - the motion stream being in order - motion and LK of every earlier frame.  The tracker-
        // output and key-point rings are at least as deep (static_asserts in common.cuh), and LK(n) and the detection
        // of frame n both start after pyramid(n), so the same wait guards them.  The rings are deeper than the
        // pipeline is long (gray -> motion takes ~150 us, 5 - 6 frame periods), so the wait never binds.
        // Taken at the first frame of each group of VS_GUARD_GROUP, on the motion kernel the LAST frame of the group
        // needs (one stream-wait call per four frames instead of one per frame: the enqueue loop is host-bound).
        const int guard = frame_no - VS_PYR_SLOTS + VS_GUARD_GROUP;
        if (frame_no % VS_GUARD_GROUP == 0 && guard >= 1 && evB_set_[guard & (VS_EV_RING - 1)])
            CUDA_TRY(cudaStreamWaitEvent(sp(), evB_[guard & (VS_EV_RING - 1)], 0));
    }
    { StageScope t(this, VS_STAGE_GRAY, sp());
      launch_gray_resize(d_lanes_, n_lanes_, src, W_, H_, e.stride, cur, sp(), aw_, ah_); }       // :449-450
    { StageScope t(this, VS_STAGE_PYRDOWN, sp());
      launch_pyrdown(d_lanes_, n_lanes_, cur, sp(), aw_, ah_); }
    if (multi_) {
        CUDA_TRY(cudaEventRecord(evP_[frame_no & (VS_EV_RING - 1)], sp()));
        // LK(n) and LK(n+1) are independent (key points are not advanced between detections, Appendix B Q4): they
        // alternate between two tracking streams, so the tracker is not a serial chain
        CUDA_TRY(cudaStreamWaitEvent(sa(frame_no), evP_[frame_no & (VS_EV_RING - 1)], 0));
        if (c_pending_[kp_slot]) CUDA_TRY(cudaStreamWaitEvent(sa(frame_no), evC_[kp_slot], 0));   // both frames after a detection
    }
    { StageScope t(this, VS_STAGE_LK, sa(frame_no));
      launch_pyr_lk(d_lanes_, n_lanes_, prev, cur, frame_no <= 2 ? cap_first_ : cap_redetect_, kp_slot, lk_slot, sa(frame_no), lk_tma_); }   // :611-619
    launches_ += 3;
    if (multi_ && split_motion_) {
        // the frame-independent half of the motion step (status filter, RANSAC, refit) runs right behind LK on the
        // tracking stream; only trajectory + smoothing + set-up stay on the sequential motion stream
        { StageScope t(this, VS_STAGE_MOTION, sa(frame_no));
          launch_motion(d_lanes_, n_lanes_, step_info(-1), 1, sa(frame_no)); }
        launches_ += 1;
    }
    if (multi_) {
        CUDA_TRY(cudaEventRecord(evA_[frame_no & (VS_EV_RING - 1)], sa(frame_no)));
        evA_set_[frame_no & (VS_EV_RING - 1)] = true;
        CUDA_TRY(cudaStreamWaitEvent(sm(), evA_[frame_no & (VS_EV_RING - 1)], 0));
    }

    const bool adaptive = p_.adaptive_smoothing != 0;
    int pop_index = -1;
    if (!adaptive) {
        int gate = clampi(smoothing_radius_, 5, 35);                                  // :383
        *will_pop = (int)queue_.size() >= gate;
        if (*will_pop) pop_index = queue_.front().index;
    }
    if (pop_index >= 0) VS_TRY(setup_slot_guard());
    { StageScope t(this, VS_STAGE_MOTION, sm());
      launch_motion(d_lanes_, n_lanes_, step_info(pop_index), (multi_ && split_motion_) ? 2 : 0, sm()); }  // :629-688 (+ :783-908)
    launches_ += 1;
    if (multi_) {
        // one event serves both consumers of this kernel: the tracker-slot guard of LK(frame_no + VS_LK_SLOTS) and the
        // warp of the output this step set up (fewer stream-semaphore operations per frame)
        CUDA_TRY(cudaEventRecord(evB_[frame_no & (VS_EV_RING - 1)], sm())); evB_set_[frame_no & (VS_EV_RING - 1)] = true;
        if (pop_index >= 0) CUDA_TRY(cudaStreamWaitEvent(stream_, evB_[frame_no & (VS_EV_RING - 1)], 0));
    }

    ++detect_counter_;
    if (detect) VS_TRY(redetect(cur, frame_no, frame_no, evP_[frame_no & (VS_EV_RING - 1)]));   // pyramid-complete event: one record fewer
    if (adaptive) {
        // updateAdaptiveParameters (:691-693, :1562-1574) changes params_.smoothingRadius, which moves the
        // latency gate: the one data-dependent host decision of the path, so this mode reads it back.
        if (frame_no >= 3) {
            int nr = 0;
            CUDA_TRY(cudaMemcpyAsync(&nr, h_lanes_[0].kalman + VS_KAL_RADIUS_SLOT, sizeof(int), cudaMemcpyDeviceToHost, sm()));
            CUDA_TRY(cudaStreamSynchronize(sm()));
            smoothing_radius_ = nr;
        }
        int gate = clampi(smoothing_radius_, 5, 35);
        *will_pop = (int)queue_.size() >= gate;
        if (*will_pop) {
            VS_TRY(setup_slot_guard());
            launch_smooth_only(d_lanes_, n_lanes_, step_info(queue_.front().index), sm());
            launches_ += 1;
            VS_TRY(setup_ready());
        }
    }
    return VS_OK;
}

// The warp set-up of output k lives in LaneDev::wpb[k % 8]; its writer (motion stream) must not overtake the warp of
// output k - 8 (public stream).  Outputs are guarded in groups of four: the public stream records evW_[g & 1] after the
// last warp of group g = k / 4, and the motion stream waits for group g - 2 before the first set-up of group g - one
// record and one wait per four frames.
vs_status Engine::setup_slot_guard() {
    const int g = n_out_ >> 2;
    if (multi_ && (n_out_ & 3) == 0 && g >= 2 && evW_set_[g & 1]) CUDA_TRY(cudaStreamWaitEvent(sm(), evW_[g & 1], 0));
    return VS_OK;
}
vs_status Engine::setup_ready() {
    if (multi_) {
        CUDA_TRY(cudaEventRecord(evS_[n_out_ & 1], sm()));
        CUDA_TRY(cudaStreamWaitEvent(stream_, evS_[n_out_ & 1], 0));
    }
    return VS_OK;
}

// Output geometry of the frame at the head of the queue, and whether the caller's buffer can take it.  Called at the top of
// push()/flush(), BEFORE anything is queued, launched or popped: a too-small buffer is a recoverable caller error and
// must leave the handle exactly as it was (the call can simply be repeated with a larger buffer).
vs_status Engine::check_out_buffer(bool passthrough, uint8_t* const* outs, size_t out_stride, size_t out_capacity) const {
    const int b = p_.border_size;
    const bool grows = !passthrough && b > 0 && !p_.crop_n_zoom && !vc_on_;              // copyMakeBorder path, :981-990 (the canvas stage returns frame size)
    const int w = grows ? W_ + 2 * b : W_, h = grows ? H_ + 2 * b : H_;
    const size_t tight = (size_t)w * 3;
    if (out_stride == 0) out_stride = tight;
    if (!outs) return vs_set_error(VS_ERR_INVALID_ARG, "no output buffer");
    for (int l = 0; l < n_lanes_; ++l)
        if (!outs[l]) return vs_set_error(VS_ERR_INVALID_ARG, "no output buffer");
    if (out_stride < tight || out_stride * (size_t)(h - 1) + tight > out_capacity)
        return vs_set_error(VS_ERR_BUFFER_TOO_SMALL, "output buffer too small for the stabilized frame");
    return VS_OK;
}

// the warp half of applyNextSmoothTransform, Stabilizer.cpp:979-1137
vs_status Engine::emit(uint8_t* const* outs, size_t out_stride, size_t out_capacity, int io, int* ow, int* oh) {
    const bool host_io = io != VS_IO_DEVICE, pipe = io == VS_IO_HOST_PIPE && multi_;
    if (pipe && !sO_) {
        CUDA_TRY(cudaStreamCreateWithFlags(&sH_, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&sO_, cudaStreamNonBlocking));
    }
    QueueEntry e = queue_.front();
    queue_.pop_front();
    const bool passthrough = e.index >= n_frames_;                                    // :774-780
    const int b = p_.border_size;
    const bool canvas = vc_on_ && !passthrough;                                          // :1129-1134: replaces the warped frame
    const int mode = (passthrough || b <= 0 || canvas) ? 0 : (p_.crop_n_zoom ? 2 : 1);
    int w = W_, h = H_;
    if (mode == 1) { w = W_ + 2 * b; h = H_ + 2 * b; }
    if (mode == 2 && (W_ - 2 * b <= 0 || H_ - 2 * b <= 0)) {}                         // :1114-1115 handled below
    *ow = w; *oh = h;
    const size_t tight = (size_t)w * 3;
    if (out_stride == 0) out_stride = tight;
    if (out_stride < tight || out_stride * (size_t)(h - 1) + tight > out_capacity)
        return vs_set_error(VS_ERR_BUFFER_TOO_SMALL, "output buffer too small for the stabilized frame");
    VS_TRY(ensure_geometry(W_, H_, false, host_io, mode == 2));

    // pipelined host I/O: warp into one of VS_OUT_SLOTS staging frames, copy out on the copy-out stream
    const int oslot = pipe ? n_out_ % VS_OUT_SLOTS : 0;
    if (pipe && out_free_set_[oslot]) CUDA_TRY(cudaStreamWaitEvent(stream_, evOutFree_[oslot], 0));
    MutPtrPack dst;
    for (int l = 0; l < n_lanes_; ++l) dst.p[l] = host_io ? d_out_ + out_bytes_ * ((size_t)l * VS_OUT_SLOTS + oslot) : outs[l];
    const size_t dstride = host_io ? tight : out_stride;
    if (passthrough) {
        if (multi_ && e.in_ring) {
            // A passed-through frame has no transform, hence no pyramid -> LK -> motion chain that would order its ring copy
            // (enqueued on the pyramid or copy-in stream) before this read on the public stream: order it explicitly.  An
            // event recorded now is behind that copy.  Happens once per clip (the last frame), or for a lone first frame.
            CUDA_TRY(cudaEventRecord(evJ_[VS_TRACK_STREAMS + 2], sP_));
            CUDA_TRY(cudaStreamWaitEvent(stream_, evJ_[VS_TRACK_STREAMS + 2], 0));
            if (sH_) {
                CUDA_TRY(cudaEventRecord(evJ_[VS_TRACK_STREAMS + 3], sH_));
                CUDA_TRY(cudaStreamWaitEvent(stream_, evJ_[VS_TRACK_STREAMS + 3], 0));
            }
        }
        for (int l = 0; l < n_lanes_; ++l)
            CUDA_TRY(cudaMemcpy2DAsync(dst.p[l], dstride, e.frames[l], e.stride, tight, h, cudaMemcpyDeviceToDevice, stream_));
    } else if (canvas && canvas_.never_fills() && canvas_.geometry(W_, H_)) {
        // nothing can be filled with this canvas: the frame moves by whole pixels, taken from the set-up block on the device
        int nl = 0;
        { StageScope t(this, VS_STAGE_WARP, stream_);
          VS_TRY(canvas_.apply_async(e.frames[0], W_, H_, e.stride, h_lanes_[0].wpb[n_out_ % VS_WP_SLOTS], dst.p[0], dstride, stream_, &nl)); }
        launches_ += nl;
    } else if (canvas) {
        // The correction (dx, dy, da) of this output is read back: the stage's rectangles are integer functions of it and its
        // region logic runs on the host, as in the reference.  The warp itself is not launched - its result is discarded by
        // the reference (:1133), and so is the fade history it would update.
        WarpParams* h_wp = nullptr;
        if (!h_vc_wp_) CUDA_TRY(cudaMallocHost((void**)&h_vc_wp_, sizeof(WarpParams) + sizeof(float) * 90));
        h_wp = reinterpret_cast<WarpParams*>(h_vc_wp_);
        float* h_recent = reinterpret_cast<float*>(h_vc_wp_ + sizeof(WarpParams));
        CUDA_TRY(cudaMemcpyAsync(h_wp, h_lanes_[0].wpb[n_out_ % VS_WP_SLOTS], sizeof(WarpParams), cudaMemcpyDeviceToHost, stream_));
        int n_recent = 0;
        if (!canvas_.sized()) {                                                     // first output: transforms_ sizes the canvas
            n_recent = n_frames_ < 30 ? n_frames_ : 30;
            if (n_recent > 0)
                CUDA_TRY(cudaMemcpyAsync(h_recent, h_lanes_[0].transforms + 3 * (size_t)(n_frames_ - n_recent), sizeof(float) * 3 * n_recent,
                                         cudaMemcpyDeviceToHost, stream_));
        }
        CUDA_TRY(cudaStreamSynchronize(stream_));
        const float T3[3] = {h_wp->T[2], h_wp->T[5], h_wp->da};
        int nl = 0;
        { StageScope t(this, VS_STAGE_WARP, stream_);
          VS_TRY(canvas_.apply(e.frames[0], W_, H_, e.stride, T3, h_recent, n_recent, dst.p[0], dstride, stream_, &nl)); }
        launches_ += nl;
    } else {
        PtrPack src;
        for (int l = 0; l < n_lanes_; ++l) src.p[l] = e.frames[l];
        WarpGeom g{};
        g.src_w = W_; g.src_h = H_; g.src_stride = e.stride;
        g.mode = mode; g.border = b; g.border_mode = border_mode_;
        if (fade_) {
            // the source of the warp is the history-blended bordered frame (Stabilizer.cpp:914-978)
            float alpha = p_.fade_alpha;
            if (fade_count_ < p_.fade_duration) {
                alpha = alpha * ((float)fade_count_ / p_.fade_duration);
                ++fade_count_;
            }
            const float beta = 1.0f - alpha;
            uint8_t* hist = d_fade_;
            uint8_t* blend = d_fade_ + out_bytes_ * n_lanes_;
            launch_fade_blend(src, n_lanes_, W_, H_, e.stride, b, hist, blend, alpha, beta, !fade_hist_valid_, stream_);
            fade_hist_valid_ = true;
            for (int l = 0; l < n_lanes_; ++l) src.p[l] = blend + out_bytes_ * l;
            g.src_w = w; g.src_h = h; g.src_stride = tight;
            g.mode = 0;
            launches_ += 1;
        }
        g.out_w = w; g.out_h = h; g.out_stride = dstride;
        g.d_tmaps = d_tmaps_;
        g.wp_slot = n_out_ % VS_WP_SLOTS;
        int m2 = fade_ ? 0 : mode;
        if (mode == 2 && (W_ - 2 * b <= 0 || H_ - 2 * b <= 0)) m2 = 0;                // border larger than image
        g.mode = m2;
        std::vector<uint8_t*> scratch(n_lanes_);
        for (int l = 0; l < n_lanes_; ++l) scratch[l] = d_scratch_ ? d_scratch_ + frame_bytes_ * l : nullptr;
        { StageScope t(this, VS_STAGE_WARP, stream_);
          launches_ += launch_warp(d_lanes_, n_lanes_, src, dst, g, scratch.data(), stream_); }
        if (fade_) {                                                                  // :1070-1106
            launch_fade_update(d_fade_, dst, dstride, n_lanes_, W_, H_, b, stream_);
            launches_ += 1;
        }
    }
    if (multi_) {
        if ((n_out_ & 3) == 3) { CUDA_TRY(cudaEventRecord(evW_[(n_out_ >> 2) & 1], stream_)); evW_set_[(n_out_ >> 2) & 1] = true; }
        if (e.in_ring) {
            CUDA_TRY(cudaEventRecord(evRing_[e.slot], stream_));        // the ring slot of this frame may be refilled
            ring_ev_set_[e.slot] = true;
        }
    }
    if (pipe) {
        CUDA_TRY(cudaEventRecord(evOutReady_[oslot], stream_));
        CUDA_TRY(cudaStreamWaitEvent(sO_, evOutReady_[oslot], 0));
        { StageScope tcopy(this, VS_STAGE_D2H, sO_);
          for (int l = 0; l < n_lanes_; ++l) {
            if (out_stride == tight && dstride == tight) CUDA_TRY(cudaMemcpyAsync(outs[l], dst.p[l], tight * (size_t)h, cudaMemcpyDeviceToHost, sO_));
            else CUDA_TRY(cudaMemcpy2DAsync(outs[l], out_stride, dst.p[l], dstride, tight, h, cudaMemcpyDeviceToHost, sO_));
          } }
        CUDA_TRY(cudaEventRecord(evOutFree_[oslot], sO_));
        out_free_set_[oslot] = true;
    } else if (host_io) {
        for (int l = 0; l < n_lanes_; ++l) {
            if (out_stride == tight && dstride == tight) CUDA_TRY(cudaMemcpyAsync(outs[l], dst.p[l], tight * (size_t)h, cudaMemcpyDeviceToHost, stream_));
            else CUDA_TRY(cudaMemcpy2DAsync(outs[l], out_stride, dst.p[l], dstride, tight, h, cudaMemcpyDeviceToHost, stream_));
        }
        CUDA_TRY(cudaStreamSynchronize(stream_));
    }
    ++n_out_;
    return VS_OK;
}

// Stabilizer::stabilize, Stabilizer.cpp:258-392
vs_status Engine::push(const uint8_t* const* frames, int w, int h, size_t stride, uint8_t* const* outs, size_t out_stride,
                       size_t out_capacity, unsigned flags, int io, int* ow, int* oh, int* produced) {
    const bool host_io = io != VS_IO_DEVICE, pipe = io == VS_IO_HOST_PIPE && multi_;
    *produced = 0;
    if (pipe && !sH_) {
        CUDA_TRY(cudaStreamCreateWithFlags(&sH_, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&sO_, cudaStreamNonBlocking));
    }
    if (!frames || w <= 0 || h <= 0) return VS_OK;                                    // frame.empty() -> empty Mat (:263)
    if (w < 4 || h < 4) return vs_set_error(VS_ERR_INVALID_ARG, "frame too small");
    if (stride == 0) stride = (size_t)w * 3;
    if (stride < (size_t)w * 3) return vs_set_error(VS_ERR_INVALID_ARG, "stride smaller than a row");
    CUDA_TRY(cudaSetDevice(device_));
    const bool borrow = !host_io && (flags & VS_PUSH_BORROW);
    VS_TRY(ensure_geometry(w, h, !borrow, false, false));
    if (!first_) {
        // will this call emit a frame?  Known from counts alone (:383), except with adaptive_smoothing, where the gate
        // moves with the data: there the buffer must be good whenever the smallest possible gate (5) could open.
        const int gate = p_.adaptive_smoothing ? 5 : clampi(smoothing_radius_, 5, 35);
        if ((int)queue_.size() + 1 >= gate) VS_TRY(check_out_buffer(false, outs, out_stride, out_capacity));
    }

    QueueEntry e;
    e.index = next_index_;
    e.slot = next_index_ % ring_slots_;
    e.frames.resize(n_lanes_);
    if (borrow) {
        for (int l = 0; l < n_lanes_; ++l) e.frames[l] = frames[l];
        e.stride = stride;
    } else {
        const size_t tight = (size_t)w * 3;
        // pipelined host I/O: copy in on the copy-in stream, after the warp that last read this ring slot
        // the ring slot is refilled only after the warp that last read it (another stream) has finished
        cudaStream_t cs = pipe ? sH_ : sp();
        if (multi_ && ring_ev_set_[e.slot]) CUDA_TRY(cudaStreamWaitEvent(cs, evRing_[e.slot], 0));
        StageScope tcopy(this, VS_STAGE_H2D, cs);
        for (int l = 0; l < n_lanes_; ++l) {
            uint8_t* dst = d_ring_ + ((size_t)l * ring_slots_ + e.slot) * frame_bytes_;
            const cudaMemcpyKind kind = host_io ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
            if (stride == tight) CUDA_TRY(cudaMemcpyAsync(dst, frames[l], frame_bytes_, kind, cs));      // one flat transfer, not 1080 rows
            else CUDA_TRY(cudaMemcpy2DAsync(dst, tight, frames[l], stride, tight, h, kind, cs));
            e.frames[l] = dst;
        }
        e.stride = tight;
        e.in_ring = true;
        if (pipe) {
            cudaEvent_t ev = evH_[h_seq_++ & 7];
            CUDA_TRY(cudaEventRecord(ev, sH_));
            CUDA_TRY(cudaStreamWaitEvent(sp(), ev, 0));
        }
    }

    if (first_) {
        // first frame: 480x270 analysis image + GFTT with the user's parameters (:271-368)
        PtrPack src;
        for (int l = 0; l < n_lanes_; ++l) src.p[l] = e.frames[l];
        VS_TRY(first_frame_detect(src, w, h, e.stride));
        queue_.push_back(e);
        first_ = false;
        next_index_ = 1;
        if (host_io && !pipe) CUDA_TRY(cudaStreamSynchronize(sp()));     // the caller may reuse its frame buffer
        return VS_OK;
    }
    queue_.push_back(e);
    bool will_pop = false;
    VS_TRY(generate_transform(e, &will_pop));
    if (will_pop) {
        VS_TRY(emit(outs, out_stride, out_capacity, io, ow, oh));
        *produced = 1;
    } else if (host_io && !pipe) {
        CUDA_TRY(cudaStreamSynchronize(sp()));        // the caller may reuse its frame buffer
    }
    ++next_index_;
    return VS_OK;
}

// Stabilizer::flush, Stabilizer.cpp:394-400
vs_status Engine::flush(uint8_t* const* outs, size_t out_stride, size_t out_capacity, int io, int* ow, int* oh,
                        int* produced) {
    *produced = 0;
    if (queue_.empty()) return VS_OK;
    CUDA_TRY(cudaSetDevice(device_));
    VS_TRY(check_out_buffer(queue_.front().index >= n_frames_, outs, out_stride, out_capacity));
    VS_TRY(setup_slot_guard());
    launch_smooth_only(d_lanes_, n_lanes_, step_info(queue_.front().index), sm());
    launches_ += 1;
    VS_TRY(setup_ready());
    VS_TRY(emit(outs, out_stride, out_capacity, io, ow, oh));
    *produced = 1;
    return VS_OK;
}

// Pipelined host I/O: n frames in, up to n frames out per call.  Frame k+1's host->device copy, frame k's
// kernels and frame k-1's device->host copy overlap on the copy-in / compute / copy-out streams; the call returns
// when every output of this call is in host memory.  Results are identical to n calls of push().
vs_status Engine::push_many(const uint8_t* frames, size_t frame_step, int n, int w, int h, size_t stride, uint8_t* outs,
                            size_t out_stride, size_t out_frame_capacity, int* ow, int* oh, int* n_produced) {
    *n_produced = 0;
    if (n_lanes_ != 1) return vs_set_error(VS_ERR_INVALID_ARG, "push_many is a single-stream call");
    vs_status rc = VS_OK;
    for (int k = 0; k < n && rc == VS_OK; ++k) {
        const uint8_t* f = frames + (size_t)k * frame_step;
        uint8_t* o = outs + (size_t)(*n_produced) * out_frame_capacity;
        int produced = 0;
        rc = push(&f, w, h, stride, &o, out_stride, out_frame_capacity, 0, VS_IO_HOST_PIPE, ow, oh, &produced);
        *n_produced += produced;
    }
    vs_status rs = sync();
    return rc != VS_OK ? rc : rs;
}

// the same loop over device-resident frames; asynchronous like push_device (no synchronisation at the end)
vs_status Engine::push_many_device(const uint8_t* d_frames, size_t frame_step, int n, int w, int h, size_t stride, uint8_t* d_outs,
                                   size_t out_stride, size_t out_frame_capacity, unsigned flags, int* ow, int* oh, int* n_produced) {
    *n_produced = 0;
    if (n_lanes_ != 1) return vs_set_error(VS_ERR_INVALID_ARG, "push_many_device is a single-stream call");
    vs_status rc = VS_OK;
    for (int k = 0; k < n && rc == VS_OK; ++k) {
        const uint8_t* f = d_frames + (size_t)k * frame_step;
        uint8_t* o = d_outs + (size_t)(*n_produced) * out_frame_capacity;
        int produced = 0;
        rc = push(&f, w, h, stride, &o, out_stride, out_frame_capacity, flags, VS_IO_DEVICE, ow, oh, &produced);
        *n_produced += produced;
    }
    return rc;
}

vs_status Engine::flush_many(uint8_t* outs, size_t out_stride, size_t out_frame_capacity, int max_frames, int* ow, int* oh,
                             int* n_produced) {
    *n_produced = 0;
    if (n_lanes_ != 1) return vs_set_error(VS_ERR_INVALID_ARG, "flush_many is a single-stream call");
    vs_status rc = VS_OK;
    while (*n_produced < max_frames && !queue_.empty() && rc == VS_OK) {
        // one (width, height) per call: stop before a frame of another size (the clip's last frame comes back
        // un-warped, i.e. without the copyMakeBorder margin - Stabilizer.cpp:774-780)
        const bool passthrough = queue_.front().index >= n_frames_;
        const int b = (!passthrough && p_.border_size > 0 && !p_.crop_n_zoom && !vc_on_) ? p_.border_size : 0;
        if (*n_produced > 0 && (W_ + 2 * b != *ow || H_ + 2 * b != *oh)) break;
        uint8_t* o = outs + (size_t)(*n_produced) * out_frame_capacity;
        int produced = 0;
        rc = flush(&o, out_stride, out_frame_capacity, VS_IO_HOST_PIPE, ow, oh, &produced);
        *n_produced += produced;
    }
    vs_status rs = sync();
    return rc != VS_OK ? rc : rs;
}

vs_status Engine::reset_detect_counters() {
    CUDA_TRY(cudaMemsetAsync(d_detect_counters_, 0, sizeof(unsigned int) * 2 * n_lanes_, stream_));   // single-kernel entry points
    return VS_OK;
}

// ------------------------------------------------------------------------------ diagnostics
vs_status Engine::frame_record(int lane, int i, vs_frame_record* r) {
    if (lane < 0 || lane >= n_lanes_ || i < 0 || i >= n_frames_ || !r) return vs_set_error(VS_ERR_INVALID_ARG, "bad record index");
    VS_TRY(sync());
    CUDA_TRY(cudaMemcpy(r, h_lanes_[lane].frec + i, sizeof(*r), cudaMemcpyDeviceToHost));
    return VS_OK;
}
vs_status Engine::output_record(int lane, int i, vs_output_record* r) {
    if (lane < 0 || lane >= n_lanes_ || i < 0 || i >= n_out_ || !r) return vs_set_error(VS_ERR_INVALID_ARG, "bad record index");
    VS_TRY(sync());
    CUDA_TRY(cudaMemcpy(r, h_lanes_[lane].orec + i, sizeof(*r), cudaMemcpyDeviceToHost));
    return VS_OK;
}
vs_status Engine::frame_points(int lane, int i, float* prev, float* next, uint8_t* status, uint8_t* mask, float* det) {
    if (lane < 0 || lane >= n_lanes_ || i < 0 || i >= n_frames_) return vs_set_error(VS_ERR_INVALID_ARG, "bad record index");
    if (i < n_frames_ - log_depth_) return vs_set_error(VS_ERR_INVALID_ARG, "point log of that frame has been recycled");
    vs_frame_record r;
    VS_TRY(frame_record(lane, i, &r));
    const LaneDev& L = h_lanes_[lane];
    size_t o = (size_t)(i % log_depth_) * kp_cap_;
    if (prev && r.n_prev_pts > 0) CUDA_TRY(cudaMemcpy(prev, L.log_prev + o, sizeof(float2) * r.n_prev_pts, cudaMemcpyDeviceToHost));
    if (next && r.n_prev_pts > 0) CUDA_TRY(cudaMemcpy(next, L.log_next + o, sizeof(float2) * r.n_prev_pts, cudaMemcpyDeviceToHost));
    if (status && r.n_prev_pts > 0) CUDA_TRY(cudaMemcpy(status, L.log_status + o, r.n_prev_pts, cudaMemcpyDeviceToHost));
    if (mask && r.n_tracked > 0 && r.n_inliers >= 0) CUDA_TRY(cudaMemcpy(mask, L.log_mask + o, r.n_tracked, cudaMemcpyDeviceToHost));
    if (det && r.n_detected > 0) CUDA_TRY(cudaMemcpy(det, L.log_detected + o, sizeof(float2) * r.n_detected, cudaMemcpyDeviceToHost));
    return VS_OK;
}
vs_status Engine::first_corners(int lane, float* xy, int cap, int* n) {
    if (lane < 0 || lane >= n_lanes_ || !n) return vs_set_error(VS_ERR_INVALID_ARG, "bad lane");
    VS_TRY(sync());
    int c = 0;
    CUDA_TRY(cudaMemcpy(&c, h_lanes_[lane].first_count, sizeof(int), cudaMemcpyDeviceToHost));
    *n = c;
    if (xy && c > 0) CUDA_TRY(cudaMemcpy(xy, h_lanes_[lane].first_corners, sizeof(float2) * (c < cap ? c : cap), cudaMemcpyDeviceToHost));
    return VS_OK;
}

// ------------------------------------------------------------------------------ offline clip mode
// Motion estimation is pairwise-local: transform n needs gray(n-1), gray(n) and the
// corners detected on the latest even frame <= n-1 (or the first-frame corners while n <= 2).  A chunk
// starting at `first` therefore needs at most two leading halo frames.
int Engine::chunk_halo(int first) {
    if (first <= 2) return first;                    // start from frame 0 (first-frame quirks B-Q1 included)
    int m = (first - 1) & ~1;                        // frame whose corners transform `first` starts from
    return first - m;                                // 1 or 2
}

vs_status Engine::analyze_chunk(const uint8_t* d_frames, int w, int h, int first, int count, float* out, bool device_out, int* n_out) {
    if (n_lanes_ != 1) return vs_set_error(VS_ERR_INVALID_ARG, "clip mode uses single-lane handles");
    if (p_.adaptive_smoothing) return vs_set_error(VS_ERR_UNSUPPORTED, "adaptive_smoothing is not available in clip mode");
    if (p_.drone_high_freq_mode) return vs_set_error(VS_ERR_UNSUPPORTED, "drone_high_freq_mode filters carry state from frame to frame; not available in clip mode");
    if (vc_on_) return vs_set_error(VS_ERR_UNSUPPORTED, "enable_virtual_canvas keeps a temporal frame buffer; not available in clip mode");
    if (!d_frames || first < 0 || count <= 0) return vs_set_error(VS_ERR_INVALID_ARG, "bad chunk");
    CUDA_TRY(cudaSetDevice(device_));
    VS_TRY(clean());
    VS_TRY(ensure_geometry(w, h, false, false, false));
    while (first + count + 2 >= traj_cap_) VS_TRY(grow_trajectory());
    const size_t fb = frame_bytes_, tight = (size_t)w * 3;
    const int halo = chunk_halo(first);
    const uint8_t* base = d_frames;                  // frame (first - halo)
    auto entry = [&](int frame_index) {
        QueueEntry e;
        e.index = frame_index; e.slot = 0; e.stride = tight;
        e.frames.assign(1, base + (size_t)(frame_index - (first - halo)) * fb);
        return e;
    };
    int f = first - halo;                            // next frame to feed
    if (f == 0) {
        // frame 0: first-frame analysis (Stabilizer.cpp:271-368)
        PtrPack src; src.p[0] = entry(0).frames[0];
        VS_TRY(first_frame_detect(src, w, h, tight));
        first_ = false;
        n_frames_ = 0; detect_counter_ = 0;
        f = 1;
    } else {
        // halo: corners from the even frame m, pyramid of frame first-1
        const int m = f;
        PtrPack src; src.p[0] = entry(m).frames[0];
        launch_gray_resize(d_lanes_, 1, src, w, h, tight, m % VS_PYR_SLOTS, sp());
        if (multi_) CUDA_TRY(cudaEventRecord(evG_, sp()));
        launch_pyrdown(d_lanes_, 1, m % VS_PYR_SLOTS, sp());
        launches_ += 2;
        VS_TRY(redetect(m % VS_PYR_SLOTS, m, 0, evG_));
        if (first - 1 > m) {
            PtrPack s2; s2.p[0] = entry(first - 1).frames[0];
            launch_gray_resize(d_lanes_, 1, s2, w, h, tight, (first - 1) % VS_PYR_SLOTS, sp());
            launch_pyrdown(d_lanes_, 1, (first - 1) % VS_PYR_SLOTS, sp());
            launches_ += 2;
        }
        first_ = false;
        n_frames_ = first - 1; detect_counter_ = first - 1;      // the counters equal the frame number
        f = first;
    }
    const int first_tr = f;                          // first generateTransform call computed here
    for (; f < first + count; ++f) {
        bool pop = false;
        QueueEntry e = entry(f);
        VS_TRY(generate_transform(e, &pop));         // queue_ is empty: never pops
    }
    const int n = first + count - first_tr;
    if (n_out) *n_out = n;
    if (n > 0 && out)
        CUDA_TRY(cudaMemcpyAsync(out, h_lanes_[0].transforms + 3 * (size_t)(first_tr - 1), sizeof(float) * 3 * n,
                                 device_out ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, sm()));
    if (device_out) return join();                   // asynchronous: ordered on the public stream, no host block
    VS_TRY(sync());
    return VS_OK;
}

vs_status Engine::analyze_chunks_lockstep(const uint8_t* const* d_frames, int w, int h, int count, float* const* d_out) {
    if (p_.adaptive_smoothing) return vs_set_error(VS_ERR_UNSUPPORTED, "adaptive_smoothing is not available in clip mode");
    if (p_.drone_high_freq_mode) return vs_set_error(VS_ERR_UNSUPPORTED, "drone_high_freq_mode filters carry state from frame to frame; not available in clip mode");
    if (vc_on_) return vs_set_error(VS_ERR_UNSUPPORTED, "enable_virtual_canvas keeps a temporal frame buffer; not available in clip mode");
    if (!d_frames || !d_out || count <= 0) return vs_set_error(VS_ERR_INVALID_ARG, "bad chunk");
    for (int l = 0; l < n_lanes_; ++l)
        if (!d_frames[l] || !d_out[l]) return vs_set_error(VS_ERR_INVALID_ARG, "null chunk pointer");
    CUDA_TRY(cudaSetDevice(device_));
    VS_TRY(clean());
    VS_TRY(ensure_geometry(w, h, false, false, false));
    // every lane runs with the same LOCAL frame numbers: its chunk starts at `first` = 4 (even, past the first-frame
    // quirks), so slots, detection parity and record indices coincide across lanes
    const int first = 4, halo = 2;
    while (first + count + 2 >= traj_cap_) VS_TRY(grow_trajectory());
    const size_t fb = frame_bytes_, tight = (size_t)w * 3;
    auto entry = [&](int frame_index) {
        QueueEntry e;
        e.index = frame_index; e.slot = 0; e.stride = tight;
        e.frames.resize(n_lanes_);
        for (int l = 0; l < n_lanes_; ++l) e.frames[l] = d_frames[l] + (size_t)(frame_index - (first - halo)) * fb;
        return e;
    };
    {
        // halo: corners from the even frame m = first - 2, pyramid of frame first - 1
        const int m = first - halo;
        QueueEntry em = entry(m), e1 = entry(first - 1);
        PtrPack src, s2;
        for (int l = 0; l < n_lanes_; ++l) { src.p[l] = em.frames[l]; s2.p[l] = e1.frames[l]; }
        launch_gray_resize(d_lanes_, n_lanes_, src, w, h, tight, m % VS_PYR_SLOTS, sp());
        if (multi_) CUDA_TRY(cudaEventRecord(evG_, sp()));
        launch_pyrdown(d_lanes_, n_lanes_, m % VS_PYR_SLOTS, sp());
        launches_ += 2;
        VS_TRY(redetect(m % VS_PYR_SLOTS, m, 0, evG_));
        launch_gray_resize(d_lanes_, n_lanes_, s2, w, h, tight, (first - 1) % VS_PYR_SLOTS, sp());
        launch_pyrdown(d_lanes_, n_lanes_, (first - 1) % VS_PYR_SLOTS, sp());
        launches_ += 2;
        first_ = false;
        n_frames_ = first - 1; detect_counter_ = first - 1;      // the counters equal the (local) frame number
    }
    for (int f = first; f < first + count; ++f) {
        bool pop = false;
        QueueEntry e = entry(f);
        VS_TRY(generate_transform(e, &pop));         // queue_ is empty: never pops
    }
    for (int l = 0; l < n_lanes_; ++l)
        CUDA_TRY(cudaMemcpyAsync(d_out[l], h_lanes_[l].transforms + 3 * (size_t)(first - 1), sizeof(float) * 3 * count,
                                 cudaMemcpyDeviceToDevice, sm()));
    return join();                                   // asynchronous: ordered on the public stream, no host block
}

vs_status Engine::render_chunk(const float* all_tr, bool device_in, int n_total, const uint8_t* d_frames, int w, int h, int first,
                               int count, uint8_t* d_out, int* ow, int* oh) {
    VS_TRY(set_clip_transforms(all_tr, device_in, n_total, w, h));
    return render_prepared(d_frames, w, h, first, count, d_out, ow, oh, !device_in);
}

// The transform list of the whole clip -> transforms_ and path_ (the reference's sequential float32 running sum) of this
// handle.  Done once per clip; render_prepared() then smooths and warps any number of chunks against it.
vs_status Engine::set_clip_transforms(const float* all_tr, bool device_in, int n_total, int w, int h) {
    if (n_lanes_ != 1) return vs_set_error(VS_ERR_INVALID_ARG, "clip mode uses single-lane handles");
    if (p_.adaptive_smoothing) return vs_set_error(VS_ERR_UNSUPPORTED, "adaptive_smoothing is not available in clip mode");
    if (p_.drone_high_freq_mode) return vs_set_error(VS_ERR_UNSUPPORTED, "drone_high_freq_mode filters carry state from frame to frame; not available in clip mode");
    if (vc_on_) return vs_set_error(VS_ERR_UNSUPPORTED, "enable_virtual_canvas keeps a temporal frame buffer; not available in clip mode");
    if (!all_tr || n_total < 1) return vs_set_error(VS_ERR_INVALID_ARG, "bad clip");
    CUDA_TRY(cudaSetDevice(device_));
    if (device_in) {
        // asynchronous variant: everything below is enqueued on the public stream behind whatever the handle still has
        // in flight (the analysis of this chunk); the host state is reset without the blocking sync of clean()
        VS_TRY(join());
        queue_.clear();
        for (bool& bb : evB_set_) bb = false;
        for (bool& bb : evA_set_) bb = false;
        for (bool& bb : evW_set_) bb = false;
        for (bool& bb : c_pending_) bb = false;
        last_detect_frame_ = -100;
        first_ = true; next_index_ = 0; n_frames_ = 0; n_out_ = 0; detect_counter_ = 0;
    } else {
        VS_TRY(clean());
    }
    const int b = p_.border_size;
    const int mode = b <= 0 ? 0 : (p_.crop_n_zoom ? ((w - 2 * b > 0 && h - 2 * b > 0) ? 2 : 0) : 1);
    VS_TRY(ensure_geometry(w, h, false, false, mode == 2));
    while (n_total + 2 >= traj_cap_) VS_TRY(grow_trajectory());
    const int n_tr = n_total - 1;
    if (n_tr > 0) {
        CUDA_TRY(cudaMemcpyAsync(h_lanes_[0].transforms, all_tr, sizeof(float) * 3 * n_tr,
                                 device_in ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream_));
        launch_traj_build(d_lanes_, 1, n_tr, stream_);
        launches_ += 1;
    }
    n_frames_ = n_tr;
    clip_total_ = n_total;
    kalman_next_ = 0;
    return VS_OK;
}

vs_status Engine::render_prepared(const uint8_t* d_frames, int w, int h, int first, int count, uint8_t* d_out, int* ow, int* oh,
                                  bool host_sync) {
    const int n_total = clip_total_;
    if (n_lanes_ != 1 || n_total < 1) return vs_set_error(VS_ERR_INVALID_ARG, "no clip transforms set on this handle");
    if (!d_frames || !d_out || first < 0 || count <= 0 || first + count > n_total)
        return vs_set_error(VS_ERR_INVALID_ARG, "bad chunk");
    CUDA_TRY(cudaSetDevice(device_));
    const int b = p_.border_size;
    const int mode = b <= 0 ? 0 : (p_.crop_n_zoom ? ((w - 2 * b > 0 && h - 2 * b > 0) ? 2 : 0) : 1);
    if (w != W_ || h != H_) return vs_set_error(VS_ERR_INVALID_ARG, "frame size differs from the clip's");
    if (count > wp_batch_cap_) {
        // (stream-ordered: the warp of the previous chunk may still be reading the old array)
        if (d_wp_batch_) CUDA_TRY(cudaFreeAsync(d_wp_batch_, stream_));
        CUDA_TRY(cudaMallocAsync((void**)&d_wp_batch_, sizeof(WarpParams) * count, stream_));
        wp_batch_cap_ = count;
    }
    const int gate = clampi(smoothing_radius_, 5, 35);
    const int ow_ = mode == 1 ? w + 2 * b : w, oh_ = mode == 1 ? h + 2 * b : h;
    *ow = ow_; *oh = oh_;
    const size_t tight_in = (size_t)w * 3, tight_out = (size_t)ow_ * 3, oframe = tight_out * oh_;
    // frames with a transform are warped in one batched launch; the clip's last frame has none and is
    // passed through unchanged (Stabilizer.cpp:774-780)
    const int n_warp = (first + count == n_total) ? count - 1 : count;
    if (n_warp > 0) {
        StepInfo base = step_info(0);
        // (Kalman: the recursion continues where the previous chunk of this clip left it, else it restarts at frame 0)
        const int kal_from = (first == kalman_next_ && first > 0) ? first : 0;
        launch_smooth_batch(d_lanes_, 1, base, first, n_warp, n_total, gate, d_wp_batch_, kal_from, stream_);
        kalman_next_ = first + n_warp;
        launches_ += 1 + launch_warp_frames_mode(d_frames, w, h, tight_in, frame_bytes_, d_out, tight_out, oframe, d_wp_batch_, n_warp,
                                                 mode, b, border_mode_, d_scratch_, stream_);
    }
    if (n_warp < count) {
        // un-warped original frame, written at the top-left of its output slot (smaller than a bordered frame)
        CUDA_TRY(cudaMemcpy2DAsync(d_out + oframe * n_warp, tight_out, d_frames + frame_bytes_ * n_warp, tight_in, tight_in, h,
                                   cudaMemcpyDeviceToDevice, stream_));
    }
    n_out_ = n_warp;
    if (host_sync) CUDA_TRY(cudaStreamSynchronize(stream_));
    return VS_OK;
}
