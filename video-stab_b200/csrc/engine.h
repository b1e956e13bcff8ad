// engine.h — host side of the stabilizer: the reference's stabilize()/flush()/clean() control flow
// (Stabilizer.cpp:258-400) re-expressed as an asynchronous launch sequence on seven CUDA streams per handle
// (pyramid, two tracking, motion, two detection, output) joined by events and slot rings; see
// Engine::generate_transform and DESIGN.md section 5.
// One Engine advances n_lanes independent streams in lock-step (n_lanes == 1 is vs::Stabilizer).
#pragma once
#include <deque>
#include <string>
#include <vector>

#include "canvas.h"
#include "kernels.h"

enum { VS_IO_DEVICE = 0, VS_IO_HOST_SYNC = 1, VS_IO_HOST_PIPE = 2 };   // where push()/flush() frames live
#define VS_TRACK_STREAMS 4                // tracking streams created; LK of frame n runs on stream n % track_n_
#define VS_OUT_SLOTS 3                    // output staging frames of the pipelined host path
enum { VS_STAGE_GRAY = 0, VS_STAGE_PYRDOWN, VS_STAGE_LK, VS_STAGE_MOTION, VS_STAGE_GFTT, VS_STAGE_WARP, VS_STAGE_H2D, VS_STAGE_D2H, VS_N_STAGES };

struct QueueEntry {
    int index;                               // frameIndexQueue_
    int slot;                                // ring slot (copy mode)
    bool in_ring = false;                    // the frame lives in the device ring (not borrowed)
    std::vector<const uint8_t*> frames;      // per-lane device pointer of the queued frame
    size_t stride;
};

class Engine {
public:
    static vs_status create(const vs_params& p, int device, int n_lanes, Engine** out);
    ~Engine();

    // frames/outs: n_lanes pointers.  io: VS_IO_DEVICE (device pointers, asynchronous), VS_IO_HOST_SYNC (host
    // memory, copies + sync inside) or VS_IO_HOST_PIPE (host memory, copies on their own streams, no sync).
    vs_status push(const uint8_t* const* frames, int w, int h, size_t stride, uint8_t* const* outs, size_t out_stride,
                   size_t out_capacity, unsigned flags, int io, int* ow, int* oh, int* produced);
    vs_status flush(uint8_t* const* outs, size_t out_stride, size_t out_capacity, int io, int* ow, int* oh,
                    int* produced);
    vs_status push_many(const uint8_t* frames, size_t frame_step, int n, int w, int h, size_t stride, uint8_t* outs,
                        size_t out_stride, size_t out_frame_capacity, int* ow, int* oh, int* n_produced);
    vs_status push_many_device(const uint8_t* d_frames, size_t frame_step, int n, int w, int h, size_t stride, uint8_t* d_outs,
                               size_t out_stride, size_t out_frame_capacity, unsigned flags, int* ow, int* oh, int* n_produced);
    vs_status flush_many(uint8_t* outs, size_t out_stride, size_t out_frame_capacity, int max_frames, int* ow, int* oh,
                         int* n_produced);
    vs_status clean();
    vs_status sync();
    vs_status join();       // public stream waits for the analysis and detection streams

    cudaStream_t stream() const { return stream_; }
    uint64_t launches() const { return launches_; }
    int n_lanes() const { return n_lanes_; }
    int n_frame_records() const { return n_frames_; }
    int n_output_records() const { return n_out_; }
    vs_status frame_record(int lane, int i, vs_frame_record* r);
    vs_status output_record(int lane, int i, vs_output_record* r);
    vs_status frame_points(int lane, int i, float* prev, float* next, uint8_t* status, uint8_t* mask, float* det);
    vs_status first_corners(int lane, float* xy, int cap, int* n);

    // ---- offline clip mode (temporal chunk of a long clip; BASELINE config 5) -------------------------
    // frames before `first` that analyze_chunk needs (the pixel halo: <= 2 frames, or `first` itself if <= 2)
    static int chunk_halo(int first);
    // d_frames: frames [first - chunk_halo(first), first + count), tight rows.  Writes transforms_[n-1] for
    // the generateTransform calls n = max(first,1) .. first+count-1 to out_host (3 floats each).
    // `out` is host memory (the call returns when it is filled) or, with device_out, device memory (asynchronous: the
    // copy is ordered on the public stream, see join()).
    vs_status analyze_chunk(const uint8_t* d_frames, int w, int h, int first, int count, float* out, bool device_out, int* n_out);
    // Lock-step analysis of n_lanes() chunks at once (one launch per stage for all of them): lane l analyses `count` frames
    // starting at an EVEN frame number first_l >= 4 of the clip; d_frames[l] points at frame first_l - 2 (the halo), tight
    // rows, count + 2 frames.  Writes the `count` transforms of generateTransform calls first_l .. first_l + count - 1 of
    // lane l to d_out[l] (device memory, 3 floats each).  Asynchronous like analyze_chunk(device_out).  The frame numbers
    // themselves are not needed: motion estimation is pairwise-local, only the parity of the chunk start matters.
    vs_status analyze_chunks_lockstep(const uint8_t* const* d_frames, int w, int h, int count, float* const* d_out);
    // all_tr_host: the n_total-1 transforms of the whole clip.  d_frames: frames [first, first+count).
    // all_tr is host memory (synchronous call) or, with device_in, device memory (asynchronous on the public stream)
    vs_status render_chunk(const float* all_tr, bool device_in, int n_total, const uint8_t* d_frames, int w, int h, int first,
                           int count, uint8_t* d_out, int* ow, int* oh);
    // the two halves of render_chunk: the clip's transform list -> this handle's trajectory (once per clip), then any number
    // of chunks smoothed and warped against it
    vs_status set_clip_transforms(const float* all_tr, bool device_in, int n_total, int w, int h);
    vs_status render_prepared(const uint8_t* d_frames, int w, int h, int first, int count, uint8_t* d_out, int* ow, int* oh,
                              bool host_sync);

    // per-stage CUDA-event timing (off by default; used by bench.py and the profiles)
    void set_timing(bool on);
    void stage_time(int stage, double* ms, long long* n);
    bool timing_on() const { return timing_; }
    cudaEvent_t take_event();
    void add_pending(int stage, cudaEvent_t a, cudaEvent_t b);
    void collect_timing();
    const std::vector<float>& trace() { collect_timing(); return trace_; }

    // every stream that reads caller frames (pyramid, warp, ring / host copy-in) waits for a caller event
    vs_status wait_external(cudaEvent_t ev) {
        if (cudaStreamWaitEvent(stream_, ev, 0) != cudaSuccess) return vs_set_cuda_error(cudaGetLastError(), "cudaStreamWaitEvent", __FILE__, __LINE__);
        if (multi_ && cudaStreamWaitEvent(sP_, ev, 0) != cudaSuccess) return vs_set_cuda_error(cudaGetLastError(), "cudaStreamWaitEvent", __FILE__, __LINE__);
        if (sH_ && cudaStreamWaitEvent(sH_, ev, 0) != cudaSuccess) return vs_set_cuda_error(cudaGetLastError(), "cudaStreamWaitEvent", __FILE__, __LINE__);
        return VS_OK;
    }

    // single-op helpers behind the vs_k_* entry points (lane 0 scratch)
    const LaneDev* d_lanes() const { return d_lanes_; }
    const LaneDev& h_lane(int i) const { return h_lanes_[i]; }
    vs_status reset_detect_counters();
    int kp_capacity() const { return kp_cap_; }
    bool tracker_uses_tma() const { return lk_tma_; }

private:
    Engine() = default;
    vs_status init(const vs_params& p, int device, int n_lanes);
    vs_status alloc_fixed();
    vs_status alloc_analysis(int aw, int ah);
    vs_status ensure_geometry(int w, int h, bool need_ring, bool need_out, bool need_scratch);
    vs_status grow_trajectory();
    vs_status generate_transform(const QueueEntry& e, bool* will_pop);
    vs_status first_frame_detect(const PtrPack& src, int w, int h, size_t stride);
    vs_status redetect(int cur, int frame_no, int record_frame_no, cudaEvent_t level0_ready);
    cudaStream_t sa(int frame_no) const { return multi_ ? sA_[frame_no % track_n_] : stream_; }
    cudaStream_t sp() const { return multi_ ? sP_ : stream_; }
    cudaStream_t sm() const { return multi_ ? sM_ : stream_; }
    vs_status setup_slot_guard();
    vs_status setup_ready();
    cudaStream_t sc(int gen) const { return multi_ ? sC_[gen & 1] : stream_; }
    vs_status check_out_buffer(bool passthrough, uint8_t* const* outs, size_t out_stride, size_t out_capacity) const;
    vs_status emit(uint8_t* const* outs, size_t out_stride, size_t out_capacity, int io, int* ow, int* oh);
    StepInfo step_info(int pop_index) const;
    void free_all();

    vs_params p_{};
    int device_ = 0, n_lanes_ = 0;
    cudaStream_t stream_ = nullptr;           // public stream: output stage (warp, copies out)
    cudaStream_t sM_ = nullptr;               // motion: RANSAC, trajectory, smoothing, warp set-up
    cudaStream_t sA_[VS_TRACK_STREAMS] = {}, sC_[2] = {}, sP_ = nullptr;   // tracking (LK + motion fit, round-robin), corner detection (two generations), pyramid build
    bool multi_ = false;
    cudaEvent_t evA_[VS_EV_RING] = {}, evB_[VS_EV_RING] = {}, evP_[VS_EV_RING] = {}, evJ_[VS_TRACK_STREAMS + 4] = {}, evS_[2] = {}, evW_[2] = {}, evG_ = nullptr, evC_[VS_KP_SLOTS] = {};
    bool evB_set_[VS_EV_RING] = {}, evA_set_[VS_EV_RING] = {}, evW_set_[2] = {};
    int last_detect_frame_ = -100;
    bool c_pending_[VS_KP_SLOTS] = {};
    bool split_motion_ = true;
    bool lk_tma_ = false;                     // the tracker stages its patches with TMA (tensor maps built for every lane)
    bool lk_tma() const { return lk_tma_; }
    cudaStream_t sH_ = nullptr, sO_ = nullptr; // copy-in / copy-out streams of the pipelined host path
    cudaEvent_t evH_[8] = {}, evRing_[36] = {}, evOutReady_[VS_OUT_SLOTS] = {}, evOutFree_[VS_OUT_SLOTS] = {};
    bool ring_ev_set_[36] = {}, out_free_set_[VS_OUT_SLOTS] = {};
    unsigned h_seq_ = 0;
    int border_mode_ = 0, method_ = 0;
    int smoothing_radius_ = 30;

    // state mirrored from the reference
    bool first_ = true;
    int next_index_ = 0;
    int n_frames_ = 0;        // transforms_.size()
    int n_out_ = 0;
    int detect_counter_ = 0;  // per instance (the reference's is a process-global static)
    std::deque<QueueEntry> queue_;
    int W_ = 0, H_ = 0;

    // device memory
    std::vector<void*> allocs_;
    std::vector<LaneDev> h_lanes_;
    LaneDev* d_lanes_ = nullptr;
    unsigned int* d_detect_counters_ = nullptr;   // [lane][2] = eig_max, cand_count
    float* d_eig_generic_ = nullptr;              // eigenvalue map of the first-frame detection when block_size != 3
    int kp_cap_ = 0, cap_first_ = 0, cap_redetect_ = 0, log_depth_ = 0, traj_cap_ = 0;
    int ring_slots_ = 0;
    size_t frame_bytes_ = 0, out_bytes_ = 0;
    uint8_t* d_ring_ = nullptr;       // [lane][slot][frame]
    uint8_t* d_out_ = nullptr;        // [lane][out frame]  (host-io staging)
    uint8_t* d_scratch_ = nullptr;    // [lane][frame]      (crop+zoom first pass)
    uint8_t* d_fade_ = nullptr;       // border_type "fade": [lane][history] then [lane][blended source], bordered size
    bool fade_ = false, fade_hist_valid_ = false;
    int track_n_ = 2;                 // tracking streams in use (VS_TRACK_N=1..4)
    int aw_ = VS_AW, ah_ = VS_AH;     // analysis size the pyramids are allocated with (alloc_analysis)
    bool vc_on_ = false;              // enable_virtual_canvas (and no crop_n_zoom): canvas.h replaces the warp
    VirtualCanvas canvas_;
    uint8_t* h_vc_wp_ = nullptr;      // page-locked landing block: the output's WarpParams + up to 30 transforms
    int fade_count_ = 0;              // fadeFrameCount_
    int fade_w_ = 0, fade_h_ = 0;
    unsigned char* d_tmaps_ = nullptr;   // tensor-map scratch for the warp kernel (batches of more than 8 lanes)
    WarpParams* d_wp_batch_ = nullptr;
    int wp_batch_cap_ = 0;
    int kalman_next_ = 0;             // clip mode: the Kalman state on the device stands at output frame kalman_next_ - 1
    int clip_total_ = 0;              // frames of the clip whose transforms set_clip_transforms() installed (0: none)
    std::vector<float*> traj_bufs_;
    uint64_t launches_ = 0;

    struct Pending { int stage; cudaEvent_t a, b; };
    bool timing_ = false;
    std::vector<cudaEvent_t> event_pool_;
    std::vector<Pending> pending_;
    std::vector<float> trace_;           // (stage, start us, end us) of the last collected timing batch
    double stage_ms_[VS_N_STAGES] = {};
    long long stage_n_[VS_N_STAGES] = {};
};
