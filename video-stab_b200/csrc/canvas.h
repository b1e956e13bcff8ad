// canvas.h — "virtual canvas" output stage of vs::Stabilizer (reference src/Stabilizer.cpp:1129-1134, 2066-2443).
//
// With enable_virtual_canvas the reference discards its warped frame and returns instead the ORIGINAL frame moved by the
// integer part of the correction (the frame is pasted in the middle of a larger black canvas, a frame-sized window is cut
// out at centre - (dx, dy)), after blending older frames into the canvas' empty (gray <= 1) regions:
//   * the empty regions are the bounding rectangles (area > 100) of the external contours of the mask "gray <= 1"
//     (cv::findContours, :2224-2241);
//   * a region is filled from the most recent older frame of the temporal buffer whose motion-compensated copy covers more
//     than half of it (:2244-2273, :2401-2421), warped by cv::warpAffine(BORDER_REFLECT) with the relative motion
//     (:2423-2443), cut out, resized to the region (:2318-2350) and alpha-blended with a linear edge ramp (:2352-2399).
// Here nothing canvas-sized is ever written: one kernel computes every output pixel from the frame and, inside regions, from
// the taps of the older frame (warp, resize and blend composed per pixel, in the reference's fixed-point / float32 steps).
// When the canvas' black surround encloses the frame it is the only external contour and the one region is the whole canvas: no
// contour work at all (and with a canvas of at least twice the frame's area nothing can ever be filled: the stage is then a
// whole-pixel shift read from the device, fully asynchronous).  Otherwise the contour logic runs on the host as in the reference
// (autozoom_host.h), and only on frames that contain dark pixels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <deque>
#include <vector>
#include "../../include/vstab_b200.h"

#define VC_MAX_REGIONS 24          // regions per launch (the parameter block must stay below 4 KB); more regions = more launches

struct VcRegion {
    int x, y, w, h;                // the region, canvas coordinates
    int ix, iy, iw, ih;            // the part of the motion-compensated older frame that fills it (frame coordinates)
    double m[6];                   // inverted compensation matrix (cv::warpAffine's M)
    double sx, sy;                 // cv::resize scales iw -> w, ih -> h
    const uint8_t* src;            // the older frame (tight rows)
    float weight;
    int edge;                      // edge ramp width in pixels
    int resize;                    // 1: (iw, ih) != (w, h)
    int pad;
};
struct VcParams {
    const uint8_t* frame;          // the frame being emitted
    uint8_t* out;
    size_t frame_stride, out_stride;
    int W, H;                      // frame = output size
    int fx0, fy0;                  // where the frame sits on the canvas
    int ex, ey;                    // where the output window sits on the canvas
    int n_regions;
    int from_out;                  // 1: continue from the pixels already in `out` (regions beyond the first VC_MAX_REGIONS)
    VcRegion r[VC_MAX_REGIONS];
};

struct VcRect { int x = 0, y = 0, w = 0, h = 0; };

class VirtualCanvas {
public:
    ~VirtualCanvas();
    void configure(const vs_params& p) { p_ = p; }
    // one emitted frame: `d_frame` (W x H, `stride`) with correction (dx, dy, da) -> `d_out` (W x H).  `recent` are the last
    // (at most 30) frame-to-frame transforms (x, y, angle), needed once, when the canvas is first sized.
    // Synchronises `st` (the contour logic needs the dark-pixel count, as the reference's needs its mask).
    vs_status apply(const uint8_t* d_frame, int W, int H, size_t stride, const float T[3], const float* recent, int n_recent,
                    uint8_t* d_out, size_t out_stride, cudaStream_t st, int* launches);
    // the same for a canvas that can never be filled (never_fills()): asynchronous, the correction is read on the device
    vs_status apply_async(const uint8_t* d_frame, int W, int H, size_t stride, const struct WarpParams* d_wp, uint8_t* d_out,
                          size_t out_stride, cudaStream_t st, int* launches);
    bool never_fills() const { return sized_ && never_fills_; }
    bool geometry(int W, int H) const { return W == W_ && H == H_; }
    void reset();                  // geometry change
    float scale() const { return scale_; }
    bool sized() const { return sized_; }
    int regions_last() const { return regions_last_; }

private:
    vs_status ensure(int W, int H);
    void contour_regions(const uint8_t* h_mask, bool any_dark, std::vector<VcRect>& out);
    vs_params p_{};
    int W_ = 0, H_ = 0, cw_ = 0, ch_ = 0;
    float scale_ = 0.f, cx_ = 0.f, cy_ = 0.f;
    bool sized_ = false, never_fills_ = false;
    // temporal buffer: device ring of frames + their corrections
    uint8_t* d_ring_ = nullptr;
    int ring_slots_ = 0, ring_next_ = 0;
    struct Entry { int slot; float T[3]; };
    std::deque<Entry> buf_;
    uint8_t* d_mask_ = nullptr;    // W x H, 1 = dark
    int* d_count_ = nullptr;
    uint8_t* h_mask_ = nullptr;    // page-locked
    int* h_count_ = nullptr;
    std::vector<signed char> work_;
    std::vector<VcRect> ring_only_;
    bool ring_only_valid_ = false;
    int regions_last_ = 0;
};
