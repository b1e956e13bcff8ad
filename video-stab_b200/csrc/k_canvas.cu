// k_canvas.cu — the virtual-canvas output stage (canvas.h; reference src/Stabilizer.cpp:2066-2443).
//
// k_vc_dark     : the mask "gray <= 1" of the frame (the canvas outside the frame is black by construction) + its pixel count.
//                 HBM-bound: 3 bytes read, 1 written per pixel.
// k_vc_compose  : every output pixel = the frame pixel under the output window (or black), then, for each fill region that
//                 contains it, blended with the region's source pixel: cv::resize(INTER_LINEAR) of the cut-out of
//                 cv::warpAffine(older frame, relative motion, INTER_LINEAR, BORDER_REFLECT), all evaluated per pixel in
//                 OpenCV's fixed-point steps; the blend in float32 without contraction, as the reference's C++ loop.
//                 HBM-bound without regions (3 bytes read + 3 written per pixel); with a canvas-sized region 16 taps of the
//                 older frame per pixel, L1/L2 resident (neighbouring pixels share them).
#include <algorithm>
#include <cmath>
#include <cstring>
#include "autozoom_host.h"
#include "canvas.h"
#include "common.cuh"
#include "kernels.h"

// ---------------------------------------------------------------------------------------------------------------- device
__global__ void __launch_bounds__(256) k_vc_dark(const uint8_t* __restrict__ src, int w, int h, size_t stride, uint8_t* __restrict__ mask,
                                                 int* __restrict__ count) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    int dark = 0;
    if (x < w) {
        const uint8_t* p = src + (size_t)y * stride + 3 * (size_t)x;
        const int g = (3735 * p[0] + 19235 * p[1] + 9798 * p[2] + 16384) >> 15;      // cv::cvtColor BGR2GRAY, :2228
        dark = g <= 1;                                                                // THRESH_BINARY_INV at 1, :2232
        mask[(size_t)y * w + x] = (uint8_t)dark;
    }
    const unsigned b = __ballot_sync(0xffffffffu, dark);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, __popc(b));
}

// cv::borderInterpolate(BORDER_REFLECT)
static __device__ __forceinline__ int vc_reflect(int p, int len) {
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p - 1;
        else p = len - 1 - (p - len);
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

// one pixel of cv::warpAffine(src, M, src.size(), INTER_LINEAR, BORDER_REFLECT) (:2436-2438): 10-bit coordinates, 5-bit fractions,
// weights (32 - ax)(32 - ay) * 32 ... of the 15-bit table, rounding shift
static __device__ __forceinline__ void vc_warp_pixel(const VcRegion& R, int fw, int fh, int u, int v, int out[3]) {
    const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(R.m[1], (double)v), R.m[2]), 1024.0)) + 16;
    const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(R.m[4], (double)v), R.m[5]), 1024.0)) + 16;
    const int ad = __double2int_rn(__dmul_rn(__dmul_rn(R.m[0], (double)u), 1024.0));
    const int bd = __double2int_rn(__dmul_rn(__dmul_rn(R.m[3], (double)u), 1024.0));
    const int X = (X0 + ad) >> 5, Y = (Y0 + bd) >> 5;
    const int sx = min(max(X >> 5, -32768), 32767), sy = min(max(Y >> 5, -32768), 32767);
    const int ax = X & 31, ay = Y & 31;
    const int x0 = vc_reflect(sx, fw), x1 = vc_reflect(sx + 1, fw), y0 = vc_reflect(sy, fh), y1 = vc_reflect(sy + 1, fh);
    const int w00 = (32 - ax) * (32 - ay) * 32, w01 = ax * (32 - ay) * 32, w10 = (32 - ax) * ay * 32, w11 = ax * ay * 32;
    const size_t pitch = (size_t)fw * 3;
    const uint8_t *r0 = R.src + (size_t)y0 * pitch, *r1 = R.src + (size_t)y1 * pitch;
#pragma unroll
    for (int c = 0; c < 3; ++c)
        out[c] = (r0[3 * x0 + c] * w00 + r0[3 * x1 + c] * w01 + r1[3 * x0 + c] * w10 + r1[3 * x1 + c] * w11 + 16384) >> 15;
}

__global__ void __launch_bounds__(256) k_vc_compose(const __grid_constant__ VcParams P) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= P.W) return;
    const int cx = P.ex + x, cy = P.ey + y;                                          // canvas coordinates (:2116-2147)
    uint8_t* o = P.out + (size_t)y * P.out_stride + 3 * (size_t)x;
    int v[3];
    if (P.from_out) { v[0] = o[0]; v[1] = o[1]; v[2] = o[2]; }
    else {
        const int fx = cx - P.fx0, fy = cy - P.fy0;                                  // createVirtualCanvas, :2169-2212
        if ((unsigned)fx < (unsigned)P.W && (unsigned)fy < (unsigned)P.H) {
            const uint8_t* p = P.frame + (size_t)fy * P.frame_stride + 3 * (size_t)fx;
            v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
        } else v[0] = v[1] = v[2] = 0;
    }
    for (int k = 0; k < P.n_regions; ++k) {
        const VcRegion& R = P.r[k];
        const int px = cx - R.x, py = cy - R.y;
        if ((unsigned)px >= (unsigned)R.w || (unsigned)py >= (unsigned)R.h) continue;
        int s[3];
        if (!R.resize) vc_warp_pixel(R, P.W, P.H, R.ix + px, R.iy + py, s);
        else {                                                                       // cv::resize(INTER_LINEAR), :2341-2344
            const AxisTap th = tap_h(px, R.iw, R.sx), tv = tap_v(py, R.ih, R.sy);
            int c00[3], c01[3], c10[3], c11[3];
            vc_warp_pixel(R, P.W, P.H, R.ix + th.s0, R.iy + tv.s0, c00);
            vc_warp_pixel(R, P.W, P.H, R.ix + th.s1, R.iy + tv.s0, c01);
            vc_warp_pixel(R, P.W, P.H, R.ix + th.s0, R.iy + tv.s1, c10);
            vc_warp_pixel(R, P.W, P.H, R.ix + th.s1, R.iy + tv.s1, c11);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int h0 = c00[c] * th.a0 + c01[c] * th.a1, h1 = c10[c] * th.a0 + c11[c] * th.a1;
                s[c] = min(max(vres(h0, h1, tv.a0, tv.a1), 0), 255);
            }
        }
        // seamlessBlend, :2352-2399
        const int dist = min(min(px, py), min(R.w - px - 1, R.h - py - 1));
        float alpha = R.weight;
        if (dist < R.edge) alpha = __fmul_rn(alpha, __fdiv_rn((float)dist, (float)R.edge));
        const float na = __fsub_rn(1.0f, alpha);
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = (int)__fadd_rn(__fmul_rn(na, (float)v[c]), __fmul_rn(alpha, (float)s[c])) & 255;
    }
    o[0] = (uint8_t)v[0]; o[1] = (uint8_t)v[1]; o[2] = (uint8_t)v[2];
}

// The stage when nothing can ever be filled (canvas area at least twice the frame's, or no temporal buffer): the output is the
// frame moved by the integer part of the correction, which the kernel takes straight from the warp set-up block the motion
// kernel left on the device - no host round trip.  Window position: :2116-2133.
__global__ void __launch_bounds__(256) k_vc_shift(const uint8_t* __restrict__ frame, size_t frame_stride, uint8_t* __restrict__ out,
                                                  size_t out_stride, int W, int H, const WarpParams* __restrict__ wp, float ox, float oy,
                                                  int fx0, int fy0, int cw, int ch) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const int ex = min(max(0, (int)__fsub_rn(ox, wp->T[2])), cw - W), ey = min(max(0, (int)__fsub_rn(oy, wp->T[5])), ch - H);
    const int fx = ex + x - fx0, fy = ey + y - fy0;
    uint8_t* o = out + (size_t)y * out_stride + 3 * (size_t)x;
    if ((unsigned)fx < (unsigned)W && (unsigned)fy < (unsigned)H) {
        const uint8_t* p = frame + (size_t)fy * frame_stride + 3 * (size_t)fx;
        o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
    } else o[0] = o[1] = o[2] = 0;
}

// ------------------------------------------------------------------------------------------------------------------ host
static inline VcRect rect_and(const VcRect& a, const VcRect& b) {
    const int x1 = std::max(a.x, b.x), y1 = std::max(a.y, b.y), x2 = std::min(a.x + a.w, b.x + b.w), y2 = std::min(a.y + a.h, b.y + b.h);
    VcRect r;
    if (x2 > x1 && y2 > y1) { r.x = x1; r.y = y1; r.w = x2 - x1; r.h = y2 - y1; }
    return r;
}

VirtualCanvas::~VirtualCanvas() { reset(); }

void VirtualCanvas::reset() {
    if (d_ring_) cudaFree(d_ring_);
    if (d_mask_) cudaFree(d_mask_);
    if (d_count_) cudaFree(d_count_);
    if (h_mask_) cudaFreeHost(h_mask_);
    if (h_count_) cudaFreeHost(h_count_);
    d_ring_ = d_mask_ = h_mask_ = nullptr;
    d_count_ = h_count_ = nullptr;
    buf_.clear();
    W_ = H_ = 0;
    ring_slots_ = ring_next_ = 0;
    sized_ = false;
    scale_ = 0.f;
    ring_only_valid_ = false;
}

vs_status VirtualCanvas::ensure(int W, int H) {
    if (W == W_ && H == H_) return VS_OK;
    reset();                                           // (the reference keeps older frames of another size; not mirrored)
    W_ = W; H_ = H;
    ring_slots_ = p_.temporal_buffer_size >= 2 ? p_.temporal_buffer_size : 0;
    const size_t fb = (size_t)W * 3 * H;
    if (ring_slots_) CUDA_TRY(cudaMalloc((void**)&d_ring_, fb * ring_slots_));
    CUDA_TRY(cudaMalloc((void**)&d_mask_, (size_t)W * H));
    CUDA_TRY(cudaMalloc((void**)&d_count_, sizeof(int)));
    CUDA_TRY(cudaMallocHost((void**)&h_mask_, (size_t)W * H));
    CUDA_TRY(cudaMallocHost((void**)&h_count_, sizeof(int)));
    return VS_OK;
}

// :2224-2241 — the empty regions of the canvas: bounding rectangles (area > 100) of the external contours of "gray <= 1", in
// cv::findContours order.  Outside the frame the canvas is black, inside it the device mask decides.
void VirtualCanvas::contour_regions(const uint8_t* h_mask, bool any_dark, std::vector<VcRect>& out) {
    out.clear();
    const int step = cw_ + 2;
    const int fx0 = (int)(cx_ - W_ / 2.0f), fy0 = (int)(cy_ - H_ / 2.0f);
    work_.assign((size_t)step * (ch_ + 2), 0);
    for (int y = 0; y < ch_; ++y) {
        signed char* row = work_.data() + (size_t)(y + 1) * step + 1;
        const int fy = y - fy0;
        if ((unsigned)fy >= (unsigned)H_) { std::memset(row, 1, cw_); continue; }
        std::memset(row, 1, fx0);
        if (any_dark) std::memcpy(row + fx0, h_mask + (size_t)fy * W_, W_);
        std::memset(row + fx0 + W_, 1, cw_ - fx0 - W_);
    }
    std::vector<std::vector<azc::Pt>> contours;
    azc::find_external_contours_padded(work_.data(), cw_, ch_, contours);
    for (const auto& c : contours) {
        int x0 = c[0].x, x1 = c[0].x, y0 = c[0].y, y1 = c[0].y;
        for (const auto& p : c) { x0 = std::min(x0, p.x); x1 = std::max(x1, p.x); y0 = std::min(y0, p.y); y1 = std::max(y1, p.y); }
        VcRect r;
        r.x = x0; r.y = y0; r.w = x1 - x0 + 1; r.h = y1 - y0 + 1;                     // cv::boundingRect
        if (r.w * r.h > 100) out.push_back(r);
    }
}

vs_status VirtualCanvas::apply(const uint8_t* d_frame, int W, int H, size_t stride, const float T[3], const float* recent, int n_recent,
                               uint8_t* d_out, size_t out_stride, cudaStream_t st, int* launches) {
    { const vs_status s0 = ensure(W, H); if (s0 != VS_OK) return s0; }
    const size_t tight = (size_t)W * 3, fb = tight * H;
    // updateTemporalFrameBuffer, :2153-2167
    if (ring_slots_ && !(sized_ && never_fills_)) {
        Entry e;
        e.slot = ring_next_;
        ring_next_ = (ring_next_ + 1) % ring_slots_;
        e.T[0] = T[0]; e.T[1] = T[1]; e.T[2] = T[2];
        CUDA_TRY(cudaMemcpy2DAsync(d_ring_ + fb * e.slot, tight, d_frame, stride, tight, H, cudaMemcpyDeviceToDevice, st));
        buf_.push_back(e);
        while ((int)buf_.size() > ring_slots_) buf_.pop_front();
    }
    // canvas size, first frame only (:2072-2107): the scale is kept once it has been chosen
    if (!sized_) {
        scale_ = p_.canvas_scale_factor;
        if (p_.adaptive_canvas_size && n_recent > 0) {                                // calculateOptimalCanvasSize, :2280-2314
            float max_motion = 0.0f;
            for (int i = 0; i < n_recent; ++i) {
                const float mx = recent[3 * i], my = recent[3 * i + 1];
                const float a = mx * mx, b = my * my;
                const float mag = std::sqrt(a + b);
                max_motion = std::max(max_motion, mag);
            }
            const float factor = std::max(1.0f, max_motion / 50.0f);
            const float d = factor - 1.0f, e = d * 0.5f;
            const float opt = p_.canvas_scale_factor + e;
            scale_ = std::max(p_.min_canvas_scale, std::min(p_.max_canvas_scale, opt));
        }
        cw_ = (int)(W * scale_);
        ch_ = (int)(H * scale_);
        if (cw_ < W || ch_ < H) return vs_set_error(VS_ERR_UNSUPPORTED, "virtual canvas smaller than the frame");
        cx_ = cw_ / 2.0f;
        cy_ = ch_ / 2.0f;
        sized_ = true;
        ring_only_valid_ = false;
        // isRegionAvailable (:2401-2421) can never pass when even a whole older frame covers at most half of the canvas - the
        // one region there is whenever the surround encloses the frame - or when the buffer never holds two frames
        const bool surround = (int)(cx_ - W / 2.0f) > 0 && (int)(cy_ - H / 2.0f) > 0 && (int)(cx_ - W / 2.0f) + W < cw_ && (int)(cy_ - H / 2.0f) + H < ch_;
        never_fills_ = ring_slots_ == 0 || (surround && !((float)(W * H) / (float)(cw_ * ch_) > 0.5f));
    }
    VcParams P{};
    P.frame = d_frame; P.out = d_out; P.frame_stride = stride; P.out_stride = out_stride;
    P.W = W; P.H = H;
    const float ox = cx_ - W / 2.0f, oy = cy_ - H / 2.0f;
    P.fx0 = (int)ox; P.fy0 = (int)oy;
    {   // :2116-2133
        const float fox = ox - T[0], foy = oy - T[1];
        int ex = std::max(0, (int)fox), ey = std::max(0, (int)foy);
        ex = std::min(ex, cw_ - W);
        ey = std::min(ey, ch_ - H);
        P.ex = ex; P.ey = ey;
    }
    // blendTemporalRegions, :2214-2278
    std::vector<VcRegion> fills;
    if (buf_.size() >= 2) {
        std::vector<VcRect> dark_regions;
        const std::vector<VcRect>* regions = &ring_only_;
        const bool surround = P.fx0 > 0 && P.fy0 > 0 && P.fx0 + W < cw_ && P.fy0 + H < ch_;
        if (surround) {
            // The canvas' black surround encloses the frame: its outer border is the only EXTERNAL contour (everything dark
            // inside the frame lies in its hole), so the one region is the whole canvas whatever the frame shows.
            if (!ring_only_valid_) { ring_only_.assign(1, VcRect{0, 0, cw_, ch_}); ring_only_valid_ = true; }
        } else {
            CUDA_TRY(cudaMemsetAsync(d_count_, 0, sizeof(int), st));
            k_vc_dark<<<dim3((W + 255) / 256, H), 256, 0, st>>>(d_frame, W, H, stride, d_mask_, d_count_);
            ++*launches;
            CUDA_TRY(cudaMemcpyAsync(h_count_, d_count_, sizeof(int), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            if (*h_count_ > 0) {
                CUDA_TRY(cudaMemcpyAsync(h_mask_, d_mask_, (size_t)W * H, cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamSynchronize(st));
                contour_regions(h_mask_, true, dark_regions);
                regions = &dark_regions;
            } else if (!ring_only_valid_) {
                contour_regions(nullptr, false, ring_only_);
                ring_only_valid_ = true;
            }
        }
        const int n = (int)buf_.size();
        const VcRect frame_rect{0, 0, W, H};
        for (const VcRect& r : *regions) {
            int best = -1;
            float best_w = 0.0f;
            VcRect best_inter;
            float best_rel[3] = {0, 0, 0};
            for (int i = 0; i < n - 1; ++i) {
                const float rel[3] = {T[0] - buf_[i].T[0], T[1] - buf_[i].T[1], T[2] - buf_[i].T[2]};
                const VcRect moved{r.x + (int)rel[0], r.y + (int)rel[1], r.w, r.h};
                const VcRect inter = rect_and(moved, frame_rect);
                const float coverage = (float)(inter.w * inter.h) / (float)(r.w * r.h);       // isRegionAvailable, :2401-2421
                if (!(coverage > 0.5f)) continue;
                float tw = (float)(i + 1) / (float)(size_t)n;
                tw *= p_.canvas_blend_weight;
                if (tw > best_w) { best = i; best_w = tw; best_inter = inter; best_rel[0] = rel[0]; best_rel[1] = rel[1]; best_rel[2] = rel[2]; }
            }
            if (best < 0) continue;
            VcRegion R{};
            R.x = r.x; R.y = r.y; R.w = r.w; R.h = r.h;
            R.ix = best_inter.x; R.iy = best_inter.y; R.iw = best_inter.w; R.ih = best_inter.h;
            const float da = -best_rel[2];                                                    // applyMotionCompensation, :2423-2443
            const float M[6] = {std::cos(da), -std::sin(da), -best_rel[0], std::sin(da), std::cos(da), -best_rel[1]};
            WarpParams wp;
            warp_params_from_T(M, &wp);
            for (int k = 0; k < 6; ++k) R.m[k] = wp.m[k];
            R.resize = (R.iw != R.w || R.ih != R.h) ? 1 : 0;
            R.sx = 1.0 / ((double)R.w / (double)R.iw);
            R.sy = 1.0 / ((double)R.h / (double)R.ih);
            R.src = d_ring_ + fb * buf_[best].slot;
            R.weight = best_w;
            R.edge = std::min(p_.edge_blend_radius, std::min(r.w, r.h) / 4);                  // :2371
            fills.push_back(R);
        }
    }
    regions_last_ = (int)fills.size();
    const dim3 grid((W + 255) / 256, H);
    size_t done = 0;
    do {
        const size_t n = std::min(fills.size() - done, (size_t)VC_MAX_REGIONS);
        P.n_regions = (int)n;
        P.from_out = done > 0;
        for (size_t k = 0; k < n; ++k) P.r[k] = fills[done + k];
        k_vc_compose<<<grid, 256, 0, st>>>(P);
        ++*launches;
        done += n;
    } while (done < fills.size());
    CUDA_TRY(cudaGetLastError());
    return VS_OK;
}

vs_status VirtualCanvas::apply_async(const uint8_t* d_frame, int W, int H, size_t stride, const WarpParams* d_wp, uint8_t* d_out,
                                     size_t out_stride, cudaStream_t st, int* launches) {
    if (!sized_ || !never_fills_ || W != W_ || H != H_) return vs_set_error(VS_ERR_INVALID_ARG, "virtual canvas: not a fill-free geometry");
    const float ox = cx_ - W / 2.0f, oy = cy_ - H / 2.0f;
    k_vc_shift<<<dim3((W + 255) / 256, H), 256, 0, st>>>(d_frame, stride, d_out, out_stride, W, H, d_wp, ox, oy, (int)ox, (int)oy, cw_, ch_);
    ++*launches;
    CUDA_TRY(cudaGetLastError());
    regions_last_ = 0;
    return VS_OK;
}
