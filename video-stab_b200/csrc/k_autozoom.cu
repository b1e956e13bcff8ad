// k_autozoom.cu — the device half of vs::AutoZoomCrop::autoZoomCrop (reference src/AutoZoomCrop.cpp:102-283, SURVEY.md section 8f
// rank 2) and the call that strings both halves together.  Device: cv::cvtColor(BGR2GRAY) -> threshold(gray > 1) -> morphological
// close with the 5x5 ellipse (dilate, erode; out-of-image pixels never win, as with cv::morphologyDefaultBorderValue) -> [host:
// contour + rectangle, autozoom_host.h] -> crop + cv::warpAffine([sx 0 0; 0 sy 0], 640x360, INTER_LINEAR, BORDER_CONSTANT) through
// the stabilizer's warp kernels (k_warp.cu).  The reference's first mask (THRESH_BINARY_INV + close, AutoZoomCrop.cpp:118-126) is
// never read again and is not computed.
#include "autozoom_host.h"
#include "kernels.h"

#define AZ_OUT_W 640
#define AZ_OUT_H 360

// cv::getStructuringElement(MORPH_ELLIPSE, Size(5, 5)): rows 0 and 4 hold only the centre column
static __device__ __forceinline__ bool az_in_kernel(int dy, int dx) { return (dy == -2 || dy == 2) ? dx == 0 : true; }

__global__ void __launch_bounds__(256) k_az_threshold(const uint8_t* __restrict__ src, int w, int h, size_t stride, uint8_t* __restrict__ bin) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const uint8_t* p = src + (size_t)y * stride + 3 * (size_t)x;
    const int g = (3735 * p[0] + 19235 * p[1] + 9798 * p[2] + 16384) >> 15;
    bin[(size_t)y * w + x] = g > 1 ? 255 : 0;
}
template <bool DILATE>
__global__ void __launch_bounds__(256) k_az_morph(const uint8_t* __restrict__ in, int w, int h, uint8_t* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    bool r = !DILATE;                       // dilate: any set pixel under the kernel; erode: all of them (outside the image never decides)
#pragma unroll
    for (int dy = -2; dy <= 2; ++dy)
#pragma unroll
        for (int dx = -2; dx <= 2; ++dx) {
            if (!az_in_kernel(dy, dx)) continue;
            const int xx = x + dx, yy = y + dy;
            if ((unsigned)xx >= (unsigned)w || (unsigned)yy >= (unsigned)h) continue;
            const bool v = in[(size_t)yy * w + xx] != 0;
            if (DILATE) r = r || v;
            else r = r && v;
        }
    out[(size_t)y * w + x] = r ? 255 : 0;
}

// mask (0 / 255, w x h) -> the padded 0 / 1 image ((h + 2) x (w + 2), zero frame) the border follower works on
__global__ void __launch_bounds__(256) k_az_pad01(const uint8_t* __restrict__ mask, int w, int h, signed char* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w + 2) return;
    const bool in = x >= 1 && x <= w && y >= 1 && y <= h;
    out[(size_t)y * (w + 2) + x] = (in && mask[(size_t)(y - 1) * w + (x - 1)]) ? 1 : 0;
}

void launch_content_mask(const uint8_t* d_bgr, int w, int h, size_t stride, uint8_t* d_mask, uint8_t* d_scratch, cudaStream_t st) {
    const dim3 g((w + 255) / 256, h);
    k_az_threshold<<<g, 256, 0, st>>>(d_bgr, w, h, stride, d_mask);
    k_az_morph<true><<<g, 256, 0, st>>>(d_mask, w, h, d_scratch);
    k_az_morph<false><<<g, 256, 0, st>>>(d_scratch, w, h, d_mask);
}

// One call = the reference's autoZoomCrop on a device frame.  The mask makes one trip to the host (as in the reference) for the
// contour logic; everything else is stream-ordered.
vs_status auto_zoom_crop_device(const uint8_t* d_bgr, int w, int h, size_t stride, uint8_t* d_out, size_t out_stride, size_t out_capacity,
                                int* ow, int* oh, cudaStream_t st) {
    if (!d_bgr || !d_out || w < 4 || h < 4 || !ow || !oh) return vs_set_error(VS_ERR_INVALID_ARG, "auto zoom-crop: bad argument");
    if (stride == 0) stride = (size_t)w * 3;
    const size_t padded = (size_t)(w + 2) * (h + 2);
    // Scratch kept per host thread and device (a call per frame must not pay for allocations): the mask, the padded 0/1 image,
    // one warp set-up block, and the page-locked landing buffer of the padded image.  Grown on demand, never shrunk.
    struct Scratch { int dev = -1; size_t cap = 0; uint8_t *d_mask = nullptr, *d_scratch = nullptr; WarpParams* d_wp = nullptr; signed char* h_work = nullptr; };
    static thread_local Scratch sc;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return vs_set_cuda_error(cudaGetLastError(), "cudaGetDevice", __FILE__, __LINE__);
    if (sc.dev != dev || padded > sc.cap) {
        if (sc.d_mask) cudaFree(sc.d_mask);
        if (sc.d_scratch) cudaFree(sc.d_scratch);
        if (sc.d_wp) cudaFree(sc.d_wp);
        if (sc.h_work) cudaFreeHost(sc.h_work);
        sc = Scratch();
        cudaError_t ea = cudaMalloc((void**)&sc.d_mask, padded);
        if (ea == cudaSuccess) ea = cudaMalloc((void**)&sc.d_scratch, padded);
        if (ea == cudaSuccess) ea = cudaMalloc((void**)&sc.d_wp, sizeof(WarpParams));
        if (ea == cudaSuccess) ea = cudaMallocHost((void**)&sc.h_work, padded);
        if (ea != cudaSuccess) return vs_set_cuda_error(ea, "auto zoom-crop scratch", __FILE__, __LINE__);
        sc.dev = dev;
        sc.cap = padded;
    }
    uint8_t *d_mask = sc.d_mask, *d_scratch = sc.d_scratch;
    WarpParams* d_wp = sc.d_wp;
    signed char* h_work = sc.h_work;
    cudaError_t e = cudaSuccess;
    vs_status rc = VS_OK;
    if (e == cudaSuccess) {
        launch_content_mask(d_bgr, w, h, stride, d_mask, d_scratch, st);
        k_az_pad01<<<dim3((w + 2 + 255) / 256, h + 2), 256, 0, st>>>(d_mask, w, h, reinterpret_cast<signed char*>(d_scratch));
        e = cudaMemcpyAsync(h_work, d_scratch, padded, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    if (e == cudaSuccess) {
        azc::Rect r;
        const bool found = azc::crop_rect_from_padded(h_work, w, h, &r);
        // no contour, or an empty crop: the reference returns the frame itself (AutoZoomCrop.cpp:149-152, 233-245)
        const bool unchanged = !found || r.width <= 0 || r.height <= 0;
        const int dw = unchanged ? w : AZ_OUT_W, dh = unchanged ? h : AZ_OUT_H;
        const size_t tight = (size_t)dw * 3;
        if (out_stride == 0) out_stride = tight;
        *ow = dw; *oh = dh;
        if (out_stride < tight || out_stride * (size_t)(dh - 1) + tight > out_capacity)
            rc = vs_set_error(VS_ERR_BUFFER_TOO_SMALL, "auto zoom-crop: output buffer too small");
        else if (unchanged)
            e = cudaMemcpy2DAsync(d_out, out_stride, d_bgr, stride, (size_t)w * 3, h, cudaMemcpyDeviceToDevice, st);
        else {
            const double sx = 640.0 / r.width, sy = 360.0 / r.height;                  // AutoZoomCrop.cpp:246-252
            const float T[6] = {(float)sx, 0.f, 0.f, 0.f, (float)sy, 0.f};
            WarpParams wp;
            warp_params_from_T(T, &wp);
            e = cudaMemcpyAsync(d_wp, &wp, sizeof(wp), cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) {
                const uint8_t* crop = d_bgr + (size_t)r.y * stride + 3 * (size_t)r.x;
                launch_warp_matrices(crop, r.width, r.height, stride, 0, d_out, dw, dh, out_stride, 0, d_wp, 1, st);
                e = cudaGetLastError();
                if (e == cudaSuccess) e = cudaStreamSynchronize(st);                   // `wp` is a stack object read by the copy
            }
        }
    }
    if (e != cudaSuccess) return vs_set_cuda_error(e, "auto zoom-crop", __FILE__, __LINE__);
    return rc;
}
