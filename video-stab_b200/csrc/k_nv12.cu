// k_nv12.cu — the pixel formats either side of the path when the frames never leave the device:
// a hardware decoder hands over NV12 surfaces (a Y plane and an interleaved half-resolution UV plane), an encoder takes NV12
// back, and vs_stabilizer_push_device works on packed BGR.  The reference converts on the CPU inside its GStreamer pipelines
// (`videoconvert ! video/x-raw,format=BGR` before the appsink, examples/vsg.cpp:91-134, and back before the encoder, :230-311);
// here the two conversions are device kernels so that decode -> stabilize -> encode stays in HBM.
// Arithmetic: OpenCV's 8-bit BT.601 limited-range fixed point (20 fractional bits), bit-exact with
// cv::cvtColor(COLOR_YUV2BGR_NV12) and cv::cvtColor(COLOR_BGR2YUV_I420) + interleaving of U and V (tests/test_gpu_kernels.py):
//   C = max(Y - 16, 0) * 1220542
//   B = sat((C + 2116026 (U-128)                  + 2^19) >> 20)
//   G = sat((C -  409993 (U-128) - 852492 (V-128) + 2^19) >> 20)
//   R = sat((C + 1673527 (V-128)                  + 2^19) >> 20)
//   Y = (269484 R + 528482 G + 102760 B + (16 << 20) + 2^19) >> 20
//   U = (-155188 R - 305135 G + 460324 B + (128 << 20) + 2^19) >> 20      of the TOP-LEFT pixel of every 2x2 block
//   V = ( 460324 R - 385875 G -  74448 B + (128 << 20) + 2^19) >> 20      (OpenCV does not average the block)
// Both kernels are HBM-bound byte work: one thread = a 4x2 pixel block (aligned 32-bit loads / stores); frames whose width is
// not a multiple of 4 or whose pointers / strides are not 4-byte aligned take a per-2x2-block path.
#include "kernels.h"

static __device__ __forceinline__ int sat8(int v) { return min(max(v, 0), 255); }

static __device__ __forceinline__ uint32_t yuv_to_bgr(int y, int u, int v) {          // -> B | G << 8 | R << 16
    const int c = max(y - 16, 0) * 1220542 + (1 << 19);
    const int b = sat8((c + 2116026 * u) >> 20);
    const int g = sat8((c - 409993 * u - 852492 * v) >> 20);
    const int r = sat8((c + 1673527 * v) >> 20);
    return (uint32_t)b | ((uint32_t)g << 8) | ((uint32_t)r << 16);
}
static __device__ __forceinline__ int bgr_to_y(int b, int g, int r) {
    return (269484 * r + 528482 * g + 102760 * b + (16 << 20) + (1 << 19)) >> 20;
}
static __device__ __forceinline__ int bgr_to_u(int b, int g, int r) {
    return (-155188 * r - 305135 * g + 460324 * b + (128 << 20) + (1 << 19)) >> 20;
}
static __device__ __forceinline__ int bgr_to_v(int b, int g, int r) {
    return (460324 * r - 385875 * g - 74448 * b + (128 << 20) + (1 << 19)) >> 20;
}

// 4 pixels [B G R -] -> 12 packed bytes
static __device__ __forceinline__ void store_bgr4(uint8_t* p, const uint32_t (&px)[4]) {
    uint32_t* w = reinterpret_cast<uint32_t*>(p);
    w[0] = __byte_perm(px[0], px[1], 0x4210);
    w[1] = __byte_perm(px[1], px[2], 0x5421);
    w[2] = __byte_perm(px[2], px[3], 0x6542);
}

__global__ void __launch_bounds__(256) k_nv12_to_bgr(const uint8_t* __restrict__ yp, size_t ys, const uint8_t* __restrict__ uvp, size_t uvs,
                                                    int w, int h, uint8_t* __restrict__ bgr, size_t bs, int vec) {
    const int bx = blockIdx.x * blockDim.x + threadIdx.x, by = blockIdx.y;             // 4x2 block
    const int x = 4 * bx, y = 2 * by;
    if (x >= w || y >= h) return;
    if (vec) {
        const uint32_t y0 = *reinterpret_cast<const uint32_t*>(yp + (size_t)y * ys + x);
        const uint32_t y1 = (y + 1 < h) ? *reinterpret_cast<const uint32_t*>(yp + (size_t)(y + 1) * ys + x) : 0u;
        const uint32_t uv = *reinterpret_cast<const uint32_t*>(uvp + (size_t)by * uvs + x);
        const int u0 = (int)(uv & 255u) - 128, v0 = (int)((uv >> 8) & 255u) - 128;
        const int u1 = (int)((uv >> 16) & 255u) - 128, v1 = (int)(uv >> 24) - 128;
        uint32_t r0[4], r1[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int u = k < 2 ? u0 : u1, v = k < 2 ? v0 : v1;
            r0[k] = yuv_to_bgr((int)((y0 >> (8 * k)) & 255u), u, v);
            r1[k] = yuv_to_bgr((int)((y1 >> (8 * k)) & 255u), u, v);
        }
        store_bgr4(bgr + (size_t)y * bs + 3 * x, r0);
        if (y + 1 < h) store_bgr4(bgr + (size_t)(y + 1) * bs + 3 * x, r1);
    } else {
        for (int k = 0; k < 4 && x + k < w; ++k) {
            const uint8_t* q = uvp + (size_t)by * uvs + ((x + k) & ~1);
            const int u = (int)q[0] - 128, v = (int)q[1] - 128;
            for (int dy = 0; dy < 2 && y + dy < h; ++dy) {
                const uint32_t p = yuv_to_bgr(yp[(size_t)(y + dy) * ys + x + k], u, v);
                uint8_t* o = bgr + (size_t)(y + dy) * bs + 3 * (x + k);
                o[0] = (uint8_t)p; o[1] = (uint8_t)(p >> 8); o[2] = (uint8_t)(p >> 16);
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_bgr_to_nv12(const uint8_t* __restrict__ bgr, size_t bs, int w, int h, uint8_t* __restrict__ yp,
                                                    size_t ys, uint8_t* __restrict__ uvp, size_t uvs, int vec) {
    const int bx = blockIdx.x * blockDim.x + threadIdx.x, by = blockIdx.y;             // 4x2 block
    const int x = 4 * bx, y = 2 * by;
    if (x >= w || y >= h) return;
    if (vec) {
        uint32_t yy[2] = {0u, 0u}, uv = 0u;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            if (y + dy >= h) break;
            const uint32_t* s = reinterpret_cast<const uint32_t*>(bgr + (size_t)(y + dy) * bs + 3 * x);
            const uint32_t a = s[0], b = s[1], c = s[2];             // B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
            const int B[4] = {(int)(a & 255u), (int)(a >> 24), (int)((b >> 16) & 255u), (int)((c >> 8) & 255u)};
            const int G[4] = {(int)((a >> 8) & 255u), (int)(b & 255u), (int)(b >> 24), (int)((c >> 16) & 255u)};
            const int R[4] = {(int)((a >> 16) & 255u), (int)((b >> 8) & 255u), (int)(c & 255u), (int)(c >> 24)};
#pragma unroll
            for (int k = 0; k < 4; ++k) yy[dy] |= (uint32_t)bgr_to_y(B[k], G[k], R[k]) << (8 * k);
            if (dy == 0)
                uv = (uint32_t)bgr_to_u(B[0], G[0], R[0]) | ((uint32_t)bgr_to_v(B[0], G[0], R[0]) << 8) |
                     ((uint32_t)bgr_to_u(B[2], G[2], R[2]) << 16) | ((uint32_t)bgr_to_v(B[2], G[2], R[2]) << 24);
        }
        *reinterpret_cast<uint32_t*>(yp + (size_t)y * ys + x) = yy[0];
        if (y + 1 < h) *reinterpret_cast<uint32_t*>(yp + (size_t)(y + 1) * ys + x) = yy[1];
        *reinterpret_cast<uint32_t*>(uvp + (size_t)by * uvs + x) = uv;
    } else {
        for (int k = 0; k < 4 && x + k < w; ++k) {
            for (int dy = 0; dy < 2 && y + dy < h; ++dy) {
                const uint8_t* s = bgr + (size_t)(y + dy) * bs + 3 * (x + k);
                yp[(size_t)(y + dy) * ys + x + k] = (uint8_t)bgr_to_y(s[0], s[1], s[2]);
                if (dy == 0 && (k & 1) == 0) {
                    uint8_t* q = uvp + (size_t)by * uvs + x + k;
                    q[0] = (uint8_t)bgr_to_u(s[0], s[1], s[2]);
                    q[1] = (uint8_t)bgr_to_v(s[0], s[1], s[2]);
                }
            }
        }
    }
}

static inline bool a4(const void* p, size_t s) { return ((uintptr_t)p % 4 == 0) && (s % 4 == 0); }

void launch_nv12_to_bgr(const uint8_t* y, size_t ys, const uint8_t* uv, size_t uvs, int w, int h, uint8_t* bgr, size_t bs, cudaStream_t st) {
    const int vec = (w % 4 == 0 && a4(y, ys) && a4(uv, uvs) && a4(bgr, bs)) ? 1 : 0;
    dim3 grid(((w + 3) / 4 + 255) / 256, (h + 1) / 2);
    k_nv12_to_bgr<<<grid, 256, 0, st>>>(y, ys, uv, uvs, w, h, bgr, bs, vec);
}
void launch_bgr_to_nv12(const uint8_t* bgr, size_t bs, int w, int h, uint8_t* y, size_t ys, uint8_t* uv, size_t uvs, cudaStream_t st) {
    const int vec = (w % 4 == 0 && a4(y, ys) && a4(uv, uvs) && a4(bgr, bs)) ? 1 : 0;
    dim3 grid(((w + 3) / 4 + 255) / 256, (h + 1) / 2);
    k_bgr_to_nv12<<<grid, 256, 0, st>>>(bgr, bs, w, h, y, ys, uv, uvs, vec);
}
