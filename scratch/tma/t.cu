// minimal TMA probe: 3-D u32 tensor map over a packed BGR clip, one 112x42x1 box into shared memory
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#define BW 112
#define BH 42
__global__ void k(const __grid_constant__ CUtensorMap tm, uint32_t* out, int c0, int c1, int c2, int mode, const uint8_t* g) {
    extern __shared__ __align__(1024) unsigned char sm[];
    uint32_t* raw = (uint32_t*)sm;
    unsigned long long* mbar = (unsigned long long*)(sm + BW * BH * 4);
    uint32_t s_raw = (uint32_t)__cvta_generic_to_shared(raw), s_mbar = (uint32_t)__cvta_generic_to_shared(mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s_mbar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (mode == 0) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(s_mbar) : "memory");
        } else if (mode == 1) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s_mbar), "r"(448) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(s_raw), "l"(g), "r"(448), "r"(s_mbar) : "memory");
        } else {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s_mbar), "r"(BW * BH * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     :: "r"(s_raw), "l"(&tm), "r"(s_mbar), "r"(c0), "r"(c1), "r"(c2) : "memory");
        }
    }
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n"
                 :: "r"(s_mbar), "r"(0) : "memory");
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = raw[i];
}
int main() {
    const int W = 1920, H = 1080, N = 3;
    size_t stride = (size_t)W * 3, frame = stride * H;
    std::vector<uint8_t> h(frame * N);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t* d; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    uint32_t* dout; cudaMalloc(&dout, BW * BH * 4);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t ce = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry point: %d %d %p\n", (int)ce, (int)q, p);
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)W * 3 / 4, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[2] = {stride, frame};
    cuuint32_t box[3] = {BW, BH, 1}, es[3] = {1, 1, 1};
    CUresult r = ((EncFn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    int smem = BW * BH * 4 + 64;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int tests[4][3] = {{30, 100, 1}, {-6, -3, 0}, {1400, 1060, 2}, {0, 0, 0}};
    std::vector<uint32_t> o(BW * BH);
    for (int mode = 0; mode < 2; ++mode) {
        k<<<1, 128, smem>>>(tm, dout, 0, 0, 0, mode, d);
        printf("mode %d: %s\n", mode, cudaGetErrorString(cudaDeviceSynchronize()));
    }
    for (auto& t : tests) {
        k<<<1, 128, smem>>>(tm, dout, t[0], t[1], t[2], 2, d);
        cudaError_t e = cudaDeviceSynchronize();
        printf("launch c=(%d,%d,%d): %s\n", t[0], t[1], t[2], cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int y = 0; y < BH; ++y)
            for (int x = 0; x < BW; ++x) {
                int gx = t[0] + x, gy = t[1] + y;
                uint32_t exp = 0;
                if (gx >= 0 && gx < W * 3 / 4 && gy >= 0 && gy < H) memcpy(&exp, &h[(size_t)t[2] * frame + (size_t)gy * stride + 4 * (size_t)gx], 4);
                if (o[y * BW + x] != exp) ++bad;
            }
        printf("  mismatches: %ld\n", bad);
    }
    return 0;
}
