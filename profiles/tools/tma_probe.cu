// TMA probe 2: argv selects configuration.  usage: t2 rank dtype(0=u8,1=u32) boxw boxh
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tm_param, const CUtensorMap* tm_glob, uint8_t* out, int c0, int c1, int c2, int rank, int bytes) {
    const CUtensorMap* tmp = tm_glob ? tm_glob : &tm_param;
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned long long* mbar = (unsigned long long*)(sm + 32768);
    uint32_t s_raw = (uint32_t)__cvta_generic_to_shared(sm), s_mbar = (uint32_t)__cvta_generic_to_shared(mbar);
    if (threadIdx.x == 0) printf("smem %x mbar %x tm %p; ", s_raw, s_mbar, tmp);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s_mbar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s_mbar), "r"(bytes) : "memory");
        if (rank == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         :: "r"(s_raw), "l"(tmp), "r"(s_mbar), "r"(c0), "r"(c1) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         :: "r"(s_raw), "l"(tmp), "r"(s_mbar), "r"(c0), "r"(c1), "r"(c2) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n"
                 :: "r"(s_mbar), "r"(0) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char** argv) {
    int rank = atoi(argv[1]), dt = atoi(argv[2]), bw = atoi(argv[3]), bh = atoi(argv[4]);
    const int W = 1920, H = 1080, N = 3;
    size_t stride = (size_t)W * 3, frame = stride * H;
    std::vector<uint8_t> h(frame * N);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t* d; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    uint8_t* dout; cudaMalloc(&dout, 32768);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    int es_b = dt ? 4 : 1;
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)W * 3 / es_b, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[2] = {stride, frame};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
    CUresult r = ((EncFn)p)(&tm, dt ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    int bytes = bw * bh * es_b;
    printf("rank %d dtype %d box %dx%d (%d bytes): encode %d; ", rank, dt, bw, bh, bytes, (int)r);
    int smem = 32768 + 64;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int c0 = argc > 6 ? atoi(argv[6]) : 0, c1 = argc > 7 ? atoi(argv[7]) : 0, c2 = rank == 3 ? 1 : 0;
    CUtensorMap* dtm = nullptr;
    if (argc > 5) { cudaMalloc(&dtm, 128); cudaMemcpy(dtm, &tm, 128, cudaMemcpyHostToDevice); }
    { const unsigned long long* w = (const unsigned long long*)&tm; printf("desc %016llx %016llx %016llx %016llx; ", w[0], w[1], w[2], w[3]); }
    k<<<1, 128, smem>>>(tm, dtm, dout, c0, c1, c2, rank, bytes);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run: %s; ", cudaGetErrorString(e));
    if (e != cudaSuccess) { printf("\n"); return 1; }
    std::vector<uint8_t> o(bytes);
    cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int y = 0; y < bh; ++y)
        for (int x = 0; x < bw * es_b; ++x)
            if (o[(size_t)y * bw * es_b + x] != h[(size_t)c2 * frame + (size_t)(c1 + y) * stride + (size_t)c0 * es_b + x]) ++bad;
    printf("mismatches %ld\n", bad);
    return 0;
}
