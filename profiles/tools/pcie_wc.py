"""PCIe copy rates with default page-locked memory vs write-combined page-locked input buffers (cudaHostAllocWriteCombined),
per-frame 1080p copies issued back to back on two streams (the pattern of vs_stabilizer_push_many)."""
import ctypes as C, time
import torch
rt = C.CDLL("libcudart.so.12") if True else None
def host_alloc(n, flags):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(flags)) == 0
    return p.value
FB, N = 1920 * 1080 * 3, 64
dev_in = torch.empty(N * FB, dtype=torch.uint8, device="cuda")
dev_out = torch.empty(N * FB, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
def run(src, dst, duplex):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for rep in range(3):
        for k in range(N):
            rt.cudaMemcpyAsync(dev_in.data_ptr() + k * FB, src + k * FB, FB, 1, C.c_void_p(s1.cuda_stream))
            if duplex:
                rt.cudaMemcpyAsync(dst + k * FB, dev_out.data_ptr() + k * FB, FB, 2, C.c_void_p(s2.cuda_stream))
    torch.cuda.synchronize()
    return 3 * N / (time.perf_counter() - t0)
plain_in, wc_in, out = host_alloc(N * FB, 0), host_alloc(N * FB, 4), host_alloc(N * FB, 0)
C.memset(plain_in, 1, N * FB); C.memset(wc_in, 1, N * FB)
for name, src in (("default pinned", plain_in), ("write-combined", wc_in)):
    for duplex in (False, True):
        r = max(run(src, out, duplex) for _ in range(3))
        print(f"{name:15s} {'H2D + D2H' if duplex else 'H2D only '}: {r:7.0f} frames/s  ({r * FB / 1e9:.1f} GB/s per direction)")
