"""Host-buffer path (vs_stabilizer_push_many): device time of the copy-in and copy-out of every frame (stage timing, CUDA events on
the copy streams) next to the throughput, against plain duplex copies of the same frames."""
import ctypes as C, os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")     # before the CUDA context exists (as bench.py does)
import numpy as np, torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__; __graft_entry__.build()
import video_stab_b200 as vsb
import synthclip
from video_stab_b200._capi import lib

W, H, n = 1920, 1080, 64
fb = W * H * 3
clip = torch.from_numpy(synthclip.make_clip(W, H, n, 2000)).pin_memory()
outs = torch.empty((n, H, W, 3), dtype=torch.uint8).pin_memory()
ow, oh, pr = C.c_int(), C.c_int(), C.c_int()
for timing in (False, True):
    st = vsb.Stabilizer(vsb.Parameters(smoothingRadius=15))
    def call():
        assert lib.vs_stabilizer_push_many(st._h, clip.data_ptr(), fb, n, W, H, W * 3, outs.data_ptr(), W * 3, fb, C.byref(ow), C.byref(oh), C.byref(pr)) == 0
    for _ in range(4):
        call()
    if timing:
        st.set_timing(True)
    t0 = time.perf_counter()
    reps = 16
    for _ in range(reps):
        call()
    dt = time.perf_counter() - t0
    print(f"timing {'on ' if timing else 'off'}: {reps * n / dt:7.0f} frames/s ({1e6 * dt / (reps * n):6.1f} us per frame)")
    if timing:
        for k, v in st.stage_times().items():
            if v["count"]:
                print(f"   {k:12s} {1e3 * v['ms'] / v['count']:7.1f} us x {v['count']}")
