"""Host-buffer throughput of vs_stabilizer_push_many as a function of the frames per call (the per-call drain costs one
pipeline latency: the last input's copy, analysis, warp and copy-out cannot overlap anything)."""
import os, sys, time, ctypes as C
import numpy as np, torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__; __graft_entry__.build()
import video_stab_b200 as vsb
import synthclip
from video_stab_b200._capi import lib
W, H = 1920, 1080
fb = W * H * 3
base = synthclip.make_clip(W, H, 64, 2000)
order = list(range(64)) + list(range(62, 0, -1))
N = 512
seq = torch.from_numpy(np.stack([base[order[k % len(order)]] for k in range(N)])).pin_memory()
outs = torch.empty((N, H, W, 3), dtype=torch.uint8).pin_memory()
ow, oh, pr = C.c_int(), C.c_int(), C.c_int()
for per_call in (16, 64, 128, 512):
    st = vsb.Stabilizer(vsb.Parameters(smoothingRadius=15))
    def run():
        for a in range(0, N, per_call):
            assert lib.vs_stabilizer_push_many(st._h, seq[a].data_ptr(), fb, per_call, W, H, W * 3, outs[a].data_ptr(), W * 3, fb,
                                               C.byref(ow), C.byref(oh), C.byref(pr)) == 0
    run()
    best = 0
    for _ in range(3):
        t0 = time.perf_counter(); run(); best = max(best, N / (time.perf_counter() - t0))
    print(f"{per_call:4d} frames per call: {best:7.0f} frames/s")
