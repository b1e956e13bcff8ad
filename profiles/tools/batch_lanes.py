"""Aggregate frames/s and per-stage kernel times of one StabilizerBatch as a function of its lane count (config 4 sharding puts
64 / N streams on a GPU): python profiles/tools/batch_lanes.py [lanes ...]"""
import ctypes as C
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
import video_stab_b200 as vsb
from video_stab_b200._capi import lib
import synthclip

W, H, L = 1920, 1080, 24
fb = W * H * 3
dev = torch.device("cuda", 0)
loop = list(range(L)) + list(range(L - 2, 0, -1))
for S in [int(a) for a in sys.argv[1:]] or [8, 16, 32, 64]:
    lanes = []
    for s in range(S):
        fr = synthclip.DeviceClip(W, H, L, 2000 + s, dev).frames(0, L)
        lanes.append(fr[torch.tensor(loop, device=dev)].contiguous())
    out = torch.empty((S, 8, H, W, 3), dtype=torch.uint8, device=dev)
    sb = vsb.StabilizerBatch(vsb.Parameters(smoothingRadius=15), S)
    PA = C.c_void_p * S
    it = [PA(*[lanes[i][f].data_ptr() for i in range(S)]) for f in range(len(loop))]
    ot = [PA(*[out[i, k].data_ptr() for i in range(S)]) for k in range(8)]
    ow, oh, pr = C.c_int(), C.c_int(), C.c_int()

    def steps(n, pos=[0]):
        for _ in range(n * 8):
            assert lib.vs_batch_push_device(sb._h, it[pos[0] % len(loop)], W, H, W * 3, ot[pos[0] % 8], W * 3, fb, 1, C.byref(ow), C.byref(oh), C.byref(pr)) == 0
            pos[0] += 1
    steps(4)
    sb.sync()
    t0 = time.perf_counter()
    steps(20)
    t_host = time.perf_counter() - t0
    sb.sync()
    dt = time.perf_counter() - t0
    sb.set_timing(True)
    steps(3)
    st = sb.stage_times()
    sb.set_timing(False)
    us = {k: round(v["ms"] / v["count"] * 1e3, 1) for k, v in st.items() if v["count"]}
    print(f"lanes {S:2d}: {20 * 8 * S / dt:9.0f} frames/s  {dt / 160 * 1e6:6.1f} us per lock-step frame (host enqueue {t_host / 160 * 1e6:5.1f})  stage us {us}", flush=True)
    del sb, lanes, out
    torch.cuda.empty_cache()
