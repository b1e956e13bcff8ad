"""8 streams on one GPU (config 4 at N = 8) as G lock-step groups of 8 / G lanes, interleaved from one host thread:
python profiles/tools/batch_groups.py [total_lanes]"""
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
TOT = int(sys.argv[1]) if len(sys.argv) > 1 else 8
lanes_all = []
for s in range(TOT):
    fr = synthclip.DeviceClip(W, H, L, 2000 + s, dev).frames(0, L)
    lanes_all.append(fr[torch.tensor(loop, device=dev)].contiguous())
for G in (1, 2, 4):
    S = TOT // G
    groups = []
    for g in range(G):
        lanes = lanes_all[g * S:(g + 1) * S]
        out = torch.empty((S, 8, H, W, 3), dtype=torch.uint8, device=dev)
        sb = vsb.StabilizerBatch(vsb.Parameters(smoothingRadius=15), S)
        PA = C.c_void_p * S
        it = [PA(*[lanes[i][f].data_ptr() for i in range(S)]) for f in range(len(loop))]
        ot = [PA(*[out[i, k].data_ptr() for i in range(S)]) for k in range(8)]
        groups.append((sb, it, ot, out))
    ow, oh, pr = C.c_int(), C.c_int(), C.c_int()
    pos = [0]

    def steps(n):
        for _ in range(n * 8):
            for sb, it, ot, _ in groups:
                assert lib.vs_batch_push_device(sb._h, it[pos[0] % len(loop)], W, H, W * 3, ot[pos[0] % 8], W * 3, fb, 1, C.byref(ow), C.byref(oh), C.byref(pr)) == 0
            pos[0] += 1
    steps(4)
    for g in groups:
        g[0].sync()
    t0 = time.perf_counter()
    steps(20)
    th = time.perf_counter() - t0
    for g in groups:
        g[0].sync()
    dt = time.perf_counter() - t0
    print(f"{TOT} lanes as {G} group(s) of {S}: {20 * 8 * TOT / dt:9.0f} frames/s  ({dt / 160 * 1e6:6.1f} us per frame step, host enqueue {th / 160 * 1e6:5.1f})", flush=True)
    del groups
    torch.cuda.empty_cache()
