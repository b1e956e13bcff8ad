"""Clip-mode trajectory build and smoothing at config 5's length (18 000 frames), per smoothing method: device time of
vs_clip_set_transforms_device (k_traj_build) and of vs_clip_render_prepared_device on tiny frames (k_smooth_batch dominates)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
import video_stab_b200 as vsb  # noqa: E402
from video_stab_b200 import offline  # noqa: E402

n, w, h = 18000, 64, 64
dev = torch.device("cuda", 0)
rng = np.random.default_rng(3)
tr = np.stack([rng.normal(0.5, 3, n - 1), rng.normal(0, 3, n - 1), rng.normal(0, 0.004, n - 1)], 1).astype(np.float32)
d_tr = torch.from_numpy(tr).to(dev)
frames = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device=dev)
out = torch.empty_like(frames)
for method in ("box", "gaussian", "kalman"):
    st = vsb.Stabilizer(vsb.Parameters(smoothingRadius=15, smoothingMethod=method))
    ext = torch.cuda.ExternalStream(st.stream, device=dev)
    ow, oh = C.c_int(), C.c_int()
    res = []
    for rep in range(3):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(ext)
        offline.check(offline.lib.vs_clip_set_transforms_device(st._h, d_tr.data_ptr(), n, w, h))
        e[1].record(ext)
        for c in range(8):                                      # 8 chunks rendered in order, as one rank of config 5 would
            first, count = offline.chunk_bounds(n, 8, c)
            offline.check(offline.lib.vs_clip_render_prepared_device(st._h, frames[first].data_ptr(), w, h, first, count,
                                                                     out[first].data_ptr(), C.byref(ow), C.byref(oh)))
        e[2].record(ext)
        st.sync()
        res.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
    print(f"{method:9s} set_transforms (k_traj_build) {min(r[0] for r in res):7.3f} ms   smooth + warp of 8 chunks {min(r[1] for r in res):7.3f} ms")
