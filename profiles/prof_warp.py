"""Stand-alone driver for profiling the batched warp kernel (64 x 1080p frames per launch)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
import video_stab_b200 as vsb  # noqa: E402

N, H, W = 64, 1080, 1920
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
g = torch.Generator(device="cuda").manual_seed(0)
src = torch.randint(0, 256, (N, H, W, 3), dtype=torch.uint8, device="cuda", generator=g)
dst = torch.empty_like(src)
rng = np.random.default_rng(7)
T = np.zeros((N, 2, 3), np.float32)
for i in range(N):
    da = np.float32(rng.normal(0, 0.004))
    T[i] = [[np.cos(da), -np.sin(da), rng.normal(0, 3)], [np.sin(da), np.cos(da), rng.normal(0, 3)]]
cur = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    vsb.kernels.warp_affine(src, T, out=dst, stream=cur)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    vsb.kernels.warp_affine(src, T, out=dst, stream=cur)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"warp 64x1080p: {ms*1e3/N:.2f} us/frame, {2*3*W*H*N/ms/1e6:.0f} GB/s algorithmic")
