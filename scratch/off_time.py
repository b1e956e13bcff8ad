import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__; __graft_entry__.build()
import video_stab_b200 as vsb
from video_stab_b200 import offline
W, H, n = 1920, 1080, 768
base = vsb.synth.make_clip(W, H, 64, 5000)
pp = list(range(64)) + list(range(62, 0, -1))
idx = [pp[k % 126] for k in range(n)]
clip = torch.from_numpy(base).cuda()[torch.tensor(idx).cuda()]
out = torch.empty_like(clip)
fb = H * W * 3
for rep in range(2):
    for chunks in (1, 8):
        t0 = time.perf_counter()
        st = vsb.Stabilizer(vsb.Parameters(smoothingRadius=15))
        torch.cuda.synchronize(); t1 = time.perf_counter()
        parts = []
        for r in range(chunks):
            f, c = offline.chunk_bounds(n, chunks, r)
            hl = offline.halo(f)
            parts.append(offline.analyze_chunk(st, clip.data_ptr() + (f - hl) * fb, W, H, f, c))
        torch.cuda.synchronize(); t2 = time.perf_counter()
        tr = np.concatenate(parts)
        for r in range(chunks):
            f, c = offline.chunk_bounds(n, chunks, r)
            offline.render_chunk(st, tr, n, clip.data_ptr() + f * fb, W, H, f, c, out[f].data_ptr())
        torch.cuda.synchronize(); t3 = time.perf_counter()
        print(f"chunks={chunks}: create {1e3*(t1-t0):.1f} ms, analyze {1e3*(t2-t1):.1f} ms, render {1e3*(t3-t2):.1f} ms")
