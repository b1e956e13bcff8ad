import os, sys
import numpy as np, torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__; __graft_entry__.build()
import video_stab_b200 as vsb
import synthclip
from video_stab_b200 import offline
W, H = 1920, 1080
dev = torch.device("cuda", 0)
base = torch.from_numpy(synthclip.make_clip(W, H, 64, 5000)).to(dev)
pp = list(range(64)) + list(range(62, 0, -1))
for n in (256, 1024, 2048):
    clip = base[torch.tensor([pp[k % 126] for k in range(n)], device=dev)]
    p = vsb.Parameters(smoothingRadius=15)
    ref = None
    for rep, chunks in enumerate((1, 1, 2, 8)):
        out, tr = offline.stabilize_clip(clip, p, n_chunks=chunks)
        if ref is None:
            ref = (out.clone(), tr.copy())
        else:
            same_tr = np.array_equal(tr.view(np.uint32), ref[1].view(np.uint32))
            bad = np.nonzero((tr.view(np.uint32) != ref[1].view(np.uint32)).any(axis=1))[0]
            print(n, chunks, "transforms equal", same_tr, "first diffs", bad[:8], "frames equal", bool(torch.equal(out, ref[0])))
    del clip, out, ref
