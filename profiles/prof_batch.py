"""Stand-alone driver for profiling the 64-stream lock-step batch (config 4)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
import video_stab_b200 as vsb  # noqa: E402
import synthclip

W, H, n, S = 1920, 1080, 8, 64
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 24
dev = torch.device("cuda", 0)
clips = [torch.from_numpy(synthclip.make_clip(W, H, n, 2000 + s)).to(dev) for s in range(S)]
outs = torch.empty((S, H, W, 3), dtype=torch.uint8, device=dev)
order = list(range(n)) + list(range(n - 2, 0, -1))
sb = vsb.StabilizerBatch(vsb.Parameters(smoothingRadius=15), S)
torch.cuda.synchronize()
for k in range(steps):
    i = order[k % len(order)]
    sb.push_device([c[i].data_ptr() for c in clips], W, H, W * 3, [outs[s].data_ptr() for s in range(S)], W * 3, H * W * 3, borrow=True)
sb.sync()
print("done")
