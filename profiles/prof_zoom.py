"""Stand-alone driver for profiling the fused warp + crop + zoom kernel (config 3: 3840x2160, cropNZoom, b = 30)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
import video_stab_b200 as vsb  # noqa: E402
import synthclip

W, H, n = 3840, 2160, 12
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda", 0)
clip = torch.from_numpy(synthclip.make_clip(W, H, n, 3000)).to(dev)
out = torch.empty((n, H, W, 3), dtype=torch.uint8, device=dev)
st = vsb.Stabilizer(vsb.Parameters(smoothingRadius=5, cropNZoom=True, borderSize=30))
order = list(range(n)) + list(range(n - 2, 0, -1))
for k in range(frames):
    i = order[k % len(order)]
    st.push_device(clip[i].data_ptr(), W, H, W * 3, out[k % n].data_ptr(), W * 3, H * W * 3, borrow=True)
st.sync()
print("done", st.counts())
