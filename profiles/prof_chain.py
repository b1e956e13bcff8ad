"""Stand-alone driver for profiling the latency chain of a single 1080p stream (k_gray_half, k_pyrdown2, k_pyr_lk, k_motion,
k_eig_nms, k_select, k_warp_tma): 48 frames through vs_stabilizer_push_many_device.  VS_SINGLE_STREAM=1 serialises the handle's
streams (what ncu does anyway)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_stab_b200 as vsb
import synthclip

W, H, n = 1920, 1080, 48
dev = torch.device("cuda", 0)
clip = torch.from_numpy(synthclip.make_clip(W, H, n, 2000)).to(dev)
out = torch.empty_like(clip)
torch.cuda.synchronize()
st = vsb.Stabilizer(vsb.Parameters(smoothingRadius=15))
fb = H * W * 3
k = st.push_many_device(clip.data_ptr(), fb, n, W, H, W * 3, out.data_ptr(), W * 3, fb, borrow=True)
st.sync()
print("frames", n, "outputs", k, "launches", st.launch_count())
