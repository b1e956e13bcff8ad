"""Stand-alone driver for profiling RollCorrection (k_roll.cu) on device frames: python profiles/prof_roll.py [width height frames]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_stab_b200 as vsb
import synthclip

W, H, n = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (3840, 2160, 12)
dev = torch.device("cuda", 0)
clip = torch.from_numpy(synthclip.horizon_clip(W, H, n, 3000)).to(dev)
out = torch.empty_like(clip)
roll = vsb.RollCorrection(vsb.RollParameters(angleFilterMin=-70.0, angleFilterMax=70.0, angleDecay=0.98))
torch.cuda.synchronize()
s = torch.cuda.current_stream()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(n):
        roll.correct_device(clip[k].data_ptr(), W, H, W * 3, out[k].data_ptr(), W * 3, s.cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    print(f"{W}x{H}: {e0.elapsed_time(e1) / n * 1e3:.1f} us per frame", roll.state())
