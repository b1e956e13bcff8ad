"""Duplex PCIe copies of 1080p frames (H2D and D2H streams, page-locked host memory) alone, and with the stabilizer's kernels
running beside them on device-resident frames: does compute slow the copies, or does the host-buffer pipeline lose time elsewhere?"""
import os, sys, time, threading
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np, torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__; __graft_entry__.build()
import video_stab_b200 as vsb
import synthclip

W, H, n = 1920, 1080, 64
fb = W * H * 3
host_in = torch.from_numpy(synthclip.make_clip(W, H, n, 2000)).pin_memory()
host_out = torch.empty((n, H, W, 3), dtype=torch.uint8).pin_memory()
dev_in = torch.empty((n, H, W, 3), dtype=torch.uint8, device="cuda")
dev_out = torch.randint(0, 255, (n, H, W, 3), dtype=torch.uint8, device="cuda")
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

def duplex(frames):
    t0 = time.perf_counter()
    for k in range(frames):
        with torch.cuda.stream(s_in):
            dev_in[k % n].copy_(host_in[k % n], non_blocking=True)
        with torch.cuda.stream(s_out):
            host_out[k % n].copy_(dev_out[k % n], non_blocking=True)
    s_in.synchronize(); s_out.synchronize()
    return frames / (time.perf_counter() - t0)

duplex(256)
print(f"duplex copies alone: {duplex(2048):8.0f} frames/s each way")

clip = torch.from_numpy(synthclip.make_clip(W, H, n, 2001)).cuda()
seq = clip[torch.tensor(list(range(n)) + list(range(n - 2, 0, -1)), device="cuda")].contiguous()
outd = torch.empty_like(clip)
st = vsb.Stabilizer(vsb.Parameters(smoothingRadius=15))
stop = False
count = [0]
def compute():
    pos = 0
    while not stop:
        a = pos % seq.shape[0]
        m = min(32, seq.shape[0] - a)
        st.push_many_device(seq[a].data_ptr(), fb, m, W, H, W * 3, outd.data_ptr(), W * 3, fb, borrow=True)
        st.sync()
        pos += m; count[0] += m
th = threading.Thread(target=compute); th.start()
time.sleep(0.2)
c0, t0 = count[0], time.perf_counter()
r = duplex(2048)
c1, t1 = count[0], time.perf_counter()
stop = True; th.join()
print(f"duplex copies beside a device-resident stream ({(c1 - c0) / (t1 - t0):.0f} frames/s of kernels): {r:8.0f} frames/s each way")
