import os, sys, time
if len(sys.argv) > 1: os.environ["CUDA_DEVICE_MAX_CONNECTIONS"] = sys.argv[1]
import numpy as np, torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__; __graft_entry__.build()
import video_stab_b200 as vsb
import synthclip
W, H, n = 1920, 1080, 64
clip = torch.from_numpy(synthclip.make_clip(W, H, n, 2000)).cuda()
out = torch.empty_like(clip)
order = list(range(n)) + list(range(n - 2, 0, -1))
st = vsb.Stabilizer(vsb.Parameters(smoothingRadius=15))
pos = 0
def step(k):
    global pos
    for _ in range(k):
        i = order[pos % len(order)]
        st.push_device(clip[i].data_ptr(), W, H, W * 3, out[pos % n].data_ptr(), W * 3, H * W * 3, borrow=True)
        pos += 1
step(200); st.sync()
for rep in range(3):
    t0 = time.perf_counter(); step(640); t1 = time.perf_counter(); st.sync(); t2 = time.perf_counter()
    print(f"host enqueue {1e6*(t1-t0)/640:.1f} us/frame, total {1e6*(t2-t0)/640:.1f} us/frame")
# pure python/ctypes overhead of the pointer computations
t0 = time.perf_counter()
for _ in range(640):
    i = order[pos % len(order)]; a = clip[i].data_ptr(); b = out[pos % n].data_ptr(); pos += 1
print(f"python pointer math {1e6*(time.perf_counter()-t0)/640:.1f} us/frame")
