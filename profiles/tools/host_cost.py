"""Host enqueue cost per frame of the in-library loop (vs_stabilizer_push_many_device): multi-stream engine vs the
single-stream verification mode (no events), enqueue time vs total time."""
import os, sys, time
import numpy as np, torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__; __graft_entry__.build()
import video_stab_b200 as vsb
import synthclip
W, H, n = 1920, 1080, 64
fb = W * H * 3
clip = torch.from_numpy(synthclip.make_clip(W, H, n, 2000)).cuda()
order = list(range(n)) + list(range(n - 2, 0, -1))
seq = clip[torch.tensor(order, device="cuda")].contiguous()
out = torch.empty_like(clip)
torch.cuda.synchronize()
for mode in ("multi", "single"):
    if mode == "single":
        os.environ["VS_SINGLE_STREAM"] = "1"
    else:
        os.environ.pop("VS_SINGLE_STREAM", None)
    st = vsb.Stabilizer(vsb.Parameters(smoothingRadius=15))
    pos = 0
    def run(k):                       # walks the ping-pong sequence without a jump, like bench.py
        global pos
        done = 0
        while done < k:
            a = pos % len(order)
            m = min(k - done, len(order) - a, n)
            st.push_many_device(seq[a].data_ptr(), fb, m, W, H, W * 3, out.data_ptr(), W * 3, fb, borrow=True)
            done += m; pos += m
    run(256); st.sync()
    for rep in range(3):
        t0 = time.perf_counter(); run(1024); t1 = time.perf_counter(); st.sync(); t2 = time.perf_counter()
        print(f"{mode:6s}: host enqueue {1e6 * (t1 - t0) / 1024:5.1f} us/frame, total {1e6 * (t2 - t0) / 1024:5.1f} us/frame")
