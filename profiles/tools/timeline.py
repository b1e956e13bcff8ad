"""Per-stream timeline of the single-stream loop from the handle's stage events (vs_stabilizer_trace).
Note: the event records themselves add host and device overhead, so the period is longer than in the untimed loop."""
import os, sys
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
st.set_timing(True)
step(int(os.environ.get("TL_FRAMES", "40")))
tr = st.trace()
names = ["gray", "pyrdown", "lk", "motion", "gftt", "warp"]
t0 = tr[len(tr) // 2, 1]
if not os.environ.get("VS_TRACE_STAGE"):
    for s, a, b in tr[len(tr) // 2: len(tr) // 2 + 40]:
        print(f"{names[int(s)]:8s} {a - t0:8.1f} -> {b - t0:8.1f}  ({b - a:5.1f} us)")
for k, nm in enumerate(names):
    sel = tr[tr[:, 0] == k]
    if len(sel) > 2:
        d = sel[:, 2] - sel[:, 1]
        gap = sel[1:, 1] - sel[:-1, 2]
        print(f"{nm:8s} n={len(sel):3d} duration mean {d.mean():6.1f} min {d.min():6.1f} max {d.max():6.1f} us, period {np.mean(np.diff(sel[:, 1])):6.1f} us, "
              f"idle gap before next mean {gap.mean():6.1f} us")
