"""Copy / kernel timeline of the host-buffer path (vs_stabilizer_push_many) from CUPTI through torch.profiler."""
import json, os, sys, tempfile, ctypes as C
import numpy as np, torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__; __graft_entry__.build()
import video_stab_b200 as vsb
import synthclip
from video_stab_b200._capi import lib
from torch.profiler import profile, ProfilerActivity
W, H, n = 1920, 1080, 64
fb = W * H * 3
clip = torch.from_numpy(synthclip.make_clip(W, H, n, 2000)).pin_memory()
outs = torch.empty((n, H, W, 3), dtype=torch.uint8).pin_memory()
st = vsb.Stabilizer(vsb.Parameters(smoothingRadius=15))
ow, oh, pr = C.c_int(), C.c_int(), C.c_int()
def call():
    assert lib.vs_stabilizer_push_many(st._h, clip.data_ptr(), fb, n, W, H, W * 3, outs.data_ptr(), W * 3, fb, C.byref(ow), C.byref(oh), C.byref(pr)) == 0
for _ in range(3):
    call()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    call(); call()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in e]
ev.sort(key=lambda e: e["ts"])
span = ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]
print(f"{len(ev)} activities over {span:.0f} us = {span / (2 * n):.1f} us per frame ({2 * n / span * 1e6:.0f} frames/s)")
for kind in ("HtoD", "DtoH"):
    cp = [e for e in ev if e["cat"] == "gpu_memcpy" and kind in e["name"]]
    d = np.array([e["dur"] for e in cp]); gaps = np.array([b["ts"] - (a["ts"] + a["dur"]) for a, b in zip(cp[:-1], cp[1:])])
    print(f"{kind}: n={len(cp)} duration mean {d.mean():.1f} us ({fb / d.mean() / 1e3:.1f} GB/s) min {d.min():.1f} max {d.max():.1f}; gap to next mean {gaps.mean():.1f} us, "
          f"gaps > 20 us: {int((gaps > 20).sum())} (total {gaps[gaps > 20].sum():.0f} us), busy {d.sum() / span * 100:.0f}%")
mid = len(ev) // 2
base = ev[mid]["ts"]
for e in ev[mid: mid + 40]:
    print(f"{e['ts'] - base:8.1f} {e['ts'] + e['dur'] - base:8.1f}  s{e['args'].get('stream')}  {e['name'].split('(')[0].replace('void ', '')[:24]:24s} {e['dur']:6.1f}")
