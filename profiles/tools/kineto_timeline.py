"""Kernel timeline of the single-stream loop from CUPTI (through torch.profiler, which records every kernel of the
process, ours included): per-kernel duration in situ, per-stream busy time, and the dependency gaps.
    python profiles/tools/kineto_timeline.py [frames] > gpurun_out/timeline.txt"""
import json, os, sys, tempfile
import numpy as np, torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__; __graft_entry__.build()
import video_stab_b200 as vsb
import synthclip
from torch.profiler import profile, ProfilerActivity
W, H, n = 1920, 1080, 64
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 120
clip = torch.from_numpy(synthclip.make_clip(W, H, n, 2000)).cuda()
out = torch.empty_like(clip)
order = list(range(n)) + list(range(n - 2, 0, -1))
seq = clip[torch.tensor(order, device="cuda")].contiguous()
torch.cuda.synchronize()
st = vsb.Stabilizer(vsb.Parameters(smoothingRadius=15))
pos = 0
fb = W * H * 3
def step(k):                      # the loop runs inside the library (vs_stabilizer_push_many_device): no Python per frame
    global pos
    while k > 0:
        a = pos % len(order)
        m = min(k, len(order) - a, n)
        st.push_many_device(seq[a].data_ptr(), fb, m, W, H, W * 3, out.data_ptr(), W * 3, fb, borrow=True)
        pos += m; k -= m
step(252); st.sync()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(frames); st.sync()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in e]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
span = ev[-1]["ts"] + ev[-1]["dur"] - t0
print(f"{len(ev)} device activities over {span:.0f} us = {span / frames:.1f} us per frame")
by = {}
for e in ev:
    nm = e["name"].split("(")[0].replace("void ", "")[:28]
    by.setdefault(nm, []).append(e["dur"])
for nm, d in sorted(by.items(), key=lambda kv: -sum(kv[1])):
    print(f"  {nm:28s} n={len(d):4d} mean {np.mean(d):6.1f} us  min {min(d):6.1f}  max {max(d):6.1f}  busy/frame {sum(d) / frames:6.1f}")
streams = {}
for e in ev:
    streams.setdefault(e["args"].get("stream"), []).append(e)
for s, es in streams.items():
    print(f"stream {s}: {len(es)} activities, busy {sum(x['dur'] for x in es) / span * 100:.0f}% : {sorted(set(x['name'].split('(')[0][:20] for x in es))}")
mid = len(ev) // 2
print("--- timeline excerpt (us from excerpt start; stream; kernel; duration)")
base = ev[mid]["ts"]
for e in ev[mid: mid + 60]:
    print(f"{e['ts'] - base:8.1f} {e['ts'] + e['dur'] - base:8.1f}  s{e['args'].get('stream')}  {e['name'].split('(')[0].replace('void ', '')[:24]:24s} {e['dur']:6.1f}")
