"""A small end-to-end pass for compute-sanitizer (racecheck / initcheck / memcheck): the seven-stream single-handle engine on device
frames (copy mode and borrowed), the pipelined host path, a 3-stream lock-step batch, crop-n-zoom and a bordered warp, clip mode,
roll correction and auto zoom-crop.  Small frames: the tools slow kernels down 10-100x.

    compute-sanitizer --tool racecheck python profiles/tools/sanitize_run.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import video_stab_b200 as vsb
import synthclip

w, h, n = 640, 360, 28
fb = w * h * 3
clip = synthclip.make_clip(w, h, n, 77)
d = torch.from_numpy(clip).cuda()
out = torch.zeros((n + 4, h + 48, w + 48, 3), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()


def run(params, borrow):
    st = vsb.Stabilizer(params)
    b = params.borderSize if (params.borderSize > 0 and not params.cropNZoom) else 0
    cap = (w + 2 * b) * (h + 2 * b) * 3
    k = 0
    for i in range(n):
        if st.push_device(d[i].data_ptr(), w, h, w * 3, out[k].data_ptr(), 0, cap, borrow=borrow) is not None:
            k += 1
    while st.flush_device(out[k].data_ptr(), 0, cap) is not None:
        k += 1
    st.sync()
    return k


print("plain", run(vsb.Parameters(smoothingRadius=6), True), run(vsb.Parameters(smoothingRadius=6), False))
print("cropzoom", run(vsb.Parameters(smoothingRadius=5, cropNZoom=True, borderSize=16), True))
print("border", run(vsb.Parameters(smoothingRadius=5, borderType="reflect", borderSize=12), True))
print("gaussian", run(vsb.Parameters(smoothingRadius=8, smoothingMethod="gaussian"), True))
st = vsb.Stabilizer(vsb.Parameters(smoothingRadius=6))
got = st.stabilize_many(np.ascontiguousarray(clip[:20]))
print("host pipe", len(got), len(st.flush_many(32)))
S = 3
batch = vsb.StabilizerBatch(vsb.Parameters(smoothingRadius=5), S)
ob = torch.zeros((S, n, h, w, 3), dtype=torch.uint8, device="cuda")
k = 0
for i in range(16):
    if batch.push_device([d[(i + s) % n].data_ptr() for s in range(S)], w, h, w * 3, [ob[s, k].data_ptr() for s in range(S)], w * 3, fb, borrow=True) is not None:
        k += 1
batch.sync()
print("batch", k)
o2, tr = vsb.offline.stabilize_clip(d, vsb.Parameters(smoothingRadius=6), n_chunks=3)
print("clip mode", tuple(o2.shape), tr.shape)
roll = vsb.RollCorrection(vsb.RollParameters(angleFilterMin=-70.0, angleFilterMax=70.0))
hz = synthclip.horizon_clip(w, h, 4, 5)
for f in hz:
    r = roll.autoCorrectRoll(f)
print("roll", roll.state())
z = vsb.AutoZoomCrop.autoZoomCrop(synthclip.black_corner_frame(w, h, 3, 4.0, (5.0, -7.0)))
print("zoom", z.shape)
torch.cuda.synchronize()
print("done")
