"""Rates of the virtual-canvas stage at 1080p: the stage on its own (vs_canvas_apply_device) with the default 1.5x canvas (no
fills: a shifted copy) and with a 1.3x canvas (every pixel blended with a warped + resized older frame), and the live stabilizer
with enable_virtual_canvas.  Usage: python profiles/tools/canvas_rate.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import video_stab_b200 as vsb  # noqa: E402
import synthclip  # noqa: E402

W, H, N = 1920, 1080, 200
clip = torch.from_numpy(np.maximum(synthclip.make_clip(W, H, 24, 3), 6)).cuda()
rng = np.random.default_rng(0)
corr = np.stack([rng.uniform(-30, 30, N), rng.uniform(-20, 20, N), rng.uniform(-0.02, 0.02, N)], axis=1).astype(np.float32)
out = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
for name, kw in (("canvas 1.5x (no fill)", dict()), ("canvas 1.3x (whole-canvas fill)", dict(canvasScaleFactor=1.3, adaptiveCanvasSize=False)),
                 ("canvas 1.0x (dark objects)", dict(canvasScaleFactor=1.0, adaptiveCanvasSize=False))):
    vc = vsb.VirtualCanvas(vsb.Parameters(enableVirtualCanvas=True, **kw))
    src = clip.clone()
    if "1.0x" in name:
        src[:, 300:360, 500:640] = 0
        src[:, 700:720, 100:900] = 0
    for k in range(40):
        vc.apply_device(src[k % 24].data_ptr(), W, H, W * 3, corr[k], out.data_ptr(), W * 3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(N):
        vc.apply_device(src[k % 24].data_ptr(), W, H, W * 3, corr[k], out.data_ptr(), W * 3)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name}: {N / dt:.0f} frames/s ({dt / N * 1e6:.0f} us/frame), regions filled {vc.info()['regions_filled']}")
for name, kw in (("stabilizer + canvas 1.5x", dict()), ("stabilizer + canvas 1.3x", dict(canvasScaleFactor=1.3, adaptiveCanvasSize=False)),
                 ("stabilizer, no canvas", None)):
    P = vsb.Parameters(smoothingRadius=15, **({} if kw is None else dict(enableVirtualCanvas=True, **kw)))
    st = vsb.Stabilizer(P)
    outs = torch.zeros((N + 64, H, W, 3), dtype=torch.uint8, device="cuda")
    fb = W * H * 3
    seq = clip[torch.arange(N + 64, device="cuda") % 24].contiguous()
    st.push_many_device(seq.data_ptr(), fb, 64, W, H, W * 3, outs.data_ptr(), W * 3, fb, borrow=True)
    st.sync()
    t0 = time.perf_counter()
    k = st.push_many_device(seq[64].data_ptr(), fb, N, W, H, W * 3, outs.data_ptr(), W * 3, fb, borrow=True)
    st.sync()
    dt = time.perf_counter() - t0
    print(f"{name}: {k / dt:.0f} frames/s ({dt / k * 1e6:.0f} us/frame)")
