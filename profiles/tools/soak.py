"""Soak test of the asynchronous multi-stream engine: many repetitions of long clips through the device path
(push_many_device / push_device, borrowed and ring-copied), each compared frame for frame (CRC) and transform for
transform with a pass in which every kernel is serialised on ONE stream (VS_SINGLE_STREAM=1).  A missing dependency
between the handle's streams would show up as a difference."""
import os, sys, zlib
import numpy as np, torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__; __graft_entry__.build()
import video_stab_b200 as vsb
import synthclip

def crcs(t):
    return [zlib.crc32(f.tobytes()) for f in t.cpu().numpy()]

def run(W, H, n, params, reps, mode):
    clip = torch.from_numpy(synthclip.make_clip(W, H, 48, 777)).cuda()
    order = list(range(48)) + list(range(46, 0, -1))
    seq = clip[torch.tensor([order[k % len(order)] for k in range(n)], device="cuda")].contiguous()
    torch.cuda.synchronize()
    b = params.borderSize if (params.borderSize > 0 and not params.cropNZoom) else 0
    fb, ob = W * H * 3, (W + 2 * b) * (H + 2 * b) * 3
    want = None
    for rep in range(reps + 1):
        out = torch.zeros((n, ob), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        # pass 0 is the reference: the same calls with every kernel serialised on one stream (VS_SINGLE_STREAM)
        if rep == 0:
            os.environ["VS_SINGLE_STREAM"] = "1"
        else:
            os.environ.pop("VS_SINGLE_STREAM", None)
        st = vsb.Stabilizer(params)
        if mode == "many":
            k = st.push_many_device(seq.data_ptr(), fb, n, W, H, W * 3, out.data_ptr(), (W + 2 * b) * 3, ob, borrow=True)
        else:
            k = 0
            for i in range(n):
                if st.push_device(seq[i].data_ptr(), W, H, W * 3, out[k].data_ptr(), (W + 2 * b) * 3, ob, borrow=(mode == "borrow")) is not None:
                    k += 1
        while k < n and st.flush_device(out[k].data_ptr(), (W + 2 * b) * 3, ob) is not None:
            k += 1
        st.sync()
        got = crcs(out[:k])
        recs = [tuple(st.frame_record(i).transform) for i in range(0, st.counts()[0], 7)]
        if want is None:
            want = (got, recs)
        else:
            bad = [i for i in range(min(len(got), len(want[0]))) if got[i] != want[0][i]]
            assert len(got) == len(want[0]) and not bad, f"{W}x{H} {mode} rep {rep}: frames {bad[:8]} differ"
            assert recs == want[1], f"{W}x{H} {mode} rep {rep}: transforms differ"
    print(f"ok {W}x{H} n={n} reps={reps} mode={mode} outputs={len(want[0])}", flush=True)

P = vsb.Parameters
run(1920, 1080, 400, P(smoothingRadius=15), 4, "many")
run(1920, 1080, 300, P(smoothingRadius=15), 3, "borrow")
run(1920, 1080, 200, P(smoothingRadius=5), 3, "ring")
run(1280, 720, 400, P(smoothingRadius=30, borderType="reflect", borderSize=20), 3, "many")
run(3840, 2160, 80, P(smoothingRadius=5, cropNZoom=True, borderSize=30), 3, "many")
run(1280, 720, 300, P(smoothingRadius=8, smoothingMethod="kalman", horizonLock=True), 3, "many")
print("soak passed")
