import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__; __graft_entry__.build()
import video_stab_b200 as vsb
from oracle.stabilizer_ref import Parameters as RP, run_clip
w, h, b, dur, n = 644, 362, 10, 4, 40
clip = vsb.synth.make_clip(w, h, n, 91)
kw = dict(smoothingRadius=5, borderType="fade", borderSize=b, fadeDuration=dur, fadeAlpha=0.25)
ref, rst = run_clip(clip, RP(**kw))
st = vsb.Stabilizer(vsb.Parameters(**kw))
outs = [o for o in (st.stabilize(f) for f in clip) if o is not None]
while True:
    o = st.flush()
    if o is None: break
    outs.append(o)
for i, (a, r) in enumerate(zip(outs, ref)):
    d = np.abs(a.astype(int) - r.astype(int))
    if d.max() > 0:
        ys, xs, cs = np.nonzero(d)
        print(i, "max", d.max(), "count", len(ys), "rows", ys.min(), ys.max(), "cols", xs.min(), xs.max())
        y0, x0 = ys[0], xs[0]
        print("   got", a[y0, max(x0-2,0):x0+3].tolist(), "\n   ref", r[y0, max(x0-2,0):x0+3].tolist())
for i in range(len(rst.output_records)):
    To = rst.output_records[i].T
    if To is None: continue
    Tg = np.array(st.output_record(i).T, np.float32).reshape(2, 3)
    if not np.array_equal(To.view(np.uint32), Tg.view(np.uint32)):
        print("T differs at output", i, (To - Tg).tolist())
import torch, cv2
cv2.setUseOptimized(False)
T = np.array(st.output_record(33).T, np.float32).reshape(2, 3)
print("T", T.tolist())
rng = np.random.default_rng(1)
img = rng.integers(0, 256, (382, 664, 3), dtype=np.uint8)
refw = cv2.warpAffine(img, T, (664, 382), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
got = vsb.kernels.warp_affine(torch.from_numpy(img).cuda(), T[None]).cpu().numpy()
d = np.abs(got.astype(int) - refw.astype(int))
ys, xs, cs = np.nonzero(d)
print("kernel-level mismatches", len(ys), list(zip(ys[:10], xs[:10], cs[:10])), d.max())
# fixed-point coordinates of the bad pixel
M = T.astype(np.float64); D = 1.0 / (M[0,0]*M[1,1] - M[0,1]*M[1,0])
A11 = M[1,1]*D; A22 = M[0,0]*D; m = np.array([[A11, -M[0,1]*D, 0],[-M[1,0]*D, A22, 0]])
m[0,2] = -m[0,0]*M[0,2] - m[0,1]*M[1,2]; m[1,2] = -m[1,0]*M[0,2] - m[1,1]*M[1,2]
for (y, x) in sorted(set(zip(ys.tolist(), xs.tolist())))[:4]:
    X0 = int(np.rint((m[0,1]*y + m[0,2])*1024)) + 16; Y0 = int(np.rint((m[1,1]*y + m[1,2])*1024)) + 16
    X = (X0 + int(np.rint(m[0,0]*x*1024))) >> 5; Y = (Y0 + int(np.rint(m[1,0]*x*1024))) >> 5
    print("pixel", (y, x), "sx,sy", X >> 5, Y >> 5, "ax,ay", X & 31, Y & 31, "got", got[y, x], "ref", refw[y, x])
