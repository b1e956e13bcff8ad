import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__; __graft_entry__.build()
import video_stab_b200 as vsb
import cv2
cv2.setUseOptimized(False)
def run(clip, params):
    st = vsb.Stabilizer(params)
    outs=[]
    for f in clip:
        o = st.stabilize(f)
        if o is not None: outs.append(o)
    while True:
        o = st.flush()
        if o is None: break
        outs.append(o)
    return outs, st
clip = vsb.synth.make_clip(640,360,30,2000)
p = vsb.Parameters(smoothingRadius=6)
a, sa = run(clip,p); b, sb = run(clip,p)
nf,_ = sa.counts()
for i in range(nf):
    pa, pb = sa.frame_points(i), sb.frame_points(i)
    ra, rb = sa.frame_record(i), sb.frame_record(i)
    for k in ('prev','next','status','inlier_mask','detected'):
        x,y = pa[k],pb[k]
        if (x is None) != (y is None) or (x is not None and not np.array_equal(x,y)):
            print('frame',i,'differs in',k, None if x is None else x.shape, None if y is None else y.shape)
            if k=='detected' and x is not None and y is not None:
                n=min(len(x),len(y)); d=np.nonzero((x[:n]!=y[:n]).any(axis=1))[0]
                print('  first diff idx',d[:5], x[d[:3]], y[d[:3]])
                g = vsb.kernels.gray_pyramid(torch.from_numpy(clip[i+1]).cuda())[0]
                ref = cv2.goodFeaturesToTrack(g.cpu().numpy(),200,0.02,15.0,None,blockSize=3).reshape(-1,2)
                print('  a==ref',np.array_equal(x,ref),'b==ref',np.array_equal(y,ref))
            break
for i,(x,y) in enumerate(zip(a,b)):
    if not np.array_equal(x,y): print('out',i,'differs', np.abs(x.astype(int)-y).max())
# repeated gftt on same image
g = vsb.kernels.gray_pyramid(torch.from_numpy(clip[5]).cuda())[0]
ref = cv2.goodFeaturesToTrack(g.cpu().numpy(),200,0.02,15.0,None,blockSize=3).reshape(-1,2)
bad=0
for t in range(50):
    got = vsb.kernels.good_features(g,200,0.02,15.0)
    if not np.array_equal(got,ref): bad+=1
print('gftt repeat mismatches',bad,'/50')
