"""Generates tests/golden/*.npz from the oracle (cv2 4.13.0, setUseOptimized(False)).

TEST INFRASTRUCTURE ONLY.  Run:  python -m oracle.make_golden
The reference holds no golden vectors for this path; these are outputs of
the reference's own dependency (OpenCV) driven by the restated host logic, on the seeded
synthetic clips of SURVEY.md §8d.  Summaries only (corner lists, LK points, inlier masks,
transforms, matrices, frame CRCs) so the fixtures stay small.
"""
from __future__ import annotations

import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.stabilizer_ref import Parameters, run_clip  # noqa: E402
import synthclip as synth  # noqa: E402

CASES = {
    # BASELINE.json configs[0]: 1280x720, 300 frames, defaults (GFTT 200, LK 3 lvls, smooth win 30)
    "cfg1_720p_default": dict(w=1280, h=720, n=300, seed=1234, params=Parameters()),
    # configs[1]: 1080p live stream, smoothing radius 15
    "cfg2_1080p_r15": dict(w=1920, h=1080, n=48, seed=2000, params=Parameters(smoothingRadius=15)),
    # configs[2] (Stabilizer part): 4K, cropNZoom + borderSize 30
    "cfg3_4k_cropzoom": dict(w=3840, h=2160, n=10, seed=3000,
                             params=Parameters(smoothingRadius=5, cropNZoom=True, borderSize=30)),
    # border path: copyMakeBorder reflect, output grows by 2b
    "border_reflect_720p": dict(w=1280, h=720, n=12, seed=77,
                                params=Parameters(smoothingRadius=5, borderType="reflect", borderSize=24)),
    "gaussian_720p": dict(w=1280, h=720, n=40, seed=78,
                          params=Parameters(smoothingRadius=10, smoothingMethod="gaussian", gaussianSigma=2.0)),
    "kalman_hlock_720p": dict(w=1280, h=720, n=40, seed=79,
                              params=Parameters(smoothingRadius=10, smoothingMethod="kalman", horizonLock=True)),
}


def crc(a: np.ndarray) -> int:
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def pack(recs_pts, width=200):
    n = len(recs_pts)
    out = np.full((n, width, 2), np.nan, np.float32)
    cnt = np.zeros(n, np.int32)
    for i, p in enumerate(recs_pts):
        if p is None:
            cnt[i] = -1
            continue
        cnt[i] = len(p)
        out[i, :len(p)] = p
    return out, cnt


def build(name, w, h, n, seed, params):
    clip = synth.make_clip(w, h, n, seed)
    outs, st = run_clip(clip, params)
    fr = st.frame_records
    prev, prev_n = pack([r.prev_pts for r in fr])
    nxt, _ = pack([r.next_pts for r in fr])
    det, det_n = pack([r.detected for r in fr])
    status = np.zeros((len(fr), 200), np.uint8)
    mask = np.zeros((len(fr), 200), np.uint8)
    mask_n = np.zeros(len(fr), np.int32)
    for i, r in enumerate(fr):
        status[i, :len(r.status)] = r.status
        if r.inlier_mask is None:
            mask_n[i] = -1
        else:
            mask_n[i] = len(r.inlier_mask)
            mask[i, :len(r.inlier_mask)] = r.inlier_mask
    o = st.output_records
    T = np.stack([np.zeros((2, 3), np.float32) if r.T is None else r.T for r in o])
    data = dict(
        input_crc=np.array([crc(f) for f in clip], np.uint32),
        first_corners=st.first_corners,
        prev_pts=prev, prev_n=prev_n, next_pts=nxt, status=status,
        inlier_mask=mask, inlier_n=mask_n, detected=det, detected_n=det_n,
        affine=np.stack([np.full((2, 3), np.nan) if r.affine is None else r.affine for r in fr]),
        transforms=np.stack([r.transform for r in fr]),
        path=np.stack([r.path for r in fr]),
        out_index=np.array([r.index for r in o], np.int32),
        out_passthrough=np.array([r.T is None for r in o], np.uint8),
        out_radius=np.array([r.radius for r in o], np.int32),
        out_intent=np.array([r.intent for r in o], np.int32),
        out_smoothed=np.stack([r.smoothed for r in o]),
        out_T=T,
        out_shape=np.array([f.shape for f in outs], np.int32),
        out_crc=np.array([crc(f) for f in outs], np.uint32),
        # a thin slice of real pixels so a CRC mismatch can be localised
        out_row=np.stack([f[f.shape[0] // 2, 100:260, :] for f in outs]),
    )
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **data)
    print(name, "frames", n, "outputs", len(outs), os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    only = sys.argv[1:]
    for k, v in CASES.items():
        if only and k not in only:
            continue
        build(k, **v)
