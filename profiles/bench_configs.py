"""Throughput of the other BASELINE configs on one B200 (device-resident frames), plus PCIe copy bandwidth.
Not the bench contract (bench.py is): these are the parity-test configurations, timed for the profiles.
    python profiles/bench_configs.py [cfg3] [cfg4] [cfg5] [pcie]"""
import json
import os
import sys
import time

import numpy as np

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
import video_stab_b200 as vsb  # noqa: E402
import synthclip

dev = torch.device("cuda", 0)
which = set(sys.argv[1:]) or {"cfg3", "cfg4", "cfg5", "pcie"}
res = {}


def pingpong(n):
    return list(range(n)) + list(range(n - 2, 0, -1))


def timed(fn, sync):
    sync()
    t0 = time.perf_counter()
    fn()
    sync()
    return time.perf_counter() - t0


if "pcie" in which:
    a = torch.empty(64 * 1920 * 1080 * 3, dtype=torch.uint8).pin_memory()
    b = torch.empty_like(a).pin_memory()
    da, db = torch.empty_like(a, device=dev), torch.empty_like(a, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(2):
        da.copy_(a, non_blocking=True); b.copy_(db, non_blocking=True)
    torch.cuda.synchronize()
    t = timed(lambda: da.copy_(a, non_blocking=True), torch.cuda.synchronize)
    res["h2d_GBs"] = a.numel() / t / 1e9
    t = timed(lambda: b.copy_(db, non_blocking=True), torch.cuda.synchronize)
    res["d2h_GBs"] = a.numel() / t / 1e9

    def both():
        with torch.cuda.stream(s1):
            da.copy_(a, non_blocking=True)
        with torch.cuda.stream(s2):
            b.copy_(db, non_blocking=True)
    t = timed(both, torch.cuda.synchronize)
    res["bidir_each_GBs"] = a.numel() / t / 1e9
    del a, b, da, db

if "cfg3" in which:
    W, H, n = 3840, 2160, 24
    clip = torch.from_numpy(synthclip.make_clip(W, H, n, 3000)).to(dev)
    out = torch.empty((n, H, W, 3), dtype=torch.uint8, device=dev)
    order = pingpong(n)
    st = vsb.Stabilizer(vsb.Parameters(smoothingRadius=15, cropNZoom=True, borderSize=30))
    pos = 0

    def run(k):
        global pos
        for _ in range(k):
            i = order[pos % len(order)]
            st.push_device(clip[i].data_ptr(), W, H, W * 3, out[pos % n].data_ptr(), W * 3, H * W * 3, borrow=True)
            pos += 1
    run(60)
    t = timed(lambda: run(240), st.sync)
    res["cfg3_4k_cropzoom_fps"] = 240 / t
    st.set_timing(True); run(48); res["cfg3_stage_us"] = {k: (v["ms"] / v["count"] * 1e3 if v["count"] else None) for k, v in st.stage_times().items()}
    del st, clip, out

if "cfg4" in which:
    W, H, n, S = 1920, 1080, 12, 64
    clips = [torch.from_numpy(synthclip.make_clip(W, H, n, 2000 + s)).to(dev) for s in range(S)]
    outs = torch.empty((S, H, W, 3), dtype=torch.uint8, device=dev)
    order = pingpong(n)
    sb = vsb.StabilizerBatch(vsb.Parameters(smoothingRadius=15), S)
    pos = 0

    def runb(k):
        global pos
        for _ in range(k):
            i = order[pos % len(order)]
            sb.push_device([c[i].data_ptr() for c in clips], W, H, W * 3, [outs[s].data_ptr() for s in range(S)], W * 3,
                           H * W * 3, borrow=True)
            pos += 1
    runb(40)
    t = timed(lambda: runb(100), sb.sync)
    res["cfg4_64x1080p_aggregate_fps"] = 100 * S / t
    sb.set_timing(True); runb(20); res["cfg4_stage_us_per_step"] = {k: (v["ms"] / v["count"] * 1e3 if v["count"] else None) for k, v in sb.stage_times().items()}
    del sb, clips, outs

if "cfg5" in which:
    W, H, n = 1920, 1080, 768
    base = synthclip.make_clip(W, H, 64, 5000)
    idx = [pingpong(64)[k % 126] for k in range(n)]
    clip = torch.from_numpy(base).to(dev)[torch.tensor(idx, device=dev)]
    out = torch.empty_like(clip)
    for chunks in (1, 8):
        vsb.offline.stabilize_clip(clip[:64], vsb.Parameters(smoothingRadius=15), out=out[:64], n_chunks=1)
        t = timed(lambda: vsb.offline.stabilize_clip(clip, vsb.Parameters(smoothingRadius=15), out=out, n_chunks=chunks),
                  torch.cuda.synchronize)
        res[f"cfg5_offline_{n}f_{chunks}chunk_fps"] = n / t

print(json.dumps(res, indent=1))
