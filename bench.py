#!/usr/bin/env python
"""bench.py — headline benchmark of the stabilization hot path (BASELINE.json metric:
"1080p stabilized frames/s at 1/2/4/8 B200; warp kernel % of HBM peak").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference's own CPU path (oracle/_ref: Stabilizer.cpp itself)

N = 1 (BASELINE configs[1]): one 1920x1080 live stream, smoothing radius 15.  A *step* is one pass of the hot path over
256 consecutive frames of that stream, read in place from a 126-frame sequence resident in HBM (784 MB > the 126 MB L2).
N > 1 (BASELINE configs[3]): 64 concurrent 1080p streams sharded s mod N over the ranks, each rank advancing its
64/N streams in lock-step through one StabilizerBatch; a step advances every stream by 32 frames (2048 frames per step
in total, fixed as N grows: strong scaling).  No collective on this data path.  The same figure for 64 streams on ONE
GPU is in `config4` of the N = 1 line, so the strong-scaling curve has its base point.

Every line also carries `config5` (BASELINE configs[4]): an 18 000-frame 1080p clip, generated on the device, cut into
temporal chunks over the ranks: per-chunk analysis -> ONE NCCL all-gather of the transforms (12 bytes per frame, device
to device, inside the timed region) -> path rebuild, smoothing and warp of the rank's own frames; before the line is
printed rank 0 streams the same clip through a single handle and requires the stitched transforms to be bit-equal.

  value   frames/s with the frames already resident in HBM (device-pointer API, frames borrowed in place)
  e2e     frames/s through the reference-shaped host API (host frames in, host frames out; H2D and D2H copies inside
          the timed region), with the plain-copy ceiling of the same bytes measured beside it
  roofline  the warp kernel (the kernel the metric names): algorithmic bytes 2*3*W*H per frame divided by its average
          launch duration, 64 frames per launch, timed alone with CUDA events
  cpu_baseline  the reference's own Stabilizer.cpp (oracle/_ref, OpenCV = the cv2 wheel) on this box's host cores
"""
from __future__ import annotations

import argparse
import json
import os

# Hardware work queues of the process (must be set before the CUDA context exists).  Measured on this pool (profiles/r02_summary.md):
# with 32 queues two lock-step groups on one GPU (N = 8: 14 streams) run 5 % faster than with the default 8, but page-locked
# host <-> device copies - the plain duplex ceiling itself - lose 8 - 24 % and become erratic; one live stream and the 64-stream
# batch do not care.  So: the default at N = 1, where the host-buffer figure is the headline; 32 when a rank drives several groups.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else "8")
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1920, 1080
FRAMES_PER_STEP = 256          # config 2: frames of the single stream per step
SEQ_FRAMES = 64                # distinct generated frames; laid out ping-pong (126 frames) so the sequence loops without a jump
SMOOTHING_RADIUS = 15
C4_STREAMS = 64                # config 4: concurrent streams in total
C4_FRAMES = 32                 # frames every stream advances per step  (64 x 32 = 2048 frames per step, all ranks together: ~3 ms at N = 8)
C4_CLIP = 24                   # generated frames per stream (ping-pong: 46-frame loop)
C5_FRAMES = 18000              # config 5: 10 minutes at 30 fps
C5_CHUNKS = 8
WARP_BATCH = 64                # frames per launch of the roofline measurement
METRIC = "stabilized_1080p_frames_per_s"
UNIT = "frames/s"


def workload_config(n_gpus: int) -> dict:
    """`config` of the JSON line — the same dict for this repo's arm and for the reference arm."""
    if n_gpus == 1:
        return {"workload": "1920x1080 single live stream, smoothing radius 15 (BASELINE configs[1])",
                "frames_per_step": FRAMES_PER_STEP, "gftt": "200 pts every 2nd frame", "lk": "15x15, 3 levels"}
    return {"workload": f"{C4_STREAMS} concurrent 1920x1080 streams, smoothing radius 15, sharded s mod N over the GPUs "
                        "(BASELINE configs[3]); no collective on the data path",
            "frames_per_step": C4_STREAMS * C4_FRAMES, "streams": C4_STREAMS, "frames_per_stream_per_step": C4_FRAMES,
            "gftt": "200 pts every 2nd frame", "lk": "15x15, 3 levels"}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks and throttle reasons while the timed region runs: NVML from a thread (a sample every ~4 ms,
    so even a 30 ms region at N = 8 is covered), nvidia-smi -lms as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index
        self.nvml = None
        self.stop_flag = False
        self.sm, self.mask, self.sm_max = [], 0, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.sm.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mask |= int(reasons(self.h))
            except Exception:
                pass
            time.sleep(0.004)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nvml:
            self.stop_flag = True
            self.t.join(timeout=1)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max,
                    "reasons": [n for n, b in self.BITS if self.mask & b], "samples": len(sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvidia-smi"}


def _pingpong(n):
    """frame order that loops without a jump: 0..n-1, n-2..1"""
    return list(range(n)) + list(range(n - 2, 0, -1))


# ------------------------------------------------------------------------------------------- reference arm
def _make_reference(use_ref: bool):
    """Factory for the CPU arm: the reference's own Stabilizer.cpp (oracle/_ref, kind "reference") when the compiled
    library is present, else its Python restatement (kind "port", bit-identical: tests/test_ref_pin.py)."""
    from oracle.stabilizer_ref import Parameters
    params = Parameters(smoothingRadius=SMOOTHING_RADIUS)
    if use_ref:
        from oracle import ref_lib

        def make():
            st = ref_lib.RefStabilizer(params, use_optimized=True, record=False)
            st.stabilize = st.stabilize_nocopy          # timing path: no harness copies of the input or the output frame
            return st
        return make
    from oracle.stabilizer_ref import StabilizerRef
    return lambda: StabilizerRef(params, use_optimized=True)


def _reference_kind():
    import oracle
    use_ref = oracle.reference_available()
    import cv2
    what = ("the reference's own src/Stabilizer.cpp compiled unmodified (oracle/_ref), OpenCV calls served by cv2 "
            if use_ref else "Stabilizer.cpp host logic restated in Python (oracle/stabilizer_ref.py) over cv2 ")
    return use_ref, ("reference" if use_ref else "port"), what + f"{cv2.__version__}, optimized paths on"


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path, all host threads, on this arm's config.  Rank 0 only."""
    if rank != 0:
        return
    import cv2
    import synthclip
    cores = os.cpu_count() or 1
    cv2.setNumThreads(cores)
    use_ref, kind, what = _reference_kind()
    make = _make_reference(use_ref)
    n = args.gpus
    if n == 1:
        clip = synthclip.make_clip(W, H, 32, seed=2000)
        order = _pingpong(len(clip))
        st = make()
        k = 0
        frames_per_step = FRAMES_PER_STEP

        def step():
            nonlocal k
            for _ in range(FRAMES_PER_STEP):
                st.stabilize(clip[order[k % len(order)]])
                k += 1
    else:
        # 64 independent streams (8 seeded clips x 8 start offsets), every stream advanced by C4_FRAMES per step
        clips = [synthclip.make_clip(W, H, 32, seed=2000 + s) for s in range(8)]
        order = _pingpong(32)
        sts = [make() for _ in range(C4_STREAMS)]
        pos = [s // 8 for s in range(C4_STREAMS)]
        frames_per_step = C4_STREAMS * C4_FRAMES

        def step():
            for s, st in enumerate(sts):
                c = clips[s % 8]
                for _ in range(C4_FRAMES):
                    st.stabilize(c[order[pos[s] % len(order)]])
                    pos[s] += 1
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    fps = args.steps * frames_per_step / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": n, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak" if n == 1 else "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(n),
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{args.steps} steps x {frames_per_step} frames 1080p; {what}, {cores} threads; one CPU process "
                                   "(rank 0) whatever N is"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(seconds_budget=12.0):
    import cv2
    import synthclip
    cores = os.cpu_count() or 1
    use_ref, kind, what = _reference_kind()
    make = _make_reference(use_ref)
    cv2.setNumThreads(cores)
    clip = synthclip.make_clip(W, H, 32, seed=2000)
    order = _pingpong(len(clip))
    st = make()
    for k in range(20):
        st.stabilize(clip[order[k % len(order)]])
    n, t0 = 0, time.perf_counter()
    while n < 4000 and time.perf_counter() - t0 < seconds_budget:
        st.stabilize(clip[order[(20 + n) % len(order)]])
        n += 1
    dt = time.perf_counter() - t0
    # the same path on ONE host thread (SURVEY.md section 8d asks for both), a shorter sample
    cv2.setNumThreads(1)
    st1 = make()
    for k in range(5):
        st1.stabilize(clip[order[k % len(order)]])
    n1, t1 = 0, time.perf_counter()
    while n1 < 400 and time.perf_counter() - t1 < 4.0:
        st1.stabilize(clip[order[(5 + n1) % len(order)]])
        n1 += 1
    dt1 = time.perf_counter() - t1
    cv2.setNumThreads(cores)
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{n} frames 1080p (~{dt:.0f} s) after 20 warm-up frames; {what}, {cores} threads, wall clock",
            "single_thread": {"value": n1 / dt1, "unit": UNIT, "cores": 1, "sample": f"{n1} frames (~{dt1:.0f} s), cv2.setNumThreads(1)"}}


def bind_to_gpu_numa_node(index: int, local_world: int):
    """Pin this process to host cores near its GPU: the cores of the GPU's NUMA node when sysfs reports one, else (a
    single-node VM) an even share of the cores per local rank, so that eight enqueue threads do not migrate over each
    other.  Page-locked buffers allocated afterwards are first-touched there."""
    try:
        before = sorted(os.sched_getaffinity(0))
        node, cpus = -1, []
        try:
            import pynvml
            pynvml.nvmlInit()
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
            bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
            if len(bus.split(":")[0]) == 8:
                bus = bus[4:]                               # 00000000:1b:00.0 -> 0000:1b:00.0
            node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
            if node >= 0:
                for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                    a, _, b = part.partition("-")
                    cpus += list(range(int(a), int(b or a) + 1))
        except Exception:
            node = -1
        allowed = sorted(set(cpus) & set(before)) if node >= 0 else before
        if local_world > 1:
            # an even, disjoint share of the allowed cores per local rank
            share = max(1, len(allowed) // local_world)
            mine = allowed[(index % local_world) * share:(index % local_world) * share + share] or allowed
        else:
            mine = allowed
        os.sched_setaffinity(0, mine)
        return {"node": node if node >= 0 else None, "cpus": len(mine), "first_cpu": mine[0], "_restore": before}
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args, rank, world, local_rank):
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    numa = bind_to_gpu_numa_node(local_rank, local_world)
    all_cpus = numa.pop("_restore") if numa else None
    import ctypes as C

    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__
    __graft_entry__.build()
    import synthclip
    import video_stab_b200 as vsb
    from video_stab_b200 import offline
    from video_stab_b200._capi import lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    parts = set(args.parts.split(","))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(*vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def allsum(*vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(x) for x in t]

    frame_bytes = H * W * 3
    params = vsb.Parameters(smoothingRadius=SMOOTHING_RADIUS)
    ow, oh, produced = C.c_int(), C.c_int(), C.c_int()
    line = {}
    clocks = None

    # ============================================================ config 2: one live stream per GPU (headline at N = 1)
    c2 = None
    if "config2" in parts and (world == 1 or args.config2_at_all_n):
        # the 64 generated frames laid out in ping-pong order (126 frames = 784 MB > L2): a step is a run of consecutive
        # frames, read in place, and the sequence loops without a jump
        clip_h = synthclip.make_clip(W, H, SEQ_FRAMES, seed=2000 + rank)
        order = _pingpong(SEQ_FRAMES)
        clip_d = torch.from_numpy(clip_h).to(dev)
        seq_d = clip_d[torch.tensor(order, device=dev)].contiguous()
        out_d = torch.empty((FRAMES_PER_STEP, H, W, 3), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        st = vsb.Stabilizer(params, device=local_rank)
        ext = torch.cuda.ExternalStream(st.stream, device=dev)
        pos = 0

        def step():
            # one step = the stabilize() loop over the next 256 frames of the sequence, one C-ABI call per run of
            # consecutive frames (frames read in place; outputs land in a 256-frame device ring)
            nonlocal pos
            done = 0
            while done < FRAMES_PER_STEP:
                a = pos % len(order)
                k = min(FRAMES_PER_STEP - done, len(order) - a)
                st.push_many_device(seq_d[a].data_ptr(), frame_bytes, k, W, H, W * 3, out_d[done].data_ptr(), W * 3, frame_bytes, borrow=True)
                pos += k
                done += k

        for _ in range(args.warmup):
            step()
        st.sync()
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        l0 = st.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(args.steps):
            step()
        st.join()                      # the public stream waits for the handle's analysis / detection streams
        e1.record(ext)
        st.sync()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = st.launch_count() - l0
        clocks = sampler.stop()
        # stage breakdown: a second pass with per-stage CUDA events on the library's streams
        st.set_timing(True)
        step()
        stages = st.stage_times()
        st.set_timing(False)
        stage_us = {k: (v["ms"] / v["count"] * 1e3 if v["count"] else None) for k, v in stages.items()}
        # per-frame latency of a LIVE stream: one frame pushed, then waited for - the whole
        # gray -> pyramid -> LK -> motion -> warp chain with nothing to overlap with
        lat = []
        for i in range(48):
            a = (pos + i) % len(order)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            st.push_device(seq_d[a].data_ptr(), W, H, W * 3, out_d[0].data_ptr(), W * 3, frame_bytes, borrow=True)
            st.sync()
            lat.append((time.perf_counter() - t0) * 1e6)
        pos += 48
        lat.sort()
        ms_max, = allmax(ms)
        c2 = {"value": args.steps * FRAMES_PER_STEP * world / (ms_max * 1e-3), "ms_per_step": ms_max / args.steps,
              "launches": int(launches), "stage_us_per_launch_group": stage_us,
              "latency_us_push_to_frame": {"median": lat[len(lat) // 2], "p90": lat[int(len(lat) * 0.9)],
                                           "what": "vs_stabilizer_push_device of ONE frame followed by a host wait, device frames (no copies)"},
              "api": "vs_stabilizer_push_many_device (borrowed device frames; the per-frame stabilize() loop runs inside the library)",
              "l2_policy": "inputs larger than L2 (126-frame sequence = 784 MB, frames read in place)"}
        del st, out_d, seq_d

    # ============================================================ config 4: 64 streams sharded s mod N, lock-step batches
    c4 = None
    if "config4" in parts:
        my_streams = [s for s in range(args.c4_streams) if s % world == rank]
        S = len(my_streams)
        loop = _pingpong(C4_CLIP)
        lanes = []
        for s in my_streams:                           # seeds 2000 .. 2063, generated on the device
            gen = synthclip.DeviceClip(W, H, C4_CLIP, 2000 + s, dev)
            fr = gen.frames(0, C4_CLIP)
            lanes.append(fr[torch.tensor(loop, device=dev)].contiguous())       # (46, H, W, 3)
            del gen, fr
        out4 = torch.empty((S, C4_FRAMES, H, W, 3), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        # The rank's streams advance in lock-step GROUPS, one StabilizerBatch (one launch per stage) per group.  With few streams per
        # GPU (N = 8: eight) two groups of four keep the SMs busier than one group of eight (profiles/tools/batch_groups.py:
        # 84.7 k -> 90.2 k frames/s on one GPU, both groups fed from one host thread); the host's enqueue time per lock-step
        # frame (32 - 42 us) is then the limit, so each group gets its own host thread (VS_C4_GROUPS overrides the count).
        n_groups = 2 if (S <= 8 and S % 2 == 0 and S >= 4) else 1
        if os.environ.get("VS_C4_GROUPS"):
            n_groups = max(1, min(S, int(os.environ["VS_C4_GROUPS"])))
            while S % n_groups:
                n_groups -= 1
        per = S // n_groups
        groups = []
        for g in range(n_groups):
            sbg = vsb.StabilizerBatch(params, per, device=local_rank)
            PA = C.c_void_p * per
            idx = range(g * per, (g + 1) * per)
            in_tab = [PA(*[lanes[i][f].data_ptr() for i in idx]) for f in range(len(loop))]       # pointer tables built once: the timed
            out_tab = [PA(*[out4[i, k].data_ptr() for i in idx]) for k in range(C4_FRAMES)]       # loop is one C call per lock-step frame
            groups.append((sbg, in_tab, out_tab))
        ext4 = torch.cuda.ExternalStream(groups[0][0].stream, device=dev)
        pos4 = 0

        def push_group(gi, first_pos, nsteps):
            # one host thread per lock-step group: handles are independent, the C call releases the GIL
            sbg, in_tab, out_tab = groups[gi]
            o_w, o_h, prod = C.c_int(), C.c_int(), C.c_int()
            p = first_pos
            for _ in range(nsteps):
                for k in range(C4_FRAMES):
                    rc = lib.vs_batch_push_device(sbg._h, in_tab[p % len(loop)], W, H, W * 3, out_tab[k], W * 3, frame_bytes, 1,
                                                  C.byref(o_w), C.byref(o_h), C.byref(prod))
                    assert rc == 0, lib.vs_last_error()
                    p += 1

        def run4(nsteps):
            nonlocal pos4
            if n_groups == 1:
                push_group(0, pos4, nsteps)
            else:
                import threading
                th = [threading.Thread(target=push_group, args=(gi, pos4, nsteps)) for gi in range(n_groups)]
                for t in th:
                    t.start()
                for t in th:
                    t.join()
            pos4 += nsteps * C4_FRAMES

        def join4():
            # every group's work ordered before the timing event on the first group's public stream
            for sbg, _, _ in groups:
                sbg.join()
            for sbg, _, _ in groups[1:]:
                evj = torch.cuda.Event()
                evj.record(torch.cuda.ExternalStream(sbg.stream, device=dev))
                ext4.wait_event(evj)

        run4(max(args.warmup, 3))
        for sbg, _, _ in groups:
            sbg.sync()
        barrier()
        sampler4 = ClockSampler(local_rank)
        sampler4.start()
        l0 = sum(g[0].launch_count() for g in groups)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        join4()
        e0.record(ext4)
        run4(args.steps)
        join4()
        e1.record(ext4)
        for sbg, _, _ in groups:
            sbg.sync()
        barrier()
        ms4 = e0.elapsed_time(e1)
        launches4 = sum(g[0].launch_count() for g in groups) - l0
        clocks4 = sampler4.stop()
        ms4_max, = allmax(ms4)
        l4_sum, = allsum(launches4)
        c4 = {"value": args.steps * args.c4_streams * C4_FRAMES / (ms4_max * 1e-3), "unit": UNIT, "ms_per_step": ms4_max / args.steps,
              "streams": args.c4_streams, "streams_per_gpu": S, "frames_per_step": args.c4_streams * C4_FRAMES, "n_gpus": world,
              "launches": int(l4_sum), "clocks": clocks4, "scaling": "strong",
              "sharding": f"stream s on GPU s mod N; {n_groups} lock-step group(s) of {per} streams per rank (one StabilizerBatch each, one launch per stage per group)",
              "groups_per_gpu": n_groups,
              "l2_policy": f"inputs larger than L2 ({S} streams x 46-frame loop = {S * 46 * frame_bytes / 1e9:.1f} GB per GPU, frames read in place)"}
        if world > 1:
            clocks = clocks4
        del groups, out4, lanes
        torch.cuda.empty_cache()

    # ============================================================ e2e: host frames in / host frames out, copies timed
    e2e = None
    if "e2e" in parts:
        clip_h = synthclip.make_clip(W, H, SEQ_FRAMES, seed=2000 + rank)
        order = _pingpong(SEQ_FRAMES)
        pin_in = torch.from_numpy(clip_h).pin_memory()
        seq = torch.stack([pin_in[i] for i in order]).pin_memory()           # 126 frames in ping-pong order
        if world == 1:
            # config 2 through vs_stabilizer_push_many (the stabilize() loop of the reference's file apps as one call; copy-in,
            # compute and copy-out overlap inside the call) and, beside it, the strictly synchronous per-frame vs_stabilizer_push
            n_e2e = FRAMES_PER_STEP
            pin_out = torch.empty((n_e2e, H, W, 3), dtype=torch.uint8).pin_memory()
            hs = [vsb.Stabilizer(params, device=local_rank)]
            epos = [0]

            def e2e_step():
                done = 0
                while done < n_e2e:
                    a = epos[0] % len(order)
                    k = min(n_e2e - done, len(order) - a)
                    rc = lib.vs_stabilizer_push_many(hs[0]._h, seq[a].data_ptr(), frame_bytes, k, W, H, W * 3, pin_out[done].data_ptr(),
                                                     W * 3, frame_bytes, C.byref(ow), C.byref(oh), C.byref(produced))
                    assert rc == 0, lib.vs_last_error()
                    done += k
                    epos[0] += k
            frames_e2e = n_e2e
            api = "vs_stabilizer_push_many: page-locked host frames in, host frames out (copy-in / compute / copy-out streams overlap inside the call)"
        else:
            # config 4 through host buffers: the rank's 64/N streams are independent handles, each fed C4_FRAMES host frames
            # per step through vs_stabilizer_push_many (all streams of a rank read the same page-locked 126-frame sequence at
            # different offsets; the handles are independent)
            S = len([s for s in range(C4_STREAMS) if s % world == rank])
            pin_out = torch.empty((C4_FRAMES, H, W, 3), dtype=torch.uint8).pin_memory()
            hs = [vsb.Stabilizer(params, device=local_rank) for _ in range(S)]
            epos = [3 * i for i in range(S)]

            def e2e_step():
                for i, hnd in enumerate(hs):
                    done = 0
                    while done < C4_FRAMES:
                        a = epos[i] % len(order)
                        k = min(C4_FRAMES - done, len(order) - a)
                        rc = lib.vs_stabilizer_push_many(hnd._h, seq[a].data_ptr(), frame_bytes, k, W, H, W * 3, pin_out[done].data_ptr(),
                                                         W * 3, frame_bytes, C.byref(ow), C.byref(oh), C.byref(produced))
                        assert rc == 0, lib.vs_last_error()
                        done += k
                        epos[i] += k
            frames_e2e = S * C4_FRAMES
            api = (f"vs_stabilizer_push_many on the rank's {S} independent stream handles, {C4_FRAMES} page-locked host frames in and out "
                   "per stream per step")
        e2e_steps = max(1, args.steps if world == 1 else max(2, args.steps // 2))
        for _ in range(3):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        for hnd in hs:
            hnd.sync()
        e2e_s = time.perf_counter() - t0
        barrier()
        sync_fps = None
        if world == 1:
            st3 = vsb.Stabilizer(params, device=local_rank)
            pin_out1 = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
            for rep in range(2):
                t0 = time.perf_counter()
                for i in range(64):
                    rc = lib.vs_stabilizer_push(st3._h, pin_in[order[i]].data_ptr(), W, H, W * 3, pin_out1.data_ptr(), W * 3, frame_bytes,
                                                C.byref(ow), C.byref(oh), C.byref(produced))
                    assert rc == 0
                st3.sync()
                sync_fps = 64 / (time.perf_counter() - t0)
            del st3
        # the ceiling of the same bytes: plain duplex copies (one H2D + one D2H stream), nothing else running
        dbuf = torch.empty((8, H, W, 3), dtype=torch.uint8, device=dev)
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        pin_o8 = torch.empty((8, H, W, 3), dtype=torch.uint8).pin_memory()
        barrier()
        t0 = time.perf_counter()
        reps = 24
        for r in range(reps):
            with torch.cuda.stream(s_in):
                dbuf.copy_(seq[(8 * r) % 112:(8 * r) % 112 + 8], non_blocking=True)
            with torch.cuda.stream(s_out):
                pin_o8.copy_(dbuf, non_blocking=True)
        torch.cuda.synchronize()
        copy_s = time.perf_counter() - t0
        barrier()
        e2e_max, copy_max = allmax(e2e_s, copy_s)
        fr_sum, = allsum(frames_e2e)
        e2e = {"value": e2e_steps * fr_sum / e2e_max, "unit": UNIT,
               "h2d_bytes_per_step": int(frame_bytes * fr_sum), "d2h_bytes_per_step": int(frame_bytes * fr_sum),
               "api": api, "steps": e2e_steps,
               "pcie_gb_per_s_each_way_per_gpu": e2e_steps * frames_e2e * frame_bytes / e2e_s / 1e9,
               "copy_ceiling": {"value": reps * 8 * world / copy_max, "unit": UNIT,
                                "what": "the same frames moved by plain duplex cudaMemcpyAsync (H2D and D2H streams) on every rank at "
                                        "once, no kernels: the host-memory / PCIe ceiling of the host-buffer API at this N",
                                "gb_per_s_each_way_per_gpu": reps * 8 * frame_bytes / copy_s / 1e9}}
        if sync_fps:
            e2e["sync_per_frame_push"] = {"value": sync_fps, "unit": UNIT, "api": "vs_stabilizer_push: one host frame in, one host frame out per call"}
        del hs, seq, pin_in, pin_out, dbuf
        torch.cuda.empty_cache()

    # ============================================================ roofline of the warp kernel and of the pyramid build
    roof = roof_pyr = None
    if "roofline" in parts:
        clip_r = torch.from_numpy(synthclip.make_clip(W, H, 8, seed=2000 + rank)).to(dev)
        clip_d = clip_r[torch.arange(WARP_BATCH, device=dev) % 8].contiguous()
        out_d = torch.empty((WARP_BATCH, H, W, 3), dtype=torch.uint8, device=dev)
        T = np.zeros((WARP_BATCH, 2, 3), np.float32)
        rng = np.random.default_rng(7)
        for i in range(WARP_BATCH):
            da = np.float32(rng.normal(0, 0.004))
            T[i] = [[np.cos(da), -np.sin(da), rng.normal(0, 3)], [np.sin(da), np.cos(da), rng.normal(0, 3)]]
        cur = torch.cuda.current_stream()
        for _ in range(3):
            vsb.kernels.warp_affine(clip_d, T, out=out_d, stream=cur.cuda_stream)
        torch.cuda.synchronize()
        reps = 10
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        for _ in range(reps):
            vsb.kernels.warp_affine(clip_d, T, out=out_d, stream=cur.cuda_stream)
        w1.record()
        torch.cuda.synchronize()
        warp_ms = w0.elapsed_time(w1) / reps
        traffic = None                                  # dram read+write bytes of that launch from the committed ncu capture
        try:
            cap = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_warp_ncu.json"))
            traffic = json.load(open(os.path.join(ROOT, "profiles", cap[-1])))["traffic_bytes_per_launch"]
        except Exception:
            pass
        alg_bytes = 2 * 3 * W * H * WARP_BATCH
        peak, peak_src = _peaks()
        achieved = alg_bytes / (warp_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "k_warp_quad (cv::warpAffine, 64 frames per launch, timed alone; 796 MB per launch > L2)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": warp_ms,
                "frac_of_nominal_8000": achieved / 8000.0}
        # pyramid build (north-star kernel 1): gray + both pyrDown levels of 64 frames in one lock-step launch pair
        sbp = vsb.StabilizerBatch(params, WARP_BATCH, device=local_rank)
        sb_ext = torch.cuda.ExternalStream(sbp.stream, device=dev)
        ptrs = [clip_d[i].data_ptr() for i in range(WARP_BATCH)]
        for _ in range(3):
            sbp.build_pyramids(ptrs, W, H, W * 3)
        sbp.sync()

        def timed_levels(p):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(sb_ext)
            for _ in range(reps):
                sbp.build_levels(ptrs, W, H, W * 3, p)
            b.record(sb_ext)
            sbp.sync()
            return a.elapsed_time(b) / reps
        pyr_ms, gray_ms, down_ms = timed_levels(3), timed_levels(1), timed_levels(2)
        pyr_bytes = (3 * W * H + 960 * 540 + 480 * 270 + 240 * 135) * WARP_BATCH
        gray_bytes = (3 * W * H + 960 * 540) * WARP_BATCH
        pyr_gbs = pyr_bytes / (pyr_ms * 1e-3) / 1e9
        gray_gbs = gray_bytes / (gray_ms * 1e-3) / 1e9
        roof_pyr = {"bound": "hbm", "kernel": "k_gray_half + k_pyrdown2 (resize + gray + 2 pyrDown levels, 64 frames per launch pair, timed alone)",
                    "achieved": pyr_gbs, "peak": peak, "unit": "GB/s", "frac": pyr_gbs / peak,
                    "algorithmic_bytes_per_launch": pyr_bytes, "ms_per_launch": pyr_ms,
                    "level0": {"kernel": "k_gray_half alone (resize + gray: the HBM-bound half, 3*W*H read + 518 400 written per frame)",
                               "achieved": gray_gbs, "frac": gray_gbs / peak, "algorithmic_bytes_per_launch": gray_bytes, "ms_per_launch": gray_ms},
                    "pyrdown": {"kernel": "k_pyrdown2 alone (level 0 -> levels 1, 2: 0.69 MB per frame, L2-resident, latency / issue bound)",
                                "ms_per_launch": down_ms}}
        del sbp, clip_d, out_d, clip_r
        torch.cuda.empty_cache()

    # ============================================================ config 3: one 4K stream, RollCorrection -> Stabilizer (crop-n-zoom) -> AutoZoomCrop
    c3 = None
    if "config3" in parts and world == 1:
        W3, H3, n3, ring3 = 3840, 2160, 24, 24
        fb3 = H3 * W3 * 3
        clip3 = torch.from_numpy(synthclip.horizon_clip(W3, H3, n3, 3000)).to(dev)
        loop3 = _pingpong(n3)
        rolled = torch.empty((ring3, H3, W3, 3), dtype=torch.uint8, device=dev)
        out3 = torch.empty((ring3, H3, W3, 3), dtype=torch.uint8, device=dev)
        zoomed = torch.empty((360, 640, 3), dtype=torch.uint8, device=dev)
        rp = vsb.RollParameters(angleFilterMin=-70.0, angleFilterMax=70.0, angleDecay=0.98)       # examples/config.yaml roll_correction:
        sp3 = vsb.Parameters(smoothingRadius=SMOOTHING_RADIUS, cropNZoom=True, borderSize=30)
        s_roll = torch.cuda.Stream(dev)

        def run3(frames, with_roll, with_stab, with_zoom):
            roll = vsb.RollCorrection(rp, device=local_rank)
            st3 = vsb.Stabilizer(sp3, device=local_rank)
            ext3 = torch.cuda.ExternalStream(st3.stream, device=dev)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            produced = 0
            lead = 24                                      # untimed: the handles allocate their 4K buffers on their first frames and at the first output (frame 15)
            for k in range(frames + lead):
                if k == lead:
                    st3.sync()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                src = clip3[loop3[k % len(loop3)]]
                cur_in = src
                if with_roll:
                    dst = rolled[k % ring3]
                    roll.correct_device(src.data_ptr(), W3, H3, W3 * 3, dst.data_ptr(), W3 * 3, s_roll.cuda_stream)
                    cur_in = dst
                if with_stab:
                    if with_roll:
                        ev = torch.cuda.Event()
                        ev.record(s_roll)
                        st3.wait_event(ev.cuda_event)
                    got = st3.push_device(cur_in.data_ptr(), W3, H3, W3 * 3, out3[k % ring3].data_ptr(), W3 * 3, fb3, borrow=True)
                    if got is not None:
                        produced += 1
                        if with_zoom:
                            st3.join()
                            ev2 = torch.cuda.Event()
                            ev2.record(ext3)
                            s_roll.wait_event(ev2)
                            vsb.AutoZoomCrop.crop_device(out3[k % ring3].data_ptr(), W3, H3, W3 * 3, zoomed.data_ptr(), 640 * 3, 640 * 360 * 3, s_roll.cuda_stream)
                elif with_zoom:
                    vsb.AutoZoomCrop.crop_device(cur_in.data_ptr(), W3, H3, W3 * 3, zoomed.data_ptr(), 640 * 3, 640 * 360 * 3, s_roll.cuda_stream)
            st3.sync()
            torch.cuda.synchronize()
            return frames / (time.perf_counter() - t0)
        run3(24, True, True, True)                         # warm-up (allocations, first-use set-up)
        nfr = 96
        c3 = {"workload": "3840x2160 single stream (BASELINE configs[2]): RollCorrection (roll_correction: of examples/config.yaml) -> Stabilizer "
                          "(radius 15, crop_n_zoom, border 30) -> AutoZoomCrop (640x360 out, as the reference hard-codes); device frames",
              "frames": nfr, "unit": UNIT,
              "stabilizer_only": run3(nfr, False, True, False),
              "roll_only": run3(nfr, True, False, False),
              "roll_then_stabilizer": run3(nfr, True, True, False),
              "roll_stabilizer_autozoom": run3(nfr, True, True, True),
              "autozoom_only": run3(nfr, False, False, True),
              "note": "AutoZoomCrop makes one device->host trip per frame for the contour / rectangle logic, like the reference (AutoZoomCrop.cpp:141-147)"}
        del clip3, rolled, out3, zoomed
        torch.cuda.empty_cache()

    # ============================================================ config 5: one long clip, temporal chunks + all-gather
    c5 = None
    if "config5" in parts:
        n_total, chunks = args.clip_frames, max(C5_CHUNKS, world)
        chunks -= chunks % world
        # lock-step analysis: as many equal, even-length chunks per rank as the clip divides into (<= 64): they advance together,
        # one launch per stage for all of them (VS_C5_LOCKSTEP=0: the chunk-by-chunk, frame-by-frame analysis of round 1)
        cpr = offline.lockstep_chunking(n_total, world) if os.environ.get("VS_C5_LOCKSTEP", "1") != "0" else None
        if cpr:
            chunks = world * cpr
        per_rank = chunks // world
        mine = list(range(rank * per_rank, (rank + 1) * per_rank))
        need = sum(offline.chunk_bounds(n_total, chunks, c)[1] + 2 for c in mine) * frame_bytes
        ring_frames = max(offline.chunk_bounds(n_total, chunks, c)[1] for c in mine)
        free = torch.cuda.mem_get_info(dev)[0]
        ok_mem, = allmax(0.0 if need + ring_frames * frame_bytes + (12 << 30) < free else 1.0)
        if ok_mem != 0.0:                                  # never drive the box out of memory: fall back to a shorter clip
            n_total = 4096
            if cpr:
                cpr = offline.lockstep_chunking(n_total, world)
                chunks = world * cpr if cpr else chunks
                per_rank = chunks // world
                mine = list(range(rank * per_rank, (rank + 1) * per_rank))
            ring_frames = max(offline.chunk_bounds(n_total, chunks, c)[1] for c in mine)
        gen = synthclip.DeviceClip(W, H, n_total, 5000, dev)
        chunk_frames = {}
        for c in mine:
            first, count = offline.chunk_bounds(n_total, chunks, c)
            if count > 0:
                chunk_frames[c] = gen.frames(first - offline.halo(first), first + count)
        out_ring = torch.empty((ring_frames, H, W, 3), dtype=torch.uint8, device=dev)
        st5 = vsb.Stabilizer(params, device=local_rank)
        n_lock = sum(1 for c in mine if offline.chunk_bounds(n_total, chunks, c)[0] >= 4) if cpr else 0
        sb5 = vsb.StabilizerBatch(params, n_lock, device=local_rank) if n_lock >= 2 else None
        ext5 = torch.cuda.ExternalStream(st5.stream, device=dev)
        torch.cuda.synchronize()
        full = None
        times = []
        for it in range(3):                                # one warm-up pass, two timed passes (max over ranks of each)
            barrier()
            l0 = st5.launch_count() + (sb5.launch_count() if sb5 else 0)
            t0 = torch.cuda.Event(enable_timing=True)
            t1 = torch.cuda.Event(enable_timing=True)
            t0.record(torch.cuda.current_stream(dev))
            full = offline.stabilize_rank_chunks(st5, chunk_frames, n_total, chunks, mine, out_ring, batch=sb5)
            st5.join()
            tail = torch.cuda.Event()
            tail.record(ext5)
            torch.cuda.current_stream(dev).wait_event(tail)
            t1.record(torch.cuda.current_stream(dev))
            st5.sync()
            torch.cuda.synchronize()
            barrier()
            if it:
                times.append(allmax(t0.elapsed_time(t1))[0])
            l5 = st5.launch_count() + (sb5.launch_count() if sb5 else 0) - l0
        # bit-equality with ONE handle streaming the whole clip (rank 0; blocks generated on the fly)
        equal = None
        if rank == 0:
            ref = vsb.Stabilizer(params, device=local_rank)
            blk = 250
            scratch = torch.empty((blk, H, W, 3), dtype=torch.uint8, device=dev)
            ring = torch.empty((blk + 40, H, W, 3), dtype=torch.uint8, device=dev)
            for a in range(0, n_total, blk):
                b = min(a + blk, n_total)
                gen.frames(a, b, out=scratch)
                torch.cuda.synchronize()
                ref.push_many_device(scratch.data_ptr(), frame_bytes, b - a, W, H, W * 3, ring.data_ptr(), W * 3, frame_bytes, borrow=False)
                ref.sync()
            nf, _ = ref.counts()
            assert nf == n_total - 1
            tr = np.empty((nf, 3), np.float32)
            for i in range(nf):
                tr[i] = list(ref.frame_record(i).transform)
            equal = bool(np.array_equal(tr.view(np.uint32), full.cpu().numpy().view(np.uint32)))
            assert equal, "config 5: stitched transforms differ from the single-handle run"
            del ref, scratch, ring
        ms5 = min(times)
        c5 = {"value": n_total / (ms5 * 1e-3), "unit": UNIT, "n_frames": n_total, "ms": ms5, "ms_all": times, "chunks": chunks,
              "chunks_per_gpu": per_rank, "n_gpus": world, "launches_per_gpu": int(l5),
              "collective": {"op": "all_gather_into_tensor (NCCL)" if world > 1 else "none (one rank)", "inside_timed_region": True,
                             "bytes_per_rank": int(per_rank * offline.stitch_index(n_total, chunks)[0] * 12),
                             "payload": "per-chunk transforms, 12 bytes per frame, device to device"},
              "transforms_bit_equal_to_single_handle_stream": equal,
              "lockstep_lanes_per_gpu": n_lock if sb5 else 0,
              "what": "analyse own chunks (in lock-step: one launch per stage for all of a rank's chunks; the clip's first chunk frame by "
                      "frame) -> one all-gather of the transforms -> rebuild path in the reference's float32 order once -> smooth + warp "
                      "own frames chunk by chunk; inputs resident in HBM, outputs into a device ring of one chunk"}
        del st5, sb5, chunk_frames, out_ring, gen
        torch.cuda.empty_cache()

    if rank == 0:
        head = c2 if world == 1 else c4
        if head is None:
            head = c4 or c2 or {"value": None, "ms_per_step": None, "launches": 0}
        cfg = workload_config(world)
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak" if world == 1 else "strong",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": cfg,
            "notes": {"api": (c2 or {}).get("api") if world == 1 else "vs_batch_push_device (borrowed device frames, one call per lock-step frame)",
                      "l2_policy": head.get("l2_policy"), "host_cpu_binding": numa},
            "e2e": e2e, "gpu_launches": int(head.get("launches", 0)), "clocks": clocks,
            "roofline": roof, "roofline_pyramid": roof_pyr, "config3": c3, "config4": c4, "config5": c5,
        }
        if c2 is not None:
            line["stage_us_per_launch_group"] = c2["stage_us_per_launch_group"]
            line["latency_us_push_to_frame"] = c2["latency_us_push_to_frame"]
            if world > 1:
                line["config2_per_gpu"] = c2
        if world == 1 and not args.no_cpu_baseline:
            if all_cpus:                                   # the CPU baseline gets every host core back (all threads)
                for tid in os.listdir("/proc/self/task"):
                    try:
                        os.sched_setaffinity(int(tid), all_cpus)
                    except OSError:
                        pass
            line["cpu_baseline"] = cpu_baseline_sample()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parts", default="config2,config4,e2e,roofline,config3,config5",
                    help="comma list of config2, config4, e2e, roofline, config3, config5 (profiling runs select one)")
    ap.add_argument("--clip-frames", type=int, default=C5_FRAMES)
    ap.add_argument("--c4-streams", type=int, default=C4_STREAMS, help="profiling only: total streams of config 4 (default 64)")
    ap.add_argument("--config2-at-all-n", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
