#!/usr/bin/env python
"""bench.py — headline benchmark of the stabilization hot path (BASELINE.json metric:
"1080p stabilized frames/s at 1/2/4/8 B200; warp kernel % of HBM peak").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # reference CPU path (oracle over cv2)

Workload at N=1 (BASELINE configs[1]): one 1920x1080 live stream, smoothing radius 15.  A *step* is one
pass of the hot path over a batch of 64 consecutive frames of that stream (398 MB of input, larger than
the 126 MB L2, so no frame is served from cache between steps).  At N>1 every rank runs the same
workload on its own stream (config 4's sharding: independent streams, one per GPU, no collective).

  value   frames/s with the clip already resident in HBM (device-pointer API, frames borrowed in place)
  e2e     frames/s through the reference-shaped host API (vs_stabilizer_push: host frame in, host frame
          out; the H2D and D2H copies are inside the timed region)
  roofline  the warp kernel (the kernel the metric names): algorithmic bytes 2*3*W*H per frame divided
          by its average launch duration, measured with CUDA events on the library's stream
  cpu_baseline  the oracle (reference host logic over cv2) on this box's host cores, bounded sample
"""
from __future__ import annotations

import argparse
import json
import os

# A handle overlaps its stages on six CUDA streams (eight with the pipelined host path); the default of 8 hardware
# queues would alias them with the other streams of this process and add false dependencies.  Must be set before
# the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1920, 1080
FRAMES_PER_STEP = 64
SMOOTHING_RADIUS = 15
METRIC = "stabilized_1080p_frames_per_s"
UNIT = "frames/s"


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
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path: the oracle restatement of Stabilizer.cpp over
    the real OpenCV (cv2), all host threads.  Rank 0 only."""
    if rank != 0:
        return
    import cv2
    import numpy as np
    from oracle.stabilizer_ref import Parameters, StabilizerRef
    import synthclip as synth
    cores = os.cpu_count() or 1
    cv2.setNumThreads(cores)
    sample = 16                                    # frames per step (bounded sample of the 64-frame step)
    clip = synth.make_clip(W, H, 32, seed=2000)
    order = _pingpong(len(clip))
    st = StabilizerRef(Parameters(smoothingRadius=SMOOTHING_RADIUS), use_optimized=True)
    k = 0

    def step():
        nonlocal k
        for _ in range(sample):
            st.stabilize(clip[order[k % len(order)]])
            k += 1
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    fps = args.steps * sample / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "1920x1080 single live stream, smoothing radius 15 (BASELINE configs[1])",
                   "frames_per_step": sample, "note": "bounded sample of the 64-frame step"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x {sample} frames 1080p, oracle (Stabilizer.cpp host logic restated in "
                                   f"Python over cv2 {cv2.__version__}, IPP/AVX paths on, {cores} threads)"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(seconds_budget=12.0):
    import cv2
    from oracle.stabilizer_ref import Parameters, StabilizerRef
    import synthclip as synth
    cores = os.cpu_count() or 1
    cv2.setNumThreads(cores)
    clip = synth.make_clip(W, H, 32, seed=2000)
    order = _pingpong(len(clip))
    st = StabilizerRef(Parameters(smoothingRadius=SMOOTHING_RADIUS), use_optimized=True)
    for k in range(20):
        st.stabilize(clip[order[k % len(order)]])
    n, t0 = 0, time.perf_counter()
    while n < 2000 and time.perf_counter() - t0 < seconds_budget:
        st.stabilize(clip[order[(20 + n) % len(order)]])
        n += 1
    dt = time.perf_counter() - t0
    # the same path on ONE host thread (SURVEY.md section 8d asks for both), a shorter sample
    cv2.setNumThreads(1)
    st1 = StabilizerRef(Parameters(smoothingRadius=SMOOTHING_RADIUS), use_optimized=True)
    for k in range(5):
        st1.stabilize(clip[order[k % len(order)]])
    n1, t1 = 0, time.perf_counter()
    while n1 < 400 and time.perf_counter() - t1 < 4.0:
        st1.stabilize(clip[order[(5 + n1) % len(order)]])
        n1 += 1
    dt1 = time.perf_counter() - t1
    cv2.setNumThreads(cores)
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} frames 1080p (~{dt:.0f} s) after 20 warm-up frames, oracle (Stabilizer.cpp host logic restated over cv2 "
                      f"{cv2.__version__}, optimized paths on, {cores} threads), wall clock",
            "single_thread": {"value": n1 / dt1, "unit": UNIT, "cores": 1, "sample": f"{n1} frames (~{dt1:.0f} s), cv2.setNumThreads(1)"}}


def bind_to_gpu_numa_node(index: int):
    """Pin this process to the host cores of the NUMA node the GPU hangs off (sysfs; no numactl in the image), so the
    page-locked frame buffers allocated afterwards are first-touched on that node and the enqueue thread stays near the
    GPU.  Matters for the host-buffer (e2e) leg at N > 1, where every rank streams ~2 x 40 GB/s over PCIe."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]                               # 00000000:1b:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        before = os.sched_getaffinity(0)
        allowed = sorted(set(cpus) & set(before))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return {"node": node, "cpus": len(allowed), "_restore": sorted(before)}
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args, rank, world, local_rank):
    numa = bind_to_gpu_numa_node(local_rank)
    all_cpus = numa.pop("_restore") if numa else None
    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__
    __graft_entry__.build()
    import synthclip
    import video_stab_b200 as vsb
    from video_stab_b200._capi import lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # synthetic clip, resident in HBM: the 64 generated frames laid out in ping-pong order (126 frames = 784 MB > L2), so a
    # step is a run of consecutive frames and the sequence loops without a jump
    clip_h = synthclip.make_clip(W, H, FRAMES_PER_STEP, seed=2000 + rank)
    order = _pingpong(FRAMES_PER_STEP)
    clip_d = torch.from_numpy(clip_h).to(dev)
    seq_d = clip_d[torch.tensor(order, device=dev)].contiguous()
    out_d = torch.empty((FRAMES_PER_STEP, H, W, 3), dtype=torch.uint8, device=dev)
    frame_bytes = H * W * 3
    torch.cuda.synchronize()

    params = vsb.Parameters(smoothingRadius=SMOOTHING_RADIUS)
    st = vsb.Stabilizer(params, device=local_rank)
    ext = torch.cuda.ExternalStream(st.stream, device=dev)
    pos = 0

    def step():
        # one step = the stabilize() loop over the next 64 frames of the sequence, enqueued by one C-ABI call per
        # consecutive run (frames are read in place; outputs land in a 64-frame device ring)
        nonlocal pos
        done = 0
        while done < FRAMES_PER_STEP:
            a = pos % len(order)
            k = min(FRAMES_PER_STEP - done, len(order) - a)
            st.push_many_device(seq_d[a].data_ptr(), frame_bytes, k, W, H, W * 3, out_d.data_ptr(), W * 3, frame_bytes, borrow=True)
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

    # ---- stage breakdown: a second pass of the same step with per-stage CUDA events on the library's stream
    st.set_timing(True)
    for _ in range(2):
        step()
    stages = st.stage_times()
    st.set_timing(False)
    stage_us = {k: (v["ms"] / v["count"] * 1e3 if v["count"] else None) for k, v in stages.items()}

    # ---- e2e: host frames in / host frames out through the reference-shaped host API, copies inside the timed
    #      region.  Headline: vs_stabilizer_push_many (the stabilize() loop of the reference's file apps as one call,
    #      one step = one call over the step's 64 frames; copy-in, compute and copy-out overlap).  Also reported:
    #      the strictly synchronous per-frame vs_stabilizer_push (one frame in, one frame out per call).
    import ctypes as C
    pin_in = torch.from_numpy(clip_h).pin_memory()
    seq = torch.stack([pin_in[i] for i in order]).pin_memory()           # 126 frames in ping-pong order
    pin_outs = torch.empty((FRAMES_PER_STEP, H, W, 3), dtype=torch.uint8).pin_memory()
    ow, oh, produced = C.c_int(), C.c_int(), C.c_int()
    st2 = vsb.Stabilizer(params, device=local_rank)
    epos = 0

    def e2e_step():
        nonlocal epos
        a = epos % len(order)
        nfr = min(FRAMES_PER_STEP, len(order) - a)
        done = 0
        while done < FRAMES_PER_STEP:                                      # wrap around the ping-pong sequence
            k = min(nfr, FRAMES_PER_STEP - done)
            rc = lib.vs_stabilizer_push_many(st2._h, seq[a].data_ptr(), frame_bytes, k, W, H, W * 3, pin_outs[done].data_ptr(),
                                             W * 3, frame_bytes, C.byref(ow), C.byref(oh), C.byref(produced))
            assert rc == 0
            done += k
            a = (a + k) % len(order)
            nfr = len(order) - a
        epos += FRAMES_PER_STEP

    st3 = vsb.Stabilizer(params, device=local_rank)
    pin_out1 = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
    spos = 0

    def sync_step():
        nonlocal spos
        for _ in range(FRAMES_PER_STEP):
            i = order[spos % len(order)]
            rc = lib.vs_stabilizer_push(st3._h, pin_in[i].data_ptr(), W, H, W * 3, pin_out1.data_ptr(), W * 3, frame_bytes,
                                        C.byref(ow), C.byref(oh), C.byref(produced))
            assert rc == 0
            spos += 1

    e2e_steps = max(1, args.steps)
    e2e_s = float("nan")
    sync_s = float("nan")
    if not args.no_e2e:
        for _ in range(max(1, min(args.warmup, 3))):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        st2.sync()
        e2e_s = time.perf_counter() - t0
        barrier()
        sync_step()
        t0 = time.perf_counter()
        for _ in range(2):
            sync_step()
        st3.sync()
        sync_s = (time.perf_counter() - t0) / 2
        barrier()

    # ---- roofline of the warp kernel: batched launch (64 frames per launch), timed alone with CUDA events
    T = np.zeros((FRAMES_PER_STEP, 2, 3), np.float32)
    rng = np.random.default_rng(7)
    for i in range(FRAMES_PER_STEP):
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
    alg_bytes = 2 * 3 * W * H * FRAMES_PER_STEP
    peak, peak_src = _peaks()
    achieved = alg_bytes / (warp_ms * 1e-3) / 1e9

    # ---- pyramid build (north-star kernel 1): gray + both pyrDown levels of 64 frames in one lock-step launch pair,
    #      timed alone.  Algorithmic bytes per frame: 3*W*H read + 960*540*(1 + 1/4 + 1/16) written = 6 901 200.
    sb = vsb.StabilizerBatch(params, FRAMES_PER_STEP, device=local_rank)
    sb_ext = torch.cuda.ExternalStream(sb.stream, device=dev)
    ptrs = [clip_d[i].data_ptr() for i in range(FRAMES_PER_STEP)]
    for _ in range(3):
        sb.build_pyramids(ptrs, W, H, W * 3)
    sb.sync()
    def timed_levels(parts):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(sb_ext)
        for _ in range(reps):
            sb.build_levels(ptrs, W, H, W * 3, parts)
        b.record(sb_ext)
        sb.sync()
        return a.elapsed_time(b) / reps
    pyr_ms = timed_levels(3)
    gray_ms = timed_levels(1)
    down_ms = timed_levels(2)
    pyr_bytes = (3 * W * H + 960 * 540 + 480 * 270 + 240 * 135) * FRAMES_PER_STEP
    gray_bytes = (3 * W * H + 960 * 540) * FRAMES_PER_STEP
    pyr_gbs = pyr_bytes / (pyr_ms * 1e-3) / 1e9
    gray_gbs = gray_bytes / (gray_ms * 1e-3) / 1e9
    del sb

    # max over ranks
    t = torch.tensor([ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_max = float(t[0]), float(t[1])
    total_frames = args.steps * FRAMES_PER_STEP * world
    value = total_frames / (ms_max * 1e-3)
    e2e_value = e2e_steps * FRAMES_PER_STEP * world / e2e_max

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "1920x1080 single live stream per GPU, smoothing radius 15 (BASELINE configs[1]); "
                                   "at N>1 one independent stream per GPU (configs[3] sharding, no collective)",
                       "frames_per_step": FRAMES_PER_STEP, "gftt": "200 pts every 2nd frame", "lk": "15x15, 3 levels",
                       "l2_policy": "inputs larger than L2 (126-frame sequence = 784 MB, frames read in place)",
                       "api": "vs_stabilizer_push_many_device (borrowed device frames; the per-frame stabilize() loop runs inside the "
                              "library, one call per run of consecutive frames)",
                       "host_numa_binding": numa},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": frame_bytes * FRAMES_PER_STEP,
                    "d2h_bytes_per_step": frame_bytes * FRAMES_PER_STEP,
                    "api": "vs_stabilizer_push_many: 64 page-locked host frames in, 64 host frames out per call "
                           "(copy-in / compute / copy-out streams overlap inside the call)",
                    "steps": e2e_steps,
                    "sync_per_frame_push": {"value": FRAMES_PER_STEP / sync_s, "unit": UNIT,
                                            "api": "vs_stabilizer_push: one host frame in, one host frame out per call"}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "stage_us_per_launch_group": stage_us,
            "roofline": {"bound": "hbm", "kernel": "k_warp_tma (cv::warpAffine, 64 frames per launch, timed alone)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                         "ms_per_launch": warp_ms, "frac_of_nominal_8000": achieved / 8000.0},
            "roofline_pyramid": {"bound": "hbm", "kernel": "k_gray_half + k_pyrdown2 (resize + gray + 2 pyrDown levels, 64 frames "
                                                           "per launch pair, timed alone)",
                                 "achieved": pyr_gbs, "peak": peak, "unit": "GB/s", "frac": pyr_gbs / peak,
                                 "algorithmic_bytes_per_launch": pyr_bytes, "ms_per_launch": pyr_ms,
                                 "level0": {"kernel": "k_gray_half alone (resize + gray: the HBM-bound half, 3*W*H read + 518 400 written "
                                                      "per frame)", "achieved": gray_gbs, "frac": gray_gbs / peak,
                                            "algorithmic_bytes_per_launch": gray_bytes, "ms_per_launch": gray_ms},
                                 "pyrdown": {"kernel": "k_pyrdown2 alone (level 0 -> levels 1, 2: 0.69 MB per frame, L2-resident, "
                                                       "latency / issue bound)", "ms_per_launch": down_ms}},
        }
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
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
