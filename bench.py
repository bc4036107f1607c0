#!/usr/bin/env python3
"""Benchmark of the foveation hot path: SAT log-rectilinear encode + decode, frames/s.

One "step" = one batch of B synthetic equirect RGB0 frames, each with its own gaze, through
EncodeFrameGPU -> SampleFrameRectGPU -> InterpolateFrameRectGPU (run_satlogrectilinear.cc:926-943)
via the C ABI of libfov360.so.

  value     frames/s with the frames already resident in HBM (CUDA events on the library's stream)
  e2e       frames/s through the same C-ABI calls with HOST buffers: every frame is copied
            host->device from pinned memory and its un-warped result device->host inside the
            timed region (video_server.cc:297-299,342-345 are blocking cl::copy; here the copies
            are pipelined over a few contexts = in-order queues)
  roofline  dominant kernel: algorithmic bytes / its CUDA-event duration vs the measured HBM peak
  cpu_baseline  the reference's own kernels (oracle/_ref) or the oracle port on the host cores

Multi-GPU (torchrun): frames are independent, every rank runs the same per-GPU batch on its own
device (weak scaling, no collective on the data path); timing is max over ranks.

`--impl reference` times the reference's CPU implementation of the same pipeline instead.
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (W, H)  - reduced size follows 16*ceil(dim/1.8/16) (run_satlogrectilinear.cc:113-114)
    "1080p": (1920, 1080),
    "4k": (3840, 1920),
    "8k": (7680, 3840),
}


def reduced(dim: int) -> int:
    import math

    return 16 * math.ceil(dim / 1.8 / 16)


def algorithmic_bytes(W, H, ow, oh):
    """per-frame algorithmic bytes of each stage."""
    sat = 16 * W * H
    sample = 12 * (ow + 1) * (oh + 1) + 4 * ow * oh
    interp = 4 * ow * oh + 4 * W * H
    return {"sat": sat, "sample": sample, "interp": interp, "total": sat + sample + interp}


def synth_frame(W, H, seed):
    """Natural-image-like RGB0 frame (low-passed noise + fine noise), padding byte 0."""
    rng = np.random.default_rng(seed)
    small = rng.integers(0, 256, size=(H // 32 + 2, W // 32 + 2, 3), dtype=np.uint8)
    img = np.repeat(np.repeat(small, 32, axis=0), 32, axis=1)[:H, :W].astype(np.int16)
    img += rng.integers(-24, 25, size=(H, W, 3), dtype=np.int16)
    out = np.zeros((H, W, 4), np.uint8)
    out[..., :3] = np.clip(img, 0, 255).astype(np.uint8)
    return out


def gaze_trace(steps, batch, seed=1):
    return np.random.default_rng(seed).random((steps, batch, 2)).astype(np.float32)


def peak_hbm_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons = index, False, [], set()
        self.max_mhz = None
        try:
            import pynvml as nv

            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)  # the first query is slow: pay it here
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's kernels on the host (oracle/_ref) or the oracle port
# ------------------------------------------------------------------------------------------------
def load_cpu_oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O

    try:
        if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libfovref.so")) or O.ref_available():
            orc, kind = O.Oracle("ref"), "reference"
        else:
            orc, kind = O.Oracle("port"), "port"
    except Exception:
        orc, kind = O.Oracle("port"), "port"
    # all host cores the process may use (torchrun exports OMP_NUM_THREADS=1 for its workers)
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    orc.set_threads(cores)
    return orc, kind


def cpu_pipeline_once(orc, frame, ow, oh, cx, cy, grid):
    H, W, _ = frame.shape
    sat = orc.sat_encode(frame)
    red = orc.sat_sample_rect(sat, ow, oh, cx, cy, grid=grid)
    return orc.sat_interpolate_rect(red, W, H, cx, cy)


def cpu_baseline(W, H, ow, oh, budget_s=12.0, max_frames=64):
    orc, kind = load_cpu_oracle()
    cores = orc.get_threads()
    frame = synth_frame(W, H, 1)
    grid = orc.sat_create_grid(ow, oh, W, H)
    gz = gaze_trace(max_frames + 1, 1, seed=2)[:, 0]
    cpu_pipeline_once(orc, frame, ow, oh, 0.5, 0.5, grid)  # warm-up (page faults, OpenMP pool)
    n, t0 = 0, time.perf_counter()
    while n < max_frames and (n == 0 or time.perf_counter() - t0 < budget_s):
        cpu_pipeline_once(orc, frame, ow, oh, float(gz[n, 0]), float(gz[n, 1]), grid)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "frames/s", "cores": cores, "kind": kind,
            "sample": "%d frame(s) of %dx%d -> %dx%d encode+sample+interpolate in %.1f s, "
                      "OpenMP over NDRange rows" % (n, W, H, ow, oh, dt)}


def run_reference_arm(args, W, H, ow, oh, rank):
    if rank != 0:
        return
    orc, kind = load_cpu_oracle()
    cores = orc.get_threads()
    frame = synth_frame(W, H, 1)
    grid = orc.sat_create_grid(ow, oh, W, H)
    gz = gaze_trace(args.steps + args.warmup, 1, seed=1)[:, 0]
    for i in range(args.warmup):
        cpu_pipeline_once(orc, frame, ow, oh, float(gz[i, 0]), float(gz[i, 1]), grid)
    t0 = time.perf_counter()
    for i in range(args.steps):
        j = args.warmup + i
        cpu_pipeline_once(orc, frame, ow, oh, float(gz[j, 0]), float(gz[j, 1]), grid)
    dt = time.perf_counter() - t0
    fps = args.steps / dt
    sample = "1 frame of %dx%d -> %dx%d per step (bounded sample of the batch), %d host threads" % (
        W, H, ow, oh, cores)
    line = {
        "impl": "reference", "metric": "frames/s SAT log-rect encode+decode", "value": fps,
        "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": args.workload_name(W, H, ow, oh), "frames_per_step": 1},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def pinned_array(m, nbytes):
    p = C.c_void_p()
    m._check(m.lib.fov_host_alloc(m.ctx, C.byref(p), nbytes))
    arr = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(p.value))
    return arr, p.value


def run_gpu_arm(args, W, H, ow, oh, rank, world, local_rank):
    import torch

    fov = importlib.import_module("foveated-360-video_b200")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the fov360 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    B, K, Wm = args.batch, args.steps, args.warmup
    m = fov.OpenCLManager(local_rank)
    m.InitializeContext()
    dec = fov.SATDecoder(m)
    dec.InitializeGrid(ow, oh, W, H)
    stream = torch.cuda.ExternalStream(m.stream, device=local_rank)

    fb, sb, rb = 4 * W * H, 12 * W * H, 4 * ow * oh
    frames = np.stack([synth_frame(W, H, 100 * rank + f) for f in range(B)])
    src = m.upload(frames)
    sat = m.Buffer(B * sb)
    red = m.Buffer(B * rb)
    full = m.Buffer(B * fb)
    m.memset(red, 0, B * rb)
    gaze = gaze_trace(K + Wm, B, seed=1 + rank)

    def step(i):
        fov.FoveateFramesGPU(m, B, full, fb, red, rb, sat, sb, src, fb, W, H, 4 * W, ow, oh, gaze[i])

    def barrier():
        m.Finish()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    # sanity: SAT corner of frame 0 == channel sums mod 2^32 (catches a dead kernel, costs nothing)
    step(0)
    last = m.copy_to_host(np.empty(3, np.uint32), sat, src_offset=sb - 12)
    want = frames[0][..., :3].reshape(-1, 3).astype(np.uint64).sum(axis=0) % (1 << 32)
    if [int(v) for v in last] != [int(v) for v in want]:
        raise SystemExit("bench.py: SAT checksum mismatch - refusing to time a wrong kernel")

    for i in range(Wm):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = m.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(K):
        step(Wm + i)
    e1.record(stream)
    barrier()
    clocks = sampler.result()
    launches = m.launch_count - launches0
    # whole-job rate = all frames of all ranks / the slowest rank's device time
    ms = fov.sharding.reduce_max_seconds(e0.elapsed_time(e1), dist, "cuda")
    frames_total = fov.sharding.reduce_sum_int(B * K, dist, "cuda")
    fps = frames_total / (ms * 1e-3)

    # ---- per-kernel pass (same K steps, every launch bracketed by CUDA events on the stream) ----
    m.profile_reset()
    m.profile(True)
    for i in range(K):
        step(Wm + i)
    totals = m.profile_totals()
    m.profile(False)
    ab = algorithmic_bytes(W, H, ow, oh)
    kernel_bytes = {  # algorithmic bytes per FRAME attributed to each kernel
        "sat_scan": ab["sat"], "sat_onepass": ab["sat"], "sat_reduce": 4 * W * H, "sat_carry": 0,
        "sat_sample_rect": ab["sample"], "sat_interpolate_rect": ab["interp"],
    }
    peak, peak_src = peak_hbm_gbs()
    step_ms = sum(v[0] for v in totals.values()) / K
    kernels = {}
    for name, (tot, cnt) in sorted(totals.items()):
        per_launch_ms = tot / max(cnt, 1)
        gbs = kernel_bytes.get(name, 0) * B / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms else 0
        kernels[name] = {"ms_per_launch": round(per_launch_ms, 4), "launches": cnt,
                         "share": round(tot / K / step_ms, 4), "alg_gbs": round(gbs, 1)}
    dom = max(totals, key=lambda k: totals[k][0])
    traffic = None
    try:  # measured DRAM bytes per launch of the dominant kernel, from the committed ncu capture
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as fh:
            t = json.load(fh)[args.workload][str(B)][dom]
            traffic = t["dram_read"] + t["dram_write"]
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": kernels[dom]["alg_gbs"], "peak": peak,
        "unit": "GB/s", "frac": round(kernels[dom]["alg_gbs"] / peak, 4), "traffic": traffic,
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": kernel_bytes.get(dom, 0) * B,
        "pipeline": {"bytes_per_frame": ab["total"],
                     "achieved": round(ab["total"] * fps / world / 1e9, 1),
                     "frac": round(ab["total"] * fps / world / 1e9 / peak, 4)},
        "kernels": kernels,
    }

    # ---- end to end through the C ABI with host buffers ----------------------------------------
    e2e = run_e2e(args, fov, local_rank, W, H, ow, oh, frames, gaze, dist, world)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(W, H, ow, oh)

    if rank == 0:
        line = {
            "metric": "frames/s SAT log-rect encode+decode", "value": fps, "unit": "frames/s",
            "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": {"workload": args.workload_name(W, H, ow, oh), "frames_per_step_per_gpu": B,
                       "gaze": "per-frame uniform random [0,1]^2, new every step",
                       "l2": "inputs larger than L2: per step %.0f MB of frames + %.0f MB of SAT "
                             "stream through a 126 MB L2" % (B * fb / 1e6, B * sb / 1e6),
                       "parallelism": "frames sharded over %d GPU(s), no collective" % world},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    m.close()


def run_e2e(args, fov, device, W, H, ow, oh, frames, gaze, dist, world):
    """Frames start in pinned HOST memory; each is copied in, foveated, and its un-warped result
    copied back out - the offline runner's loop (run_satlogrectilinear.cc:915-949) with the blocking
    cl::copy calls replaced by stream-ordered copies on `depth` contexts so PCIe and HBM overlap."""
    B, K, Wm = args.batch, args.steps, args.warmup
    depth = max(1, min(args.e2e_depth, B))
    fb, sb, rb = 4 * W * H, 12 * W * H, 4 * ow * oh
    lanes = []
    for _ in range(depth):
        m = fov.OpenCLManager(device)
        m.InitializeContext()
        enc, dec = fov.SATEncoder(m), fov.SATDecoder(m)
        dec.InitializeGrid(ow, oh, W, H)
        lane = {"m": m, "enc": enc, "dec": dec, "src": m.Buffer(fb), "sat": m.Buffer(sb),
                "red": m.Buffer(rb), "full": m.Buffer(fb)}
        m.memset(lane["red"], 0, rb)
        lanes.append(lane)
    m0 = lanes[0]["m"]
    hin, _ = pinned_array(m0, B * fb)
    hout, _ = pinned_array(m0, B * fb)
    hin[:] = frames.reshape(-1)
    lib = m0.lib

    def step(i):
        for f in range(B):
            ln = lanes[f % depth]
            m = ln["m"]
            cx, cy = float(gaze[i, f, 0]), float(gaze[i, f, 1])
            m._check(lib.fov_memcpy_h2d_async(m.ctx, ln["src"].ptr, hin.ctypes.data + f * fb, fb))
            ln["enc"].EncodeFrameGPU(ln["sat"], ln["src"], W, H, 4 * W)
            ln["dec"].SampleFrameRectGPU(ln["red"], ow, oh, 4 * ow, ln["sat"], W, H, cx, cy)
            ln["dec"].InterpolateFrameRectGPU(ln["full"], W, H, 4 * W, ln["red"], ow, oh, 4 * ow,
                                              cx, cy)
            m._check(lib.fov_memcpy_d2h_async(m.ctx, hout.ctypes.data + f * fb, ln["full"].ptr, fb))

    def sync():
        for ln in lanes:
            ln["m"].Finish()
        if dist:
            dist.barrier()

    for i in range(min(Wm, 3)):
        step(i)
    sync()
    t0 = time.perf_counter()
    for i in range(K):
        step(Wm + i)
    sync()
    dt = fov.sharding.reduce_max_seconds(time.perf_counter() - t0, dist, "cuda")
    checksum = int(hout[: 4 * W].astype(np.uint32).sum())  # the result really is on the host
    for ln in lanes:
        for k in ("src", "sat", "red", "full"):
            ln[k].free()
    lib.fov_host_free(m0.ctx, hin.ctypes.data)
    lib.fov_host_free(m0.ctx, hout.ctypes.data)
    for ln in lanes:
        ln["m"].close()
    return {"value": world * B * K / dt, "unit": "frames/s", "h2d_bytes_per_step": B * fb,
            "d2h_bytes_per_step": B * fb, "pipeline_depth": depth, "host_checksum": checksum,
            "timing": "host wall clock around K steps, all streams synchronised on both sides"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="8k", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=16,
                    help="frames per step per GPU (SURVEY 8(d): cfg3 is a batch of 16 frames)")
    ap.add_argument("--e2e-depth", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    W, H = WORKLOADS[args.workload]
    ow, oh = reduced(W), reduced(H)
    args.workload_name = lambda W, H, ow, oh: (
        "%s equirect %dx%d -> %dx%d log-rect, SAT encode + sample_rect + interpolate_rect "
        "(BASELINE.json configs[%d])" % (args.workload, W, H, ow, oh,
                                        {"1080p": 0, "4k": 1, "8k": 2}[args.workload]))
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, W, H, ow, oh, rank)
        return
    run_gpu_arm(args, W, H, ow, oh, rank, world, local_rank)


if __name__ == "__main__":
    main()
