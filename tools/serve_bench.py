#!/usr/bin/env python3
"""BASELINE configs[4]: concurrent 4K streams with independent per-frame gaze
traces, stream s pinned to GPU s % G, the streams of one GPU foveated by one batched call per frame
time, host wall clock with the stream synchronised after every frame time (the same measurement
`bench.py` reports as `configs.serving_4k_streams`; this tool adds gaze-trace files and the encoder
surface).  Run under torchrun for G > 1; a single process plays rank 0 of `--world` GPUs.

    python tools/serve_bench.py [--streams 64] [--world 8] [--frames 300]
    python -m torch.distributed.run --nproc-per-node 8 tools/serve_bench.py
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


gaze_walk = bench.gaze_walk


ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=64)
ap.add_argument("--world", type=int, default=8, help="GPUs the streams are spread over")
ap.add_argument("--frames", type=int, default=300)
ap.add_argument("--workload", default="4k")
ap.add_argument("--gaze-file", default=None,
                help="GazeViewPoints trace (frame,<n>,forward,..,eye,..) replayed by every stream, "
                     "stream s starting at record 7*s; default: per-stream random walks")
ap.add_argument("--encoder-surface", action="store_true",
                help="also convert every reduced buffer to NV12 (what the server hands to NVENC)")
args = ap.parse_args()

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", args.world)) if "RANK" in os.environ else args.world
dist = None
if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", 1)) > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl" if os.environ.get("FOV_DIST", "nccl") == "nccl" else "gloo")
fov = importlib.import_module("foveated-360-video_b200")
mine = fov.sharding.streams_for_rank(args.streams, world, rank)
n = len(mine)
W, H = bench.WORKLOADS[args.workload]
ow, oh = bench.reduced(W), bench.reduced(H)
m = fov.OpenCLManager(int(os.environ.get("LOCAL_RANK", 0)))
m.InitializeContext()
fb, sb, rb = 4 * W * H, 12 * W * H, 4 * ow * oh
base = bench.synth_frame(W, H, 0)
frames = np.stack([np.roll(base, 131 * s, axis=1) for s in mine])
src, sat, red, full = m.upload(frames), m.Buffer(n * sb), m.Buffer(n * rb), m.Buffer(n * fb)
m.memset(red, 0, n * rb)
if args.gaze_file:
    rec = fov.GazeViewPoints(args.gaze_file).gaze_array()
    assert len(rec) > 0, "no records in " + args.gaze_file
    idx = np.arange(args.frames + 3)
    traces = np.stack([np.clip(rec[(idx + 7 * s) % len(rec)], 0.0, 1.0) for s in mine], axis=1)
else:
    traces = np.stack([gaze_walk(s, args.frames + 3) for s in mine], axis=1)  # [frame][stream][2]
conv = fov.VideoFrameConverter(m)
ny, nuv = (m.Buffer(n * ow * oh), m.Buffer(n * ow * oh // 2)) if args.encoder_surface else (None, None)


def frame_time(t):
    fov.FoveateFramesGPU(m, n, full, fb, red, rb, sat, sb, src, fb, W, H, 4 * W, ow, oh, traces[t])
    if args.encoder_surface:
        conv.RGB0ToNV12Frames(n, ny, ow * oh, ow, nuv, ow * oh // 2, ow, red, rb, 4 * ow, ow, oh)


import time  # noqa: E402

for t in range(3):
    frame_time(t)
m.Finish()
if dist:
    dist.barrier()
# Host wall clock; the stream is synchronised after every frame time (a server delivers each frame
# before it takes the next one), so launch gaps and the wait itself are inside the number.
t0 = time.perf_counter()
for t in range(args.frames):
    frame_time(3 + t)
    m.Finish()
sec = time.perf_counter() - t0
sec = fov.sharding.reduce_max_seconds(sec, dist, "cuda" if dist else None)
if rank == 0:
    ranks = int(os.environ.get("WORLD_SIZE", 1)) if dist else 1
    print(json.dumps({
        "workload": "%d concurrent %s streams over %d GPU(s), stream s -> GPU s %% %d, %d frames each, "
                    "%s gaze%s" % (args.streams, args.workload, world, world, args.frames,
                                   "trace-file" if args.gaze_file else "random-walk",
                                   ", NV12 encoder surface" if args.encoder_surface else ""),
        "gpus_measured": ranks, "streams_per_gpu": n,
        "frame_time_ms": round(sec / args.frames * 1e3, 4),
        "fps_per_stream": round(args.frames / sec, 1),
        "frames_per_s_measured_gpus": round(ranks * n * args.frames / sec, 1)}))
m.close()
if dist:
    dist.destroy_process_group()
