#!/usr/bin/env python3
"""Minimal driver for ncu: a few batched foveation steps at one resolution, nothing else.

    python tools/profile_step.py [--workload 8k] [--batch 2] [--steps 2]
"""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="8k")
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--pad-zero", action="store_true",
                help="FOV_OPT_REDUCED_PAD_ZERO: sample_rect writes whole pixels (buffer cleared below)")
args = ap.parse_args()

fov = importlib.import_module("foveated-360-video_b200")
W, H = bench.WORKLOADS[args.workload]
ow, oh = bench.reduced(W), bench.reduced(H)
m = fov.OpenCLManager(0)
m.InitializeContext()
B = args.batch
fb, sb, rb = 4 * W * H, 12 * W * H, 4 * ow * oh
frames = np.stack([bench.synth_frame(W, H, f) for f in range(B)])
src, sat, red, full = m.upload(frames), m.Buffer(B * sb), m.Buffer(B * rb), m.Buffer(B * fb)
m.memset(red, 0, B * rb)
m.set_option(m.OPT_REDUCED_PAD_ZERO, args.pad_zero)
gaze = bench.gaze_trace(args.steps, B, seed=1)
for i in range(args.steps):
    fov.FoveateFramesGPU(m, B, full, fb, red, rb, sat, sb, src, fb, W, H, 4 * W, ow, oh, gaze[i])
m.Finish()
print("ok", m.launch_count, "launches")
m.close()
