#!/usr/bin/env python3
"""BASELINE configs[3]: the no-SAT log-polar ImageSampler path (sample_logpolar + blur + interpolate)
beside the SAT log-rectilinear path, single frames at one resolution, per-kernel CUDA-event times.

    python tools/logpolar_bench.py [--workload 4k] [--steps 50]
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
ap.add_argument("--workload", default="4k")
ap.add_argument("--steps", type=int, default=50)
args = ap.parse_args()
fov = importlib.import_module("foveated-360-video_b200")
W, H = bench.WORKLOADS[args.workload]
ow, oh = bench.reduced(W), bench.reduced(H)
m = fov.OpenCLManager(0)
m.InitializeContext()
enc, dec, img = fov.SATEncoder(m), fov.SATDecoder(m), fov.ImageSampler(m)
frame = bench.synth_frame(W, H, 0)
src, sat = m.upload(frame), m.Buffer(12 * W * H)
red, blur, full = m.Buffer(4 * ow * oh), m.Buffer(4 * ow * oh), m.Buffer(4 * W * H)
m.memset(red, 0, 4 * ow * oh)
gaze = bench.gaze_trace(args.steps + 3, 1, seed=1)[:, 0]


def logpolar(i):
    cx, cy = float(gaze[i, 0]), float(gaze[i, 1])
    img.SampleFrameLogPolarGPU(red, ow, oh, 4 * ow, src, W, H, 4 * W, cx, cy)
    img.ApplyLogPolarGaussianBlur(blur, ow, oh, 4 * ow, red)
    img.InterpolateFrameLogPolarGPU(full, W, H, 4 * W, blur, ow, oh, 4 * ow, cx, cy)


def logrect(i):
    cx, cy = float(gaze[i, 0]), float(gaze[i, 1])
    enc.EncodeFrameGPU(sat, src, W, H, 4 * W)
    dec.SampleFrameRectGPU(red, ow, oh, 4 * ow, sat, W, H, cx, cy)
    dec.InterpolateFrameRectGPU(full, W, H, 4 * W, red, ow, oh, 4 * ow, cx, cy)


for name, fn in (("log-polar (ImageSampler)", logpolar), ("log-rect (SAT)", logrect)):
    for i in range(3):
        fn(i)
    m.profile_reset()
    m.profile(True)
    for i in range(args.steps):
        fn(3 + i)
    tot = m.profile_totals()
    m.profile(False)
    ms = sum(v[0] for v in tot.values()) / args.steps
    print("%s %s: %.4f ms/frame  %.0f fps   " % (args.workload, name, ms, 1e3 / ms)
          + "  ".join("%s %.4f" % (k, v[0] / v[1]) for k, v in sorted(tot.items())))
m.close()
