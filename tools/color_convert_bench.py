#!/usr/bin/env python3
"""Device timing of RGB0 -> YUV420P / NV12 on the reduced buffers of a batch (CUDA events inside
the library).

    python tools/color_convert_bench.py [--workload 8k] [--batch 16] [--steps 50]
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
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--steps", type=int, default=50)
args = ap.parse_args()

fov = importlib.import_module("foveated-360-video_b200")
W, H = bench.WORKLOADS[args.workload]
ow, oh = bench.reduced(W), bench.reduced(H)
m = fov.OpenCLManager(0)
m.InitializeContext()
conv = fov.VideoFrameConverter(m)
B = args.batch
rng = np.random.default_rng(0)
red = m.upload(rng.integers(0, 256, (B, oh, ow, 4), dtype=np.uint8))
y, u, v, uv = m.Buffer(B * ow * oh), m.Buffer(B * ow * oh // 4), m.Buffer(B * ow * oh // 4), \
    m.Buffer(B * ow * oh // 2)


def planar():
    conv.RGB0ToYUV420PFrames(B, y, ow * oh, ow, u, v, ow * oh // 4, ow // 2, red, 4 * ow * oh, 4 * ow,
                             ow, oh)


def nv12():
    conv.RGB0ToNV12Frames(B, y, ow * oh, ow, uv, ow * oh // 2, ow, red, 4 * ow * oh, 4 * ow, ow, oh)


for _ in range(3):
    planar()
    nv12()
m.profile_reset()
m.profile(True)
for _ in range(args.steps):
    planar()
    nv12()
tot = m.profile_totals()
m.profile(False)
peak, _ = bench.peak_hbm_gbs()
bytes_per_launch = B * ow * oh * 5.5  # 4 B/px read + 1.5 B/px written
for k, (ms, cnt) in sorted(tot.items()):
    t = ms / cnt
    print("%s %dx%d x%d: %.4f ms/launch, %.0f GB/s algorithmic (%.1f%% of the measured copy peak)" % (
        k, ow, oh, B, t, bytes_per_launch / t / 1e6, 100 * bytes_per_launch / t / 1e6 / peak))
m.close()
